"""Transition store for the A2C update (reference wurm/rl/trajectory_store.py:4-89: same methods and properties).

`values`, `log_probs` and `entropies` carry autograd history, so they are kept as the tensors the policy produced and
stacked on demand, exactly like the reference.  The fields without gradient -- `states`, `actions`, `rewards`, `dones` --
can live in a preallocated device RING instead (`TrajectoryStore(capacity=T)`): one (T, num_envs, ...) buffer per field,
`append` writing time slice t in place, the properties returning views of the first `len(store)` slices (no torch.stack,
no per-update allocation).  With `state_slot()` the env's step kernel renders its observation STRAIGHT into the ring
(`env.step(a, obs_out=store.state_slot(shape))`): the observation is written to HBM once,
where the learner will read it.
"""
import torch

_FIELDS = ('state', 'action', 'log_prob', 'reward', 'value', 'done', 'entropy', 'hidden_state')
_RING_FIELDS = ('state', 'action', 'reward', 'done')


class TrajectoryStore(object):
    """Each property returns a tensor of shape (num_steps, num_envs, ...)."""

    def __init__(self, capacity: int = None):
        self.capacity = capacity
        self._ring = {}
        self.clear()

    def state_slot(self, shape, dtype=torch.float32, device='cuda'):
        """The (num_envs, ...) time slice of the `states` ring that the NEXT append(state=...) fills: hand it to the env as
        `obs_out=` so that the step kernel writes the observation there, then pass the returned observation to append()."""
        if self.capacity is None:
            raise RuntimeError('state_slot() needs a TrajectoryStore(capacity=T)')
        device = torch.empty(0, device=device).device             # 'cuda' -> 'cuda:<current>': what the tensors will report
        return self._slot('state', len(self._lists['state']), torch.Size(shape), dtype, device)

    def _slot(self, name, t, shape, dtype, device):
        if t >= self.capacity:
            raise RuntimeError(f'TrajectoryStore(capacity={self.capacity}) is full: clear() it after every update')
        buf = self._ring.get(name)
        if buf is None or buf.shape[1:] != shape or buf.dtype != dtype or buf.device != device:
            buf = torch.empty((self.capacity,) + tuple(shape), dtype=dtype, device=device)
            self._ring[name] = buf
        return buf[t]

    def append(self, state=None, action=None, log_prob=None, reward=None, value=None, done=None, entropy=None,
               hidden_state=None):
        """Adds a transition; each argument is a (num_envs, 1) tensor (entropy: a scalar), omitted ones are skipped."""
        given = dict(state=state, action=action, log_prob=log_prob, reward=reward, value=value, done=done, entropy=entropy,
                     hidden_state=hidden_state)
        for name in _FIELDS:
            x = given[name]
            if x is None:
                continue
            if self.capacity is not None and name in _RING_FIELDS and not x.requires_grad:
                slot = self._slot(name, len(self._lists[name]), x.shape, x.dtype, x.device)
                if slot.data_ptr() != x.data_ptr():          # (already in place when the kernel rendered into state_slot())
                    slot.copy_(x)
                x = slot
            self._lists[name].append(x)

    def clear(self):
        self._lists = {name: [] for name in _FIELDS}

    def __len__(self):
        return max(len(v) for v in self._lists.values())

    def _stack(self, name):
        items = self._lists[name]
        buf = self._ring.get(name)
        if buf is not None and items and all(x.data_ptr() == buf[t].data_ptr() for t, x in enumerate(items)):
            return buf[:len(items)]                          # the ring already IS the stacked tensor
        return torch.stack(items)

    states = property(lambda self: self._stack('state'))
    actions = property(lambda self: self._stack('action'))
    log_probs = property(lambda self: self._stack('log_prob'))
    rewards = property(lambda self: self._stack('reward'))
    values = property(lambda self: self._stack('value'))
    dones = property(lambda self: self._stack('done'))
    entropies = property(lambda self: self._stack('entropy'))
    hidden_state = property(lambda self: self._stack('hidden_state'))
