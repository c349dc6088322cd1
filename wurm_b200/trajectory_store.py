"""Transition store for the A2C update (reference wurm/rl/trajectory_store.py:4-89: same methods and properties).

Host-side bookkeeping only: `values`, `log_probs` and `entropies` carry autograd history, so they are
kept as the tensors the policy produced and stacked on demand, exactly like the reference.  `rewards`
and `dones` are what `wurm_b200.rl.returns` scans in one launch.
"""
import torch

_FIELDS = ('state', 'action', 'log_prob', 'reward', 'value', 'done', 'entropy', 'hidden_state')


class TrajectoryStore(object):
    """Each property returns a tensor of shape (num_steps, num_envs, ...)."""

    def __init__(self):
        self.clear()

    def append(self, state=None, action=None, log_prob=None, reward=None, value=None, done=None, entropy=None,
               hidden_state=None):
        """Adds a transition; each argument is a (num_envs, 1) tensor (entropy: a scalar), omitted ones are skipped."""
        given = dict(state=state, action=action, log_prob=log_prob, reward=reward, value=value, done=done, entropy=entropy,
                     hidden_state=hidden_state)
        for name in _FIELDS:
            if given[name] is not None:
                self._lists[name].append(given[name])

    def clear(self):
        self._lists = {name: [] for name in _FIELDS}

    def __len__(self):
        return max(len(v) for v in self._lists.values())

    def _stack(self, name):
        return torch.stack(self._lists[name])

    states = property(lambda self: self._stack('state'))
    actions = property(lambda self: self._stack('action'))
    log_probs = property(lambda self: self._stack('log_prob'))
    rewards = property(lambda self: self._stack('reward'))
    values = property(lambda self: self._stack('value'))
    dones = property(lambda self: self._stack('done'))
    entropies = property(lambda self: self._stack('entropy'))
    hidden_state = property(lambda self: self._stack('hidden_state'))
