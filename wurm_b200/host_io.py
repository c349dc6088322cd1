"""Stepping an env from HOST action buffers with results delivered to HOST buffers.

The env state lives on the GPU; what crosses PCIe every step is the actions (host -> device) and the
per-env results a host-side consumer needs (device -> host: rewards and done flags; on request
(`return_actions=True`) also SingleSnake's sanitised actions -- the reference sanitises the DEVICE tensor it
is given in place, single_snake.py:222, and so does the kernel with the uploaded copy; bringing that back
costs 8 of 13.6 MB per step at 2^20 envs and makes two GPUs behind one PCIe switch copy-bound).
`HostStepper` double-buffers both directions in pinned memory and puts each direction on its own copy
stream, so the copies of step t+1 / t-1 overlap the kernels of step t:

    stepper = HostStepper(env)
    tickets = []
    for actions in host_action_batches:          # pinned CPU tensors (or dicts of them for MultiSnake)
        tickets.append(stepper.submit(actions))  # H2D copy, step kernel, reset kernel, D2H copies: all async
        if len(tickets) > stepper.depth:
            result = tickets.pop(0).wait()       # .reward / .done (/.actions) are pinned host tensors, .obs stays on the device

Calling `env.step(host_tensor)` directly also works (the env copies on the compute stream); it is the
un-pipelined form of the same thing.
"""
import torch




class Ticket(object):
    """Result of one submitted step; host tensors are valid after wait()."""

    def __init__(self, event, fields):
        self._event = event
        self._fields = fields

    def wait(self):
        self._event.synchronize()
        return self

    def __getattr__(self, name):
        try:
            return self._fields[name]
        except KeyError:
            raise AttributeError(name)


class HostStepper(object):
    def __init__(self, env, depth: int = 2, auto_reset: bool = True, return_actions: bool = False):
        self.env = env
        self.depth = depth
        self.auto_reset = auto_reset
        self.return_actions = return_actions
        self.multi = hasattr(env, 'num_snakes')
        self.device = torch.device(env.device) if not isinstance(env.device, torch.device) else env.device
        if self.device.index is None:
            self.device = torch.device('cuda', torch.cuda.current_device())
        self.h2d = torch.cuda.Stream(self.device)
        self.d2h = torch.cuda.Stream(self.device)
        self._slot = 0
        self._dev_actions = [None] * (depth + 1)
        self._host_out = [None] * (depth + 1)
        self._slot_free = [None] * (depth + 1)      # event: the slot's previous D2H copies are done
        self.h2d_bytes_per_step = 0
        self.d2h_bytes_per_step = 0

    def _buffers(self, slot, actions):
        if self._dev_actions[slot] is None:
            N = self.env.num_envs
            pin = dict(pin_memory=True)
            if self.multi:
                K = self.env.num_snakes
                self._dev_actions[slot] = {a: torch.empty_like(t, device=self.device) for a, t in actions.items()}
                self._host_out[slot] = dict(reward=torch.empty((N, K), dtype=torch.float32, **pin),
                                            done=torch.empty((N, K), dtype=torch.bool, **pin),
                                            all_done=torch.empty(N, dtype=torch.bool, **pin))
                self.h2d_bytes_per_step = sum(t.numel() * t.element_size() for t in actions.values())
                self.d2h_bytes_per_step = N * K * 5 + N
            else:
                self._dev_actions[slot] = torch.empty_like(actions, device=self.device)
                self._host_out[slot] = dict(reward=torch.empty((N, 1), dtype=torch.float32, **pin),
                                            done=torch.empty((N, 1), dtype=torch.bool, **pin))
                act = actions.numel() * actions.element_size()
                if self.return_actions:
                    self._host_out[slot]['actions'] = torch.empty_like(actions, **pin)
                self.h2d_bytes_per_step = act
                self.d2h_bytes_per_step = N * 5 + (act if self.return_actions else 0)
        return self._dev_actions[slot], self._host_out[slot]

    def submit(self, actions) -> Ticket:
        """Enqueues H2D(actions) -> env.step -> [env.reset(done)] -> D2H(results); returns immediately."""
        slot = self._slot
        self._slot = (self._slot + 1) % (self.depth + 1)
        dev_actions, host_out = self._buffers(slot, actions)
        compute = torch.cuda.current_stream(self.device)
        if self._slot_free[slot] is not None:
            self.h2d.wait_event(self._slot_free[slot])       # the slot's host/device buffers are being reused
        with torch.cuda.stream(self.h2d):
            if self.multi:
                for a, t in actions.items():
                    dev_actions[a].copy_(t, non_blocking=True)
            else:
                dev_actions.copy_(actions, non_blocking=True)
            uploaded = torch.cuda.Event()
            uploaded.record(self.h2d)
        compute.wait_event(uploaded)
        fused = False
        if self.multi:
            fused = self.auto_reset                  # MultiSnake: the fused launch never loses (profiles/r01_final_*.json)
            obs, rewards, dones, info = self.env.step(dev_actions, auto_reset=fused)
            reward_t = self.env.rewards.view(self.env.num_envs, self.env.num_snakes)
            done_t = self.env._step_dones
            env_done = dones['__all__']
        else:
            # one fused step+reset launch: the warp re-creates its finished envs while their sectors are still in L2;
            # far ahead where the pair is launch-bound, a little ahead at 2^20 envs (profiles/r01_final_*.json)
            fused = self.auto_reset and getattr(self.env, 'supports_fused_reset', False)
            obs, reward_t, done_t, info = self.env.step(dev_actions, auto_reset=True) if fused else self.env.step(dev_actions)
            env_done = done_t
        stepped = torch.cuda.Event()
        stepped.record(compute)
        if self.auto_reset and not fused:
            self.env.reset(env_done, return_observations=False)
        self.d2h.wait_event(stepped)
        with torch.cuda.stream(self.d2h):
            host_out['reward'].copy_(reward_t, non_blocking=True)
            if self.multi:
                host_out['done'].copy_(done_t, non_blocking=True)
                host_out['all_done'].copy_(env_done, non_blocking=True)
            else:
                host_out['done'].copy_(done_t, non_blocking=True)
                if self.return_actions:
                    host_out['actions'].copy_(dev_actions, non_blocking=True)  # sanitised in place by the kernel
            done_ev = torch.cuda.Event()
            done_ev.record(self.d2h)
        for t in (reward_t, done_t, env_done):
            if t is not None:
                t.record_stream(self.d2h)            # keep the caching allocator from recycling them early
        self._slot_free[slot] = done_ev
        fields = dict(host_out)
        fields['obs'] = obs                          # stays on the device: it is the policy's input
        return Ticket(done_ev, fields)
