"""Stepping an env from HOST action buffers with results delivered to HOST buffers.

The env state lives on the GPU; what crosses PCIe every step is the actions (host -> device) and the
per-env results a host-side consumer needs (device -> host).  `HostStepper` double-buffers both directions in
pinned memory and puts each direction on its own copy stream, so the copies of step t+1 / t-1 overlap the
kernels of step t:

    stepper = HostStepper(env)
    tickets = []
    for actions in host_action_batches:          # pinned CPU tensors (or dicts of them for MultiSnake)
        tickets.append(stepper.submit(actions))  # H2D copy, fused step+reset launch, D2H copies: all async
        if len(tickets) > stepper.depth:
            result = tickets.pop(0).wait()       # .reward / .done are host tensors, .obs stays on the device

Bytes per env-step.  The reference's types cost 13: int64 actions up (8), fp32 reward + done flag down (5) -- at
2^20 SingleSnake envs and 0.3 ms per step that is 40 GB/s per GPU, and eight ranks saturate the host (round 1: e2e
scaling efficiency 0.48 at 8 GPUs).  So:
  * actions are uploaded in the dtype the caller submits; the kernels also take uint8 (`action_bytes` 1), so a host
    policy that emits uint8 actions uploads ONE byte per env-step;
  * SingleSnake results come back as ONE packed byte per env (`compact=True`, the default: bit 0 done, bit 1 self
    collision, bit 2 edge collision, bits 3-4 reward; include/wurm_b200.h WURM_PACKED_*); `ticket.done`, `.reward`,
    `.self_collision`, `.edge_collision` are decoded from it on the host on first access (`compact=False` copies the
    fp32 reward and the bool done flag instead, 5 bytes);
  * `return_actions=True` also brings SingleSnake's sanitised actions back (the reference sanitises the DEVICE tensor
    it is given in place, single_snake.py:222, and so does the kernel with the uploaded copy);
  * `return_obs=True` also copies the step's observation into pinned host memory (`ticket.obs_host`), for a policy that
    runs on the host -- by far the largest item (300 B per env for `partial_2`), PCIe-bound by construction.

Calling `env.step(host_tensor)` directly also works (the env copies on the compute stream); it is the
un-pipelined form of the same thing.
"""
import torch




class Ticket(object):
    """Result of one submitted step; host tensors are valid after wait().  Fields listed in `derive` are decoded from
    the packed result byte on first access."""

    def __init__(self, event, fields, derive=None):
        self._event = event
        self._fields = fields
        self._derive = derive or {}

    def wait(self):
        self._event.synchronize()
        return self

    def __getattr__(self, name):
        fields = self.__dict__['_fields']
        if name in fields:
            return fields[name]
        derive = self.__dict__['_derive']
        if name in derive:
            fields[name] = derive[name](fields['packed'])
            return fields[name]
        raise AttributeError(name)


_DECODE = {
    'done': lambda p: (p & _lib_consts()[0]).bool().unsqueeze(-1),
    'self_collision': lambda p: (p & _lib_consts()[1]).bool(),
    'edge_collision': lambda p: (p & _lib_consts()[2]).bool(),
    'reward': lambda p: ((p >> _lib_consts()[3]) & 3).float().unsqueeze(-1),
}


def _lib_consts():
    from . import _lib
    return _lib.PACKED_DONE, _lib.PACKED_SELF, _lib.PACKED_EDGE, _lib.PACKED_REWARD_SHIFT


class HostStepper(object):
    def __init__(self, env, depth: int = 2, auto_reset: bool = True, return_actions: bool = False, compact: bool = True,
                 return_obs: bool = False):
        self.env = env
        self.depth = depth
        self.auto_reset = auto_reset
        self.return_actions = return_actions
        self.return_obs = return_obs
        self.multi = hasattr(env, 'num_snakes')
        self.compact = compact and not self.multi and hasattr(env, '_hints')     # the packed byte is SingleSnake's
        self.device = torch.device(env.device) if not isinstance(env.device, torch.device) else env.device
        if self.device.index is None:
            self.device = torch.device('cuda', torch.cuda.current_device())
        self.h2d = torch.cuda.Stream(self.device)
        self.d2h = torch.cuda.Stream(self.device)
        self._slot = 0
        self._dev_actions = [None] * (depth + 1)
        self._dev_packed = [None] * (depth + 1)
        self._host_out = [None] * (depth + 1)
        self._slot_free = [None] * (depth + 1)      # event: the slot's previous D2H copies are done
        self.h2d_bytes_per_step = 0
        self.d2h_bytes_per_step = 0
        self._action_dtype = None

    def describe(self):
        what = 'one packed result byte per env' if self.compact else 'fp32 rewards + done flags'
        return (f'HostStepper: pinned double-buffered H2D actions ({self._action_dtype}) / D2H {what}'
                f"{' + observations' if self.return_obs else ''}{' + sanitised actions' if self.return_actions else ''}"
                ' on copy streams around the fused step+reset launch')

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
                self._action_dtype = str(next(iter(actions.values())).dtype).replace('torch.', '')
            else:
                self._dev_actions[slot] = torch.empty_like(actions, device=self.device)
                act = actions.numel() * actions.element_size()
                self._action_dtype = str(actions.dtype).replace('torch.', '')
                if self.compact:
                    self._dev_packed[slot] = torch.empty(N, dtype=torch.uint8, device=self.device)
                    self._host_out[slot] = dict(packed=torch.empty(N, dtype=torch.uint8, **pin))
                    self.d2h_bytes_per_step = N
                else:
                    self._host_out[slot] = dict(reward=torch.empty((N, 1), dtype=torch.float32, **pin),
                                                done=torch.empty((N, 1), dtype=torch.bool, **pin))
                    self.d2h_bytes_per_step = N * 5
                if self.return_actions:
                    self._host_out[slot]['actions'] = torch.empty_like(actions, **pin)
                    self.d2h_bytes_per_step += act
                self.h2d_bytes_per_step = act
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
            kw = dict(packed_out=self._dev_packed[slot]) if self.compact else {}
            obs, reward_t, done_t, info = self.env.step(dev_actions, auto_reset=True, **kw) if fused else self.env.step(dev_actions, **kw)
            env_done = done_t
        stepped = torch.cuda.Event()
        stepped.record(compute)
        if self.auto_reset and not fused:
            self.env.reset(env_done, return_observations=False)
        self.d2h.wait_event(stepped)
        with torch.cuda.stream(self.d2h):
            if self.multi:
                host_out['reward'].copy_(reward_t, non_blocking=True)
                host_out['done'].copy_(done_t, non_blocking=True)
                host_out['all_done'].copy_(env_done, non_blocking=True)
            else:
                if self.compact:
                    host_out['packed'].copy_(self._dev_packed[slot], non_blocking=True)
                else:
                    host_out['reward'].copy_(reward_t, non_blocking=True)
                    host_out['done'].copy_(done_t, non_blocking=True)
                if self.return_actions:
                    host_out['actions'].copy_(dev_actions, non_blocking=True)  # sanitised in place by the kernel
            if self.return_obs:
                obs_list = list(obs.values()) if isinstance(obs, dict) else [obs]
                if 'obs_host' not in host_out:
                    host_out['obs_host'] = [torch.empty(o.shape, dtype=o.dtype, pin_memory=True) for o in obs_list]
                    self.d2h_bytes_per_step += sum(o.numel() * o.element_size() for o in obs_list)
                for dst, o in zip(host_out['obs_host'], obs_list):
                    dst.copy_(o, non_blocking=True)
                    o.record_stream(self.d2h)
            done_ev = torch.cuda.Event()
            done_ev.record(self.d2h)
        for t in (reward_t, done_t, env_done):
            if t is not None:
                t.record_stream(self.d2h)            # keep the caching allocator from recycling them early
        self._slot_free[slot] = done_ev
        fields = dict(host_out)
        if self.return_obs and not isinstance(obs, dict):
            fields['obs_host'] = host_out['obs_host'][0]
        fields['obs'] = obs                          # stays on the device: it is the policy's input
        return Ticket(done_ev, fields, _DECODE if self.compact else None)
