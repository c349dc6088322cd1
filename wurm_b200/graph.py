"""CUDA-graph replay of the step/reset pair for launch-bound batch sizes.

At the reference's README scale (512 envs, BASELINE config 1) one env step moves ~1 MB: the GPU work
takes a few microseconds and the cost is launching it (two kernel launches, a dozen tensor
allocations, the Python between them).  `GraphedStepper` captures

    obs, reward, done, info = env.step(actions);  env.reset(done, return_observations=False)

once into a CUDA graph and replays it: one `cudaGraphLaunch` per env step (for SingleSnake the pair is
additionally fused into a single kernel, `wurm_single_step_reset`).  The kernels read their
Philox call counter as `step + *step_dev` (include/wurm_b200.h); the graph bumps a device word private to the
stepper after every replay (and the env's host counter moves with it), so replays draw fresh random numbers, a
graphed rollout is bit-identical to the same rollout stepped call by call, and direct `env.step()` / `env.reset()`
calls may be interleaved with replays without ever reusing a (seed, counter, env) triple.  Construction warms the
kernels up with real steps but restores the env (state, statistics, counters) afterwards.

    stepper = GraphedStepper(env, actions)      # `actions`: device tensor (dict of tensors for MultiSnake)
    for t in range(T):
        actions.copy_(policy(stepper.obs))      # fill the static input in place
        obs, reward, done, info = stepper.step()    # static outputs: overwritten by the next replay

With `host_io=True` the graph also carries the PCIe copies: actions are read from a static pinned host buffer
(`stepper.host_actions`) at the head of the graph, rewards and done flags land in pinned host buffers
(`stepper.host_reward`, `stepper.host_done`, for MultiSnake also `stepper.host_all_done`) at its tail -- a host-side
policy pays one graph launch and one stream synchronisation per env step instead of three copies, two launches
and the stream choreography between them:

    stepper = GraphedStepper(env, actions, host_io=True)
    for t in range(T):
        stepper.host_actions.copy_(host_policy(...))    # dict of pinned tensors for MultiSnake
        stepper.step_host()                             # replay + synchronise
        reward, done = stepper.host_reward, stepper.host_done
"""
import torch


class GraphedStepper(object):
    def __init__(self, env, actions, auto_reset: bool = True, warmup: int = 2, host_io: bool = False):
        self.env = env
        self.actions = actions
        self.multi = hasattr(env, 'num_snakes')
        self.auto_reset = auto_reset
        self.host_io = host_io
        dev = env._dev
        self._dev = dev
        if host_io:
            pin = dict(pin_memory=True)
            N = env.num_envs
            if self.multi:
                K = env.num_snakes
                self.host_actions = {a: torch.empty(t.shape, dtype=t.dtype, **pin) for a, t in actions.items()}
                for a, t in actions.items():
                    self.host_actions[a].copy_(t)
                self.host_reward = torch.empty((N, K), dtype=torch.float32, **pin)
                self.host_done = torch.empty((N, K), dtype=torch.bool, **pin)
                self.host_all_done = torch.empty(N, dtype=torch.bool, **pin)
            else:
                self.host_actions = torch.empty(actions.shape, dtype=actions.dtype, **pin)
                self.host_actions.copy_(actions)
                self.host_reward = torch.empty((N, 1), dtype=torch.float32, **pin)
                self.host_done = torch.empty((N, 1), dtype=torch.bool, **pin)
        # Warm-up outside the capture (first-launch work: function attributes, allocator pools) runs REAL steps, so the
        # env's state, statistics, hints and call counter are snapshotted before and restored after: constructing a
        # GraphedStepper leaves the env exactly as it found it.
        if hasattr(env, '_snapshot_names'):          # (compact resident state: the records ARE the state)
            names = env._snapshot_names()
        else:
            names = ('envs', 'done', '_hints', '_stats', '_status')
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            if warmup > 0:
                saved = {n: getattr(env, n).clone() for n in names if hasattr(env, n)}
                saved_draws = env._draws
                for _ in range(warmup):
                    self._one(capturing=False)
                for n, t in saved.items():
                    getattr(env, n).copy_(t)
                env._draws = saved_draws
                if hasattr(env, '_adopt_state'):
                    env._adopt_state()              # the restored hints describe the restored state
                if hasattr(env, '_sync_shadow'):
                    env._shadow_ok = False          # ... the shadow records (dense MultiSnake) describe the warm-up's
                del saved
            if hasattr(env, '_sync_shadow'):
                env._sync_shadow()                  # records derived now: the captured launch is the record-loading one
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        # Call counters.  The captured launches carry the host counter of the capture moment baked in (`_base`) and add
        # a device word to it -- PRIVATE to this stepper, so direct env.step()/reset() calls (which add the env's own,
        # always-zero word) are unaffected by replays.  Every replay leaves the device word advanced by the ticks it
        # consumed and advances the env's host counter by the same amount, so a graphed rollout draws exactly what the
        # same rollout stepped call by call draws; if direct calls were made in between, the device word is re-aimed
        # at the env's counter before the next replay (one tiny fill), so the two paths never share a (seed, counter, env).
        self._ticks = 2 if auto_reset else 1
        self._addend = torch.zeros(1, dtype=torch.int64, device=dev)
        self._mirror = 0                            # host mirror of the device word
        self._base = env._draws + 1                 # the counter the captured step uses
        env_word, env._draws_dev = env._draws_dev, self._addend
        try:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.outputs = self._one(capturing=True)
        finally:
            env._draws_dev = env_word
        env._draws = self._base - 1                 # the capture enqueued nothing: its host ticks are handed back
        self.obs = self.outputs[0]

    def _one(self, capturing):
        env = self.env
        fused = self.auto_reset and getattr(env, 'supports_fused_reset', False)
        if self.host_io:                                # head of the graph: this step's actions, pinned host -> device
            if self.multi:
                for a, t in self.actions.items():
                    t.copy_(self.host_actions[a], non_blocking=True)
            else:
                self.actions.copy_(self.host_actions, non_blocking=True)
        if self.multi:
            obs, rewards, dones, info = env.step(self.actions, auto_reset=fused)
            out = (obs, rewards, dones, info)
            done = dones['__all__']
        else:
            obs, reward, done, info = env.step(self.actions, auto_reset=True) if fused else env.step(self.actions)
            out = (obs, reward, done, info)
        if self.auto_reset and not fused:
            env.reset(done, return_observations=False)
        if self.host_io:                                # tail: the step's results, device -> pinned host
            if self.multi:
                self.host_reward.copy_(env.rewards.view(env.num_envs, env.num_snakes), non_blocking=True)
                self.host_done.copy_(env._step_dones, non_blocking=True)
                self.host_all_done.copy_(done, non_blocking=True)
            else:
                self.host_reward.copy_(reward, non_blocking=True)
                self.host_done.copy_(done, non_blocking=True)
        if capturing:
            self._addend.add_(self._ticks)          # one tick per step, one per reset
        return out

    def _replay(self):
        env = self.env
        if hasattr(env, '_before_replay'):          # a replay changes the state behind the Python wrapper's back: fold caller
            env._before_replay()                    # edits in (compact state) / drop hints and shadow records the caller outdated
        want = env._draws + 1 - self._base          # the captured step must run with the env's next counter value
        if want != self._mirror:                    # direct env.step()/reset() calls were made since the last replay
            self._addend.fill_(want)
            self._mirror = want
        self.graph.replay()
        self._mirror += self._ticks
        env._draws += self._ticks

    def step(self):
        """One env step (+ reset of finished envs).  Returns the static output tensors of the captured step."""
        self._replay()
        return self.outputs

    def step_host(self):
        """host_io=True: one env step fed from `host_actions`; returns once `host_reward` / `host_done` are valid."""
        if not self.host_io:
            raise RuntimeError('GraphedStepper was built without host_io=True')
        self._replay()
        torch.cuda.current_stream(self._dev).synchronize()
        return self.outputs
