"""CUDA-graph replay of the step/reset pair for launch-bound batch sizes.

At the reference's README scale (512 envs, BASELINE config 1) one env step moves ~1 MB: the GPU work
takes a few microseconds and the cost is launching it (two kernel launches, a dozen tensor
allocations, the Python between them).  `GraphedStepper` captures

    obs, reward, done, info = env.step(actions);  env.reset(done, return_observations=False)

once into a CUDA graph and replays it: one `cudaGraphLaunch` per env step (for SingleSnake the pair is
additionally fused into a single kernel, `wurm_single_step_reset`).  The kernels read their
Philox call counter as `step + *step_dev` (include/wurm_b200.h); the graph bumps the device word
after every replay, so replays draw fresh random numbers and a graphed rollout is bit-identical to the
same rollout stepped call by call.

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
        dev = env.envs.device if not self.multi else env.foods.device
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
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):                 # first-launch work (function attributes, allocator pools)
                self._one(capturing=False)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.outputs = self._one(capturing=True)
        # The capture enqueued nothing: the first replay IS the step whose host-side counters were baked in,
        # and every replay ends by moving the device-side addend on by the ticks one replay consumes.
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
            env._draws_dev.add_(2 if self.auto_reset else 1)      # one tick per step, one per reset
        return out

    def step(self):
        """One env step (+ reset of finished envs).  Returns the static output tensors of the captured step."""
        self.graph.replay()
        return self.outputs

    def step_host(self):
        """host_io=True: one env step fed from `host_actions`; returns once `host_reward` / `host_done` are valid."""
        if not self.host_io:
            raise RuntimeError('GraphedStepper was built without host_io=True')
        self.graph.replay()
        torch.cuda.current_stream(self._dev).synchronize()
        return self.outputs
