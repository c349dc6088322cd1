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
"""
import torch


class GraphedStepper(object):
    def __init__(self, env, actions, auto_reset: bool = True, warmup: int = 2):
        self.env = env
        self.actions = actions
        self.multi = hasattr(env, 'num_snakes')
        self.auto_reset = auto_reset
        dev = env.envs.device if not self.multi else env.foods.device
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
        if self.multi:
            obs, rewards, dones, info = env.step(self.actions, auto_reset=fused)
            out = (obs, rewards, dones, info)
            done = dones['__all__']
        else:
            obs, reward, done, info = env.step(self.actions, auto_reset=True) if fused else env.step(self.actions)
            out = (obs, reward, done, info)
        if self.auto_reset and not fused:
            env.reset(done, return_observations=False)
        if capturing:
            env._draws_dev.add_(2 if self.auto_reset else 1)      # one tick per step, one per reset
        return out

    def step(self):
        """One env step (+ reset of finished envs).  Returns the static output tensors of the captured step."""
        self.graph.replay()
        return self.outputs
