"""CPU cost of one MultiSnake.step() / reset() call (tiny env: the GPU work is negligible, the loop is host-bound).
    python scripts/time_host_overhead.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wurm_b200.envs import MultiSnake, SingleSnake

for K in (4, 16):
    env = MultiSnake(num_envs=64, num_snakes=K, size=25, observation_mode='partial_4', device='cuda', seed=1)
    acts = {f'agent_{k}': torch.randint(0, 8, (64,), device='cuda') for k in range(K)}
    for fused in (False, True):
        for _ in range(50):
            o, r, d, i = env.step(acts, auto_reset=fused)
            if not fused:
                env.reset(d['__all__'], return_observations=False)
        torch.cuda.synchronize()
        n = 1000
        t0 = time.perf_counter()
        for _ in range(n):
            o, r, d, i = env.step(acts, auto_reset=fused)
            if not fused:
                env.reset(d['__all__'], return_observations=False)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        print(f'MultiSnake K={K} {"step(auto_reset=True)" if fused else "step; reset"}: {1e6 * (t1 - t0) / n:.1f} us of host time per env-step')
env = SingleSnake(num_envs=64, size=9, observation_mode='partial_2', device='cuda', seed=1)
a = torch.randint(0, 4, (64,), device='cuda')
for _ in range(50):
    o, r, d, i = env.step(a); env.reset(d, return_observations=False)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(1000):
    o, r, d, i = env.step(a); env.reset(d, return_observations=False)
t1 = time.perf_counter()
print(f'SingleSnake step; reset: {1e3 * (t1 - t0):.1f} us of host time per env-step')
