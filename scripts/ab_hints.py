"""Scratch A/B: C2 step+reset with and without the (head cell, size) hints."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wurm_b200 import SingleSnake
N = 1 << 20
for use_hints in (True, False, True, False):
    env = SingleSnake(num_envs=N, size=9, observation_mode='partial_2', device='cuda', seed=1)
    if not use_hints:
        env._hints = None
    acts = [torch.randint(0, 4, (N,), device='cuda') for _ in range(16)]
    for t in range(20):
        o, r, d, i = env.step(acts[t % 16]); env.reset(d, return_observations=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(300):
        o, r, d, i = env.step(acts[t % 16]); env.reset(d, return_observations=False)
    e1.record(); torch.cuda.synchronize()
    print(f'hints={use_hints}: {e0.elapsed_time(e1)/300:.4f} ms/step')
