"""Scratch: ms per step+reset of SingleSnake for one geometry:  python scripts/time_single.py N S mode"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wurm_b200.envs import SingleSnake
N, S, mode = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
env = SingleSnake(num_envs=N, size=S, observation_mode=mode, device='cuda', seed=1)
acts = [torch.randint(0, 4, (N,), device='cuda') for _ in range(8)]
for t in range(20):
    o, r, d, i = env.step(acts[t % 8]); env.reset(d, return_observations=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for t in range(200):
    o, r, d, i = env.step(acts[t % 8]); env.reset(d, return_observations=False)
e1.record(); torch.cuda.synchronize()
print(f'{e0.elapsed_time(e1) / 200:.4f} ms/step')
