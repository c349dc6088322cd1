"""Scratch timing of the SingleSnake step/reset loop (not the bench contract; see bench.py)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wurm_b200.envs import SingleSnake

for N, S, mode in [(1 << 20, 9, 'partial_2'), (1 << 16, 36, 'default'), (512, 9, 'partial_2')]:
    env = SingleSnake(num_envs=N, size=S, observation_mode=mode, device='cuda', seed=1)
    T = 30
    acts = torch.randint(0, 4, (T, N), device='cuda')
    for t in range(5):
        o, r, d, i = env.step(acts[t]); env.reset(d, return_observations=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(5, T):
        o, r, d, i = env.step(acts[t]); env.reset(d, return_observations=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (T - 5)
    bytes_per = 2 * 3 * S * S * 4 + o[0].numel() * 4 + 23
    print(f'N={N} S={S} {mode}: {ms:.3f} ms/step  {N / ms * 1e3:.3e} env-steps/s  {N * bytes_per / ms / 1e6:.0f} GB/s algorithmic')
    # step only
    e0.record()
    for t in range(5, T):
        o, r, d, i = env.step(acts[t])
    e1.record(); torch.cuda.synchronize()
    print(f'   step only: {e0.elapsed_time(e1) / (T - 5):.3f} ms')
