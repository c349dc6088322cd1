"""SimpleGridworld (G1: size 7, 2^20 envs, default observations): step; reset loop time.  WURM_GRID_NO_TILE=1 selects the
round-1 lane-group kernel instead of the tile kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wurm_b200.envs import SimpleGridworld
for S, N, mode in ((5, 1 << 20, 'default'), (6, 1 << 20, 'default'), (7, 1 << 20, 'default'), (7, 1 << 20, 'raw'), (7, 1 << 20, 'positions'), (8, 1 << 20, 'default')):
    env = SimpleGridworld(num_envs=N, size=S, observation_mode=mode, device='cuda', start_location=(S // 2, S // 2), seed=1)
    pool = [torch.randint(0, 4, (N,), device='cuda') for _ in range(16)]
    for t in range(20):
        o, r, d, i = env.step(pool[t % 16]); env.reset(d, return_observations=False)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ks = []
    a.record()
    for t in range(300):
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record(); o, r, d, i = env.step(pool[t % 16]); k1.record(); ks.append((k0, k1))
        env.reset(d, return_observations=False)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 300
    print(f'S={S} {mode}: {ms:.4f} ms/step ({N / ms * 1e3:.4g} env-steps/s), step kernel {sum(x.elapsed_time(y) for x, y in ks) / 300:.4f} ms', flush=True)
    del env, pool
