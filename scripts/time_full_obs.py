import os, sys
sys.path.insert(0, os.getcwd())
import torch
from wurm_b200.envs import MultiSnake
for state in ('dense', 'compact'):
    for (E, K, S) in ((1 << 14, 16, 64), (1 << 14, 4, 25)):
        env = MultiSnake(num_envs=E, num_snakes=K, size=S, observation_mode='full', device='cuda', seed=1, state=state)
        pool = [{f'agent_{k}': torch.randint(0, 8, (E,), device='cuda') for k in range(K)} for _ in range(8)]
        for t in range(10): env.step(pool[t % 8], auto_reset=True)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for t in range(100): env.step(pool[t % 8], auto_reset=True)
        b.record(); torch.cuda.synchronize()
        print(f'{state} K={K} S={S} full: {a.elapsed_time(b) / 100:.4f} ms/step', flush=True)
        del env, pool; torch.cuda.empty_cache()
