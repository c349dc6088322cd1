import os, sys
sys.path.insert(0, os.getcwd())
import torch
from wurm_b200.envs import MultiSnake
for (E, K, S) in ((1 << 15, 16, 64), (1 << 16, 4, 25)):
    for state in ('compact', 'dense'):
        for mode in ('partial_4', None):
            env = MultiSnake(num_envs=E, num_snakes=K, size=S, observation_mode='partial_4', device='cuda', seed=1, state=state)
            env.observation_mode = mode
            pool = [{f'agent_{k}': torch.randint(0, 8, (E,), device='cuda') for k in range(K)} for _ in range(8)]
            for t in range(10): env.step(pool[t % 8], auto_reset=True)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for t in range(200): env.step(pool[t % 8], auto_reset=True)
            b.record(); torch.cuda.synchronize()
            print(f'K={K} S={S} {state} obs={mode}: {a.elapsed_time(b) / 200:.4f} ms/step', flush=True)
            del env, pool; torch.cuda.empty_cache()
