"""Dense vs compact resident state: fused step+reset launch time at the BASELINE geometries.
    python scripts/time_compact.py [C4|C5|C2|C3] [steps] [dense|compact|both]
Tuning overrides (WURM_MULTI_COMPACT_THREADS, ...) are read by the library from the environment."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wurm_b200.envs import MultiSnake, SingleSnake

which = sys.argv[1] if len(sys.argv) > 1 else 'C5'
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
geo = {'C4': (1 << 16, 4, 25, 'partial_4'), 'C5': (1 << 15, 16, 64, 'partial_4'), 'C2': (1 << 20, 1, 9, 'partial_2'),
       'C3': (1 << 16, 1, 36, 'default')}[which]
E, K, S, mode = geo
which_states = sys.argv[3] if len(sys.argv) > 3 else 'both'
for state in (('dense', 'compact') if which_states == 'both' else (which_states,)):
    if K == 1:
        env = SingleSnake(num_envs=E, size=S, observation_mode=mode, device='cuda', seed=1, state=state)
        pool = [torch.randint(0, 4, (E,), device='cuda') for _ in range(16)]
    else:
        env = MultiSnake(num_envs=E, num_snakes=K, size=S, observation_mode=mode, device='cuda', seed=1, state=state)
        pool = [{f'agent_{k}': torch.randint(0, 8, (E,), device='cuda') for k in range(K)} for _ in range(16)]
    for t in range(20):
        env.step(pool[t % 16], auto_reset=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for t in range(steps):
        env.step(pool[t % 16], auto_reset=True)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    print(f'{which} {state:8s} {ms:.4f} ms/step  {E / ms * 1e3:.4g} env-steps/s  stats {env.stats()}', flush=True)
    del env, pool
    torch.cuda.empty_cache()
