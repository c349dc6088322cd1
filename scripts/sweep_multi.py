"""Scratch: CTA-size sweep of the MultiSnake step kernel for a few geometries (set WURM_MULTI_THREADS per process)."""
import sys, os, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch
    from wurm_b200 import MultiSnake
    E, K, S = map(int, sys.argv[1:4])
    env = MultiSnake(num_envs=E, num_snakes=K, size=S, observation_mode='partial_4', device='cuda', seed=1)
    acts = [{f'agent_{k}': torch.randint(0, 8, (E,), device='cuda') for k in range(K)} for _ in range(8)]
    for t in range(10):
        o, r, d, i = env.step(acts[t % 8]); env.reset(d['__all__'], return_observations=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(100):
        o, r, d, i = env.step(acts[t % 8]); env.reset(d['__all__'], return_observations=False)
    e1.record(); torch.cuda.synchronize()
    print(f'{e0.elapsed_time(e1)/100:.4f} ms/step')
else:
    for E, K, S in [(1 << 16, 4, 25), (1 << 14, 10, 36), (1 << 15, 2, 12), (1 << 13, 8, 48), (1 << 15, 16, 64)]:
        for thr in (32, 64, 128, 256):
            out = subprocess.run([sys.executable, __file__, str(E), str(K), str(S)], env=dict(os.environ, WURM_MULTI_THREADS=str(thr)),
                                 capture_output=True, text=True).stdout.strip()
            print(f'E={E} K={K} S={S} threads={thr}: {out}', flush=True)
