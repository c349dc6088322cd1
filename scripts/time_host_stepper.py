"""CPU cost of one HostStepper.submit (launch-bound batch: the GPU work is negligible) and e2e at 2^20 envs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wurm_b200 import SingleSnake, HostStepper

for N in (512, 1 << 20):
    env = SingleSnake(num_envs=N, size=9, observation_mode='partial_2', device='cuda', seed=1)
    pool = [torch.randint(0, 4, (N,)).to(torch.uint8).pin_memory() for _ in range(16)]
    st = HostStepper(env, depth=2)
    tickets = []
    for t in range(50):
        tickets.append(st.submit(pool[t % 16]))
        if len(tickets) > 2: tickets.pop(0).wait()
    torch.cuda.synchronize()
    T = 2000 if N == 512 else 500
    t0 = time.perf_counter()
    for t in range(T):
        tickets.append(st.submit(pool[t % 16]))
        if len(tickets) > 2: tickets.pop(0).wait()
    while tickets: tickets.pop(0).wait()
    dt = time.perf_counter() - t0
    print(f'N={N}: {dt / T * 1e6:.1f} us per submit+wait  ({N * T / dt:.4g} env-steps/s)', flush=True)
    # the same without the stepper: direct fused launches on resident actions
    dev_pool = [p.to('cuda') for p in pool]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(T):
        env.step(dev_pool[t % 16], auto_reset=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f'N={N}: {dt / T * 1e6:.1f} us per env.step(auto_reset=True)', flush=True)
