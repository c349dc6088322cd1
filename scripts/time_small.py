"""Scratch: launch-bound batch sizes (C1 = 512 envs) stepped call by call vs through one CUDA graph."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wurm_b200 import SingleSnake, GraphedStepper

for N in (512, 4096, 32768):
    env = SingleSnake(num_envs=N, size=9, observation_mode='partial_2', device='cuda', seed=1)
    acts = torch.randint(0, 4, (64, N), device='cuda')
    def plain(t):
        o, r, d, i = env.step(acts[t % 64]); env.reset(d, return_observations=False)
    for t in range(20): plain(t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    T = 500
    e0.record()
    for t in range(T): plain(t)
    e1.record(); torch.cuda.synchronize()
    ms_plain = e0.elapsed_time(e1) / T
    static = acts[0].clone()
    st = GraphedStepper(env, static)
    for t in range(20): st.step()
    torch.cuda.synchronize()
    e0.record()
    for t in range(T):
        static.copy_(acts[t % 64]); st.step()
    e1.record(); torch.cuda.synchronize()
    ms_graph = e0.elapsed_time(e1) / T
    print(f'N={N}: call-by-call {ms_plain*1e3:.1f} us/step ({N/ms_plain*1e3:.3e} env-steps/s)   graphed {ms_graph*1e3:.1f} us/step ({N/ms_graph*1e3:.3e} env-steps/s)')
