"""The bench loop (env.step(a); env.reset(done)) of one workload for a few steps -- the target of ncu captures.
    python scripts/profile_step.py C2 [dense|compact] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

key = sys.argv[1]
state = sys.argv[2] if len(sys.argv) > 2 else 'dense'
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 14
ad = bench.make_adapter(key, torch.device('cuda', 0), 1234, 0, state=state)
for t in range(steps):
    obs, reward, done = ad.step(t)
    ad.reset(done)
torch.cuda.synchronize()
print('ran', key, state, steps, 'steps; kernel', ad.kernel)
