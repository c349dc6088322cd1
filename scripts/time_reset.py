import sys, os, torch, json
sys.path.insert(0, os.getcwd())
import bench
res = {}
for key in ('C4', 'C5'):
    ad = bench.make_adapter(key, torch.device('cuda', 0), 1234, 0)
    for t in range(20):
        obs, r, d = ad.step(t); ad.reset(d)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 300
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    torch.cuda.synchronize(); a.record()
    for t in range(K):
        obs, r, d = ad.step(t)
        evs[t][0].record(); ad.reset(d); evs[t][1].record()
    b.record(); torch.cuda.synchronize()
    res[key] = dict(ms_per_step=a.elapsed_time(b) / K, reset_ms=sum(x.elapsed_time(y) for x, y in evs) / K)
    del ad
print(json.dumps(res))
