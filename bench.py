"""Benchmark of the wurm_b200 hot path: env-steps/s of the batched env step on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload C2] [--configs C1,C3,C4,C5|none]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One "step" is one pass of the hot path over one batch of synthetic actions, in the reference's own
benchmark shape (reference tests/test_single_snake_env.py:23-34, experiments/speeds.py:28-42):

    obs, reward, done, info = env.step(actions[t]);  env.reset(done)

At N GPUs every rank owns an independent slice of `num_envs` environments (weak scaling: the
environments do not interact); the only collective is one all-reduce of the
episode-statistics counters after the timed region.  Rank 0 prints ONE JSON line.

The line's top level is the HEADLINE workload (C2 = BASELINE.json configs[1], the config the metric is quoted
on), timed over exactly --steps steps.  `configs` holds one sub-record per other BASELINE config (C1, C3, C4, C5:
same fields, each timed for at least MIN_SECONDS independent of --steps); under --gpus N every sub-record is the
N-rank weak-scaling figure, so C5 is BASELINE configs[4] itself: 32 768 envs per GPU sharded across the ranks.

  value      whole-job env-steps/s, inputs (actions) already resident in HBM, CUDA-event timed, max over ranks
  sustained  the same loop repeated until at least MIN_SECONDS have been timed (a longer-running check of `value`)
  e2e        same loop through the public host-buffer API (wurm_b200.HostStepper): every step copies that step's
             actions from pinned host memory, and copies that step's results (rewards and done flags) back to
             pinned host memory, all inside the timed region (copies of neighbouring steps overlap the kernels;
             observations stay on the device: they are the policy's input).  `e2e_with_obs` additionally brings
             the step's observation back to pinned host memory (what a HOST-side policy would need).
  roofline   for the dominant kernel (the step kernel): algorithmic bytes per launch (SURVEY.md section 8d:
             read state + write state + write observation + per-env vectors) / its average launch duration,
             measured live with CUDA events around every step launch of the timed region, against the
             measured copy bandwidth in MEASURED_PEAKS.json; `traffic` is the dram bytes per launch from the last
             `ncu --set full` capture and is reported only while wurm_b200/csrc/ is byte-identical to the
             sources that capture was taken from (profiles/traffic.json records their hash)
  cpu_baseline  the REFERENCE ITSELF (unmodified oscarknagg/wurm from baseline/_ref or /root/reference, under the
             five compatibility shims of oracle/reference_loader.py), device='cpu', all host threads, on a bounded
             sample of the same workload (num_envs reduced so that a batched step stays
             under ~1 s; CPU throughput is flat in num_envs), rank 0, N=1 only.  `cpu_baseline_port` is the C/OpenMP
             restatement (oracle/wurm_oracle.c) on the same cores; `reference_cuda` the same unmodified reference
             code with device='cuda' (stock ATen kernels on this B200).
  compact_state  the same workload on an env built with state='compact' (an opt-in extension: the env lives in HBM as
             one small record per cell instead of the reference's dense fp32 tensors, which are materialised on
             attribute access; bit-identical results) -- the headline `value` stays on the reference layout
  dense_scan_state  MultiSnake workloads only: the reference layout WITHOUT the shadow records the default keeps beside the
             tensors (state='dense_scan': every step streams the fp32 tensors) -- what the shadow buys, in the same line

`--impl reference` times the reference's own PyTorch CPU implementation alone and prints the same line shape.
"""
import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (env, size, num_envs per GPU, observation mode, num_snakes)  -- BASELINE.json configs[1], [2], [0], [3], [4]
    'C2': ('SingleSnake', 9, 1 << 20, 'partial_2', 1),
    'C3': ('SingleSnake', 36, 1 << 16, 'default', 1),
    'C1': ('SingleSnake', 9, 512, 'partial_2', 1),
    'C4': ('MultiSnake', 25, 1 << 16, 'partial_4', 4),
    'C5': ('MultiSnake', 64, 1 << 15, 'partial_4', 16),
    'C5F': ('MultiSnake', 64, 1 << 14, 'full', 16),            # SURVEY section 8d variant: class-default 'full' observations
    'G1': ('SimpleGridworld', 7, 1 << 20, 'default', 1),      # next-row env, reference test size
}
BASELINE_CONFIGS = {'C1': 'configs[0]', 'C2': 'configs[1]', 'C3': 'configs[2]', 'C4': 'configs[3]', 'C5': 'configs[4]'}
# num_envs of the reference's own PyTorch implementation on the CPU: BASELINE.md section 4.3 (a batched step under ~1 s)
REFERENCE_CPU_ENVS = {'C1': 512, 'C2': 16384, 'C3': 4096, 'C4': 1024, 'C5': 64, 'C5F': 32, 'G1': 16384}
ACTION_POOL = 16        # pre-generated action tensors cycled through by the timed loop
GRAPHED_E2E_MAX_ENVS = 16384   # at or below: e2e goes through GraphedStepper(host_io=True), the API for launch-bound sizes
MIN_SECONDS = 0.3       # sub-records and `sustained` are timed for at least this long


def workload_name(key, n_envs=None):
    env, S, N, mode, K = WORKLOADS[key]
    N = n_envs or N
    if env == 'MultiSnake':
        return (f'{env} {K} snakes size={S} num_envs={N} {mode} obs, constructor-default rules, random actions in [0,8), '
                f"step+reset(done['__all__'])")
    return f'{env} size={S} num_envs={N} {mode} obs, random actions, step+reset(done)'


def algorithmic_bytes_per_env_step(key, obs_elems_per_env):
    """read state + write state + write obs + per-env vectors
    (single: actions r/w 16, reward 4, done 1, info 2; multi: ~42.5 B per agent)."""
    env, S, _, _, K = WORKLOADS[key]
    if env == 'MultiSnake':
        return 2 * (1 + 2 * K) * S * S * 4 + obs_elems_per_env * 4 + (170 * K) // 4
    if env == 'SimpleGridworld':
        return 2 * 2 * S * S * 4 + obs_elems_per_env * 4 + 13
    return 2 * 3 * S * S * 4 + obs_elems_per_env * 4 + 23


class SingleAdapter(object):
    """Uniform loop interface over the env classes (GPU arm)."""
    kernel = 'single_tile_kernel<G,STEP=true>'

    def __init__(self, key, dev, seed, rank, state='dense'):
        import torch
        from wurm_b200.envs import SingleSnake
        _, S, N, mode, _ = WORKLOADS[key]
        self.N, self.torch = N, torch
        self.env = SingleSnake(num_envs=N, size=S, observation_mode=mode, device=dev, seed=seed, state=state)
        if state == 'compact':
            self.kernel = 'single_compact_kernel<G,STEP=true>'
        elif S % 2 == 0 and S >= 16 and (mode in ('default', 'one_channel') or not 20 < S < 32):
            self.kernel = 'single_body_kernel<G>'       # even sizes from 16 up step on body-only tiles
        g = torch.Generator(device=dev).manual_seed(4321 + rank)
        self.pool = [torch.randint(0, 4, (N,), device=dev, generator=g) for _ in range(ACTION_POOL)]
        self.action_desc = f'int64 randint(0,4), pool of {ACTION_POOL} pre-generated tensors per rank'
        self.loop_desc = 'obs,reward,done,info = env.step(a); env.reset(done, return_observations=False)'

    def step(self, t):
        obs, reward, done, info = self.env.step(self.pool[t % ACTION_POOL])
        return obs, reward, done

    def reset(self, done):
        self.env.reset(done, return_observations=False)

    def fused_step(self, t):
        return self.env.step(self.pool[t % ACTION_POOL], auto_reset=True)

    def obs_elems(self, obs):
        return obs[0].numel()

    def host_pool(self):
        # what a host-side policy hands over: uint8 actions in pinned memory (one byte per env-step over PCIe)
        return [p.to(self.torch.uint8).cpu().pin_memory() for p in self.pool]


class GridAdapter(SingleAdapter):
    kernel = 'grid_tile_kernel<4,STEP=true>'    # grids up to 64 cells (size 8: grid_small_kernel); larger: grid_env_kernel<32,true>

    def __init__(self, key, dev, seed, rank, state='dense'):
        import torch
        from wurm_b200.envs import SimpleGridworld
        _, S, N, mode, _ = WORKLOADS[key]
        self.N, self.torch = N, torch
        self.env = SimpleGridworld(num_envs=N, size=S, observation_mode=mode, device=dev, start_location=(S // 2, S // 2),
                                   seed=seed)
        g = torch.Generator(device=dev).manual_seed(4321 + rank)
        self.pool = [torch.randint(0, 4, (N,), device=dev, generator=g) for _ in range(ACTION_POOL)]
        self.action_desc = f'int64 randint(0,4), pool of {ACTION_POOL} pre-generated tensors per rank'
        self.loop_desc = 'obs,reward,done,info = env.step(a); env.reset(done, return_observations=False)'

    fused_step = None


class MultiAdapter(object):
    kernel = 'multi_env_kernel<STEP=true>'

    def __init__(self, key, dev, seed, rank, state='dense'):
        import torch
        from wurm_b200.envs import MultiSnake
        _, S, N, mode, K = WORKLOADS[key]
        self.N, self.K, self.torch = N, K, torch
        self.env = MultiSnake(num_envs=N, num_snakes=K, size=S, observation_mode=mode, device=dev, seed=seed, state=state)
        # state='dense' (the default): the reference's fp32 tensors are the state and the library shadows them with its own
        # records, which the steady-state step loads (and verifies against the tensors) instead of streaming the tensors
        self.kernel = {'dense': 'multi_env_kernel<STEP=true,COMPACT=true,SHADOW=true>', 'dense_scan': 'multi_env_kernel<STEP=true>',
                       'compact': 'multi_env_kernel<STEP=true,COMPACT=true>'}[state]
        g = torch.Generator(device=dev).manual_seed(4321 + rank)
        self.pool = [{f'agent_{k}': torch.randint(0, 8, (N,), device=dev, generator=g) for k in range(K)}
                     for _ in range(ACTION_POOL)]
        self.action_desc = f'int64 randint(0,8) per agent, pool of {ACTION_POOL} pre-generated dicts per rank'
        self.loop_desc = ("obs,rewards,dones,info = env.step(actions); "
                          "env.reset(dones['__all__'], return_observations=False)")

    def step(self, t):
        obs, rewards, dones, info = self.env.step(self.pool[t % ACTION_POOL])
        return obs, rewards, dones['__all__']

    def reset(self, done):
        self.env.reset(done, return_observations=False)

    def fused_step(self, t):
        return self.env.step(self.pool[t % ACTION_POOL], auto_reset=True)

    def obs_elems(self, obs):
        return sum(o[0].numel() for o in obs.values())

    def host_pool(self):
        return [{a: t.to(self.torch.uint8).cpu().pin_memory() for a, t in d.items()} for d in self.pool]


def make_adapter(key, dev, seed, rank, state='dense'):
    cls = {'MultiSnake': MultiAdapter, 'SimpleGridworld': GridAdapter}.get(WORKLOADS[key][0], SingleAdapter)
    return cls(key, dev, seed, rank, state)


def compact_bytes_per_env(key):
    """Bytes of the compact resident state per env (records + per-env / per-snake side vectors)."""
    env, S, _, _, K = WORKLOADS[key]
    if env == 'MultiSnake':
        return ((S * S + 3) & ~3) * 4 + K * (2 + 1 + 8 + 1 + 6)
    return ((S * S + 7) & ~7) * 2 + 8


def measure_compact(ctx, key, K, W, obs_elems):
    """The same workload with state='compact' (records resident in HBM instead of the reference's fp32 tensors; opt-in,
    bit-identical results, tensors materialised on attribute access).  Returns the sub-record on rank 0."""
    import torch
    import torch.distributed as dist
    from wurm_b200 import HostStepper
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    _, S, N, mode, _ = WORKLOADS[key]
    ad = make_adapter(key, dev, 1234 + rank, rank, state='compact')

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for t in range(W):
        obs, reward, done = ad.step(t)
        ad.reset(done)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    a.record()
    for t in range(K):
        ev[t][0].record()
        obs, reward, done = ad.step(t)
        ev[t][1].record()
        ad.reset(done)
    b.record()
    barrier()
    two_ms = a.elapsed_time(b)
    kernel_ms = statistics.mean(x.elapsed_time(y) for x, y in ev)
    for t in range(W):
        ad.fused_step(t)
    barrier()
    a.record()
    for t in range(K):
        ad.fused_step(t)
    b.record()
    barrier()
    fused_ms = a.elapsed_time(b)
    host_pool = ad.host_pool()
    stepper = HostStepper(ad.env, depth=2)
    Ke = max(10, K // 2)
    tickets = []
    for t in range(4):
        tickets.append(stepper.submit(host_pool[t % ACTION_POOL]))
    while tickets:
        tickets.pop(0).wait()
    barrier()
    a.record()
    for t in range(Ke):
        tickets.append(stepper.submit(host_pool[t % ACTION_POOL]))
        if len(tickets) > stepper.depth:
            tickets.pop(0).wait()
    while tickets:
        tickets.pop(0).wait()
    b.record(stepper.d2h)
    barrier()
    e2e_ms = a.elapsed_time(b)
    h2d, d2h, api = stepper.h2d_bytes_per_step, stepper.d2h_bytes_per_step, stepper.describe()
    ad.env.check_status()
    t = torch.tensor([two_ms, fused_ms, kernel_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    two_ms, fused_ms, kernel_ms, e2e_ms = t.tolist()
    kernel_name = ad.kernel
    del ad, stepper, host_pool
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    peak, _ = measured_peak_gbs()
    traffic, traffic_src = profiled_traffic(key, 'compact')
    algorithmic = algorithmic_bytes_per_env_step(key, obs_elems) * N
    moved = (2 * compact_bytes_per_env(key) + obs_elems * 4 + 16) * N      # records read + written (upper bound) + obs + vectors
    return {'value': world * N * K / (two_ms * 1e-3), 'unit': 'env-steps/s', 'ms_per_step': two_ms / K, 'steps': K,
            'loop': 'env.step(a); env.reset(done, return_observations=False) on an env built with state=\'compact\'',
            'fused_step_reset': {'value': world * N * K / (fused_ms * 1e-3), 'ms_per_step': fused_ms / K,
                                 'loop': 'env.step(a, auto_reset=True)'},
            'e2e': {'value': world * N * Ke / (e2e_ms * 1e-3), 'unit': 'env-steps/s', 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h, 'ms_per_step': e2e_ms / Ke, 'api': api},
            'resident_state_bytes_per_env': compact_bytes_per_env(key),
            'roofline': {'bound': 'hbm', 'kernel': kernel_name, 'kernel_ms': kernel_ms,
                         'achieved': algorithmic / (kernel_ms * 1e-3) / 1e9, 'peak': peak, 'unit': 'GB/s',
                         'frac': algorithmic / (kernel_ms * 1e-3) / 1e9 / peak, 'bytes_per_launch': algorithmic,
                         'compact_bytes_per_launch': moved, 'traffic': traffic, 'traffic_source': traffic_src,
                         'physical_frac': (traffic / (kernel_ms * 1e-3) / 1e9 / peak) if traffic else None,
                         'compact_gbs': moved / (kernel_ms * 1e-3) / 1e9, 'compact_frac': moved / (kernel_ms * 1e-3) / 1e9 / peak,
                         'note': 'frac keeps the SURVEY 8(d) denominator (the reference\'s dense fp32 layout) over the step '
                                 'kernel\'s launch time; compact_* count the bytes this layout has to move (records in + out, '
                                 'observation, per-env vectors)'}}



def measure_dense_scan(ctx, key, K, W, obs_elems):
    """MultiSnake with state='dense_scan': the reference's tensors WITHOUT the shadow records -- every step streams them (the
    round-1 / early round-2 kernel, kept for callers who write the state through raw pointers).  Reported beside the default so
    that what the shadow buys is visible in the same line."""
    import torch
    import torch.distributed as dist
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    N = WORKLOADS[key][2]
    ad = make_adapter(key, dev, 1234 + rank, rank, state='dense_scan')

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for t in range(W):
        obs, reward, done = ad.step(t)
        ad.reset(done)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    a.record()
    for t in range(K):
        ev[t][0].record()
        obs, reward, done = ad.step(t)
        ev[t][1].record()
        ad.reset(done)
    b.record()
    barrier()
    ms = a.elapsed_time(b)
    kernel_ms = statistics.mean(x.elapsed_time(y) for x, y in ev)
    ad.env.check_status()
    t = torch.tensor([ms, kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, kernel_ms = t.tolist()
    kernel_name = ad.kernel
    del ad
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    peak, _ = measured_peak_gbs()
    traffic, traffic_src = profiled_traffic(key, 'dense_scan')
    algorithmic = algorithmic_bytes_per_env_step(key, obs_elems) * N
    return {'value': world * N * K / (ms * 1e-3), 'unit': 'env-steps/s', 'ms_per_step': ms / K, 'steps': K,
            'loop': "env.step(a); env.reset(done, return_observations=False) on an env built with state='dense_scan'",
            'roofline': {'bound': 'hbm', 'kernel': kernel_name, 'kernel_ms': kernel_ms,
                         'achieved': algorithmic / (kernel_ms * 1e-3) / 1e9, 'peak': peak, 'unit': 'GB/s',
                         'frac': algorithmic / (kernel_ms * 1e-3) / 1e9 / peak, 'bytes_per_launch': algorithmic,
                         'traffic': traffic, 'traffic_source': traffic_src,
                         'physical_frac': (traffic / (kernel_ms * 1e-3) / 1e9 / peak) if traffic else None}}


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '50', '-i', str(self.gpu)], stdout=subprocess.PIPE, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()

    def samples_since(self, t0):
        return sum(1 for (ts, _) in self.rows if ts >= t0)

    def summary(self, t0, t1):
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        rows = [r for (t, r) in self.rows if t0 <= t <= t1] or [r for (_, r) in self.rows[-3:]]
        for r in rows:
            f = [x.strip() for x in r.split(',')]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, f[2:]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def measured_peak_gbs():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


KERNEL_SOURCES = {
    # the files a workload's step kernel is compiled from: what a profiled traffic figure is valid for
    'single': ('common.cuh', 'host_util.h', 'single_device.cuh', 'single_snake.cu', 'single_compact.cu'),
    'multi': ('common.cuh', 'host_util.h', 'multi_snake.cu'),
    'grid': ('common.cuh', 'host_util.h', 'gridworld.cu'),
}


def kernel_family(key):
    return {'SingleSnake': 'single', 'MultiSnake': 'multi', 'SimpleGridworld': 'grid'}[WORKLOADS[key][0]]


def csrc_hash(family=None):
    """Content hash of kernel sources: all of wurm_b200/csrc, or the files one kernel family is compiled from."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, 'wurm_b200', 'csrc')
    names = KERNEL_SOURCES[family] if family else sorted(n for n in os.listdir(d) if n.endswith(('.cu', '.cuh', '.h')))
    for name in names:
        h.update(name.encode())
        h.update(open(os.path.join(d, name), 'rb').read())
    return h.hexdigest()[:16]


def profiled_traffic(key, state='dense'):
    """dram bytes per launch of the step kernel from the last `ncu --set full` capture (profiles/traffic.json), and
    where it came from.  Refused (None + reason) when the kernel sources changed since that capture."""
    path = os.path.join(ROOT, 'profiles', 'traffic.json')
    if not os.path.exists(path):
        return None, 'profiles/traffic.json missing'
    rec = json.load(open(path)).get(key if state == 'dense' else key + ':' + state)
    if not isinstance(rec, dict):
        return None, 'no capture recorded for this workload'
    now = csrc_hash(kernel_family(key))
    if rec.get('csrc_hash') != now:
        return None, (f"stale: captured from kernel sources {rec.get('csrc_hash')} ({rec.get('source')}), "
                      f'the library is built from {now}')
    return rec['dram_bytes_per_launch'], f"{rec.get('source')} (csrc {rec.get('csrc_hash')})"


# ------------------------------------------------------------------------------------------------
# CPU legs: the reference itself (PyTorch, unmodified, shims) and the C/OpenMP port (oracle)
# ------------------------------------------------------------------------------------------------
def time_cpu_port(key, n_envs, steps, warmup, threads):
    """Times the oracle's step + observe + reset loop on `threads` host threads.  Returns (env-steps/s, s)."""
    import numpy as np
    from oracle import oracle as orc      # the checker; allowed here as a timed CPU baseline only
    os.environ['OMP_NUM_THREADS'] = str(threads)
    env_name, S, _, mode, K = WORKLOADS[key]
    rng = np.random.default_rng(0)
    if env_name == 'MultiSnake':
        cfg = orc.multi_cfg(n_envs, K, S)
        st = orc.MultiState(n_envs, K, S)
        orc.multi_reset(cfg, st, np.ones(n_envs, np.uint8), None, seed=1234, step=0)
        pool = [rng.integers(0, 8, (n_envs, K)).astype(np.int64) for _ in range(ACTION_POOL)]
        t0 = None
        for t in range(warmup + steps):
            if t == warmup:
                t0 = time.perf_counter()
            out = orc.multi_step(cfg, st, pool[t % ACTION_POOL], None, seed=1234, step=2 * t + 1)
            obs, _ = orc.multi_observe(cfg, st, mode)
            orc.multi_reset(cfg, st, out['all_done'], None, seed=1234, step=2 * t + 2)
        dt = time.perf_counter() - t0
        return n_envs * steps / dt, dt
    if env_name == 'SimpleGridworld':
        start = (S // 2, S // 2)
        state = np.zeros((n_envs, 2, S, S), np.float32)
        orc.grid_reset(state, np.ones(n_envs, np.uint8), start, None, seed=1234, step=0)
        pool = [rng.integers(0, 4, n_envs).astype(np.int64) for _ in range(ACTION_POOL)]
        t0 = None
        for t in range(warmup + steps):
            if t == warmup:
                t0 = time.perf_counter()
            r, d = orc.grid_step(state, pool[t % ACTION_POOL], None, seed=1234, step=2 * t + 1)
            obs = orc.grid_observe(state, mode)
            orc.grid_reset(state, d, start, None, seed=1234, step=2 * t + 2)
        dt = time.perf_counter() - t0
        return n_envs * steps / dt, dt
    state = np.zeros((n_envs, 3, S, S), np.float32)
    orc.single_reset(state, np.ones(n_envs, np.uint8), None, seed=1234, step=0)
    pool = [rng.integers(0, 4, n_envs).astype(np.int64) for _ in range(ACTION_POOL)]
    t0 = None
    for t in range(warmup + steps):
        if t == warmup:
            t0 = time.perf_counter()
        r, d, sc, ec = orc.single_step(state, pool[t % ACTION_POOL], None, seed=1234, step=2 * t + 1)
        obs, _ = orc.single_observe(state, mode)
        orc.single_reset(state, d, None, seed=1234, step=2 * t + 2)
    dt = time.perf_counter() - t0
    return n_envs * steps / dt, dt


def port_sample_size(key):
    _, S, N, _, K = WORKLOADS[key]
    per_env = (3 if K == 1 else 1 + 2 * K) * S * S * 4
    if per_env > (1 << 16):
        return min(N, max(64, (1 << 26) // per_env // 64 * 64))
    return min(N, max(512, ((1 << 26) // per_env) // 1024 * 1024))    # ~64 MiB of state


def time_reference_torch(key, n_envs, steps, warmup, device):
    """The UNMODIFIED reference (wurm.envs.*) in its own benchmark loop (tests/test_single_snake_env.py:23-34,
    experiments/speeds.py:28-42): obs, r, d, info = env.step(a); env.reset(d).  Must run in a process of its own:
    oracle/reference_loader.py patches torch process-wide to restore torch-1.1 semantics.  Returns a dict."""
    import torch
    from oracle import reference_loader as rl          # loader of the reference; allowed here as the timed baseline only
    ref = rl.load(record=False, device=device)
    env_name, S, _, mode, K = WORKLOADS[key]
    g = torch.Generator().manual_seed(4321)
    sync = (lambda: torch.cuda.synchronize()) if device != 'cpu' else (lambda: None)
    torch.manual_seed(1234)
    if env_name == 'MultiSnake':
        env = ref.MultiSnake(num_envs=n_envs, num_snakes=K, size=S, observation_mode=mode, device=device)
        pool = [{f'agent_{k}': torch.randint(0, 8, (n_envs,), generator=g).to(device) for k in range(K)} for _ in range(ACTION_POOL)]

        def one(t):
            obs, rewards, dones, info = env.step(pool[t % ACTION_POOL])
            env.reset(dones['__all__'], return_observations=False)
        loop = "obs,rewards,dones,info = env.step(actions); env.reset(dones['__all__'], return_observations=False)"
    else:
        cls = ref.SimpleGridworld if env_name == 'SimpleGridworld' else ref.SingleSnake
        kw = dict(start_location=(S // 2, S // 2)) if env_name == 'SimpleGridworld' else {}
        env = cls(num_envs=n_envs, size=S, observation_mode=mode, device=device, **kw)
        pool = [torch.randint(0, 4, (n_envs,), generator=g).to(device) for _ in range(ACTION_POOL)]

        def one(t):
            obs, reward, done, info = env.step(pool[t % ACTION_POOL].clone())
            env.reset(done)
        loop = 'obs,reward,done,info = env.step(a); env.reset(done)'
    for t in range(warmup):
        one(t)
    sync()
    t0 = time.perf_counter()
    for t in range(steps):
        one(warmup + t)
    sync()
    dt = time.perf_counter() - t0
    return {'value': n_envs * steps / dt, 'seconds': dt, 'num_envs': n_envs, 'steps': steps, 'warmup': warmup,
            'ms_per_step': dt / steps * 1e3, 'loop': loop, 'device': device, 'reference_path': rl.REFERENCE_PATH}


def reference_leg(keys, device, steps, warmup, sizes=None, timeout=600):
    """Runs time_reference_torch for `keys` in ONE child process (the shims must not leak into this one) and
    returns {key: dict}; a key that failed maps to {'unavailable': reason}."""
    cmd = [sys.executable, os.path.abspath(__file__), '--reference-child', ','.join(keys), '--ref-device', device,
           '--steps', str(steps), '--warmup', str(warmup)]
    if sizes:
        cmd += ['--ref-envs', ','.join(str(sizes[k]) for k in keys)]
    env = dict(os.environ)
    for var in ('RANK', 'LOCAL_RANK', 'WORLD_SIZE', 'MASTER_ADDR', 'MASTER_PORT'):
        env.pop(var, None)
    try:
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout, env=env)
    except subprocess.TimeoutExpired:
        return {k: {'unavailable': f'timed out after {timeout} s'} for k in keys}
    for line in reversed(out.stdout.splitlines()):
        if line.startswith('{'):
            return json.loads(line)
    tail = (out.stderr or out.stdout).strip().splitlines()[-1:] or ['no output']
    return {k: {'unavailable': f'child failed (rc {out.returncode}): {tail[0][:200]}'} for k in keys}


def reference_child(args):
    """--reference-child: the body of reference_leg's child process; prints one JSON dict."""
    import torch
    keys = args.reference_child.split(',')
    sizes = [int(x) for x in args.ref_envs.split(',')] if args.ref_envs else None
    device = args.ref_device
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    results = {}
    for i, key in enumerate(keys):
        n = sizes[i] if sizes else (REFERENCE_CPU_ENVS[key] if device == 'cpu' else WORKLOADS[key][2])
        tried = []
        while True:
            try:
                r = time_reference_torch(key, n, args.steps, args.warmup, device)
                r['threads'] = torch.get_num_threads()
                r['cpu_count'] = os.cpu_count()
                if tried:
                    r['larger_sizes_failed'] = tried
                results[key] = r
                break
            except Exception as exc:                 # the reference on a device / size it cannot handle: say so
                tried.append({'num_envs': n, 'error': f'{type(exc).__name__}: {str(exc)[:160]}'})
                if device != 'cpu':
                    torch.cuda.empty_cache()
                if n <= 64 or device == 'cpu':
                    results[key] = {'unavailable': tried[-1]['error'], 'tried': tried}
                    break
                n //= 4
    print(json.dumps(results), flush=True)


def cpu_model():
    try:
        for line in open('/proc/cpuinfo'):
            if line.startswith('model name'):
                return line.split(':', 1)[1].strip()
    except OSError:
        pass
    return 'unknown'


def reference_baseline_record(r, key):
    """cpu_baseline-shaped record from a time_reference_torch result."""
    if 'unavailable' in r:
        return {'unavailable': r['unavailable'], 'kind': 'reference'}
    return {'value': r['value'], 'unit': 'env-steps/s', 'cores': r['threads'], 'kind': 'reference',
            'sample': (f"unmodified reference ({r['reference_path']}) under the oracle/reference_loader.py shims, device={r['device']}, "
                       f"{r['num_envs']} envs x {r['steps']} steps (+{r['warmup']} warm-up) of {workload_name(key, r['num_envs'])}; "
                       f"loop: {r['loop']}; torch threads {r['threads']} of {r['cpu_count']} cpus ({cpu_model()}); {r['seconds']:.1f} s"),
            'num_envs': r['num_envs'], 'ms_per_step': r['ms_per_step']}


def run_reference(args):
    """--impl reference: the reference's own PyTorch implementation on all host cores (rank 0 only)."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    key = args.workload
    subs = [k for k in parse_configs(args) if k != key]
    steps = max(3, min(args.steps, 20))
    warmup = max(1, min(args.warmup, 5))
    res = reference_leg([key] + subs, 'cpu', steps, warmup)
    head = res[key]
    if 'unavailable' in head:
        print(json.dumps({'impl': 'reference', 'unavailable': head['unavailable']}), flush=True)
        return
    rec = reference_baseline_record(head, key)
    threads = os.cpu_count() or 1
    n_port = port_sample_size(key)
    port_value, port_dt = time_cpu_port(key, n_port, max(3, min(args.steps, 50)), 2, threads)
    line = {
        'impl': 'reference', 'metric': 'env-steps/sec', 'value': head['value'], 'unit': 'env-steps/s', 'n_gpus': args.gpus,
        'steps': steps, 'warmup': warmup, 'ms_per_step': head['ms_per_step'], 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(key), 'baseline_config': BASELINE_CONFIGS.get(key),
                   'sample': rec['sample'], 'num_envs_sampled': head['num_envs']},
        'cpu_baseline': rec,
        'cpu_baseline_port': {'value': port_value, 'unit': 'env-steps/s', 'cores': threads, 'kind': 'port',
                              'sample': f'{n_port} envs, oracle/wurm_oracle.c (C/OpenMP restatement) on {threads} threads, {port_dt:.1f} s'},
        'e2e': {'value': head['value'], 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
        'configs': {k: ({'workload': workload_name(k), 'baseline_config': BASELINE_CONFIGS.get(k),
                         'value': res[k]['value'], 'unit': 'env-steps/s', 'ms_per_step': res[k]['ms_per_step'],
                         'cpu_baseline': reference_baseline_record(res[k], k)} if 'unavailable' not in res[k] else res[k])
                    for k in subs},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class Ctx(object):
    pass


def measure_config(ctx, key, K, W, min_seconds, exact_steps):
    """One workload on this rank's GPU (all ranks call this in lockstep).  Returns the record on rank 0, None elsewhere.
    exact_steps: time exactly K steps (the headline contract); otherwise repeat the K-step block until min_seconds."""
    import torch
    import torch.distributed as dist
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    _, S, N, mode, _ = WORKLOADS[key]
    ad = make_adapter(key, dev, 1234 + rank, rank)
    env = ad.env

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def reduce_max(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    # ---- device-resident throughput (value) + per-launch step-kernel time (roofline) ----
    for t in range(W):
        obs, reward, done = ad.step(t)
        ad.reset(done)

    def timed_block(steps, with_kernel_events):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)] if with_kernel_events else None
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        start.record()
        for t in range(steps):
            if ev:
                ev[t][0].record()
            obs, reward, done = ad.step(t)
            if ev:
                ev[t][1].record()
            ad.reset(done)
        stop.record()
        barrier()
        kern = statistics.mean(a.elapsed_time(b) for a, b in ev) if ev else None
        return start.elapsed_time(stop), kern, obs

    if not exact_steps:                      # calibrate the block length to the minimum timed duration
        ms, _, _ = timed_block(max(5, min(K, 20)), False)
        per = reduce_max([ms])[0] / max(5, min(K, 20))
        K = int(min(max(min_seconds * 1e3 / max(per, 1e-3), 20), 20000))
    t_wall0 = time.time()
    ms_total, step_kernel_ms, obs = timed_block(K, True)
    t_wall1 = time.time()
    sustained = None
    if exact_steps and ms_total < min_seconds * 1e3:
        reps = int(min(max(min_seconds * 1e3 / max(ms_total / K, 1e-3), K), 20000))
        ms_sus, _, _ = timed_block(reps, False)
        sustained = (reps, ms_sus)
        t_wall1 = time.time()
    # a short timed region may end before nvidia-smi has produced three rows: keep the same loop running, untimed,
    # until it has, so that the clocks are sampled under this very load
    t_extra = 0
    while ctx.sampler.samples_since(t_wall0) < 3 and time.time() - t_wall1 < 1.5:
        o2, r2, d2 = ad.step(t_extra)
        ad.reset(d2)
        t_extra += 1
        if t_extra % 16 == 0:
            torch.cuda.synchronize(dev)
    if t_extra:
        torch.cuda.synchronize(dev)
        t_wall1 = time.time()
    obs_elems = ad.obs_elems(obs)

    # ---- supplementary: the fused step+reset fast path (one launch per step) ----
    fused_ms = 0.0
    if getattr(ad, 'fused_step', None) is not None:
        for t in range(W):
            ad.fused_step(t)
        f_start, f_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        f_start.record()
        for t in range(K):
            ad.fused_step(t)
        f_stop.record()
        barrier()
        fused_ms = f_start.elapsed_time(f_stop)

    # ---- C3 as BASELINE.json words it: "RGB obs + feedforward A2C rollout"  ----
    # the env's observation feeds a 2 x 64 MLP (experiments.main.FeedforwardAgent, plain torch: the policy is outside the hot
    # path) whose sampled actions drive the next step: obs -> policy -> Categorical.sample -> env.step(auto_reset=True)
    policy_ms = 0.0
    if key == 'C3':
        from torch.distributions import Categorical
        from experiments.main import FeedforwardAgent
        model = FeedforwardAgent(num_actions=4, num_layers=2, hidden_units=64, num_inputs=obs_elems).to(dev)
        state = obs
        with torch.no_grad():
            for t in range(W + K):
                if t == W:
                    p_start, p_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    barrier()
                    p_start.record()
                probs, _ = model(state)
                action = Categorical(probs).sample()
                state, _, _, _ = env.step(action, auto_reset=True)
            p_stop.record()
            barrier()
        policy_ms = p_start.elapsed_time(p_stop)
        del model, state

    # ---- end to end through the public API with host buffers ----
    from wurm_b200 import HostStepper, GraphedStepper
    host_pool = ad.host_pool()
    Ke = max(10, K // 2)
    e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    graphed_e2e = N <= GRAPHED_E2E_MAX_ENVS and getattr(env, 'supports_fused_reset', False)
    graphed_ms = 0.0
    e2e_obs_ms, e2e_obs_bytes = 0.0, 0
    if graphed_e2e:
        # supplementary: the same step+reset through one CUDA-graph launch per step, actions resident on the device
        first = ad.pool[0]
        static_dev = {a: t.clone() for a, t in first.items()} if isinstance(first, dict) else first.clone()
        gs = GraphedStepper(env, static_dev)
        g_start, g_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        g_start.record()
        for t in range(K):
            src = ad.pool[t % ACTION_POOL]
            if isinstance(src, dict):
                for a, v in src.items():
                    static_dev[a].copy_(v)
            else:
                static_dev.copy_(src)
            gs.step()
        g_stop.record()
        barrier()
        graphed_ms = g_start.elapsed_time(g_stop)
        del gs
        # launch-bound sizes: the copies ride inside the CUDA graph (GraphedStepper(host_io=True)); per step the host
        # writes that step's actions into the graph's pinned input, replays, synchronises and reads the results
        first = host_pool[0]
        static = {a: t.to(dev) for a, t in first.items()} if isinstance(first, dict) else first.to(dev)
        stepper = GraphedStepper(env, static, host_io=True)

        def put(src):
            if isinstance(src, dict):
                for a, t in src.items():
                    stepper.host_actions[a].copy_(t)
            else:
                stepper.host_actions.copy_(src)
        for t in range(4):
            put(host_pool[t % ACTION_POOL]); stepper.step_host()
        barrier()
        e_start.record()
        checksum = 0.0
        for t in range(Ke):
            put(host_pool[t % ACTION_POOL])
            stepper.step_host()
            checksum += float(stepper.host_reward[0].sum())      # the host reads the step's result
        e_stop.record()
        barrier()
        e2e_ms = e_start.elapsed_time(e_stop)
        h2d = sum(t.numel() * t.element_size() for t in (first.values() if isinstance(first, dict) else [first]))
        d2h = stepper.host_reward.numel() * 4 + stepper.host_done.numel() + (stepper.host_all_done.numel() if hasattr(stepper, 'host_all_done') else 0)
        e2e_api = 'GraphedStepper(host_io=True): one CUDA-graph launch per step carrying H2D actions, fused step+reset, D2H rewards + done flags'
        del stepper
    else:
        def pipelined(stepper):
            tickets = []
            for t in range(4):
                tickets.append(stepper.submit(host_pool[t % ACTION_POOL]))
            while tickets:
                tickets.pop(0).wait()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for t in range(Ke):
                tickets.append(stepper.submit(host_pool[t % ACTION_POOL]))
                if len(tickets) > stepper.depth:
                    tickets.pop(0).wait()
            while tickets:
                tickets.pop(0).wait()
            b.record(stepper.d2h)
            barrier()
            return a.elapsed_time(b)
        stepper = HostStepper(env, depth=2)
        e2e_ms = pipelined(stepper)
        h2d, d2h = stepper.h2d_bytes_per_step, stepper.d2h_bytes_per_step
        e2e_api = stepper.describe() if hasattr(stepper, 'describe') else 'HostStepper'
        del stepper
        try:
            stepper = HostStepper(env, depth=2, return_obs=True)
        except TypeError:
            stepper = None
        if stepper is not None:
            e2e_obs_ms = pipelined(stepper)
            e2e_obs_bytes = stepper.d2h_bytes_per_step
            del stepper

    ms_total, e2e_ms, step_kernel_ms, fused_ms, graphed_ms, e2e_obs_ms, policy_ms = reduce_max(
        [ms_total, e2e_ms, step_kernel_ms, fused_ms, graphed_ms, e2e_obs_ms, policy_ms])
    if sustained is not None:
        sustained = (sustained[0], reduce_max([sustained[1]])[0])
    stats = env.stats(reduce_group=True if world > 1 else None)
    env.check_status()
    clocks = ctx.sampler.summary(t_wall0, t_wall1)
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, clocks)
        sms = [c['sm_mhz'] for c in gathered if c['sm_mhz'] is not None]
        clocks = {'sm_mhz': min(sms) if sms else None, 'sm_max_mhz': gathered[0]['sm_max_mhz'],
                  'reasons': sorted(set(sum((c['reasons'] for c in gathered), []))),
                  'samples': sum(c['samples'] for c in gathered)}
    kernel_name, action_desc, loop_desc = ad.kernel, ad.action_desc, ad.loop_desc
    del ad, env, host_pool
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    compact = None
    if ctx.compact and WORKLOADS[key][0] in ('SingleSnake', 'MultiSnake'):
        compact = measure_compact(ctx, key, K, W, obs_elems)
    dense_scan = None
    if ctx.compact and WORKLOADS[key][0] == 'MultiSnake':
        dense_scan = measure_dense_scan(ctx, key, K, W, obs_elems)
    if rank != 0:
        return None

    value = world * N * K / (ms_total * 1e-3)
    peak, peak_src = measured_peak_gbs()
    bytes_per_launch = algorithmic_bytes_per_env_step(key, obs_elems) * N
    achieved = bytes_per_launch / (step_kernel_ms * 1e-3) / 1e9
    traffic, traffic_src = profiled_traffic(key)
    rec = {
        'metric': 'env-steps/sec', 'value': value, 'unit': 'env-steps/s', 'n_gpus': world, 'steps': K, 'warmup': W,
        'ms_per_step': ms_total / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(key), 'baseline_config': BASELINE_CONFIGS.get(key), 'num_envs_per_gpu': N,
                   'num_envs_total': N * world, 'size': S, 'observation_mode': mode,
                   'actions': action_desc, 'loop': loop_desc,
                   'l2': f'inputs larger than L2: state {(bytes_per_launch - N * obs_elems * 4) / 2e6:.0f} MB + obs '
                         f'{N * obs_elems * 4 / 1e6:.0f} MB per step vs 126 MB L2' if bytes_per_launch > 2.6e8 else
                         f'working set {bytes_per_launch / 2e6:.1f} MB fits L2: launch-latency-bound config, roofline fraction reported but not targeted (SURVEY.md section 8)',
                   'parallelism': f'{world} independent env slices, NCCL all-reduce of episode stats only'},
        'clocks': clocks,
        'e2e': {'value': world * N * Ke / (e2e_ms * 1e-3), 'unit': 'env-steps/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'steps': Ke, 'ms_per_step': e2e_ms / Ke, 'api': e2e_api},
        'gpu_launches': 2 * K,
        'roofline': {'bound': 'hbm', 'kernel': kernel_name, 'achieved': achieved, 'peak': peak,
                     'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic, 'traffic_source': traffic_src,
                     'peak_source': peak_src, 'bytes_per_launch': bytes_per_launch,
                     'kernel_ms': step_kernel_ms, 'frac_of_nominal_8TBs': achieved / 8000.0,
                     'note': 'achieved = ALGORITHMIC bytes (dense fp32 state read + write + obs) / kernel time; the '
                             'kernels write back only the cells a step changed and do not stream what they already know '
                             '(SingleSnake: verified head / food hints; MultiSnake: the state tensors are shadowed by 4-byte '
                             'cell records, which the step loads and verifies against the tensors instead of streaming ~99 % '
                             'zeros), so the DRAM traffic ncu measures (traffic) is below the algorithmic '
                             'count and frac can exceed 1; physical_frac = traffic / kernel time / peak is how close '
                             'the kernel runs to the hardware',
                     'dram_gbs_from_traffic': (traffic / (step_kernel_ms * 1e-3) / 1e9) if traffic else None,
                     'physical_frac': (traffic / (step_kernel_ms * 1e-3) / 1e9 / peak) if traffic else None},
        'episode_stats': stats,
    }
    if policy_ms:
        rec['rollout_with_policy'] = {'value': world * N * K / (policy_ms * 1e-3), 'unit': 'env-steps/s', 'ms_per_step': policy_ms / K,
                                      'loop': 'probs, value = FeedforwardAgent(2 x 64)(obs); a = Categorical(probs).sample(); '
                                              'obs, r, d, info = env.step(a, auto_reset=True)   (policy in plain torch, no grad)'}
    if compact is not None:
        rec['compact_state'] = compact
    if dense_scan is not None:
        rec['dense_scan_state'] = dense_scan
    if sustained is not None:
        rec['sustained'] = {'value': world * N * sustained[0] / (sustained[1] * 1e-3), 'unit': 'env-steps/s', 'steps': sustained[0],
                            'seconds': sustained[1] * 1e-3}
    if e2e_obs_ms:
        rec['e2e_with_obs'] = {'value': world * N * Ke / (e2e_obs_ms * 1e-3), 'unit': 'env-steps/s', 'd2h_bytes_per_step': e2e_obs_bytes,
                               'h2d_bytes_per_step': h2d, 'ms_per_step': e2e_obs_ms / Ke,
                               'note': 'as e2e, plus the D2H copy of the step\'s observation (a host-side policy\'s input)'}
    if fused_ms:
        rec['fused_step_reset'] = {'value': world * N * K / (fused_ms * 1e-3), 'unit': 'env-steps/s',
                                   'ms_per_step': fused_ms / K, 'gpu_launches': K,
                                   'loop': 'env.step(actions, auto_reset=True)  (one launch per step)'}
    if graphed_ms:
        rec['graphed_step_reset'] = {'value': world * N * K / (graphed_ms * 1e-3), 'unit': 'env-steps/s',
                                     'ms_per_step': graphed_ms / K,
                                     'loop': 'GraphedStepper(env, actions).step()  (one CUDA-graph launch per step; plus '
                                             'the device-side copy of the step\'s actions into the static input)'}
    return rec


def parse_configs(args):
    if args.configs in ('none', ''):
        return []
    if args.configs == 'auto':
        return [k for k in ('C1', 'C3', 'C4', 'C5') if k != args.workload] if args.workload == 'C2' else []
    return [k for k in args.configs.split(',') if k]


def run_gpu(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit('launch with torch.distributed.run --nproc-per-node N for --gpus N > 1')
    torch.cuda.set_device(local_rank)
    ctx = Ctx()
    ctx.dev = torch.device('cuda', local_rank)
    ctx.world, ctx.rank = world, rank
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=ctx.dev)
    ctx.compact = not args.no_compact
    ctx.sampler = ClockSampler(local_rank)
    ctx.sampler.start()                     # nvidia-smi takes a moment to produce its first row: start before the warm-up

    key = args.workload
    K, W = args.steps, max(3, args.warmup)
    line = measure_config(ctx, key, K, W, MIN_SECONDS, exact_steps=True)
    subs = {}
    for sub in parse_configs(args):
        if sub == key:
            continue
        rec = measure_config(ctx, sub, min(K, 200), W, MIN_SECONDS, exact_steps=False)
        if rec is not None:
            for drop in ('metric', 'unit', 'higher_is_better', 'scaling', 'vs_baseline', 'dtype', 'data', 'n_gpus'):
                rec.pop(drop, None)
            subs[sub] = rec
    ctx.sampler.stop()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    line['configs'] = subs
    if world == 1 and not args.no_cpu_baseline:
        # the reference itself on this box's host cores (a child process: its shims patch torch process-wide) ...
        keys = [key] + list(subs)
        res = reference_leg(keys, 'cpu', 20, 5)
        line['cpu_baseline'] = reference_baseline_record(res[key], key)
        for k in subs:
            subs[k]['cpu_baseline'] = reference_baseline_record(res[k], k)
        # ... the same unmodified reference code on this GPU (stock ATen kernels) ...
        if not args.no_reference_cuda:
            res = reference_leg(keys, 'cuda', 10, 3, timeout=900)
            for k in keys:
                r = res[k]
                out = r if 'unavailable' in r else {
                    'value': r['value'], 'unit': 'env-steps/s', 'num_envs': r['num_envs'], 'ms_per_step': r['ms_per_step'],
                    'steps': r['steps'], 'kind': 'reference', 'device': 'cuda',
                    'note': 'the unmodified reference (PyTorch ops, its own loop incl. its host syncs) on this B200'}
                if r.get('larger_sizes_failed'):
                    out['larger_sizes_failed'] = r['larger_sizes_failed']
                (line if k == key else subs[k])['reference_cuda'] = out
        # ... and the C/OpenMP restatement of the algorithm (the parity oracle) as a second CPU figure
        n = port_sample_size(key)
        threads = os.cpu_count() or 1
        rate, _ = time_cpu_port(key, n, 3, 1, threads)                    # calibrate, then ~8 s of CPU work
        cpu_steps = int(min(max(8.0 * rate / n, 5), 2000))
        cpu_value, dt = time_cpu_port(key, n, cpu_steps, 2, threads)
        line['cpu_baseline_port'] = {'value': cpu_value, 'unit': 'env-steps/s', 'cores': threads, 'kind': 'port',
                                     'sample': f'{n} envs x {cpu_steps} steps of the same workload (step+observe+reset), '
                                               f'oracle/wurm_oracle.c with OpenMP on {threads} threads, {dt:.1f} s'}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=1000)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='wurm_b200', choices=['wurm_b200', 'reference'])
    ap.add_argument('--workload', default='C2', choices=sorted(WORKLOADS))    # C1..C5 = BASELINE.json configs
    ap.add_argument('--configs', default='auto', help="sub-records: 'auto' (C1,C3,C4,C5 beside the default headline), 'none', or a comma list")
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-reference-cuda', action='store_true')
    ap.add_argument('--no-compact', action='store_true', help="skip the state='compact' sub-records")
    ap.add_argument('--reference-child', default=None, help=argparse.SUPPRESS)
    ap.add_argument('--ref-device', default='cpu', help=argparse.SUPPRESS)
    ap.add_argument('--ref-envs', default=None, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.reference_child:
        reference_child(args)
    elif args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
