"""CPU, world size 2 (gloo): the rank plumbing of the multi-GPU path -- env slices, per-rank seeds, the
episode-statistics all-reduce (the only collective the path has) -- and bench.py's reference arm under
torchrun (rank 0 alone runs and prints)."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wurm_b200.distributed import env_slice, rank_seed, all_reduce_stats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_env_slices_partition_the_batch():
    for total, world in [(1 << 20, 8), (1000, 3), (5, 8), (32768 * 8, 8)]:
        slices = [env_slice(total, r, world) for r in range(world)]
        assert slices[0][0] == 0
        assert sum(c for _, c in slices) == total
        for (s0, c0), (s1, _) in zip(slices, slices[1:]):
            assert s0 + c0 == s1
        assert max(c for _, c in slices) - min(c for _, c in slices) <= 1


def test_rank_seeds_differ():
    assert len({rank_seed(1234, r) for r in range(8)}) == 8


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    totals = torch.tensor([10 + rank, 1, 2 * rank, 0, 7], dtype=torch.int64)      # one rank's kernel counters
    all_reduce_stats(totals)
    out[rank] = totals.tolist()
    dist.destroy_process_group()


def test_stats_all_reduce_gloo_world_size_2():
    world, port = 2, 29000 + os.getpid() % 2000
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert out[0] == out[1] == [21, 2, 2, 0, 14]


def test_bench_reference_arm_under_torchrun():
    """`bench.py --impl reference` launched like the driver does at N=2: exactly one JSON line, from rank 0."""
    port = 31000 + os.getpid() % 2000
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', str(port), os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2',
           '--steps', '2', '--warmup', '1', '--workload', 'C1']
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line['impl'] == 'reference' and line['n_gpus'] == 2 and line['value'] > 0
    for key in ('metric', 'unit', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'dtype', 'data', 'config',
                'cpu_baseline', 'e2e'):
        assert key in line
    # the arm times the reference ITSELF (PyTorch, under the loader's shims); the C port rides along as a second figure
    assert line['cpu_baseline']['kind'] == 'reference' and line['cpu_baseline_port']['kind'] == 'port'
    assert line['e2e']['h2d_bytes_per_step'] == 0


_RANK_SCRIPT = '''
import sys, torch
sys.path.insert(0, %r)
from wurm_b200.distributed import init_from_env, env_slice, rank_seed, all_reduce_stats, finish
ranks = init_from_env('cpu')                          # gloo: the same plumbing the drivers use on GPUs with NCCL
start, count = env_slice(1001, ranks.rank, ranks.world_size)
totals = torch.tensor([count, ranks.rank, 1, 0, 0], dtype=torch.int64)
all_reduce_stats(totals)
if ranks.is_main:
    print('TOTALS', totals.tolist(), ranks.world_size, rank_seed(3, 0) != rank_seed(3, 1))
finish(ranks)
'''


def test_driver_rank_plumbing_under_torchrun_gloo(tmp_path):
    """init_from_env + env_slice + all_reduce_stats + finish as experiments/main.py uses them, two ranks over gloo."""
    script = tmp_path / 'ranks.py'
    script.write_text(_RANK_SCRIPT % ROOT)
    port = 33000 + os.getpid() % 2000
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', str(port), str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith('TOTALS')]
    assert lines == ['TOTALS [1001, 1, 2, 0, 0] 2 True']
