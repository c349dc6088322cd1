"""GPU (two or more devices): the drivers under `torch.distributed.run` -- one process per GPU, one env slice and one
seed per rank, episode counters all-reduced over NCCL every LOG_INTERVAL steps, rank 0 logging (SURVEY.md section 8e;
reference experiments/main.py:264-311, multiagent.py:486-503).  Skipped on a single-GPU box."""
import os
import re
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launch(module, args, nproc=2):
    port = 29500 + os.getpid() % 1000
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(nproc), '--master-addr',
           '127.0.0.1', '--master-port', str(port), '-m', module] + args
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-3000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith('steps=')]
    assert lines, res.stdout[-2000:]
    return [dict((k, float(v)) for k, v in (f.split('=') for f in l.split('\t'))) for l in lines]


needs_two = pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')


@needs_two
def test_single_agent_driver_on_two_ranks():
    """1000 envs split 500 / 500: rank 0 prints job-wide lines whose reduced `env_steps` counter (summed over the ranks'
    device counters by the NCCL all-reduce) equals the job-wide step count, and only rank 0 prints."""
    lines = launch('experiments.main', ['--env', 'snake', '--num-envs', '1000', '--size', '9', '--agent', 'random',
                                        '--observation', 'partial_2', '--total-steps', str(1000 * 200), '--seed', '5'])
    assert len(lines) == 2                                   # LOG_INTERVAL = 100, one printer
    for i, line in enumerate(lines, start=1):
        assert line['ranks'] == 2 and line['steps'] == 1000 * 100 * i
        assert line['env_steps'] == line['steps']            # sum of the per-rank counters
        assert line['episodes'] > 0 and 3.0 <= line['avg_size'] < 6.0


@needs_two
def test_multi_agent_driver_on_two_ranks():
    lines = launch('experiments.multiagent', ['--n-envs', '257', '--n-agents', '4', '--size', '25', '--obs', 'partial_4',
                                              '--total-steps', str(257 * 100), '--seed', '2'])
    assert len(lines) == 1
    assert lines[0]['ranks'] == 2 and lines[0]['env_steps'] == lines[0]['steps'] == 257 * 100       # 129 + 128 envs
    assert lines[0]['edge_collisions'] > 0


def test_ranks_get_different_draws():
    from wurm_b200.distributed import env_slice, rank_seed
    assert [env_slice(257, r, 2) for r in range(2)] == [(0, 129), (129, 128)]
    assert len({rank_seed(7, r) for r in range(8)}) == 8


@needs_two
def test_two_devices_in_one_process():
    """The dynamic shared-memory opt-in of a kernel is a PER-DEVICE attribute: an env stepped on cuda:0 and
    then another on cuda:1 in the same process must both launch -- with tiles above the 48 KB default (size 9 `default`
    observations: 62 KB; MultiSnake size 64: > 48 KB of records) -- and agree with each other."""
    from wurm_b200.envs import SingleSnake, MultiSnake
    outs = []
    for dev in ('cuda:0', 'cuda:1'):
        env = SingleSnake(num_envs=300, size=9, observation_mode='default', device=dev, seed=4)
        g = torch.Generator().manual_seed(1)
        for _ in range(5):
            obs, reward, done, _ = env.step(torch.randint(0, 4, (300,), generator=g).to(dev), auto_reset=True)
        env.check_status()
        outs.append((obs.cpu(), env.envs.cpu()))
        menv = MultiSnake(num_envs=8, num_snakes=16, size=64, observation_mode='partial_4', device=dev, seed=4)
        acts = {f'agent_{k}': torch.zeros(8, dtype=torch.long, device=dev) for k in range(16)}
        menv.step(acts, auto_reset=True)
        menv.check_consistency()
        cenv = SingleSnake(num_envs=300, size=36, observation_mode='default', device=dev, seed=4, state='compact')
        cenv.step(torch.zeros(300, dtype=torch.long, device=dev), auto_reset=True)
        cenv.check_consistency()
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
