"""Multi-GPU plumbing: one process per GPU, one independent env slice per rank.

Environments never interact, so the data path has NO collective; the only
exchange is the all-reduce of the episode-statistics counters the step kernels accumulate.
"""
import os

import torch


def env_slice(total_envs: int, rank: int, world_size: int):
    """(first env, number of envs) owned by `rank` when `total_envs` are split as evenly as possible."""
    base, extra = divmod(total_envs, world_size)
    count = base + (1 if rank < extra else 0)
    start = rank * base + min(rank, extra)
    return start, count


def rank_seed(seed: int, rank: int) -> int:
    """Per-rank Philox seed: ranks must not replay each other's draws."""
    return (seed + 0x9E3779B97F4A7C15 * (rank + 1)) % (1 << 62)


def all_reduce_stats(totals: torch.Tensor, group=None) -> torch.Tensor:
    """Sums the (fields,) int64 counter vector over the ranks of `group` (NCCL for CUDA tensors,
    gloo for CPU tensors).  A no-op without an initialised process group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(totals, op=dist.ReduceOp.SUM, group=group)
    return totals


class Ranks(object):
    """What a driver needs to know about its place in a one-process-per-GPU job."""

    def __init__(self, rank, world_size, local_rank, device):
        self.rank, self.world_size, self.local_rank, self.device = rank, world_size, local_rank, device

    @property
    def is_main(self):
        return self.rank == 0


def init_from_env(device: str = 'cuda') -> Ranks:
    """Reads RANK / WORLD_SIZE / LOCAL_RANK (set by `torch.distributed.run`); with more than one rank binds this process
    to its GPU (`cuda:LOCAL_RANK`) and initialises the default process group (NCCL for CUDA, gloo otherwise).  A plain
    `python -m experiments.main` run has no such variables and gets Ranks(0, 1, 0, device) without any group."""
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        if torch.device(device).type == 'cuda':
            torch.cuda.set_device(local)
            device = f'cuda:{local}'
            if not dist.is_initialized():
                dist.init_process_group('nccl', device_id=torch.device(device))
        elif not dist.is_initialized():
            dist.init_process_group('gloo')
    return Ranks(rank, world, local, device)


def finish(ranks: Ranks):
    import torch.distributed as dist
    if ranks.world_size > 1 and dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()
