"""Multi-GPU plumbing: one process per GPU, one independent env slice per rank.

Environments never interact, so the data path has NO collective; the only
exchange is the all-reduce of the episode-statistics counters the step kernels accumulate.
"""
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
