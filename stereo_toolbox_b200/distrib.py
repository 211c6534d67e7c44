"""Multi-GPU plumbing of the hot path: the batch/image dimension shards across ranks with NO data-path
collective; torch.distributed is only used to agree on timings and counters.
Works with NCCL (GPUs) and gloo (CPU tests)."""
from __future__ import annotations

import os
from typing import Sequence, Tuple

import torch


def env_rank() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment; (0, 1, 0) when launched plainly."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_range(total: int, rank: int, world: int) -> range:
    """Contiguous slice of `total` work units owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return range(lo, lo + base + (1 if rank < rem else 0))


def reduce_stats(times_ms: Sequence[float], counts: Sequence[float], device="cpu"):
    """Slowest rank's time for each timed region and the job-wide sum of each counter."""
    import torch.distributed as dist
    t = torch.tensor(list(times_ms), dtype=torch.float64, device=device)
    c = torch.tensor(list(counts), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return t.tolist(), c.tolist()
