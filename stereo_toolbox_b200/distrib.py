"""Multi-GPU plumbing of the hot path: the batch/image dimension shards across ranks with NO data-path
collective in inference; torch.distributed is only used to agree on timings and counters.
Training adds the one real exchange step of the path: the gradient all-reduce (sum, then / world) that the reference
gets from DistributedDataParallel (trainer/trainer_torchrun.py:116-121) -- ``FlatGradAllReduce`` below, one NCCL
all-reduce over NVLink / NVSwitch of a single flat fp32 buffer that the parameters' ``.grad`` tensors are views of.
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


class FlatGradAllReduce:
    """All parameter gradients live in ONE flat fp32 buffer (each ``p.grad`` is a view into it, autograd accumulates in
    place), so the data-parallel exchange is a single large all-reduce (20.9 MB for PSMNet, 27.6 MB for GwcNet_GC):
    one launch-latency, NVLS-friendly message instead of ~300 small ones, and no flatten / unflatten copies."""

    def __init__(self, params, device=None):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        device = device or self.params[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=device)
        self._views = []
        off = 0
        for p in self.params:
            assert p.dtype == torch.float32
            self._views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self._attach()

    def _attach(self):
        for p, v in zip(self.params, self._views):
            p.grad = v

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def zero_(self):
        """Zero the gradients IN PLACE and keep every ``p.grad`` a view of the flat buffer (use this, or
        ``optimizer.zero_grad(set_to_none=False)``, instead of the default ``optimizer.zero_grad()``)."""
        self.flat.zero_()
        self._attach()

    def repack_(self) -> int:
        """Make the flat buffer hold the current gradients again and re-attach the views.  The reference trainer's
        ``optimizer.zero_grad()`` (trainer/trainer_torchrun.py:286-301) defaults to ``set_to_none=True``, which drops the
        views; autograd then allocates fresh ``.grad`` tensors and the flat buffer would be reduced as all zeros while the
        ranks silently diverge.  A parameter whose ``.grad`` no longer aliases its slice is copied in (``None``: the slice is
        zeroed -- DDP would also contribute zeros for a parameter that received no gradient).  Returns how many were stale."""
        stale = 0
        for p, v in zip(self.params, self._views):
            g = p.grad
            if g is not None and g.data_ptr() == v.data_ptr() and g.shape == v.shape:
                continue
            stale += 1
            if g is None:
                v.zero_()
            else:
                v.copy_(g)
            p.grad = v
        return stale

    def allreduce_(self):
        """sum over ranks, then / world (what DDP does); no-op for a single process.  Gradients that were detached from
        the flat buffer since the last step (``zero_grad(set_to_none=True)``) are packed back first."""
        import torch.distributed as dist
        self.repack_()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(dist.get_world_size())
        return self.flat
