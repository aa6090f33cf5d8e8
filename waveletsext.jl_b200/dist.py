"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL on GPUs / gloo in CPU tests).

The transforms are independent per signal, so the batch (last Julia dimension = first tensor dimension) is
sharded contiguously across ranks with NO data-path collective.  The only exchange on the hot path is the
all-reduce of the small per-position state of the JBB / LSDB cost trees (bestbasis.py).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

__all__ = ["shard_range", "shard", "is_dist", "rank", "world_size", "allreduce_sum", "allreduce_min", "allreduce_max",
           "total_count", "broadcast_from_first"]


def is_dist(group=None) -> bool:
    return dist.is_available() and dist.is_initialized()


def rank(group=None) -> int:
    return dist.get_rank(group) if is_dist(group) else 0


def world_size(group=None) -> int:
    return dist.get_world_size(group) if is_dist(group) else 1


def shard_range(N: int, r: int | None = None, R: int | None = None, group=None):
    """contiguous slab of signals owned by rank r of R: [lo, hi)"""
    r = rank(group) if r is None else r
    R = world_size(group) if R is None else R
    base, rem = divmod(N, R)
    lo = r * base + min(r, rem)
    return lo, lo + base + (1 if r < rem else 0)


def shard(x: torch.Tensor, group=None) -> torch.Tensor:
    lo, hi = shard_range(x.shape[0], group=group)
    return x[lo:hi]


def _reduce(t: torch.Tensor, op, group):
    if is_dist(group) and world_size(group) > 1:
        if t.is_contiguous():
            dist.all_reduce(t, op=op, group=group)
        else:
            c = t.contiguous()
            dist.all_reduce(c, op=op, group=group)
            t.copy_(c)
    return t


def allreduce_sum(t, group=None):
    return _reduce(t, dist.ReduceOp.SUM, group)


def allreduce_min(t, group=None):
    return _reduce(t, dist.ReduceOp.MIN, group)


def allreduce_max(t, group=None):
    return _reduce(t, dist.ReduceOp.MAX, group)


def total_count(nlocal: int, device, group=None) -> int:
    if not (is_dist(group) and world_size(group) > 1):
        return int(nlocal)
    t = torch.tensor([nlocal], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())


def broadcast_from_first(t: torch.Tensor, group=None) -> torch.Tensor:
    """value held by rank 0 (the owner of the first signal of the global batch)"""
    if is_dist(group) and world_size(group) > 1:
        t = t.contiguous()
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    return t
