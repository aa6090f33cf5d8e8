"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL on GPUs / gloo in CPU tests).

The transforms are independent per signal, so the batch (last Julia dimension = first tensor dimension) is
sharded contiguously across ranks with NO data-path collective.  The only exchange on the hot path is the
all-reduce of the small per-position state of the JBB / LSDB cost trees (bestbasis.py).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

__all__ = ["comm", "destroy_comms", "exchange_unique_id", "shard_range", "shard", "is_dist", "rank", "world_size", "allreduce_sum", "allreduce_min", "allreduce_max",
           "total_count", "broadcast_from_first", "allgather_parts", "dd_sum_host"]


_COMMS: dict = {}


def comm(device, group=None):
    """libwx_b200's own NCCL communicator for this process group (``wx_comm_t*`` as an integer; ``None`` = single GPU).
    Created on first use: rank 0 asks the library for an NCCL unique id (``wx_comm_unique_id``), the 128 bytes travel over
    the existing torch.distributed group (any backend -- this is the only thing torch.distributed carries for the best-basis
    path), every rank calls ``wx_comm_init_rank`` on its device.  Collective: all ranks of the group must call it together."""
    if not (is_dist(group) and world_size(group) > 1):
        return None
    import ctypes as C
    from . import _lib
    device = torch.device(device)
    key = (id(group) if group is not None else 0, device.index)
    c = _COMMS.get(key)
    if c is None:
        ident = exchange_unique_id(group)
        h = C.c_void_p()
        with torch.cuda.device(device):
            _lib.call("wx_comm_init_rank", C.byref(h), ident, rank(group), world_size(group))
        c = _COMMS[key] = h.value
    return c


def exchange_unique_id(group=None) -> bytes:
    """rank 0 of the group asks libwx_b200 for an NCCL unique id (``wx_comm_unique_id``, 128 bytes) and every rank receives it
    over the existing process group (any backend).  Collective."""
    import ctypes as C
    from . import _lib
    ident = [None]
    if rank(group) == 0:
        buf = (C.c_ubyte * 128)()
        _lib.call("wx_comm_unique_id", buf)
        ident[0] = bytes(buf)
    dist.broadcast_object_list(ident, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    assert isinstance(ident[0], bytes) and len(ident[0]) == 128
    return ident[0]


def destroy_comms() -> None:
    from . import _lib
    for c in _COMMS.values():
        _lib.call("wx_comm_destroy", c)
    _COMMS.clear()


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


def allgather_parts(pair: torch.Tensor, group=None) -> torch.Tensor:
    """all-gather of a (2, count) double-double buffer (hi row, lo row) -> (world, 2, count), rank order.  The caller
    sums the parts in that fixed order in double-double arithmetic (``wx_dd_sum`` on the device, ``dd_sum_host`` in the
    CPU tests): the result is exact to ~1e-32, so it does not depend on how the batch was sharded."""
    pair = pair.contiguous()
    if not (is_dist(group) and world_size(group) > 1):
        return pair.unsqueeze(0)
    out = torch.empty((world_size(group),) + tuple(pair.shape), dtype=pair.dtype, device=pair.device)
    if pair.is_cuda:
        dist.all_gather_into_tensor(out, pair, group=group)
    else:                                                  # gloo (CPU tests)
        dist.all_gather(list(out.unbind(0)), pair, group=group)
    return out


def allgather_host_vector(v, group=None):
    """concatenation, in rank order, of one small host vector per rank (shards may differ in length) -- the per-signal noise
    levels a ``bestTH`` summary needs (Denoising.jl:684-705).  Host-side glue, like the user's ``bestTH`` function itself."""
    import numpy as np
    v = np.ascontiguousarray(np.asarray(v, dtype=np.float64))
    if not (is_dist(group) and world_size(group) > 1):
        return v
    parts = [None] * world_size(group)
    dist.all_gather_object(parts, v, group=group)
    return np.concatenate(parts)


def dd_sum_host(parts: torch.Tensor) -> torch.Tensor:
    """(nparts, 2, count) -> (2, count): the double-double sum in index order with elementwise torch ops (each op is
    correctly rounded; no fused multiply-add is involved).  Host mirror of ``wx_dd_sum`` for the gloo tests."""
    hi = torch.zeros_like(parts[0, 0]); lo = torch.zeros_like(hi)
    for q in range(parts.shape[0]):
        bh, bl = parts[q, 0], parts[q, 1]
        s = hi + bh
        bb = s - hi
        e = (hi - (s - bb)) + (bh - bb)
        e = e + (lo + bl)
        hi = s + e
        lo = e - (hi - s)
    return torch.stack([hi, lo])
