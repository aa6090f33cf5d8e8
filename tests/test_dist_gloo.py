"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: batch sharding without a data-path collective and the
all-reduce protocol of the JBB cost tree (per-position sum / sumsq -> identical costs and tree on every rank)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import waveletsext_b200 as wx
    import oracle as O
    try:
        N, n, K = 37, 32, 4
        rng = np.random.default_rng(123)
        X = rng.standard_normal((N, K, n))                     # the same global table on every rank
        lo, hi = wx.dist.shard_range(N)
        assert (lo, hi) == ((0, 19) if rank == 0 else (19, 37))
        Xl = torch.from_numpy(X[lo:hi])
        assert torch.equal(wx.dist.shard(torch.from_numpy(X)), Xl)
        assert wx.dist.total_count(hi - lo, "cpu") == N
        # JBB protocol: local moments -> all-reduce(sum) -> costs -> tree (host)
        mom = torch.stack([Xl.sum(0).reshape(-1), (Xl * Xl).sum(0).reshape(-1)])
        wx.dist.allreduce_sum(mom)
        ref_mom = np.stack([X.sum(0).reshape(-1), (X * X).sum(0).reshape(-1)])
        assert np.allclose(mom.numpy(), ref_mom, rtol=1e-13, atol=1e-12)
        ex, ex2 = mom[0].numpy() / N, mom[1].numpy() / N
        sig = np.sqrt(ex2 - ex * ex).reshape(K, n)
        costs = []
        for lvl in range(K):
            n0 = n >> lvl
            for node in range(1 << lvl):
                costs.append(2 * np.log(sig[lvl, node * n0:(node + 1) * n0]).sum())
        costs = np.array(costs)
        ref = O.tree_costs_jbb(X)
        assert np.abs(costs - ref).max() <= 1e-11 * np.abs(ref).max()
        tree = wx.bestbasis_treeselection(costs.copy(), n)
        assert np.array_equal(tree, O.tree_select(ref, n))
        # every rank ends with the same tree
        t = torch.from_numpy(tree.astype(np.int64))
        tmin, tmax = t.clone(), t.clone()
        wx.dist.allreduce_min(tmin); wx.dist.allreduce_max(tmax)
        assert torch.equal(tmin, tmax)
        # LSDB protocol pieces: min / max reductions and the common shift broadcast from rank 0
        mn = torch.from_numpy(X[lo:hi].min(0)); mx = torch.from_numpy(X[lo:hi].max(0))
        wx.dist.allreduce_min(mn); wx.dist.allreduce_max(mx)
        assert np.array_equal(mn.numpy(), X.min(0)) and np.array_equal(mx.numpy(), X.max(0))
        first = wx.dist.broadcast_from_first(torch.from_numpy(X[lo].copy()))
        assert np.array_equal(first.numpy(), X[0])
        # LSDB sums travel as double-double pairs, all-gathered and summed in rank order: the rounded total does not depend
        # on the sharding.  Shard-local pairs here: (plain sum, 0) per half-shard, i.e. deliberately different roundings.
        flat = Xl.reshape(Xl.shape[0], -1)
        half = flat.shape[0] // 2
        local = wx.dist.dd_sum_host(torch.stack([torch.stack([flat[:half].sum(0), torch.zeros(flat.shape[1], dtype=torch.float64)]),
                                                 torch.stack([flat[half:].sum(0), torch.zeros(flat.shape[1], dtype=torch.float64)])]))
        parts = wx.dist.allgather_parts(local)
        assert parts.shape == (2, 2, flat.shape[1])
        tot = wx.dist.dd_sum_host(parts)
        import math
        exact = np.array([math.fsum(X.reshape(N, -1)[:, j]) for j in range(flat.shape[1])])
        assert np.abs(tot[0].numpy() - exact).max() <= 4e-16 * np.abs(exact).max() + 1e-15
        g = [torch.empty_like(tot) for _ in range(world)]
        dist.all_gather(g, tot)
        assert torch.equal(g[0], g[1])                       # bitwise identical on every rank
        # the library's own communicator is bootstrapped over this group: rank 0's NCCL unique id reaches every rank unchanged
        # (wx_comm_init_rank itself needs a GPU per rank: tests/mgpu_check.py)
        ident = wx.dist.exchange_unique_id()
        assert isinstance(ident, bytes) and len(ident) == 128 and any(ident)
        both = [None] * world
        dist.all_gather_object(both, ident)
        assert both[0] == both[1]
        # denoiseall's bestTH: per-signal noise levels of unequal shards, gathered in rank order
        sig = np.arange(lo, hi, dtype=np.float64) * 0.5
        allsig = wx.dist.allgather_host_vector(sig)
        assert np.array_equal(allsig, np.arange(N) * 0.5)
        q.put((rank, "ok"))
    except Exception as e:      # pragma: no cover
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_protocol():
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def test_shard_ranges_cover_batch():
    sys.path.insert(0, ROOT)
    import waveletsext_b200 as wx
    for N in (0, 1, 7, 64, 65536, 1000003):
        for R in (1, 2, 3, 8):
            spans = [wx.dist.shard_range(N, r, R) for r in range(R)]
            assert spans[0][0] == 0 and spans[-1][1] == N
            assert all(spans[i][1] == spans[i + 1][0] for i in range(R - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
