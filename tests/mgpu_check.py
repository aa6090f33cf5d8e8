"""Multi-GPU check, one process per GPU (launch with torchrun): sharding the batch changes nothing bitwise for the
transforms, and the NCCL all-reduced JBB / LSDB cost trees equal the single-GPU trees on the concatenated batch.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import waveletsext_b200 as wx
    from test_gpu_bestbasis import signals
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    wt = wx.wavelet("db4")
    n, N = 256, 1000
    X = torch.from_numpy(signals(n, N, 77)).to(dev)             # same global batch everywhere
    lo, hi = wx.dist.shard_range(N)
    xl = X[lo:hi].contiguous()
    yl = wx.wpdall(xl, wt)                                       # shard-local, no collective
    yfull = wx.wpdall(X, wt)
    assert torch.equal(yl, yfull[lo:hi]), "sharded wpdall differs from the single-GPU result"
    # the other transforms of the path shard the same way (batch = slowest dimension, no collective)
    Xs = X[:64, :128].contiguous()
    l2, h2 = wx.dist.shard_range(64)
    for fn in (lambda v: wx.swpdall(v, wt, 4), lambda v: wx.acwpdall(v, wt, 4), lambda v: wx.wptall(v, wt, 5), lambda v: wx.sdwtall(v, wt, 3)):
        assert torch.equal(fn(Xs[l2:h2].contiguous()), fn(Xs)[l2:h2]), "sharded transform differs from the single-GPU result"
    img = X[:48, :].reshape(48, 16, 16).contiguous()
    l3, h3 = wx.dist.shard_range(48)
    assert torch.equal(wx.wpdall(img[l3:h3].contiguous(), wt, 2), wx.wpdall(img, wt, 2)[l3:h3]), "sharded 2-D wpd differs"
    for method in (wx.JBB(), wx.LSDB()):
        keep = yl.clone()
        c_sh = wx.tree_costs(yl, method)                         # all-reduce inside
        assert torch.equal(yl, keep), "tree_costs modified its input"
        dist.barrier()
        # single-GPU reference on the concatenated batch: bypass the process group
        saved = wx.dist.is_dist
        wx.dist.is_dist = lambda group=None: False
        try:
            c_one = wx.tree_costs(yfull, method)
        finally:
            wx.dist.is_dist = saved
        rel = np.abs(c_sh - c_one).max() / np.abs(c_one).max()
        # JBB costs are smooth in the moments: reassociating the sum moves them by ~1e-16.  LSDB bins every sample on a
        # grid derived from the statistics; those travel as double-double pairs combined in rank order, so grid and bin
        # counts are independent of the sharding and the costs agree to rounding of the final log sums.
        tol = 1e-11 if isinstance(method, wx.JBB) else 1e-12
        assert rel <= tol, (type(method).__name__, rel)
        t_sh = wx.bestbasis_treeselection(c_sh.copy(), n)
        t_one = wx.bestbasis_treeselection(c_one.copy(), n)
        assert np.array_equal(t_sh, t_one), type(method).__name__
        # the fused driver (reduction kernels -> NCCL exchange -> costs -> selection inside libwx_b200) gives the same tree
        assert np.array_equal(wx.bestbasistree(yl, method), t_one), "fused bestbasistree differs"
        tt = torch.from_numpy(t_sh.astype(np.int64)).to(dev)
        a, b = tt.clone(), tt.clone()
        dist.all_reduce(a, op=dist.ReduceOp.MIN); dist.all_reduce(b, op=dist.ReduceOp.MAX)
        assert torch.equal(a, b), "ranks disagree on the tree"
        # downstream: best-basis coefficients and inverse, shard-local
        coef = wx.getbasiscoefall(yl, t_sh)
        xr = wx.iwptall(coef, wt, t_sh)
        assert (xr - xl).abs().max().item() <= 1e-10 * xl.abs().max().item()
        if rank == 0:
            print(f"{type(method).__name__}: sharded == single (rel {rel:.2e}), tree nodes {int(t_sh.sum())}", flush=True)
    # the raw collectives of the C ABI over the library's own communicator
    import ctypes as C
    cm = wx.dist.comm(dev)
    assert cm is not None
    st = int(torch.cuda.current_stream().cuda_stream)
    b = torch.full((1001,), float(rank + 1), dtype=torch.float64, device=dev)
    wx._lib.call("wx_allreduce", cm, b.data_ptr(), b.numel(), 0, 0, st)
    assert float(b[1000]) == world * (world + 1) / 2
    b = torch.full((5,), float(rank), dtype=torch.float32, device=dev)
    wx._lib.call("wx_allreduce", cm, b.data_ptr(), 5, 1, 2, st)
    assert float(b[0]) == world - 1
    g = torch.empty((world, 3), dtype=torch.int64, device=dev)
    mine = torch.full((3,), rank, dtype=torch.int64, device=dev)
    wx._lib.call("wx_allgather", cm, g.data_ptr(), mine.data_ptr(), 3, 2, st)
    assert g[:, 0].tolist() == list(range(world))
    bb = torch.full((4,), float(rank), dtype=torch.float64, device=dev)
    wx._lib.call("wx_broadcast", cm, bb.data_ptr(), 4, 0, world - 1, st)
    assert float(bb[0]) == world - 1
    # host pipeline: x shard (host) -> wpdall -> JBB tree of the GLOBAL batch -> coefficients (host)
    coef, th = wx.host.wpd_bestbasis_host(xl.cpu().numpy(), wt, None, wx.JBB(), device=local)
    saved = wx.dist.is_dist
    wx.dist.is_dist = lambda group=None: False
    try:
        t_one = wx.bestbasistree(yfull, wx.JBB())
    finally:
        wx.dist.is_dist = saved
    assert np.array_equal(th, t_one) and np.array_equal(coef, wx.getbasiscoefall(yl, t_one).cpu().numpy()), "host pipeline differs"
    # denoiseall: shard-local except for the bestTH summary over the noise levels of the WHOLE batch
    dw = wx.dwtall(X, wt)
    sh = wx.denoiseall(dw[lo:hi].contiguous(), "dwt", wt)
    saved = wx.dist.is_dist
    wx.dist.is_dist = lambda group=None: False
    try:
        one = wx.denoiseall(dw, "dwt", wt, bestTH=np.mean)
        assert torch.equal(sh, wx.denoiseall(dw, "dwt", wt)[lo:hi])
    finally:
        wx.dist.is_dist = saved
    shb = wx.denoiseall(dw[lo:hi].contiguous(), "dwt", wt, bestTH=np.mean)
    assert torch.equal(shb, one[lo:hi]), "sharded denoiseall(bestTH) differs from the single-GPU result"
    dist.barrier()
    wx.dist.destroy_comms()
    if rank == 0:
        print(f"mgpu_check ok on {world} GPUs", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
