"""GPU tests of the collective layer of the C ABI (wx_comm.cu): communicators, the fused best-basis drivers
(wx_tree_costs_* / wx_bestbasistree_* / wx_bestbasistree_multi_*) and the host-buffer pipeline (wx_wpd_bestbasis_host).
Reference: tree_costs / bestbasistree bestbasis/bestbasis_tree.jl:104-207, BestBasis.jl:185-217.

The 1-rank cases run on any GPU box; the 2-GPU cases (one host thread driving two devices through ncclCommInitAll, and the
one-process-per-GPU torchrun check in tests/mgpu_check.py) need two devices and are skipped otherwise."""
import ctypes as C

import numpy as np
import pytest
import torch

from test_gpu_bestbasis import signals
from test_gpu_dwt import dev

pytestmark = pytest.mark.gpu


def _tree_multi(wx, comms, shards, method, n, K, cost_kind=0, p=2.0, costs=None):
    nd = len(shards)
    arr = (C.c_void_p * nd)(*comms)
    Xp = (C.c_void_p * nd)(*[s.data_ptr() for s in shards])
    Nl = (C.c_long * nd)(*[s.shape[0] for s in shards])
    tree = np.zeros(n - 1, np.uint8)
    sfx = "f64" if shards[0].dtype == torch.float64 else "f32"
    wx._lib.call(f"wx_bestbasistree_multi_{sfx}", arr, nd, method, tree.ctypes.data, n - 1, None if costs is None else costs.ctypes.data,
                 Xp, Nl, 0, n, K, 0, cost_kind, C.c_double(p))
    return tree.astype(bool)


def test_nccl_is_resolved_and_a_one_rank_communicator_is_a_no_op(wx, cuda):
    v = C.c_int()
    wx._lib.call("wx_nccl_version", C.byref(v))
    assert v.value >= 21800
    ident = (C.c_ubyte * 128)()
    wx._lib.call("wx_comm_unique_id", ident)
    h = C.c_void_p()
    with torch.cuda.device(cuda):
        wx._lib.call("wx_comm_init_rank", C.byref(h), ident, 0, 1)
    r, w, d = C.c_int(-1), C.c_int(-1), C.c_int(-1)
    st = C.c_void_p()
    wx._lib.call("wx_comm_info", h, C.byref(r), C.byref(w), C.byref(d), C.byref(st))
    assert (r.value, w.value, d.value) == (0, 1, cuda.index or 0) and st.value
    buf = torch.arange(100, dtype=torch.float64, device=cuda)
    keep = buf.clone()
    for op in (0, 1, 2):
        wx._lib.call("wx_allreduce", h, buf.data_ptr(), 100, 0, op, 0)
    wx._lib.call("wx_broadcast", h, buf.data_ptr(), 100, 0, 0, 0)
    out = torch.empty_like(buf)
    wx._lib.call("wx_allgather", h, out.data_ptr(), buf.data_ptr(), 100, 0, 0)
    torch.cuda.synchronize()
    assert torch.equal(buf, keep) and torch.equal(out, keep)
    with pytest.raises(AssertionError):
        wx._lib.call("wx_allreduce", h, buf.data_ptr(), 100, 9, 0, 0)           # unknown dtype code
    with pytest.raises(AssertionError):
        wx._lib.call("wx_broadcast", h, buf.data_ptr(), 100, 0, 3, 0)           # root outside the communicator
    # the fused drivers with that communicator == without one
    wt = wx.wavelet("db4")
    n, N = 128, 200
    Xw = wx.wpdall(dev(signals(n, N, 3), cuda), wt)
    K = Xw.shape[1]
    for method in (0, 1):
        t0, t1 = np.zeros(n - 1, np.uint8), np.zeros(n - 1, np.uint8)
        c0, c1 = np.empty((1 << K) - 1), np.empty((1 << K) - 1)
        wx._lib.call("wx_bestbasistree_f64", None, method, t0.ctypes.data, n - 1, c0.ctypes.data, Xw.data_ptr(), 0, n, K, N, 0, 0, C.c_double(2.0), 0)
        wx._lib.call("wx_bestbasistree_f64", h, method, t1.ctypes.data, n - 1, c1.ctypes.data, Xw.data_ptr(), 0, n, K, N, 0, 0, C.c_double(2.0), 0)
        assert np.array_equal(t0, t1) and np.array_equal(c0, c1)
    wx._lib.call("wx_comm_destroy", h)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_fused_bestbasistree_equals_costs_plus_selection_and_the_oracle(wx, O, cuda, dt):
    wt = wx.wavelet("db4")
    n, N = 256, 300
    Xw = wx.wpdall(dev(signals(n, N, 11, dt), cuda), wt)
    Xh = Xw.cpu().numpy()
    for method, ref in ((wx.JBB(), O.tree_costs_jbb(Xh)), (wx.JBB(cost=wx.NormCost(1)), O.tree_costs_jbb(Xh, False, cost="norm", p=1.0)),
                        (wx.LSDB(), O.tree_costs_lsdb(Xh))):
        c = wx.tree_costs(Xw, method)
        tol = 1e-9 if isinstance(method, wx.LSDB) else 1e-12
        if dt == np.float32:
            tol = 2e-5
        assert np.abs(c - ref).max() <= tol * np.abs(ref).max()
        t = wx.bestbasistree(Xw, method)
        assert np.array_equal(t, wx.bestbasis_treeselection(c.copy(), n))
        if dt == np.float64:
            assert np.array_equal(t, O.tree_select(ref.copy(), n))
        assert wx.isvalidtree((n,), t)


def test_bestbasistree_multi_on_one_device(wx, cuda):
    """wx_comm_init_all(1) + the *_multi driver on one device == the plain driver (launches on the communicator's stream)"""
    wt = wx.wavelet("db4")
    n, N = 128, 257
    Xw = wx.wpdall(dev(signals(n, N, 5), cuda), wt)
    K = Xw.shape[1]
    comms = (C.c_void_p * 1)()
    devs = (C.c_int * 1)(cuda.index or 0)
    wx._lib.call("wx_comm_init_all", comms, 1, devs)
    torch.cuda.synchronize()
    for method, m in ((0, wx.JBB()), (1, wx.LSDB())):
        costs = np.empty((1 << K) - 1)
        t = _tree_multi(wx, [comms[0]], [Xw], method, n, K, costs=costs)
        assert np.array_equal(t, wx.bestbasistree(Xw, m))
        assert np.array_equal(costs, wx.tree_costs(Xw, m))
    with pytest.raises(AssertionError):
        bad = np.zeros(n, np.uint8)
        arr = (C.c_void_p * 1)(comms[0]); Xp = (C.c_void_p * 1)(Xw.data_ptr()); Nl = (C.c_long * 1)(N)
        wx._lib.call("wx_bestbasistree_multi_f64", arr, 1, 0, bad.ctypes.data, n, None, Xp, Nl, 0, n, K, 0, 0, C.c_double(2.0))   # wrong tree length
    wx._lib.call("wx_comm_destroy", comms[0])


def test_empty_shard_and_argument_errors(wx, cuda):
    n, K = 64, 7
    t = np.zeros(n - 1, np.uint8)
    with pytest.raises(AssertionError):          # empty batch
        wx._lib.call("wx_bestbasistree_f64", None, 0, t.ctypes.data, n - 1, None, None, 0, n, K, 0, 0, 0, C.c_double(2.0), 0)
    X = torch.randn((5, K, n), dtype=torch.float64, device=cuda)
    with pytest.raises(AssertionError):          # unknown method
        wx._lib.call("wx_bestbasistree_f64", None, 2, t.ctypes.data, n - 1, None, X.data_ptr(), 0, n, K, 5, 0, 0, C.c_double(2.0), 0)
    X1 = torch.randn((1, K, n), dtype=torch.float64, device=cuda)
    with pytest.raises(AssertionError):          # LSDB needs two signals
        wx._lib.call("wx_bestbasistree_f64", None, 1, t.ctypes.data, n - 1, None, X1.data_ptr(), 0, n, K, 1, 0, 0, C.c_double(2.0), 0)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_host_pipeline_wpd_bestbasis(wx, O, cuda, dt):
    """wx_wpd_bestbasis_host: x (host) -> wpdall -> bestbasistree -> getbasiscoefall -> coefficients (host) in one call ==
    the three device-resident calls, chunked or not"""
    wt = wx.wavelet("db4")
    n, N, L = 256, 333, 6
    x = signals(n, N, 8, dt)
    Xw = wx.wpdall(dev(x, cuda), wt, L)
    for method in (wx.JBB(), wx.LSDB()):
        tree = wx.bestbasistree(Xw, method)
        ref = wx.getbasiscoefall(Xw, tree).cpu().numpy()
        for chunk in (0, 50, 333, 1000):
            coef, t = wx.host.wpd_bestbasis_host(x, wt, L, method, chunk=chunk, device=cuda.index or 0)
            assert np.array_equal(t, tree), (type(method).__name__, chunk)
            assert np.array_equal(coef, ref)
    if dt == np.float64:
        assert np.array_equal(wx.host.wpd_bestbasis_host(x, wt, L)[1], O.tree_select(O.tree_costs_jbb(Xw.cpu().numpy()), n))
    with pytest.raises(AssertionError):
        wx.host.wpd_bestbasis_host(x, wt, 9)                         # L > maxtransformlevels


def test_two_devices_from_one_host_thread(wx, cuda):
    """ncclCommInitAll: one host thread, two devices; the sharded tree (JBB: one grouped all-reduce; LSDB: grouped
    all-gathers / all-reduces) equals the single-GPU tree of the concatenated batch, and so do the costs."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    wt = wx.wavelet("db4")
    n, N = 256, 1000
    x = signals(n, N, 77)
    X0 = wx.wpdall(dev(x, cuda), wt)
    K = X0.shape[1]
    cut = 377
    sh0 = X0[:cut].contiguous()
    sh1 = X0[cut:].to("cuda:1").contiguous()
    comms = (C.c_void_p * 2)()
    wx._lib.call("wx_comm_init_all", comms, 2, None)
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    for method, m, tol in ((0, wx.JBB(), 1e-12), (1, wx.LSDB(), 1e-12)):
        costs = np.empty((1 << K) - 1)
        t = _tree_multi(wx, [comms[0], comms[1]], [sh0, sh1], method, n, K, costs=costs)
        one = wx.tree_costs(X0, m)
        assert np.abs(costs - one).max() <= tol * np.abs(one).max(), type(m).__name__
        assert np.array_equal(t, wx.bestbasistree(X0, m)), type(m).__name__
    # raw grouped all-reduce from one thread
    a0 = torch.full((1000,), 1.5, dtype=torch.float64, device="cuda:0")
    a1 = torch.full((1000,), 2.25, dtype=torch.float64, device="cuda:1")
    wx._lib.call("wx_group_start")
    for c, a, d in ((comms[0], a0, 0), (comms[1], a1, 1)):
        with torch.cuda.device(d):
            wx._lib.call("wx_allreduce", c, a.data_ptr(), 1000, 0, 0, int(torch.cuda.current_stream(d).cuda_stream))
    wx._lib.call("wx_group_end")
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    assert float(a0[7]) == 3.75 and float(a1[999]) == 3.75
    for c in comms:
        wx._lib.call("wx_comm_destroy", c)
