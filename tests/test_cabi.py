"""The C-ABI library loads and exports every symbol include/wx_b200.h declares; without a GPU every compute entry point
fails loudly (no CPU fallback).  CPU only: no compute call is expected to succeed here."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported(wx):
    from waveletsext_b200 import _lib
    protos = _lib.parse_header()
    assert len(protos) >= 60
    l = C.CDLL(_lib.LIBPATH)
    missing = [n for n in protos if not hasattr(l, n)]
    assert not missing, f"libwx_b200.so lacks {missing}"
    for stem in ("wx_wpd1d", "wx_wpd2d", "wx_iwpt1d", "wx_rwt", "wx_irwt", "wx_dwt_step", "wx_idwt_step", "wx_sdwt_step", "wx_acdwt_step",
                 "wx_gather_basis", "wx_jbb_moments", "wx_lsdb_pass1", "wx_wpdall_host", "wx_noisest", "wx_surethreshold",
                 "wx_relerrorthreshold", "wx_threshold", "wx_sidwt_step", "wx_isidwt_step", "wx_ns_dwt", "wx_ns_idwt", "wx_select_features",
                 "wx_scatter_features", "wx_class_moments"):
        assert stem + "_f64" in protos and stem + "_f32" in protos
    assert _lib.lib().wx_version() >= 100


def test_library_is_sm100a_only():
    """the product is built for sm_100a and nothing else (no multi-arch dispatch)"""
    import shutil, subprocess
    so = os.path.join(ROOT, "waveletsext.jl_b200", "libwx_b200.so")
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", so], capture_output=True, text=True).stdout
    archs = {ln.split(".")[-2] for ln in out.splitlines() if ".cubin" in ln}
    assert archs == {"sm_100a"}, archs


def test_argument_validation_needs_no_gpu(wx):
    """argument errors the reference raises with @assert are detected before any CUDA work"""
    from waveletsext_b200 import _lib
    h = np.ones(8); g = np.ones(8)
    rc = _lib.lib().wx_wpd1d_f64(0, 0, 24, 4, 3, h.ctypes.data, g.ctypes.data, 8, 0)     # L > maxtransformlevels(24) = 3
    assert rc == _lib.WX_EINVAL and "maxtransformlevels" in _lib.last_error()
    rc = _lib.lib().wx_rwt_f64(0, 2, 0, 0, 0, 16, 0, 3, h.ctypes.data, g.ctypes.data, 8, 0)
    assert rc == _lib.WX_EINVAL and "L must be >= 1" in _lib.last_error()
    buf = np.zeros(8)
    rc = _lib.lib().wx_isdwt_step_shift_f64(buf.ctypes.data, buf.ctypes.data, buf.ctypes.data, 8, 0, 1, 0, h.ctypes.data, g.ctypes.data, 8, 0, 0)
    assert rc == _lib.WX_EINVAL and "sv" in _lib.last_error()
    # empty batch: nothing to do, success even without a device
    assert _lib.lib().wx_wpd1d_f64(0, 0, 16, 4, 0, h.ctypes.data, g.ctypes.data, 8, 0) == 0
    # the "next" rows validate the same way: threshold type / negative threshold (Wavelets.jl @assert t >= 0), elbows >= 1
    # (Denoising.jl:291), ns_dwt's 1 <= L <= Lmax and ispow2(n) (wavemult/transforms.jl:57-58), a noisest range outside the slab
    L = _lib.lib()
    p = buf.ctypes.data
    rc = L.wx_threshold_f64(p, p, 8, 1, 0, 0, 0, 7, 0, C.c_double(1.0), 1, 0)
    assert rc == _lib.WX_EINVAL and "threshold type" in _lib.last_error()
    rc = L.wx_threshold_f64(p, p, 8, 1, 0, 0, 0, 0, 0, C.c_double(-1.0), 1, 0)
    assert rc == _lib.WX_EINVAL and "t >= 0" in _lib.last_error()
    rc = L.wx_relerrorthreshold_f64(p, p, 8, 1, 0, 0, 1, 0)
    assert rc == _lib.WX_EINVAL and "elbows" in _lib.last_error()
    rc = L.wx_noisest_f64(p, p, 8, 4, 8, 1, 0)
    assert rc == _lib.WX_EINVAL
    rc = L.wx_ns_dwt_f64(p, p, 8, 4, 1, h.ctypes.data, g.ctypes.data, 8, 0)
    assert rc == _lib.WX_EINVAL and "1 <= L <= Lmax" in _lib.last_error()
    rc = L.wx_ns_dwt_f64(p, p, 12, 1, 1, h.ctypes.data, g.ctypes.data, 8, 0)
    assert rc == _lib.WX_EINVAL and "ispow2" in _lib.last_error()
    assert L.wx_threshold_f64(p, p, 8, 1, 0, 0, 0, 0, 0, C.c_double(1.0), 0, 0) == 0          # N = 0


def test_compute_fails_loudly_without_gpu(wx):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from waveletsext_b200 import _lib
    x = np.zeros(64); y = np.zeros(64 * 3)
    h = np.ones(2); g = np.ones(2)
    rc = _lib.lib().wx_wpd1d_f64(y.ctypes.data, x.ctypes.data, 64, 2, 1, h.ctypes.data, g.ctypes.data, 2, 0)
    assert rc == _lib.WX_ECUDA
    with pytest.raises(_lib.WxError):
        _lib.check(rc)
    with pytest.raises((_lib.WxError, RuntimeError, AssertionError)):
        wx.host.wpdall_host(np.zeros((2, 8)), wx.wavelet("haar"), device=0)


def test_collective_entry_points_validate_without_a_gpu(wx):
    """wx_comm.cu: NCCL resolves by dlopen (no link-time dependency), argument errors are reported before any CUDA / NCCL work,
    and without a device the communicator constructors and the fused best-basis drivers fail loudly"""
    import subprocess, shutil
    import torch
    from waveletsext_b200 import _lib
    L = _lib.lib()
    protos = _lib.parse_header()
    for nm in ("wx_nccl_version", "wx_comm_unique_id", "wx_comm_init_rank", "wx_comm_init_all", "wx_comm_info", "wx_comm_destroy",
               "wx_group_start", "wx_group_end", "wx_allreduce", "wx_allgather", "wx_broadcast", "wx_tree_costs_jbb_f64",
               "wx_tree_costs_jbb_f32", "wx_tree_costs_lsdb_f64", "wx_tree_costs_lsdb_f32", "wx_bestbasistree_f64", "wx_bestbasistree_f32",
               "wx_bestbasistree_multi_f64", "wx_bestbasistree_multi_f32", "wx_wpd_bestbasis_host_f64", "wx_wpd_bestbasis_host_f32"):
        assert nm in protos, nm
    # no DT_NEEDED on NCCL: a single-GPU host never needs it
    readelf = shutil.which("readelf")
    if readelf:
        needed = subprocess.run([readelf, "-d", _lib.LIBPATH], capture_output=True, text=True).stdout
        assert "nccl" not in needed.lower()
    v = C.c_int(0)
    assert L.wx_nccl_version(C.byref(v)) == 0 and v.value >= 21800                       # the libnccl.so.2 torch mapped
    ident = (C.c_ubyte * 128)()
    assert L.wx_comm_unique_id(ident) == 0 and any(bytes(ident))
    h = C.c_void_p()
    assert L.wx_comm_init_rank(C.byref(h), ident, 2, 2) == _lib.WX_EINVAL and "rank" in _lib.last_error()
    assert L.wx_comm_init_rank(None, ident, 0, 1) == _lib.WX_EINVAL
    assert L.wx_allreduce(None, None, 4, 0, 0, None) == _lib.WX_EINVAL
    assert L.wx_allgather(None, None, None, 4, 0, None) == _lib.WX_EINVAL
    assert L.wx_broadcast(None, None, 4, 0, 0, None) == _lib.WX_EINVAL
    assert L.wx_comm_destroy(None) == 0
    assert L.wx_comm_info(None, None, None, None, None) == _lib.WX_EINVAL
    tree = np.zeros(63, np.uint8)
    # argument checks of the fused drivers come before any device work
    rc = L.wx_bestbasistree_f64(None, 2, tree.ctypes.data, 63, None, None, 0, 64, 7, 0, 0, 0, C.c_double(2.0), None)
    assert rc == _lib.WX_EINVAL and "method" in _lib.last_error()
    rc = L.wx_bestbasistree_f64(None, 0, tree.ctypes.data, 62, None, None, 0, 64, 7, 0, 0, 0, C.c_double(2.0), None)
    assert rc == _lib.WX_EINVAL and "tree buffer" in _lib.last_error()
    rc = L.wx_bestbasistree_f64(None, 0, None, 63, None, None, 0, 64, 7, 0, 0, 0, C.c_double(2.0), None)
    assert rc == _lib.WX_EINVAL
    rc = L.wx_bestbasistree_multi_f64(None, 2, 0, tree.ctypes.data, 63, None, None, None, 0, 64, 7, 0, 0, C.c_double(2.0))
    assert rc == _lib.WX_EINVAL
    coef = np.zeros((2, 64)); x = np.zeros((2, 64)); hh = np.ones(2); gg = np.ones(2)
    rc = L.wx_wpd_bestbasis_host_f64(None, coef.ctypes.data, tree.ctypes.data, 63, x.ctypes.data, 64, 9, 2, hh.ctypes.data, gg.ctypes.data, 2, 0, 0,
                                      C.c_double(2.0), 0)
    assert rc == _lib.WX_EINVAL and "maxtransformlevels" in _lib.last_error()
    rc = L.wx_wpd_bestbasis_host_f64(None, coef.ctypes.data, tree.ctypes.data, 60, x.ctypes.data, 64, 3, 2, hh.ctypes.data, gg.ctypes.data, 2, 0, 0,
                                      C.c_double(2.0), 0)
    assert rc == _lib.WX_EINVAL and "tree buffer" in _lib.last_error()
    if not torch.cuda.is_available():
        X = np.zeros((4, 7, 64))
        costs = np.zeros(127)
        assert L.wx_tree_costs_jbb_f64(None, costs.ctypes.data, X.ctypes.data, 0, 64, 7, 4, 0, 0, C.c_double(2.0), None) == _lib.WX_ECUDA
        assert L.wx_comm_init_rank(C.byref(h), ident, 0, 1) == _lib.WX_ECUDA
        comms = (C.c_void_p * 1)()
        assert L.wx_comm_init_all(comms, 1, None) == _lib.WX_ECUDA
        rc = L.wx_wpd_bestbasis_host_f64(None, coef.ctypes.data, tree.ctypes.data, 63, x.ctypes.data, 64, 3, 2, hh.ctypes.data, gg.ctypes.data, 2, 0, 0,
                                          C.c_double(2.0), 0)
        assert rc == _lib.WX_ECUDA
