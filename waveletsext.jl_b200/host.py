"""Host-array entry points: the call a drop-in user with ordinary (numpy) arrays makes.  The arrays stay on the host;
libwx_b200.so chunks the batch through the GPU with overlapped copies (csrc/wx_host.cu)."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .filters import makereverseqmfpair
from .utils import maxtransformlevels

__all__ = ["wpdall_host", "wpd_bestbasis_host", "pinned_empty", "trim_scratch"]


def pinned_empty(shape, dtype):
    """page-locked host array (numpy view of a pinned torch tensor)"""
    tdt = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32}[np.dtype(dtype)]
    t = torch.empty(tuple(shape), dtype=tdt, pin_memory=True)
    return t.numpy(), t


def wpdall_host(x: np.ndarray, wt, L=None, out: np.ndarray | None = None, chunk: int = 0, device: int | None = None) -> np.ndarray:
    """``wpdall(x, wt, L)`` (dwt/dwt_all.jl:260-282) for HOST arrays: x (N, n) numpy -> (N, L+1, n) numpy."""
    assert isinstance(x, np.ndarray) and x.ndim == 2, "x must be a (N, n) numpy array"
    assert x.flags["C_CONTIGUOUS"], "x must be C-contiguous"
    sfx = {np.dtype(np.float64): "f64", np.dtype(np.float32): "f32"}[x.dtype]
    N, n = x.shape
    L = maxtransformlevels(n) if L is None else int(L)
    assert 0 <= L <= maxtransformlevels(n), "AssertionError: 0 <= L <= maxtransformlevels(x)"
    g, h = makereverseqmfpair(wt, True)
    h = np.ascontiguousarray(h, np.float64); g = np.ascontiguousarray(g, np.float64)
    if out is None:
        out = np.empty((N, L + 1, n), x.dtype)
    assert out.shape == (N, L + 1, n) and out.dtype == x.dtype and out.flags["C_CONTIGUOUS"]
    if device is None:
        device = torch.cuda.current_device()
    with torch.cuda.device(device):
        _lib.call(f"wx_wpdall_host_{sfx}", out.ctypes.data, x.ctypes.data, n, L, N, h.ctypes.data, g.ctypes.data, len(h), int(chunk))
    return out


def wpd_bestbasis_host(x: np.ndarray, wt, L=None, method=None, out: np.ndarray | None = None, chunk: int = 0, device: int | None = None,
                       group=None):
    """``wpdall`` -> ``bestbasistree(., method)`` -> ``getbasiscoefall`` for HOST arrays (the pipeline of paper/paper.md:60-118;
    dwt/dwt_all.jl:260-282, BestBasis.jl:185-217, Utils.jl:169-197): x (N, n) numpy -> (best-basis coefficients (N, n) numpy,
    tree).  One call into the library (``wx_wpd_bestbasis_host``); the packet table never leaves HBM.  With an initialised
    process group x is this rank's shard and the tree is that of the whole batch."""
    import ctypes as C
    from . import bestbasis as B
    from . import dist
    method = B.JBB() if method is None else method
    assert isinstance(method, (B.JBB, B.LSDB)) and not method.redundant, "method must be JBB() or LSDB() on a decimated table"
    assert isinstance(x, np.ndarray) and x.ndim == 2 and x.flags["C_CONTIGUOUS"], "x must be a C-contiguous (N, n) numpy array"
    sfx = {np.dtype(np.float64): "f64", np.dtype(np.float32): "f32"}[x.dtype]
    N, n = x.shape
    L = maxtransformlevels(n) if L is None else int(L)
    assert 0 <= L <= maxtransformlevels(n), "AssertionError: 0 <= L <= maxtransformlevels(x)"
    g, h = makereverseqmfpair(wt, True)
    h = np.ascontiguousarray(h, np.float64); g = np.ascontiguousarray(g, np.float64)
    if out is None:
        out = np.empty((N, n), x.dtype)
    assert out.shape == (N, n) and out.dtype == x.dtype and out.flags["C_CONTIGUOUS"]
    tree = np.zeros(n - 1, np.uint8)
    if device is None:
        device = torch.cuda.current_device()
    isj = isinstance(method, B.JBB)
    with torch.cuda.device(device):
        cm = dist.comm(torch.device("cuda", device), group)
        _lib.call(f"wx_wpd_bestbasis_host_{sfx}", cm, out.ctypes.data, tree.ctypes.data, n - 1, x.ctypes.data, n, L, N, h.ctypes.data,
                  g.ctypes.data, len(h), 0 if isj else 1, B._jbb_kind(method) if isj else 0, C.c_double(float(method.cost.p) if isj else 0.0),
                  int(chunk))
    return out, tree.astype(bool)


def trim_scratch(keep_bytes: int = 0, device: int | None = None) -> None:
    """hand the library's cached stream-ordered scratch above ``keep_bytes`` back to the driver (``wx_trim_scratch``)"""
    import torch
    with torch.cuda.device(torch.cuda.current_device() if device is None else device):
        _lib.call("wx_trim_scratch", int(keep_bytes))
