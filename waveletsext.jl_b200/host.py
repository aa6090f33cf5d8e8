"""Host-array entry points: the call a drop-in user with ordinary (numpy) arrays makes.  The arrays stay on the host;
libwx_b200.so chunks the batch through the GPU with overlapped copies (csrc/wx_host.cu)."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .filters import makereverseqmfpair
from .utils import maxtransformlevels

__all__ = ["wpdall_host", "pinned_empty", "trim_scratch"]


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


def trim_scratch(keep_bytes: int = 0, device: int | None = None) -> None:
    """hand the library's cached stream-ordered scratch above ``keep_bytes`` back to the driver (``wx_trim_scratch``)"""
    import torch
    with torch.cuda.device(torch.cuda.current_device() if device is None else device):
        _lib.call("wx_trim_scratch", int(keep_bytes))
