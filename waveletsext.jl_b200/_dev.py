"""Device-array plumbing: torch CUDA tensors are the minimal device-array type of this host mirror.

torch is used only for device memory, streams and (in dist.py) torch.distributed -- every transform
is a call into libwx_b200.so with raw device pointers.  Arrays follow "Julia memory order": a Julia
array of size (d1, ..., dk) is a C-contiguous tensor of shape (dk, ..., d1) (batch first), so the bytes
are exactly the reference's column-major layout.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def sfx(t: torch.Tensor) -> str:
    if t.dtype == torch.float64:
        return "f64"
    if t.dtype == torch.float32:
        return "f32"
    raise TypeError(f"unsupported element type {t.dtype}: the B200 path supports Float64 and Float32")


def _check(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch CUDA tensor (device array), got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: tensor lives on {t.device}; the B200 path has no CPU fallback")
    sfx(t)


def dev(t: torch.Tensor, name: str = "x") -> torch.Tensor:
    """validate a READ-ONLY device array (CUDA, Float64/Float32); a strided view is packed into a contiguous copy.
    Never use this for an argument the kernel writes: see ``out``.  No silent host fallback."""
    _check(t, name)
    if not t.is_contiguous():
        t = t.contiguous()
    return t


class out:
    """an argument the kernel WRITES (the ``!`` arguments of the reference).  A contiguous tensor is used as it is; a strided
    view (the reference's own ``wpd!`` hands sub-block views to ``dwt_step!``) is packed into a contiguous buffer that starts
    with the view's contents (some kernels accumulate or write only a coset) and ``commit()`` copies the result back into
    the caller's tensor.  ``.t`` is what the kernel gets; ``commit()`` returns the caller's tensor."""

    def __init__(self, t: torch.Tensor, name: str = "y"):
        _check(t, name)
        self.orig = t
        self.t = t if t.is_contiguous() else t.contiguous()

    def commit(self) -> torch.Tensor:
        if self.t is not self.orig:
            self.orig.copy_(self.t)
        return self.orig


def same(a: torch.Tensor, *others: torch.Tensor) -> None:
    for o in others:
        if o.dtype != a.dtype or o.device != a.device:
            raise TypeError("all arrays of one call must share element type and device")


def ptr(t) -> int:
    return int(t.data_ptr()) if t is not None else 0


def stream(t: torch.Tensor) -> int:
    return int(torch.cuda.current_stream(t.device).cuda_stream)


def taps(a) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    assert a.ndim == 1
    return a


def tree_bytes(tree) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(tree).astype(np.uint8))


def call(name: str, t: torch.Tensor, *args) -> None:
    """call wx_<name>_<f64|f32> on t's device"""
    with torch.cuda.device(t.device):
        _lib.call(f"wx_{name}_{sfx(t)}", *args)
