"""Shift-invariant transform steps and the nonstandard-form transform -- row f-4 of SURVEY.md section 8: the variants of
``dwt_step!`` / ``idwt_step!`` used by SIWT.jl (``sidwt_step!`` / ``isidwt_step!``, siwt/siwt_one_level.jl) and by WaveMult.jl
(``ns_dwt`` / ``ns_idwt``, wavemult/transforms.jl).  The SIWT node dictionary / best-basis search and the sparse-matrix
multiplication built on top of them stay with the reference.
"""
from __future__ import annotations

import torch

from . import _dev as D
from .dwt import _pair
from .utils import maxtransformlevels

__all__ = ["sidwt_step_", "isidwt_step_", "ndyad", "ns_dwt", "ns_idwt"]


def sidwt_step_(w1, w2, v, h, g, s: bool):
    """``sidwt_step!(w1, w2, v, h, g, s)`` siwt/siwt_one_level.jl:71-98"""
    v, o1, o2 = D.dev(v, "v"), D.out(w1, "w1"), D.out(w2, "w2")
    D.same(v, w1, w2)
    assert w1.numel() == w2.numel() == v.numel() // 2, "AssertionError: length(w1) == length(w2) == length(v)/2"
    assert len(h) == len(g), "AssertionError: length(h) == length(g)"
    h, g = D.taps(h), D.taps(g)
    D.call("sidwt_step", v, D.ptr(o1.t), D.ptr(o2.t), D.ptr(v), v.numel(), h.ctypes.data, g.ctypes.data, len(h), int(bool(s)), D.stream(v))
    return o1.commit(), o2.commit()


def isidwt_step_(v, w1, w2, h, g, s: bool):
    """``isidwt_step!(v, w1, w2, h, g, s)`` siwt/siwt_one_level.jl:153-184"""
    ov, w1, w2 = D.out(v, "v"), D.dev(w1, "w1"), D.dev(w2, "w2")
    D.same(v, w1, w2)
    assert w1.numel() == w2.numel() == v.numel() // 2, "AssertionError: length(w1) == length(w2) == length(v)/2"
    assert len(h) == len(g), "AssertionError: length(h) == length(g)"
    h, g = D.taps(h), D.taps(g)
    D.call("isidwt_step", v, D.ptr(ov.t), D.ptr(w1), D.ptr(w2), v.numel(), h.ctypes.data, g.ctypes.data, len(h), int(bool(s)), D.stream(v))
    return ov.commit()


def ndyad(L: int, Lmax: int, gender: bool) -> range:
    """wavemult/utils.jl:146-155 (0-based half-open range)"""
    assert L <= Lmax, "AssertionError: L <= Lmax"
    assert L >= 1, "AssertionError: L >= 1"
    k = Lmax - L
    if gender:
        return range((1 << (k + 1)) + (1 << k), 1 << (k + 2))
    return range(1 << (k + 1), (1 << (k + 1)) + (1 << k))


def _ispow2(n):
    return n > 0 and n & (n - 1) == 0


def ns_dwt(x, wt, L=None):
    """``ns_dwt(x, wt[, L])`` wavemult/transforms.jl:52-74: x (n,) -> nxw (2n,).  A batch (N, n) gives (N, 2n)."""
    x = D.dev(x, "x")
    single = x.dim() == 1
    X = x.unsqueeze(0) if single else x
    assert X.dim() == 2
    N, n = X.shape
    Lmax = maxtransformlevels(n)
    L = Lmax if L is None else int(L)
    assert 1 <= L <= Lmax, "AssertionError: 1 <= L <= Lmax"
    assert _ispow2(n), "AssertionError: ispow2(n)"
    h, g = _pair(wt)
    out = X.new_empty((N, 2 * n))
    D.call("ns_dwt", X, D.ptr(out), D.ptr(X), n, L, N, h.ctypes.data, g.ctypes.data, len(h), D.stream(X))
    return out[0] if single else out


def ns_idwt(nxw, wt, L=None):
    """``ns_idwt(nxw, wt[, L])`` wavemult/transforms.jl:120-139: nxw (2n,) -> x (n,)"""
    nxw = D.dev(nxw, "nxw")
    single = nxw.dim() == 1
    W = nxw.unsqueeze(0) if single else nxw
    assert W.dim() == 2
    N, n2 = W.shape
    Lmax = maxtransformlevels(n2) - 1
    L = Lmax if L is None else int(L)
    n = n2 // 2
    assert 1 <= L <= Lmax, "AssertionError: 1 <= L <= Lmax"
    assert _ispow2(n), "AssertionError: ispow2(n)"
    h, g = _pair(wt)
    out = W.new_empty((N, n))
    D.call("ns_idwt", W, D.ptr(out), D.ptr(W), n2, L, N, h.ctypes.data, g.ctypes.data, len(h), D.stream(W))
    return out[0] if single else out
