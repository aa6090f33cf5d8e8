"""Stationary (undecimated, a-trous) transforms: host mirror of SWT.jl, swt/swt_one_level.jl, swt/swt_all.jl.
``f!`` is spelled ``f_``.  Arrays: torch CUDA tensors in Julia memory order (see _dev.py)."""
from __future__ import annotations

import math

import torch

from . import _dev as D
from . import _rwt as R
from .utils import maxtransformlevels

__all__ = ["sdwt_step", "sdwt_step_", "isdwt_step", "isdwt_step_", "sdwt", "sdwt_", "isdwt", "isdwt_", "swpt", "swpt_",
           "iswpt", "iswpt_", "swpd", "swpd_", "iswpd", "iswpd_", "sdwtall", "swptall", "swpdall", "isdwtall", "iswptall",
           "iswpdall"]


# ---------------------------------------------------------------- single steps
def sdwt_step_(w1, w2, *rest):
    """``sdwt_step!(w1, w2, v, d, h, g)`` swt/swt_one_level.jl:99-127 / 2-D ``(w1,w2,w3,w4,v,d,h,g,temp)`` :334-370"""
    if len(rest) == 4:
        v, d, h, g = rest
        v, o1, o2 = D.dev(v, "v"), D.out(w1, "w1"), D.out(w2, "w2")
        D.same(v, w1, w2)
        assert w1.numel() == w2.numel() == v.numel(), "AssertionError: length(w1) == length(w2) == length(v)"
        assert len(h) == len(g), "AssertionError: length(h) == length(g)"
        h, g = D.taps(h), D.taps(g)
        D.call("sdwt_step", v, D.ptr(o1.t), D.ptr(o2.t), D.ptr(v), v.numel(), int(d), h.ctypes.data, g.ctypes.data, len(h), D.stream(v))
        return o1.commit(), o2.commit()
    w3, w4, v, d, h, g = rest[:6]
    return _rstep2(0, w1, w2, w3, w4, v, d, h, g)


def _rstep2(ac, w1, w2, w3, w4, v, d, h, g):
    v = D.dev(v, "v")
    ws = [D.out(w, "w") for w in (w1, w2, w3, w4)]
    D.same(v, w1, w2, w3, w4)
    assert all(w.shape == v.shape for w in (w1, w2, w3, w4)), "AssertionError: size(v) == size(w1) == size(w2) == size(w3) == size(w4)"
    nc, nr = v.shape
    h, g = D.taps(h), D.taps(g)
    D.call("rdwt_step2", v, ac, *[D.ptr(w.t) for w in ws], D.ptr(v), nr, nc, int(d), h.ctypes.data, g.ctypes.data, len(h), D.stream(v))
    return tuple(w.commit() for w in ws)


def sdwt_step(v, d, h, g):
    """``sdwt_step(v, d, h, g)`` swt/swt_one_level.jl:40-48, :323-332"""
    v = D.dev(v, "v")
    if v.dim() == 1:
        return sdwt_step_(torch.empty_like(v), torch.empty_like(v), v, d, h, g)
    return sdwt_step_(*[torch.empty_like(v) for _ in range(4)], v, d, h, g)


def isdwt_step_(v, *rest, add2out=False):
    """``isdwt_step!(v, w1, w2, d, h, g)`` (average, :257-277) / ``(v, w1, w2, d, sv, sw, h, g; add2out)`` (shift, :279-318)
    and the 2-D forms ``(v, w1..w4, d, h, g, temp)`` :395-431 / ``(v, w1..w4, d, sv, sw, h, g, temp)`` :433-469"""
    ov = D.out(v, "v")
    if v.dim() == 1:
        w1, w2 = D.dev(rest[0], "w1"), D.dev(rest[1], "w2")
        D.same(v, w1, w2)
        if len(rest) == 5:
            d, h, g = rest[2:]
            h, g = D.taps(h), D.taps(g)
            D.call("isdwt_step_avg", v, D.ptr(ov.t), D.ptr(w1), D.ptr(w2), v.numel(), int(d), h.ctypes.data, g.ctypes.data, len(h), D.stream(v))
            ov.commit()
            return None                      # the reference's average-based method returns nothing
        d, sv, sw, h, g = rest[2:7]
        h, g = D.taps(h), D.taps(g)
        D.call("isdwt_step_shift", v, D.ptr(ov.t), D.ptr(w1), D.ptr(w2), v.numel(), int(d), int(sv), int(sw), h.ctypes.data, g.ctypes.data, len(h),
               int(bool(add2out)), D.stream(v))
        return ov.commit()
    ws = [D.dev(w, "w") for w in rest[:4]]
    D.same(v, *ws)
    nc, nr = v.shape
    tail = [a for a in rest[4:] if not isinstance(a, torch.Tensor)]
    if len(tail) == 3:
        d, h, g = tail
        mode, sv, sw = 0, 0, 0
    else:
        d, sv, sw, h, g = tail[:5]
        mode = 1
    h, g = D.taps(h), D.taps(g)
    D.call("irdwt_step2", v, mode, D.ptr(ov.t), *[D.ptr(w) for w in ws], nr, nc, int(d), int(sv), int(sw), h.ctypes.data, g.ctypes.data, len(h),
           D.stream(v))
    return ov.commit()


def isdwt_step(*args):
    """``isdwt_step(w1, w2, d, h, g)`` / ``(w1, w2, d, sv, sw, h, g)`` and the 2-D forms with four children"""
    w1 = D.dev(args[0], "w1")
    v = torch.zeros_like(w1)
    isdwt_step_(v, *args)
    return v


# ---------------------------------------------------------------- single-signal trees
def _single(fn, x, *a, **k):
    return fn(x.unsqueeze(0), *a, **k)[0]


def sdwt_(xw, x, wt, L=None):
    """``sdwt!(xw, x, wt, L)`` SWT.jl:109-158"""
    R.forward(False, "dwt", x.unsqueeze(0), wt, L, xw.unsqueeze(0)); return xw
def sdwt(x, wt, L=None):
    return _single(lambda b: R.forward(False, "dwt", b, wt, L), D.dev(x, "x"))
def swpt_(xw, x, wt, L=None):
    """``swpt!(xw, x, wt, L)`` SWT.jl:439-513"""
    R.forward(False, "wpt", x.unsqueeze(0), wt, L, xw.unsqueeze(0)); return xw
def swpt(x, wt, L=None):
    return _single(lambda b: R.forward(False, "wpt", b, wt, L), D.dev(x, "x"))
def swpd_(xw, x, wt, L=None):
    """``swpd!(xw, x, wt, L)`` SWT.jl:840-902"""
    R.forward(False, "wpd", x.unsqueeze(0), wt, L, xw.unsqueeze(0)); return xw
def swpd(x, wt, L=None):
    return _single(lambda b: R.forward(False, "wpd", b, wt, L), D.dev(x, "x"))


def _check_sm_dwt(xw_single, sm, two):
    """isdwt! asserts: 1-D ``0 <= log2(sm) < L`` (SWT.jl:266), 2-D ``0 <= log2(sm) <= L`` (:293); sm = 0 fails (log2(0) = -Inf)"""
    k = xw_single.shape[0]
    L = (k - 1) // 3 if two else k - 1
    lg = math.log2(sm) if sm > 0 else -math.inf
    assert (0 <= lg <= L) if two else (0 <= lg < L), "AssertionError: 0 <= log2(sm) < L"


def isdwt_(x, xw, wt, sm=None):
    """``isdwt!(x, xw, wt[, sm])`` SWT.jl:259-358"""
    if sm is not None:
        _check_sm_dwt(xw, sm, xw.dim() == 3)
    R.inverse(False, "dwt", xw.unsqueeze(0), wt, None, sm, x.unsqueeze(0)); return x
def isdwt(xw, wt, sm=None):
    xw = D.dev(xw, "xw")
    return isdwt_(xw.new_empty(tuple(xw.shape[1:])), xw, wt, sm)
def iswpt_(x, xw, wt, sm=None):
    """``iswpt!(x, xw, wt[, sm])`` SWT.jl:613-758"""
    R.inverse(False, "wpt", xw.unsqueeze(0), wt, None, sm, x.unsqueeze(0)); return x
def iswpt(xw, wt, sm=None):
    xw = D.dev(xw, "xw")
    return iswpt_(xw.new_empty(tuple(xw.shape[1:])), xw, wt, sm)


def iswpd_(x, xw, wt, arg=None, sm=None):
    """``iswpd!(x, xw, wt, L|tree[, sm])`` SWT.jl:1035-1199"""
    tree = R.wpd_tree(R.sig_shape(x), xw.shape[0], arg)
    R.inverse(False, "wpd", xw.unsqueeze(0), wt, tree, sm, x.unsqueeze(0)); return x
def iswpd(xw, wt, arg=None, sm=None):
    xw = D.dev(xw, "xw")
    return iswpd_(xw.new_empty(tuple(xw.shape[1:])), xw, wt, arg, sm)


# ---------------------------------------------------------------- batches (swt/swt_all.jl)
def _assert_batch(x):
    assert x.dim() > 1, "AssertionError: ndims(x) > 1"
def sdwtall(x, wt, L=None):
    """swt/swt_all.jl:33-53"""
    _assert_batch(x); return R.forward(False, "dwt", x, wt, L)
def swptall(x, wt, L=None):
    """swt/swt_all.jl:156-176"""
    _assert_batch(x); return R.forward(False, "wpt", x, wt, L)
def swpdall(x, wt, L=None):
    """swt/swt_all.jl:279-296"""
    _assert_batch(x); return R.forward(False, "wpd", x, wt, L)
def isdwtall(xw, wt, sm=None):
    """swt/swt_all.jl:89-123"""
    if sm is not None:
        _check_sm_dwt(xw[0], sm, xw.dim() == 4)
    return R.inverse(False, "dwt", xw, wt, None, sm)
def iswptall(xw, wt, sm=None):
    """swt/swt_all.jl:212-246"""
    return R.inverse(False, "wpt", xw, wt, None, sm)
def iswpdall(xw, wt, arg=None, sm=None):
    """swt/swt_all.jl:343-390"""
    xw = D.dev(xw, "xw")
    assert 3 <= xw.dim() <= 4, "AssertionError: 3 <= ndims(xw) <= 4"
    shp = tuple(reversed(tuple(xw.shape[2:])))
    return R.inverse(False, "wpd", xw, wt, R.wpd_tree(shp, xw.shape[1], arg), sm)
