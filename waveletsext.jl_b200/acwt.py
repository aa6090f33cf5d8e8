"""Autocorrelation wavelet transforms: host mirror of ACWT.jl, acwt/acwt_one_level.jl, acwt/acwt_utils.jl, acwt/acwt_all.jl.
``f!`` is spelled ``f_``.  Arrays: torch CUDA tensors in Julia memory order (see _dev.py).
Unlike the reference's 1-D ``acdwt_step!`` (acwt/acwt_one_level.jl:101-106, ``h::Array{T,1}``) Float32 signals are
accepted: taps are rounded to Float32 once."""
from __future__ import annotations

import torch

from . import _dev as D
from . import _rwt as R
from .filters import autocorr, pfilter, qfilter, make_acqmfpair, make_acreverseqmfpair   # noqa: F401 (re-exported)
from .swt import _rstep2

__all__ = ["acdwt_step", "acdwt_step_", "iacdwt_step", "iacdwt_step_", "acdwt", "acdwt_", "iacdwt", "iacdwt_", "acwpt", "acwpt_",
           "iacwpt", "iacwpt_", "acwpd", "acwpd_", "iacwpd", "iacwpd_", "acdwtall", "acwptall", "acwpdall", "iacdwtall",
           "iacwptall", "iacwpdall", "autocorr", "pfilter", "qfilter", "make_acqmfpair", "make_acreverseqmfpair"]


def acdwt_step_(w1, w2, *rest):
    """``acdwt_step!(w1, w2, v, d, h, g)`` acwt/acwt_one_level.jl:101-128 / 2-D ``(w1..w4, v, d, h, g, temp)`` :240-276"""
    if len(rest) == 4:
        v, d, h, g = rest
        v, o1, o2 = D.dev(v, "v"), D.out(w1, "w1"), D.out(w2, "w2")
        D.same(v, w1, w2)
        assert w1.numel() == w2.numel() == v.numel(), "AssertionError: length(w1) == length(w2) == length(v)"
        assert len(h) == len(g), "AssertionError: length(h) == length(g)"
        h, g = D.taps(h), D.taps(g)
        D.call("acdwt_step", v, D.ptr(o1.t), D.ptr(o2.t), D.ptr(v), v.numel(), int(d), h.ctypes.data, g.ctypes.data, len(h), D.stream(v))
        return o1.commit(), o2.commit()
    w3, w4, v, d, h, g = rest[:6]
    return _rstep2(1, w1, w2, w3, w4, v, d, h, g)


def acdwt_step(v, d, h, g):
    """``acdwt_step(v, d, h, g)`` acwt/acwt_one_level.jl:44-52, :229-238"""
    v = D.dev(v, "v")
    if v.dim() == 1:
        return acdwt_step_(torch.empty_like(v), torch.empty_like(v), v, d, h, g)
    return acdwt_step_(*[torch.empty_like(v) for _ in range(4)], v, d, h, g)


def iacdwt_step_(v, *ws):
    """``iacdwt_step!(v, w1, w2)`` acwt/acwt_one_level.jl:217-224 / 2-D ``(v, w1..w4, temp)`` :288-322"""
    ov = D.out(v, "v")
    ws = [D.dev(w, "w") for w in ws[:4] if isinstance(w, torch.Tensor)]
    D.same(v, *ws)
    assert all(w.shape == v.shape for w in ws), "AssertionError: length(v) == length(w1) == length(w2)"
    if v.dim() == 1:
        D.call("iacdwt_step", v, D.ptr(ov.t), D.ptr(ws[0]), D.ptr(ws[1]), v.numel(), D.stream(v))
        return ov.commit()
    nc, nr = v.shape
    D.call("irdwt_step2", v, 2, D.ptr(ov.t), *[D.ptr(w) for w in ws[:4]], nr, nc, 0, 0, 0, 0, 0, 0, D.stream(v))
    return ov.commit()


def iacdwt_step(*ws):
    w1 = D.dev(ws[0], "w1")
    return iacdwt_step_(torch.empty_like(w1), *ws)


def _single(fn, x):
    return fn(x.unsqueeze(0))[0]


def acdwt_(xw, x, wt, L=None):
    """``acdwt!(xw, x, wt, L)`` ACWT.jl:109-157"""
    R.forward(True, "dwt", x.unsqueeze(0), wt, L, xw.unsqueeze(0)); return xw
def acdwt(x, wt, L=None):
    return _single(lambda b: R.forward(True, "dwt", b, wt, L), D.dev(x, "x"))
def acwpt_(xw, x, wt, L=None):
    """``acwpt!(xw, x, wt, L)`` ACWT.jl:427-501"""
    R.forward(True, "wpt", x.unsqueeze(0), wt, L, xw.unsqueeze(0)); return xw
def acwpt(x, wt, L=None):
    return _single(lambda b: R.forward(True, "wpt", b, wt, L), D.dev(x, "x"))
def acwpd_(xw, x, wt, L=None):
    """``acwpd!(xw, x, wt, L)`` ACWT.jl:733-793"""
    R.forward(True, "wpd", x.unsqueeze(0), wt, L, xw.unsqueeze(0)); return xw
def acwpd(x, wt, L=None):
    return _single(lambda b: R.forward(True, "wpd", b, wt, L), D.dev(x, "x"))


def iacdwt_(x, xw, wt=None):
    """``iacdwt!(x, xw[, wt])`` ACWT.jl:287-329"""
    R.inverse(True, "dwt", xw.unsqueeze(0), None, None, None, x.unsqueeze(0)); return x
def iacdwt(xw, wt=None):
    xw = D.dev(xw, "xw")
    return iacdwt_(xw.new_empty(tuple(xw.shape[1:])), xw)
def iacwpt_(x, xw, wt=None):
    """``iacwpt!(x, xw[, wt])`` ACWT.jl:581-648"""
    R.inverse(True, "wpt", xw.unsqueeze(0), None, None, None, x.unsqueeze(0)); return x
def iacwpt(xw, wt=None):
    xw = D.dev(xw, "xw")
    return iacwpt_(xw.new_empty(tuple(xw.shape[1:])), xw)


def _split_args(args):
    """the reference accepts (xw), (xw, L), (xw, wt), (xw, wt, L), (xw, tree), (xw, wt, tree): drop the optional wavelet"""
    from .filters import OrthoFilter
    rest = [a for a in args if not isinstance(a, OrthoFilter) and a is not None]
    return rest[0] if rest else None


def iacwpd_(x, xw, *args):
    """``iacwpd!(x, xw[, wt][, L|tree])`` ACWT.jl:917-1000"""
    assert tuple(x.shape) == tuple(xw.shape[1:]), "AssertionError: size(x) == size(xw)[1:end-1]"
    tree = R.wpd_tree(R.sig_shape(x), xw.shape[0], _split_args(args))
    R.inverse(True, "wpd", xw.unsqueeze(0), None, tree, None, x.unsqueeze(0)); return x
def iacwpd(xw, *args):
    xw = D.dev(xw, "xw")
    return iacwpd_(xw.new_empty(tuple(xw.shape[1:])), xw, *args)


def _assert_batch(x):
    assert x.dim() > 1, "AssertionError: ndims(x) > 1"
def acdwtall(x, wt, L=None):
    """acwt/acwt_all.jl:33-53"""
    _assert_batch(x); return R.forward(True, "dwt", x, wt, L)
def acwptall(x, wt, L=None):
    """acwt/acwt_all.jl:136-156"""
    _assert_batch(x); return R.forward(True, "wpt", x, wt, L)
def acwpdall(x, wt, L=None):
    """acwt/acwt_all.jl:239-256"""
    _assert_batch(x); return R.forward(True, "wpd", x, wt, L)
def iacdwtall(xw, wt=None):
    """acwt/acwt_all.jl:86-103"""
    return R.inverse(True, "dwt", xw, None)
def iacwptall(xw, wt=None):
    """acwt/acwt_all.jl:189-206"""
    return R.inverse(True, "wpt", xw, None)
def iacwpdall(xw, *args):
    """acwt/acwt_all.jl:300-331"""
    xw = D.dev(xw, "xw")
    assert 3 <= xw.dim() <= 4, "AssertionError: 3 <= ndims(xw) <= 4"
    shp = tuple(reversed(tuple(xw.shape[2:])))
    return R.inverse(True, "wpd", xw, None, R.wpd_tree(shp, xw.shape[1], _split_args(args)))
