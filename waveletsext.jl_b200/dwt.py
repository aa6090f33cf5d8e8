"""Decimated transforms on the B200: host mirror of DWT.jl, dwt/dwt_one_level.jl and dwt/dwt_all.jl.

Same function names, argument order and error behaviour as the reference; ``f!`` is spelled ``f_``.
Arrays are torch CUDA tensors in Julia memory order (see _dev.py): a batch of signals x(n,N) is a tensor of
shape (N, n); a packet table (n,L+1,N) is (N, L+1, n); images (m,n,N) are (N, n, m).
``wt`` is an OrthoFilter (filters.wavelet) or a raw qmf vector.  Every function is a thin wrapper over one
C-ABI entry point of libwx_b200.so -- there is no host computation of signal data and no CPU fallback.
"""
from __future__ import annotations

import torch

from . import _dev as D
from .filters import makereverseqmfpair
from .utils import maketree, maxtransformlevels, isvalidtree, isdyadic

__all__ = ["dwtall", "idwtall", "dwt_step", "dwt_step_", "idwt_step", "idwt_step_", "wpd", "wpd_", "iwpd", "iwpd_", "wpt", "wpt_",
           "iwpt", "iwpt_", "wpdall", "iwpdall", "wptall", "iwptall", "getbasiscoef", "getbasiscoefall"]


def _pair(wt):
    g, h = makereverseqmfpair(wt, True)        # g = scaling, h = detail  (DWT.jl:141)
    return D.taps(h), D.taps(g)


# ------------------------------------------------------------------ single steps
def dwt_step_(w1, w2, *rest):
    """``dwt_step!(w1, w2, v, h, g)`` dwt/dwt_one_level.jl:79-107  /  2-D
    ``dwt_step!(w1, w2, w3, w4, v, h, g, temp)`` :319-354 (``temp`` accepted and ignored: scratch is internal)."""
    if len(rest) == 3:
        v, h, g = rest
        v, o1, o2 = D.dev(v, "v"), D.out(w1, "w1"), D.out(w2, "w2")
        D.same(v, w1, w2)
        assert w1.numel() == w2.numel() == v.numel() // 2, "AssertionError: length(w1) == length(w2) == length(v)/2"
        assert len(h) == len(g), "AssertionError: length(h) == length(g)"
        h, g = D.taps(h), D.taps(g)
        D.call("dwt_step", v, D.ptr(o1.t), D.ptr(o2.t), D.ptr(v), v.numel(), h.ctypes.data, g.ctypes.data, len(h), D.stream(v))
        return o1.commit(), o2.commit()
    w3, w4, v, h, g = rest[:5]
    v = D.dev(v, "v")
    ws = [D.out(w, "w") for w in (w1, w2, w3, w4)]
    D.same(v, w1, w2, w3, w4)
    nc, nr = w1.shape
    assert all(tuple(w.shape) == (nc, nr) for w in (w1, w2, w3, w4)), "AssertionError: size(w1) == size(w2) == size(w3) == size(w4)"
    assert tuple(v.shape) == (2 * nc, 2 * nr), "AssertionError: size(w1)*2 == size(v)"
    h, g = D.taps(h), D.taps(g)
    D.call("dwt_step2", v, *[D.ptr(w.t) for w in ws], D.ptr(v), nr, nc, h.ctypes.data, g.ctypes.data, len(h), D.stream(v))
    return tuple(w.commit() for w in ws)


def dwt_step(v, h, g):
    """``dwt_step(v, h, g)`` dwt/dwt_one_level.jl:34-42 (1-D) / :283-293 (2-D)"""
    v = D.dev(v, "v")
    if v.dim() == 1:
        n = v.numel()
        return dwt_step_(v.new_empty(n // 2), v.new_empty(n // 2), v, h, g)
    nc2, nr2 = v.shape
    ws = [v.new_empty((nc2 // 2, nr2 // 2)) for _ in range(4)]
    return dwt_step_(*ws, v, h, g)


def idwt_step_(v, *rest):
    """``idwt_step!(v, w1, w2, h, g)`` :192-223  /  ``idwt_step!(v, w1, w2, w3, w4, h, g, temp)`` :401-436"""
    if len(rest) == 4:
        w1, w2, h, g = rest
        ov, w1, w2 = D.out(v, "v"), D.dev(w1, "w1"), D.dev(w2, "w2")
        D.same(v, w1, w2)
        assert w1.numel() == w2.numel() == v.numel() // 2, "AssertionError: length(w1) == length(w2) == length(v)/2"
        assert len(h) == len(g), "AssertionError: length(h) == length(g)"
        h, g = D.taps(h), D.taps(g)
        D.call("idwt_step", v, D.ptr(ov.t), D.ptr(w1), D.ptr(w2), v.numel(), h.ctypes.data, g.ctypes.data, len(h), D.stream(v))
        return ov.commit()
    w1, w2, w3, w4, h, g = rest[:6]
    ov = D.out(v, "v")
    ws = [D.dev(w, "w") for w in (w1, w2, w3, w4)]
    D.same(v, *ws)
    nc, nr = ws[0].shape
    assert all(tuple(w.shape) == (nc, nr) for w in ws), "AssertionError: size(w1) == size(w2) == size(w3) == size(w4)"
    assert tuple(v.shape) == (2 * nc, 2 * nr), "AssertionError: size(w1)*2 == size(v)"
    h, g = D.taps(h), D.taps(g)
    D.call("idwt_step2", v, D.ptr(ov.t), *[D.ptr(w) for w in ws], nr, nc, h.ctypes.data, g.ctypes.data, len(h), D.stream(v))
    return ov.commit()


def idwt_step(*args):
    """``idwt_step(w1, w2, h, g)`` / ``idwt_step(w1, w2, w3, w4, h, g)``"""
    if len(args) == 4:
        w1 = D.dev(args[0], "w1")
        return idwt_step_(w1.new_empty(2 * w1.numel()), *args)
    w1 = D.dev(args[0], "w1")
    nc, nr = w1.shape
    return idwt_step_(w1.new_empty((2 * nc, 2 * nr)), *args)


# ------------------------------------------------------------------ batched kernels (private)
def _wpd_batch(x, wt, L, y=None):
    """x (N, n) or (N, n, m) -> y (N, L+1, ...)"""
    x = D.dev(x, "x")
    h, g = _pair(wt)
    N = x.shape[0]
    if y is None:
        y = x.new_empty((N, L + 1) + tuple(x.shape[1:]))
    yo = D.out(y, "y")
    D.same(x, y)
    assert tuple(y.shape) == (N, L + 1) + tuple(x.shape[1:]), "AssertionError: size(y) == (n, L+1)"
    if x.dim() == 2:
        D.call("wpd1d", x, D.ptr(yo.t), D.ptr(x), x.shape[1], L, N, h.ctypes.data, g.ctypes.data, len(h), D.stream(x))
    else:
        _, n, m = x.shape
        D.call("wpd2d", x, D.ptr(yo.t), D.ptr(x), m, n, L, N, h.ctypes.data, g.ctypes.data, len(h), D.stream(x))
    return yo.commit()


def _tree_batch(name, x, wt, tree, y=None):
    """wpt / iwpt by tree on a batch x (N, n) or (N, n, m)"""
    x = D.dev(x, "x")
    h, g = _pair(wt)
    t = D.tree_bytes(tree)
    N = x.shape[0]
    if y is None:
        y = torch.empty_like(x)
    yo = D.out(y, "y")
    D.same(x, y)
    assert y.shape == x.shape, "AssertionError: size(y) == size(x)"
    if x.dim() == 2:
        D.call(f"{name}1d", x, D.ptr(yo.t), D.ptr(x), x.shape[1], N, t.ctypes.data, len(t), h.ctypes.data, g.ctypes.data, len(h), D.stream(x))
    else:
        _, n, m = x.shape
        D.call(f"{name}2d", x, D.ptr(yo.t), D.ptr(x), m, n, N, t.ctypes.data, len(t), h.ctypes.data, g.ctypes.data, len(h), D.stream(x))
    return yo.commit()


def _sigshape(x_single):
    """Julia-order shape of a single signal tensor"""
    return tuple(reversed(tuple(x_single.shape)))


def _tree_arg(shape_jl, arg):
    """L::Integer or tree::BitVector -> validated tree"""
    if isinstance(arg, (int,)) or arg is None:
        L = maxtransformlevels(shape_jl) if arg is None else int(arg)
        return maketree(*shape_jl, L, "full")
    assert isvalidtree(shape_jl, arg), "AssertionError: isvalidtree(x, tree)"
    return arg


# ------------------------------------------------------------------ wpd / iwpd (single signal)
def wpd_(y, x, wt, L=None):
    """``wpd!(y, x, wt, L)`` DWT.jl:131-161 (1-D), :164-209 (2-D)"""
    shp = _sigshape(x)
    L = maxtransformlevels(shp) if L is None else int(L)
    assert 0 <= L <= maxtransformlevels(shp), "AssertionError: 0 <= L <= maxtransformlevels(x)"
    assert tuple(y.shape) == (L + 1,) + tuple(x.shape), "AssertionError: size(y) == (n, L+1)"
    _wpd_batch(x.unsqueeze(0), wt, L, y.unsqueeze(0))
    return y


def wpd(x, wt, L=None):
    """``wpd(x, wt, L)`` DWT.jl:60-88"""
    x = D.dev(x, "x")
    shp = _sigshape(x)
    if x.dim() == 1:
        assert isdyadic(shp[0]), "AssertionError: isdyadic(x)"
    L = maxtransformlevels(shp) if L is None else int(L)
    assert 0 <= L <= maxtransformlevels(shp), "AssertionError: 0 <= L <= maxtransformlevels(x)"
    return _wpd_batch(x.unsqueeze(0), wt, L)[0]


def iwpd_(xh, xw, wt, arg=None):
    """``iwpd!(x, xw, wt, L|tree)`` DWT.jl:322-401"""
    xw = D.dev(xw, "xw")
    shp = _sigshape(xh)
    assert tuple(xh.shape) == tuple(xw.shape[1:]), "AssertionError: size(x,1) == size(xw,1)"
    tree = _tree_arg(shp, arg)
    _iwpd_batch(xw.unsqueeze(0), wt, tree, xh.unsqueeze(0))
    return xh


def iwpd(xw, wt, arg=None):
    """``iwpd(xw, wt, L|tree)`` DWT.jl:257-274"""
    xw = D.dev(xw, "xw")
    assert xw.dim() >= 2, "AssertionError: ndims(xw) >= 2"
    return iwpd_(xw.new_empty(tuple(xw.shape[1:])), xw, wt, arg)


def _iwpd_batch(Xw, wt, tree, x=None):
    Xw = D.dev(Xw, "xw")
    h, g = _pair(wt)
    t = D.tree_bytes(tree)
    N, K = Xw.shape[0], Xw.shape[1]
    if x is None:
        x = Xw.new_empty((N,) + tuple(Xw.shape[2:]))
    xo = D.out(x, "x")
    D.same(x, Xw)
    assert tuple(x.shape) == (N,) + tuple(Xw.shape[2:]), "AssertionError: size(x) == size(xw)[1:end-1]"
    if Xw.dim() == 3:
        m, n = 0, Xw.shape[2]
    else:
        n, m = Xw.shape[2], Xw.shape[3]
    D.call("iwpd", Xw, D.ptr(xo.t), D.ptr(Xw), m, n, K, N, t.ctypes.data, len(t), h.ctypes.data, g.ctypes.data, len(h), D.stream(Xw))
    return xo.commit()


# ------------------------------------------------------------------ wpt / iwpt (single signal)
def wpt_(y, x, wt, arg=None):
    """``wpt!(y, x, wt, L|tree)``: 2-D DWT.jl:493-548, 1-D Wavelets.jl"""
    x = D.dev(x, "x")
    tree = _tree_arg(_sigshape(x), arg)
    assert y.shape == x.shape, "AssertionError: size(y) == size(x)"
    _tree_batch("wpt", x.unsqueeze(0), wt, tree, y.unsqueeze(0))
    return y


def wpt(x, wt, arg=None):
    """``wpt(x, wt, L|tree)`` DWT.jl:440-451"""
    x = D.dev(x, "x")
    return wpt_(torch.empty_like(x), x, wt, arg)


def iwpt_(xh, xw, wt, arg=None):
    """``iwpt!(x, xw, wt, L|tree)``: 2-D DWT.jl:655-710, 1-D Wavelets.jl"""
    xw = D.dev(xw, "xw")
    tree = _tree_arg(_sigshape(xw), arg)
    assert xh.shape == xw.shape, "AssertionError: size(x) == size(xw)"
    _tree_batch("iwpt", xw.unsqueeze(0), wt, tree, xh.unsqueeze(0))
    return xh


def iwpt(xw, wt, arg=None):
    """``iwpt(xw, wt, L|tree)`` DWT.jl:594-605"""
    xw = D.dev(xw, "xw")
    return iwpt_(torch.empty_like(xw), xw, wt, arg)


# ------------------------------------------------------------------ *all (batch) functions
def wpdall(x, wt, L=None):
    """``wpdall(x, wt, L)`` dwt/dwt_all.jl:260-282 : x (N, n) or (N, n, m) -> (N, L+1, ...)"""
    x = D.dev(x, "x")
    assert x.dim() > 1, "AssertionError: ndims(x) > 1"
    shp = tuple(reversed(tuple(x.shape[1:])))
    L = maxtransformlevels(shp) if L is None else int(L)
    assert 0 <= L <= maxtransformlevels(shp), "AssertionError: 0 <= L <= maxtransformlevels(x)"
    return _wpd_batch(x, wt, L)


def iwpdall(xw, wt, arg=None):
    """``iwpdall(xw, wt, L|tree)`` dwt/dwt_all.jl:324-342"""
    xw = D.dev(xw, "xw")
    assert xw.dim() > 2, "AssertionError: ndims(xw) > 2"
    shp = tuple(reversed(tuple(xw.shape[2:])))
    return _iwpd_batch(xw, wt, _tree_arg(shp, arg))


def wptall(x, wt, arg=None):
    """``wptall(x, wt, L|tree)`` dwt/dwt_all.jl:152-166"""
    x = D.dev(x, "x")
    assert x.dim() > 1, "AssertionError: ndims(x) > 1"
    shp = tuple(reversed(tuple(x.shape[1:])))
    return _tree_batch("wpt", x, wt, _tree_arg(shp, arg))


def iwptall(xw, wt, arg=None):
    """``iwptall(xw, wt, L|tree)`` dwt/dwt_all.jl:210-225"""
    xw = D.dev(xw, "xw")
    assert xw.dim() > 1, "AssertionError: ndims(xw) > 1"
    shp = tuple(reversed(tuple(xw.shape[1:])))
    return _tree_batch("iwpt", xw, wt, _tree_arg(shp, arg))


# ------------------------------------------------------------------ basis extraction
def dwtall(x, wt, L=None):
    """``dwtall(x, wt[, L])`` dwt/dwt_all.jl:39-54: the discrete wavelet transform of every signal.  The reference loops over
    Wavelets.jl's ``dwt!``; on the device this is the tree transform along ``maketree(x, L, :dwt)`` (only the scaling node
    is split again), run by the fused all-levels kernel.  Values are PARITY UNPINNED inside the reference (its tests only
    check batch == single, test/transforms.jl:278-283)."""
    x = D.dev(x, "x")
    assert x.dim() > 1, "AssertionError: ndims(x) > 1"
    shp = tuple(reversed(tuple(x.shape[1:])))
    L = maxtransformlevels(shp) if L is None else int(L)
    assert 0 <= L <= maxtransformlevels(shp), "AssertionError: 0 <= L <= maxtransformlevels(x)"
    return _tree_batch("wpt", x, wt, maketree(*shp, L, "dwt"))


def idwtall(xw, wt, L=None):
    """``idwtall(xw, wt[, L])`` dwt/dwt_all.jl:95-110 (inverse of ``dwtall``)"""
    xw = D.dev(xw, "xw")
    assert xw.dim() > 1, "AssertionError: ndims(xw) > 1"
    shp = tuple(reversed(tuple(xw.shape[1:])))
    L = maxtransformlevels(shp) if L is None else int(L)
    assert 0 <= L <= maxtransformlevels(shp), "AssertionError: 0 <= L <= maxtransformlevels(xw)"
    return _tree_batch("iwpt", xw, wt, maketree(*shp, L, "dwt"))


def getbasiscoefall(Xw, tree):
    """``getbasiscoefall(Xw, tree)`` Utils.jl:169-197 : Xw (N, K, n[, m]) -> (N, n[, m]).  ``tree`` may also be a
    (ntree, N) boolean matrix (one tree per signal, Utils.jl:199-225); those are gathered signal by signal."""
    import numpy as np
    Xw = D.dev(Xw, "Xw")
    assert 3 <= Xw.dim() <= 4, "AssertionError: 3 <= ndims(Xw) <= 4"
    shp = tuple(reversed(tuple(Xw.shape[2:])))
    N, K = Xw.shape[0], Xw.shape[1]
    assert K - 1 <= maxtransformlevels(shp), "AssertionError: k-1 <= maxtransformlevels(x)"
    if isinstance(tree, torch.Tensor) and tree.dim() == 2:
        # one tree per signal, (N, ntree) on the device (what bestbasistreeall returns): gathered by one kernel
        assert tree.shape[0] == N, "AssertionError: m == m_t"
        t8 = tree.to(device=Xw.device, dtype=torch.uint8).contiguous()
        out = Xw.new_empty((N,) + tuple(Xw.shape[2:]))
        if Xw.dim() == 3:
            m, n = 0, Xw.shape[2]
        else:
            n, m = Xw.shape[2], Xw.shape[3]
        try:        # the library checks n_t == gettreelength, isvalidtree of every tree and the depth against K (Utils.jl:204-218)
            D.call("gather_basis_multi", Xw, D.ptr(out), D.ptr(Xw), m, n, K, N, D.ptr(t8), t8.shape[1], D.stream(Xw))
        except AssertionError as e:
            if "Not enough decomposition levels" in str(e):
                raise ValueError(str(e)) from None      # ArgumentError in the reference
            raise
        return out
    tree = np.asarray(tree, dtype=bool)
    if tree.ndim == 2:
        assert tree.shape[1] == N, "AssertionError: m == m_t"
        out = Xw.new_empty((N,) + tuple(Xw.shape[2:]))
        for i in range(N):
            out[i] = getbasiscoef(Xw[i], tree[:, i])
        return out
    assert isvalidtree(shp, tree), "AssertionError: isvalidtree(x, tree)"
    out = Xw.new_empty((N,) + tuple(Xw.shape[2:]))
    t = D.tree_bytes(tree)
    if Xw.dim() == 3:
        m, n = 0, Xw.shape[2]
    else:
        n, m = Xw.shape[2], Xw.shape[3]
    try:
        D.call("gather_basis", Xw, D.ptr(out), D.ptr(Xw), m, n, K, N, t.ctypes.data, len(t), D.stream(Xw))
    except AssertionError as e:
        if "Not enough decomposition levels" in str(e):
            raise ValueError(str(e)) from None      # ArgumentError in the reference
        raise
    return out


def getbasiscoef(Xw, tree):
    """``getbasiscoef(Xw, tree)`` Utils.jl:101-134 : Xw (K, n[, m]) -> (n[, m])"""
    Xw = D.dev(Xw, "Xw")
    assert 2 <= Xw.dim() <= 3, "AssertionError: 2 <= ndims(Xw) <= 3"
    return getbasiscoefall(Xw.unsqueeze(0), tree)[0]
