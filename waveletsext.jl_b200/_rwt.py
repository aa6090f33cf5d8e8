"""Shared implementation of the redundant transform families (stationary = SWT.jl, autocorrelation = ACWT.jl).

swt.py and acwt.py expose the reference's names on top of these helpers.  Layouts (Julia memory order):
signals x(n,N) -> (N, n); images x(n_rows, m_cols, N) -> (N, m_cols, n_rows); node tables put the node /
level axis right after the batch axis: (N, nodes, ...).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _dev as D
from .filters import makereverseqmfpair, make_acreverseqmfpair
from .utils import maxtransformlevels, maketree, isvalidtree, getdepth

MODE = {"dwt": 0, "wpt": 1, "wpd": 2}


def taps_for(ac: bool, wt):
    """(h, g) as the step kernels consume them.  Stationary: g = scaling, h = detail (SWT.jl:119).
    Autocorrelation: ``Pmf, Qmf = make_acreverseqmfpair(wt)`` and the step is called with (h, g) = (Qmf, Pmf)
    (ACWT.jl:120-131; the 2-D drivers bind g, h = P, Q and pass (h, g), i.e. the same operands)."""
    if ac:
        if wt is None:
            return np.zeros(1), np.zeros(1)
        P, Q = make_acreverseqmfpair(wt)
        return D.taps(Q), D.taps(P)
    g, h = makereverseqmfpair(wt, True)
    return D.taps(h), D.taps(g)


def ncols(mode: str, L: int, two: bool) -> int:
    if mode == "dwt":
        return 3 * L + 1 if two else L + 1
    if mode == "wpt":
        return 4 ** L if two else 1 << L
    return (4 ** (L + 1) - 1) // 3 if two else (1 << (L + 1)) - 1


def check_L(shape_jl, L):
    """the reference's argument checks (SWT.jl:114-116): ArgumentError -> ValueError"""
    if not L <= maxtransformlevels(shape_jl):
        raise ValueError("ArgumentError: Too many transform levels (length(x) < 2^L)")
    if not L >= 1:
        raise ValueError("ArgumentError: L must be >= 1")


def forward(ac: bool, mode: str, x, wt, L=None, xw=None):
    """batched forward transform: x (N, n) or (N, m, n) -> xw (N, ncols, ...)"""
    x = D.dev(x, "x")
    two = x.dim() == 3
    shp = tuple(reversed(tuple(x.shape[1:])))
    L = maxtransformlevels(shp) if L is None else int(L)
    check_L(shp, L)
    h, g = taps_for(ac, wt)
    N = x.shape[0]
    nc = ncols(mode, L, two)
    if xw is None:
        xw = x.new_empty((N, nc) + tuple(x.shape[1:]))
    xo = D.out(xw, "xw")
    D.same(x, xw)
    assert tuple(xw.shape) == (N, nc) + tuple(x.shape[1:]), "AssertionError: size(xw) does not match the transform"
    if two:
        _, cols, rows = x.shape
        m, n = rows, cols
    else:
        m, n = 0, x.shape[1]
    D.call("rwt", x, int(ac), MODE[mode], D.ptr(xo.t), D.ptr(x), m, n, L, N, h.ctypes.data, g.ctypes.data, len(h), D.stream(x))
    return xo.commit()


def inverse(ac: bool, mode: str, xw, wt, tree=None, sm=None, x=None):
    """batched inverse: xw (N, ncols, ...) -> x (N, ...).  sm None: average based (SWT) ; int: shift based."""
    xw = D.dev(xw, "xw")
    two = xw.dim() == 4
    N, nc = xw.shape[0], xw.shape[1]
    h, g = taps_for(ac, wt)
    if x is None:
        x = xw.new_empty((N,) + tuple(xw.shape[2:]))
    xo = D.out(x, "x")
    D.same(x, xw)
    assert tuple(x.shape) == (N,) + tuple(xw.shape[2:]), "AssertionError: size(x) == size(xw)[1:end-1]"
    if two:
        _, _, cols, rows = xw.shape
        m, n = rows, cols
    else:
        m, n = 0, xw.shape[2]
    t = D.tree_bytes(tree) if tree is not None else np.zeros(0, np.uint8)
    D.call("irwt", xw, int(ac), MODE[mode], D.ptr(xo.t), D.ptr(xw), m, n, nc, 0, N, t.ctypes.data if len(t) else 0, len(t),
           -1 if sm is None else int(sm), h.ctypes.data, g.ctypes.data, len(h), D.stream(xw))
    return xo.commit()


def sig_shape(x_single):
    return tuple(reversed(tuple(x_single.shape)))


def wpd_tree(shape_jl, nc, arg):
    """L::Integer | tree::BitVector | None -> tree for the wpd inverses"""
    two = len(shape_jl) == 2
    if arg is None:
        arg = maxtransformlevels(shape_jl)
    if isinstance(arg, (int, np.integer)):
        L = int(arg)
        check_L(shape_jl, L)
        return maketree(*shape_jl, L, "full")
    assert isvalidtree(shape_jl, arg), "AssertionError: isvalidtree(x, tree)"
    return np.asarray(arg, dtype=bool)
