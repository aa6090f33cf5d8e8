"""ctypes front end of the CPU oracle (oracle/libwx_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package never does (tests/test_host_logic.py checks that).

Array convention ("Julia memory order"): a Julia array of size (d1, d2, ..., dk) -- first index
fastest, batch last -- is held as a C-contiguous numpy array of shape (dk, ..., d2, d1).  So a batch of
signals x (n, N) is numpy (N, n); a packet table (n, L+1, N) is numpy (N, L+1, n); a 2-D image batch
(m, n, N) is numpy (N, n, m).  The bytes are identical to what the Julia reference holds.
`jl(A)` converts a numpy array indexed like the Julia literal (A[i1, i2, ...]) into that layout.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc) if needed; returns the library path."""
    so = os.path.join(_HERE, "libwx_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("wx_oracle.c", "wx_oracle_impl.h", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "-s", "CC=gcc"] + (["-B"] if force else []), check=True)
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


def jl(a) -> np.ndarray:
    """numpy array indexed like a Julia literal -> Julia memory order (all axes reversed)."""
    a = np.asarray(a)
    return np.ascontiguousarray(a.transpose(tuple(range(a.ndim))[::-1]))


unjl = jl   # the conversion is an involution

_CT = {"l": C.c_long, "i": C.c_int, "d": C.c_double}


def _sfx(dt) -> str:
    dt = np.dtype(dt)
    if dt == np.float64:
        return "f64"
    if dt == np.float32:
        return "f32"
    raise TypeError(f"unsupported element type {dt}")


def _call(name: str, sig: str, *args, restype=None):
    """sig: one char per argument. P = array pointer (any dtype, passed as void*), l = long, i = int,
    d = double."""
    f = getattr(lib(), name)
    cargs = []
    keep = []
    for s, a in zip(sig, args):
        if s == "P":
            if a is None:
                cargs.append(C.c_void_p(0))
            else:
                assert a.flags["C_CONTIGUOUS"], "oracle arrays must be C-contiguous"
                keep.append(a)
                cargs.append(C.c_void_p(a.ctypes.data))
        else:
            cargs.append(_CT[s](int(a) if s != "d" else float(a)))
    f.restype = restype
    return f(*cargs)


def _taps(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _tree(t) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(t).astype(np.uint8))


class OracleAssertion(AssertionError):
    """the reference's @assert would have fired"""


# ------------------------------------------------------------------ filters / index algebra
def makereverseqmfpair(q):
    q = _taps(q); g = np.empty_like(q); h = np.empty_like(q)
    _call("wx_makereverseqmfpair", "PiPP", q, len(q), g, h)
    return g, h


def make_acreverseqmfpair(q):
    q = _taps(q); P = np.empty(2 * len(q) - 1); Q = np.empty(2 * len(q) - 1)
    _call("wx_make_acreverseqmfpair", "PiPP", q, len(q), P, Q)
    return P, Q


def maxtransformlevels(n: int) -> int:
    return _call("wx_maxtransformlevels", "l", n, restype=C.c_int)


def getdepth(i: int, kind: str) -> int:
    return _call("wx_ilog2" if kind == "binary" else "wx_quaddepth", "l", i, restype=C.c_int)


def treelength2(nr, nc) -> int:
    return _call("wx_treelength2", "ll", nr, nc, restype=C.c_long)


def maketree1(n, L, s="full"):
    t = np.zeros(max(n - 1, 0), np.uint8)
    _call("wx_maketree1", "Plii", t, n, L, int(s == "dwt"))
    return t.astype(bool)


def maketree2(nr, nc, L, s="full"):
    t = np.zeros(treelength2(nr, nc), np.uint8)
    _call("wx_maketree2", "Pllii", t, nr, nc, L, int(s == "dwt"))
    return t.astype(bool)


def getleaf(tree, kind):
    t = _tree(tree)
    ar = 2 if kind == "binary" else 4
    leaf = np.zeros(ar * len(t) + 1, np.uint8)
    _call("wx_getleaf_binary" if ar == 2 else "wx_getleaf_quad", "PlP", t, len(t), leaf)
    return leaf.astype(bool)


def isvalidtree(tree, arity=2) -> bool:
    t = _tree(tree)
    return bool(_call("wx_isvalidtree", "Pli", t, len(t), arity, restype=C.c_int))


def quadrange(m, n, idx):
    out = (C.c_long * 4)()
    f = lib().wx_quadrange
    f.restype = None
    f(C.c_long(m), C.c_long(n), C.c_long(idx), C.byref(out, 0), C.byref(out, 8), C.byref(out, 16), C.byref(out, 24))
    return tuple(out)   # r0, c0, nr, nc (0-based start)


def main2depthshift(sm, L):
    sd = np.zeros(L + 1, np.int64)
    if _call("wx_main2depthshift", "Pli", sd, sm, L, restype=C.c_int) != 0:
        raise OracleAssertion("sm < 1<<L")
    return sd


# ------------------------------------------------------------------ single steps (1-D)
def dwt_step(v, h, g):
    v = np.ascontiguousarray(v); n = v.shape[-1]
    w1 = np.empty(n // 2, v.dtype); w2 = np.empty(n // 2, v.dtype)
    _call(f"wxo_dwt_step_{_sfx(v.dtype)}", "PlPlPllPPi", w1, 1, w2, 1, v, 1, n, _taps(h), _taps(g), len(h))
    return w1, w2


def idwt_step(w1, w2, h, g):
    w1 = np.ascontiguousarray(w1); w2 = np.ascontiguousarray(w2); n = 2 * len(w1)
    v = np.empty(n, w1.dtype)
    _call(f"wxo_idwt_step_{_sfx(v.dtype)}", "PlPlPllPPi", v, 1, w1, 1, w2, 1, n, _taps(h), _taps(g), len(h))
    return v


def sdwt_step(v, d, h, g):
    v = np.ascontiguousarray(v); n = len(v)
    w1 = np.empty_like(v); w2 = np.empty_like(v)
    _call(f"wxo_sdwt_step_{_sfx(v.dtype)}", "PlPlPlliPPi", w1, 1, w2, 1, v, 1, n, d, _taps(h), _taps(g), len(h))
    return w1, w2


def isdwt_step(w1, w2, d, h, g, sv=None, sw=None, v0=None, add2out=False):
    """average based when sv is None, else shift based (positions outside the coset keep v0 / zeros)."""
    w1 = np.ascontiguousarray(w1); w2 = np.ascontiguousarray(w2); n = len(w1)
    v = np.zeros(n, w1.dtype) if v0 is None else np.array(v0, dtype=w1.dtype)
    if sv is None:
        _call(f"wxo_isdwt_step_avg_{_sfx(v.dtype)}", "PlPlPlliPPi", v, 1, w1, 1, w2, 1, n, d, _taps(h), _taps(g), len(h))
    else:
        rc = _call(f"wxo_isdwt_step_shift_{_sfx(v.dtype)}", "PlPlPllillPPii", v, 1, w1, 1, w2, 1, n, d, sv, sw,
                   _taps(h), _taps(g), len(h), int(add2out), restype=C.c_int)
        if rc != 0:
            raise OracleAssertion("0 <= sv < 1<<d and sv <= sw < 1<<(d+1)")
    return v


def acdwt_step(v, d, h, g):
    v = np.ascontiguousarray(v); n = len(v)
    w1 = np.empty_like(v); w2 = np.empty_like(v)
    _call(f"wxo_acdwt_step_{_sfx(v.dtype)}", "PlPlPlliPPi", w1, 1, w2, 1, v, 1, n, d, _taps(h), _taps(g), len(h))
    return w1, w2


def iacdwt_step(w1, w2):
    w1 = np.ascontiguousarray(w1); w2 = np.ascontiguousarray(w2)
    v = np.empty_like(w1)
    _call(f"wxo_iacdwt_step_{_sfx(v.dtype)}", "PlPlPll", v, 1, w1, 1, w2, 1, len(w1))
    return v


# ------------------------------------------------------------------ single steps (2-D); arrays are (cols, rows)
def dwt_step2(v, h, g):
    v = np.ascontiguousarray(v); nc2, nr2 = v.shape
    nr, nc = nr2 // 2, nc2 // 2
    ws = [np.empty((nc, nr), v.dtype) for _ in range(4)]
    temp = np.empty_like(v)
    _call(f"wxo_dwt_step2_{_sfx(v.dtype)}", "PPPPlPlPlllPPi", *ws, nr, v, nr2, temp, nr2, nr, nc, _taps(h), _taps(g), len(h))
    return ws


def idwt_step2(w1, w2, w3, w4, h, g):
    ws = [np.ascontiguousarray(w) for w in (w1, w2, w3, w4)]
    nc, nr = ws[0].shape
    v = np.empty((2 * nc, 2 * nr), ws[0].dtype); temp = np.empty_like(v)
    _call(f"wxo_idwt_step2_{_sfx(v.dtype)}", "PlPPPPlPlllPPi", v, 2 * nr, *ws, nr, temp, 2 * nr, nr, nc, _taps(h), _taps(g), len(h))
    return v


def _rstep2(name, v, d, h, g):
    v = np.ascontiguousarray(v); nc, nr = v.shape
    ws = [np.empty_like(v) for _ in range(4)]
    temp = np.empty((2, nc, nr), v.dtype)
    _call(f"wxo_{name}_{_sfx(v.dtype)}", "PPPPPPlliPPi", *ws, v, temp, nr, nc, d, _taps(h), _taps(g), len(h))
    return ws


def sdwt_step2(v, d, h, g):
    return _rstep2("sdwt_step2", v, d, h, g)


def acdwt_step2(v, d, h, g):
    return _rstep2("acdwt_step2", v, d, h, g)


def isdwt_step2(w1, w2, w3, w4, d, h, g, sv=None, sw=None):
    ws = [np.ascontiguousarray(w) for w in (w1, w2, w3, w4)]
    nc, nr = ws[0].shape
    v = np.zeros_like(ws[0]); temp = np.zeros((2, nc, nr), v.dtype)
    if sv is None:
        _call(f"wxo_isdwt_step2_avg_{_sfx(v.dtype)}", "PPPPPPlliPPi", v, *ws, temp, nr, nc, d, _taps(h), _taps(g), len(h))
    else:
        rc = _call(f"wxo_isdwt_step2_shift_{_sfx(v.dtype)}", "PPPPPPllillPPi", v, *ws, temp, nr, nc, d, sv, sw,
                   _taps(h), _taps(g), len(h), restype=C.c_int)
        if rc != 0:
            raise OracleAssertion("0 <= sv < 1<<d and sv <= sw < 1<<(d+1)")
    return v


def iacdwt_step2(w1, w2, w3, w4):
    ws = [np.ascontiguousarray(w) for w in (w1, w2, w3, w4)]
    nc, nr = ws[0].shape
    v = np.empty_like(ws[0]); temp = np.empty((2, nc, nr), v.dtype)
    _call(f"wxo_iacdwt_step2_{_sfx(v.dtype)}", "PPPPPPll", v, *ws, temp, nr, nc)
    return v


# ------------------------------------------------------------------ decimated trees, single signal
def wpd(x, h, g, L):
    """1-D: x (n,) -> (L+1, n).  2-D: x (n, m) -> (L+1, n, m)  [Julia (m,n) -> (m,n,L+1)]"""
    x = np.ascontiguousarray(x)
    y = np.empty((L + 1,) + x.shape, x.dtype)
    if x.ndim == 1:
        _call(f"wxo_wpd1_{_sfx(x.dtype)}", "PPliPPi", y, x, x.shape[0], L, _taps(h), _taps(g), len(h))
    else:
        n, m = x.shape
        _call(f"wxo_wpd2_{_sfx(x.dtype)}", "PPlliPPi", y, x, m, n, L, _taps(h), _taps(g), len(h))
    return y


def _tree_tf(name1, name2, x, tree, h, g):
    x = np.ascontiguousarray(x); t = _tree(tree); y = np.empty_like(x)
    if x.ndim == 1:
        _call(f"wxo_{name1}_{_sfx(x.dtype)}", "PPlPlPPi", y, x, x.shape[0], t, len(t), _taps(h), _taps(g), len(h))
    else:
        n, m = x.shape
        _call(f"wxo_{name2}_{_sfx(x.dtype)}", "PPllPlPPi", y, x, m, n, t, len(t), _taps(h), _taps(g), len(h))
    return y


def wpt(x, tree, h, g):
    return _tree_tf("wpt1", "wpt2", x, tree, h, g)


def iwpt(xw, tree, h, g):
    return _tree_tf("iwpt1", "iwpt2", xw, tree, h, g)


def getbasiscoef(Xw, tree):
    Xw = np.ascontiguousarray(Xw); t = _tree(tree)
    out = np.empty(Xw.shape[1:], Xw.dtype)
    if Xw.ndim == 2:
        K, n = Xw.shape
        rc = _call(f"wxo_getbasiscoef1_{_sfx(Xw.dtype)}", "PPliPl", out, Xw, n, K, t, len(t), restype=C.c_int)
    else:
        K, n, m = Xw.shape
        rc = _call(f"wxo_getbasiscoef2_{_sfx(Xw.dtype)}", "PPlliPl", out, Xw, m, n, K, t, len(t), restype=C.c_int)
    if rc != 0:
        raise ValueError("Not enough decomposition levels in Xw.")
    return out


def iwpd(Xw, tree, h, g):
    Xw = np.ascontiguousarray(Xw); t = _tree(tree)
    out = np.empty(Xw.shape[1:], Xw.dtype)
    if Xw.ndim == 2:
        K, n = Xw.shape
        rc = _call(f"wxo_iwpd1_{_sfx(Xw.dtype)}", "PPliPlPPi", out, Xw, n, K, t, len(t), _taps(h), _taps(g), len(h), restype=C.c_int)
        if rc != 0:
            raise ValueError("Not enough decomposition levels in Xw.")
    else:
        K, n, m = Xw.shape
        _call(f"wxo_iwpd2_{_sfx(Xw.dtype)}", "PPlliPlPPi", out, Xw, m, n, K, t, len(t), _taps(h), _taps(g), len(h))
    return out


# ------------------------------------------------------------------ redundant trees, single signal
def _rfwd(kind, ac, x, L, h, g):
    x = np.ascontiguousarray(x)
    two = x.ndim == 2
    ncol = {"dwt": (3 * L + 1) if two else (L + 1),
            "wpt": (4 ** L) if two else (1 << L),
            "wpd": ((4 ** (L + 1) - 1) // 3) if two else ((1 << (L + 1)) - 1)}[kind]
    xw = np.empty((ncol,) + x.shape, x.dtype)
    if not two:
        _call(f"wxo_r{kind}1_{_sfx(x.dtype)}", "iPPliPPi", int(ac), xw, x, x.shape[0], L, _taps(h), _taps(g), len(h))
    else:
        nc, nr = x.shape
        _call(f"wxo_r{kind}2_{_sfx(x.dtype)}", "iPPlliPPi", int(ac), xw, x, nr, nc, L, _taps(h), _taps(g), len(h))
    return xw


def sdwt(x, L, h, g): return _rfwd("dwt", 0, x, L, h, g)
def swpt(x, L, h, g): return _rfwd("wpt", 0, x, L, h, g)
def swpd(x, L, h, g): return _rfwd("wpd", 0, x, L, h, g)
def acdwt(x, L, P, Q): return _rfwd("dwt", 1, x, L, Q, P)     # (h,g) = (Q,P)  ACWT.jl:131
def acwpt(x, L, P, Q): return _rfwd("wpt", 1, x, L, Q, P)
def acwpd(x, L, P, Q): return _rfwd("wpd", 1, x, L, Q, P)


def _mode(ac, sm):
    return 2 if ac else (0 if sm is None else 1)


def _rinv_levels(kind, ac, xw, L, h, g, sm):
    xw = np.ascontiguousarray(xw)
    two = xw.ndim == 3
    x = np.zeros(xw.shape[1:], xw.dtype)
    sd = None if (ac or sm is None) else main2depthshift(sm, L)
    hh = _taps(h) if h is not None else np.zeros(2); gg = _taps(g) if g is not None else np.zeros(2)
    if not two:
        rc = _call(f"wxo_ir{kind}1_{_sfx(xw.dtype)}", "iPPliPPPi", _mode(ac, sm), x, xw, xw.shape[1], L, sd, hh, gg, len(hh), restype=C.c_int)
    else:
        _, nc, nr = xw.shape
        rc = _call(f"wxo_ir{kind}2_{_sfx(xw.dtype)}", "iPPlliPPPi", _mode(ac, sm), x, xw, nr, nc, L, sd, hh, gg, len(hh), restype=C.c_int)
    if rc != 0:
        raise OracleAssertion("shift assertion")
    return x


def isdwt(xw, h, g, sm=None):
    L = (xw.shape[0] - 1) if xw.ndim == 2 else (xw.shape[0] - 1) // 3
    return _rinv_levels("dwt", 0, xw, L, h, g, sm)


def iswpt(xw, h, g, sm=None):
    L = getdepth(xw.shape[0], "binary") if xw.ndim == 2 else getdepth(3 * xw.shape[0] - 2, "quad") - 0
    if xw.ndim == 3:
        L = int(round(np.log(xw.shape[0]) / np.log(4)))
    return _rinv_levels("wpt", 0, xw, L, h, g, sm)


def iacdwt(xw):
    L = (xw.shape[0] - 1) if xw.ndim == 2 else (xw.shape[0] - 1) // 3
    return _rinv_levels("dwt", 1, xw, L, None, None, None)


def iacwpt(xw):
    L = getdepth(xw.shape[0], "binary") if xw.ndim == 2 else int(round(np.log(xw.shape[0]) / np.log(4)))
    return _rinv_levels("wpt", 1, xw, L, None, None, None)


def _rinv_tree(ac, xw, tree, h, g, sm):
    xw = np.ascontiguousarray(xw); t = _tree(tree)
    two = xw.ndim == 3
    x = np.zeros(xw.shape[1:], xw.dtype)
    L = getdepth(xw.shape[0], "quad" if two else "binary")
    sd = None if (ac or sm is None) else main2depthshift(sm, L)
    hh = _taps(h) if h is not None else np.zeros(2); gg = _taps(g) if g is not None else np.zeros(2)
    if not two:
        rc = _call(f"wxo_irwpd1_{_sfx(xw.dtype)}", "iPPllPlPPPi", _mode(ac, sm), x, xw, xw.shape[1], xw.shape[0], t, len(t), sd, hh, gg, len(hh), restype=C.c_int)
    else:
        _, nc, nr = xw.shape
        rc = _call(f"wxo_irwpd2_{_sfx(xw.dtype)}", "iPPlllPlPPPi", _mode(ac, sm), x, xw, nr, nc, xw.shape[0], t, len(t), sd, hh, gg, len(hh), restype=C.c_int)
    if rc != 0:
        raise OracleAssertion("iswpd/iacwpd assertion")
    return x


def iswpd(xw, tree, h, g, sm=None): return _rinv_tree(0, xw, tree, h, g, sm)
def iacwpd(xw, tree): return _rinv_tree(1, xw, tree, None, None, None)


# ------------------------------------------------------------------ batch drivers
def wpdall(x, q, L, nthreads=1):
    """x (N, n) or (N, n, m); q = qmf taps (the pair is rebuilt per signal like DWT.jl:141)"""
    x = np.ascontiguousarray(x); q = _taps(q)
    N = x.shape[0]
    y = np.empty((N, L + 1) + x.shape[1:], x.dtype)
    if x.ndim == 2:
        _call(f"wxo_wpdall1_{_sfx(x.dtype)}", "PPlilPii", y, x, x.shape[1], L, N, q, len(q), nthreads)
    else:
        _, n, m = x.shape
        _call(f"wxo_wpdall2_{_sfx(x.dtype)}", "PPllilPii", y, x, m, n, L, N, q, len(q), nthreads)
    return y


def wpdall_into(y, x, q, L, nthreads=1):
    """timing variant: no allocation inside (1-D only)"""
    _call(f"wxo_wpdall1_{_sfx(x.dtype)}", "PPlilPii", y, x, x.shape[1], L, x.shape[0], _taps(q), len(q), nthreads)


def iwptall(xw, q, tree, nthreads=1):
    xw = np.ascontiguousarray(xw); q = _taps(q); t = _tree(tree)
    y = np.empty_like(xw)
    _call(f"wxo_iwptall1_{_sfx(xw.dtype)}", "PPllPlPii", y, xw, xw.shape[1], xw.shape[0], t, len(t), q, len(q), nthreads)
    return y


def rwpdall(ac, x, L, h, g, nthreads=1):
    x = np.ascontiguousarray(x)
    N, n = x.shape
    xw = np.empty((N, (1 << (L + 1)) - 1, n), x.dtype)
    _call(f"wxo_rwpdall1_{_sfx(x.dtype)}", "iPPlilPPii", int(ac), xw, x, n, L, N, _taps(h), _taps(g), len(h), nthreads)
    return xw


def max_threads() -> int:
    return _call("wxo_max_threads", "", restype=C.c_int)


# ------------------------------------------------------------------ best basis
def tree_costs_jbb(X, redundant=False, cost="loglp", p=2.0):
    """X (N, K, n) or (N, K, n, m) -> costs (T)"""
    X = np.ascontiguousarray(X)
    kind = 0 if cost == "loglp" else 1
    if X.ndim == 3:
        N, K, n = X.shape
        nc = K if redundant else (1 << K) - 1
        costs = np.empty(nc, X.dtype)
        rc = _call(f"wxo_tree_costs_jbb1_{_sfx(X.dtype)}", "PPllliid", costs, X, n, K, N, int(redundant), kind, p, restype=C.c_int)
    else:
        N, K, n, m = X.shape
        nc = K if redundant else (4 ** K - 1) // 3
        costs = np.empty(nc, X.dtype)
        rc = _call(f"wxo_tree_costs_jbb2_{_sfx(X.dtype)}", "PPlllliid", costs, X, m, n, K, N, int(redundant), kind, p, restype=C.c_int)
    if rc != 0:
        raise OracleAssertion("all(sigma .>= 0)")
    return costs


def tree_costs_lsdb(X, redundant=False):
    X = np.ascontiguousarray(X)
    if X.ndim == 3:
        N, K, n = X.shape
        costs = np.empty(K if redundant else (1 << K) - 1, X.dtype)
        _call(f"wxo_tree_costs_lsdb1_{_sfx(X.dtype)}", "PPllli", costs, X, n, K, N, int(redundant))
    else:
        N, K, n, m = X.shape
        costs = np.empty(K if redundant else (4 ** K - 1) // 3, X.dtype)
        _call(f"wxo_tree_costs_lsdb2_{_sfx(X.dtype)}", "PPlllli", costs, X, m, n, K, N, int(redundant))
    if np.isnan(costs).any():                      # a position that is constant over the batch: zero range step (see diffentropy)
        raise ValueError("ArgumentError: range step cannot be zero")
    return costs


def tree_costs_bb(X, redundant=False, cost="shannon"):
    """tree_costs(X, ::BB) for ONE signal: X (K, n) or (K, n, m) -> costs (T)"""
    X = np.ascontiguousarray(X)
    kind = 0 if cost == "shannon" else 1
    if X.ndim == 2:
        K, n = X.shape
        costs = np.empty(K if redundant else (1 << K) - 1, X.dtype)
        _call(f"wxo_tree_costs_bb1_{_sfx(X.dtype)}", "PPllii", costs, X, n, K, int(redundant), kind)
    else:
        K, n, m = X.shape
        costs = np.empty(K if redundant else (4 ** K - 1) // 3, X.dtype)
        _call(f"wxo_tree_costs_bb2_{_sfx(X.dtype)}", "PPlllii", costs, X, m, n, K, int(redundant), kind)
    return costs


def diffentropy(x):
    """coefcost(x, DifferentialEntropyCost()) bestbasis_costs.jl:135-155; a constant sample (zero range step) raises like the
    reference's ``a:0.0:b`` (ArgumentError -> ValueError)"""
    x = np.ascontiguousarray(x)
    v = _call(f"wxo_diffentropy_{_sfx(x.dtype)}", "Pll", x, 1, len(x), restype=C.c_double)
    if np.isnan(v):
        raise ValueError("ArgumentError: range step cannot be zero")
    return v


def tree_select(costs, n, m=None, minmax="min"):
    costs = np.array(costs)     # copy: modified in place like the reference
    mm = 0 if minmax == "min" else 1
    if m is None:
        tree = np.zeros(n - 1, np.uint8)
        _call(f"wxo_tree_select1_{_sfx(costs.dtype)}", "PPlli", tree, costs, len(costs), n, mm)
    else:
        tree = np.zeros(treelength2(n, m), np.uint8)
        _call(f"wxo_tree_select2_{_sfx(costs.dtype)}", "PPllli", tree, costs, len(costs), n, m, mm)
    return tree.astype(bool)


# ------------------------------------------------------------------ LDB (row f-2): numpy restatement, small per-position maps
def energy_map_tf(Xw, y):
    """energy_map(Xw, y, TimeFrequency()) ldb/ldb_energymap.jl:109-141.  Xw (N, K, n[, m]) -> Gamma (nc, K, n[, m]); classes in
    order of first appearance (Julia unique)."""
    Xw = np.asarray(Xw); y = list(y)
    classes = list(dict.fromkeys(y))
    out = np.empty((len(classes),) + Xw.shape[1:], Xw.dtype)
    for i, c in enumerate(classes):
        idx = [k for k, v in enumerate(y) if v == c]
        xw = Xw[idx]
        x = xw[:, 0]                                              # level 0 = the signals
        norm_sum = np.sum(np.sum(x.reshape(len(idx), -1).astype(Xw.dtype) ** 2, axis=1))
        out[i] = np.sum(xw ** 2, axis=0) / norm_sum
    return out


def discriminant_measure(G, kind="are", p=2.0):
    """discriminant_measure(Gamma, dm) ldb/ldb_measures.jl:139-183 + pairwise measures :302-325 (time-frequency maps)"""
    G = np.asarray(G)
    nc = G.shape[0]
    Dm = np.zeros(G.shape[1:], G.dtype)
    def are(a, b):
        with np.errstate(divide="ignore", invalid="ignore"):
            v = a * np.log(a / b)
        return np.where((a == 0) | (b == 0), 0, v)
    for i in range(nc):
        for j in range(i + 1, nc):
            a, b = G[i], G[j]
            if kind == "are":
                Dm = Dm + are(a, b)
            elif kind == "sre":
                Dm = Dm + (are(a, b) + are(b, a))
            elif kind == "lp":
                Dm = Dm + (a - b) ** p
            else:
                Dm = Dm + (np.sqrt(a) - np.sqrt(b)) ** 2
    return Dm.astype(G.dtype)


def ldb_costs(DM):
    """node costs of fitdec! with top_k >= node size (LDB.jl:217-237): sum of DM over the node.  DM (K, n) or (K, n, m)"""
    DM = np.asarray(DM)
    if DM.ndim == 2:
        K, n = DM.shape
        return np.array([DM[d, j * (n >> d):(j + 1) * (n >> d)].sum() for d in range(K) for j in range(1 << d)], DM.dtype)
    K, nc_, nr = DM.shape
    out = []
    for i in range(1, (4 ** K - 1) // 3 + 1):
        d = getdepth(i, "quad")
        r0, c0, rr, cc = quadrange(nr, nc_, i)
        out.append(DM[d, c0:c0 + cc, r0:r0 + rr].sum())
    return np.array(out, DM.dtype)


# ------------------------------------------------------------------ denoising (next row f-3), single signal, plain numpy
# Wavelets.jl (dependency, compat "0.9, 0.10", not vendored in the reference tree) supplies Threshold.mad!, threshold! and
# VisuShrink; their published algorithms are restated here.  Everything below computes in Float64 on the data as given.
TH_HARD, TH_SOFT, TH_SEMISOFT, TH_STEIN = 0, 1, 2, 3


def mad(y):
    """Wavelets.Threshold.mad!: m = median!(y); y .= abs.(y .- m); median!(y)   (arithmetic in the element type)"""
    y = np.array(y, copy=True).reshape(-1)
    m = np.median(y).astype(y.dtype)
    return np.median(np.abs(y - m)).astype(y.dtype)


def finestdetailrange(n, tree, redundant=False):
    """Utils.jl:416-436 -> 0-based (start, stop) of the vector range, or the 0-based node index for redundant tables"""
    tree = np.asarray(tree).astype(bool)
    assert getdepth(len(tree), "binary") + 1 == maxtransformlevels(n)
    i, j = 1, 0
    while i <= len(tree) and tree[i - 1]:
        i = 2 * i + 1; j += 1
    return (i - 1) if redundant else (n - (n >> j), n)


def coarsestscalingrange(n, tree, redundant=False):
    """Utils.jl:351-370"""
    tree = np.asarray(tree).astype(bool)
    assert getdepth(len(tree), "binary") + 1 == maxtransformlevels(n)
    i, j = 1, 0
    while i < len(tree) and tree[i - 1]:
        i = 2 * i; j += 1
    return (i - 1) if redundant else (0, n >> j)


def noisest(x, redundant, tree=None):
    """Denoising.jl:214-232.  x: vector (n,) or table (K, n) in Julia memory order"""
    x = np.asarray(x)
    n = x.shape[-1]
    assert n & (n - 1) == 0
    if not redundant and tree is None:
        dr = x[n // 2:]
    elif not redundant:
        a, b = finestdetailrange(n, tree, False); dr = x[a:b]
    elif tree is None:
        dr = x[-1]
    else:
        dr = x[finestdetailrange(n, tree, True)]
    return float(mad(dr)) / 0.6745


def _selected(coef, redundant, tree):
    coef = np.asarray(coef)
    if not redundant:
        return coef.reshape(-1)
    if tree is None:
        return coef.reshape(-1)
    leaves = np.asarray(getleaf(tree, "binary")).astype(bool)
    return coef[:len(leaves)][leaves].reshape(-1)


def surethreshold(coef, redundant, tree=None):
    """Denoising.jl:142-166"""
    y = _selected(coef, redundant, tree).astype(np.float64)
    a = np.sort(np.abs(y)) ** 2
    b = np.cumsum(a)
    n = len(y)
    c = np.arange(n - 1, -1, -1)
    s = b + c * a
    risk = (n - 2 * np.arange(1, n + 1) + s) / n
    return float(np.sqrt(a[int(np.argmin(risk))]))


def orth2relerror(orth):
    """Denoising.jl:344-349"""
    o = np.sort(np.asarray(orth, np.float64) ** 2)[::-1]
    S = o.sum()
    return np.sqrt(np.abs(S - np.cumsum(o))) / np.sqrt(S)


def findelbow(x, y):
    """Denoising.jl:366-381 -> 0-based index"""
    v = np.array([x[-1] - x[0], y[-1] - y[0]])
    v = v / np.sqrt((v ** 2).sum())
    dx, dy = x - x[0], y - y[0]
    H = np.sqrt(dx ** 2 + dy ** 2)
    A = dx * v[0] + dy * v[1]
    O = np.sqrt(np.abs(H ** 2 - A ** 2))
    return int(np.argmax(O))


def relerrorthreshold(coef, redundant=False, tree=None, elbows=2):
    """Denoising.jl:285-328 (makeplot = false)"""
    assert elbows >= 1
    c = _selected(coef, redundant, tree).astype(np.float64)
    x = np.sort(np.abs(c))[::-1]
    r = orth2relerror(c)
    x = np.concatenate([x, [0.0]])
    r = np.concatenate([[r[0]], r])
    xmax, ymax = x.max(), r.max()
    x = x[::-1] / xmax
    y = r[::-1] / ymax
    ix = findelbow(x, y)
    for _ in range(1, elbows):
        ix = findelbow(x[:ix + 1], y[:ix + 1])
    return float(x[ix] * xmax)


def threshold(x, th, t):
    """Wavelets.jl threshold! for HardTH / SoftTH / SemiSoftTH / SteinTH (t a Float64; result stored in x's element type)"""
    assert t >= 0
    x = np.asarray(x)
    xd = x.astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        if th == TH_HARD:
            out = np.where(np.abs(xd) <= t, 0.0, xd)
        elif th == TH_SOFT:
            sh = np.abs(xd) - t
            out = np.where(sh < 0, 0.0, np.sign(xd) * sh)
        elif th == TH_SEMISOFT:
            sh = np.abs(xd) - t
            inner = np.where(sh < 0, 0.0, np.where(sh - t < 0, np.sign(xd) * sh * 2, xd))
            out = np.where(xd <= 2 * t, inner, xd)
        elif th == TH_STEIN:
            sh = 1.0 - t * t / (xd * xd)
            out = np.where(sh < 0, 0.0, xd * sh)
        else:
            raise ValueError(th)
    return out.astype(x.dtype)


def visushrink_t(n):
    """Wavelets.jl VisuShrink(n): t = sqrt(2 log n)"""
    return float(np.sqrt(2 * np.log(n)))


def denoise(x, inputtype, q, L=None, tree=None, th=TH_HARD, t=None, estnoise=noisest, smooth="regular"):
    """Denoising.jl:483-600 for one signal.  q: the orthogonal filter's qmf (None = no reconstruction); (th, t) = dnt.th, dnt.t;
    estnoise: a function (x, redundant, tree) or a number."""
    assert smooth in ("undersmooth", "regular")
    assert inputtype in ("sig", "dwt", "wpt", "sdwt", "swpd", "acdwt", "acwpd")
    x = np.asarray(x)
    n = x.shape[-1]
    if L is None: L = maxtransformlevels(n)
    if tree is None: tree = maketree1(n, L, "dwt")
    if t is None: t = visushrink_t(n)
    h = g = P = Q = None
    if q is not None:
        g, h = makereverseqmfpair(q)
        P, Q = make_acreverseqmfpair(q)
    sig = lambda red, tr: estnoise(x, red, tr) if callable(estnoise) else float(estnoise)
    if inputtype == "sig":
        x = wpt(x, maketree1(n, L, "dwt"), h, g); inputtype = "dwt"
    if inputtype == "dwt":
        s = sig(False, None)
        if smooth == "regular":
            xt = threshold(x, th, s * t)
        else:
            n0 = n >> L
            xt = np.concatenate([x[:n0], threshold(x[n0:], th, s * t)])
        return xt if q is None else iwpt(xt, maketree1(n, L, "dwt"), h, g)
    if inputtype == "wpt":
        s = sig(False, tree)
        if smooth == "regular":
            xt = threshold(x, th, s * t)
        else:
            a, b = coarsestscalingrange(n, tree, False)
            xt = np.concatenate([x[a:b], threshold(x[b:], th, s * t)])
        return xt if q is None else iwpt(xt, tree, h, g)
    if inputtype in ("sdwt", "acdwt"):
        assert x.ndim > 1
        s = sig(True, None)
        xt = x.copy()
        if smooth == "regular":
            xt = threshold(xt, th, s * t)
        else:
            xt[1:] = threshold(x[1:], th, s * t)
        if inputtype == "sdwt":
            return xt if q is None else isdwt(xt, h, g)
        return iacdwt(xt)
    assert x.ndim > 1
    s = sig(True, tree)
    leaves = np.flatnonzero(np.asarray(getleaf(tree, "binary")))
    if smooth == "undersmooth":
        leaves = leaves[leaves != coarsestscalingrange(n, tree, True)]
    xt = x.copy()
    xt[leaves] = threshold(x[leaves], th, s * t)
    if inputtype == "swpd":
        return xt if q is None else iswpd(xt, tree, h, g)
    return iacwpd(xt, tree)


def denoiseall(X, inputtype, q, L=None, tree=None, th=TH_HARD, t=None, estnoise=noisest, bestTH=None, smooth="regular"):
    """Denoising.jl:651-713.  X: (N, n) or (N, K, n); estnoise: function or a vector of N numbers; bestTH: None or a function"""
    X = np.asarray(X)
    assert X.ndim > 1
    N, n = X.shape[0], X.shape[-1]
    if L is None: L = maxtransformlevels(n)
    if tree is None: tree = maketree1(n, L, "dwt")
    if inputtype == "sig":
        g, h = makereverseqmfpair(q)
        X = np.stack([wpt(X[i], maketree1(n, L, "dwt"), h, g) for i in range(N)]); inputtype = "dwt"
    if bestTH is None:
        est = [estnoise if callable(estnoise) else estnoise[i] for i in range(N)]
    else:
        if callable(estnoise):
            red = inputtype not in ("dwt", "wpt")
            tr = None if inputtype in ("dwt", "sdwt") else tree
            # the reference passes the tree for every other input type (incl. :acdwt, Denoising.jl:692-700)
            sg = [estnoise(X[i], red, tr) for i in range(N)]
        else:
            sg = list(estnoise)
        est = [float(bestTH(np.asarray(sg, np.float64)))] * N
    return np.stack([denoise(X[i], inputtype, q, L=L, tree=tree, th=th, t=t, estnoise=est[i], smooth=smooth) for i in range(N)])


# ------------------------------------------------------------------ SIWT steps and the nonstandard form (next row f-4)
def sidwt_step(v, h, g, s):
    """sidwt_step!  siwt/siwt_one_level.jl:71-98 (literal loop restatement, 1-based indices kept)"""
    v = np.asarray(v); n = len(v); n1 = n // 2; F = len(h)
    s = int(bool(s))
    mod1 = lambda a, m: (a - 1) % m + 1
    w1 = np.empty(n1, v.dtype); w2 = np.empty(n1, v.dtype)
    h = np.asarray(h, v.dtype); g = np.asarray(g, v.dtype)
    for i in range(1, n1 + 1):
        k1 = mod1(2 * i - 1 - s, n); k2 = 2 * i - s
        a1 = g[F - 1] * v[k1 - 1]; a2 = h[0] * v[k2 - 1]
        for j in range(2, F + 1):
            k1 = k1 + 1
            if k1 > n: k1 = mod1(k1, n)
            k2 = k2 - 1
            if k2 <= 0: k2 = mod1(k2, n)
            a1 = a1 + g[F - j] * v[k1 - 1]
            a2 = a2 + h[j - 1] * v[k2 - 1]
        w1[i - 1] = a1; w2[i - 1] = a2
    return w1, w2


def isidwt_step(w1, w2, h, g, s):
    """isidwt_step!  siwt/siwt_one_level.jl:153-184"""
    w1 = np.asarray(w1); w2 = np.asarray(w2); n1 = len(w1); n = 2 * n1; F = len(h)
    s = int(bool(s))
    mod1 = lambda a, m: (a - 1) % m + 1
    h = np.asarray(h, w1.dtype); g = np.asarray(g, w1.dtype)
    v = np.empty(n, w1.dtype)
    for i in range(1, n + 1):
        l = mod1(i - s, n)
        j0 = mod1(i, 2); j1 = F - j0 + 1; j2 = mod1(i + 1, 2)
        k1 = (i + 1) >> 1; k2 = (i + 1) >> 1
        acc = g[j1 - 1] * w1[k1 - 1] + h[j2 - 1] * w2[k2 - 1]
        for j in range(j0 + 2, F + 1, 2):
            j1 = F - j + 1
            j2 = j + (1 if j % 2 else -1)
            k1 = k1 - 1
            if k1 <= 0: k1 = mod1(k1, n1)
            k2 = k2 + 1
            if k2 > n1: k2 = mod1(k2, n1)
            acc = acc + (g[j1 - 1] * w1[k1 - 1] + h[j2 - 1] * w2[k2 - 1])
        v[l - 1] = acc
    return v


def ndyad(L, Lmax, gender):
    """wavemult/utils.jl:146-155 -> 0-based half-open (start, stop)"""
    assert L <= Lmax and L >= 1
    k = Lmax - L
    if gender:
        return (1 << (k + 1)) + (1 << k), 1 << (k + 2)
    return 1 << (k + 1), (1 << (k + 1)) + (1 << k)


def ns_dwt(x, q, L=None):
    """wavemult/transforms.jl:52-74"""
    x = np.asarray(x); n = len(x)
    Lmax = maxtransformlevels(n)
    if L is None: L = Lmax
    assert 1 <= L <= Lmax
    assert n & (n - 1) == 0
    g, h = makereverseqmfpair(q)
    nxw = np.zeros(2 * n, x.dtype)
    for l in range(1, L + 1):
        v = x if l == 1 else nxw[slice(*ndyad(l - 1, Lmax, False))]
        w1, w2 = dwt_step(v, h, g)
        nxw[slice(*ndyad(l, Lmax, False))] = w1
        nxw[slice(*ndyad(l, Lmax, True))] = w2
    nxw[:1 << (Lmax - L)] = nxw[slice(*ndyad(L, Lmax, False))]
    return nxw


def ns_idwt(nxw, q, L=None):
    """wavemult/transforms.jl:120-139"""
    nxw = np.asarray(nxw)
    Lmax = maxtransformlevels(len(nxw)) - 1
    if L is None: L = Lmax
    n = len(nxw) // 2
    assert 1 <= L <= Lmax
    assert n & (n - 1) == 0
    g, h = makereverseqmfpair(q)
    x = np.zeros(n, nxw.dtype)
    x[:1 << (Lmax - L)] = nxw[:1 << (Lmax - L)]
    for l in range(L, 0, -1):
        w1 = nxw[slice(*ndyad(l, Lmax, False))] + x[:1 << (Lmax - l)]
        w2 = nxw[slice(*ndyad(l, Lmax, True))]
        x[:1 << (Lmax - l + 1)] = idwt_step(w1, w2, h, g)
    return x


# ------------------------------------------------------------------ LDB object (fitdec! / transform), numpy restatement
def ldb_costs_topk(DM, top_k):
    """node costs of fitdec! LDB.jl:217-237 incl. the top_k < node size branch (sort descending, sum the first top_k)"""
    DM = np.asarray(DM)
    def cost(v):
        v = v.reshape(-1)
        return np.sort(v)[::-1][:top_k].sum() if top_k < v.size else v.sum()
    if DM.ndim == 2:
        K, n = DM.shape
        return np.array([cost(DM[d, j * (n >> d):(j + 1) * (n >> d)]) for d in range(K) for j in range(1 << d)], DM.dtype)
    K, nc_, nr = DM.shape
    out = []
    for i in range(1, (4 ** K - 1) // 3 + 1):
        d = getdepth(i, "quad")
        r0, c0, rr, cc = quadrange(nr, nc_, i)
        out.append(cost(DM[d, c0:c0 + cc, r0:r0 + rr]))
    return np.array(out, DM.dtype)


def discriminant_power_basis(DM, tree):
    """discriminant_power(D, tree, BasisDiscriminantMeasure()) ldb/ldb_measures.jl:427-439 -> (power, order 0-based)"""
    power = getbasiscoef(np.ascontiguousarray(DM), tree)
    order = np.argsort(-power.reshape(-1), kind="stable")
    return power, order


def discriminant_power_fisher(coefs, y):
    """discriminant_power(coefs, y, FishersClassSeparability()) ldb/ldb_measures.jl:441-479.  coefs (N, n[, m])"""
    coefs = np.asarray(coefs); y = list(y)
    classes = list(dict.fromkeys(y))
    N = coefs.shape[0]
    flat = coefs.reshape(N, -1)
    E = np.empty((flat.shape[1], len(classes)), coefs.dtype); V = np.empty_like(E); Ni = np.empty(len(classes), coefs.dtype)
    for i, c in enumerate(classes):
        idx = [k for k, v in enumerate(y) if v == c]
        Ni[i] = len(idx)
        E[:, i] = flat[idx].mean(axis=0)
        V[:, i] = flat[idx].var(axis=0, ddof=1)
    Ea = E.mean(axis=1, keepdims=True)
    p = Ni / Ni.sum()
    with np.errstate(divide="ignore", invalid="ignore"):
        power = (((E - Ea * E) ** 2) @ p) / (V @ p)
    order = np.argsort(-power, kind="stable")
    return power.reshape(coefs.shape[1:]), order


def ldb_fitdec(Xw, y, kind="are", p=2.0, top_k=None, dp="basis"):
    """fitdec! LDB.jl:186-245 (TimeFrequency energy map) -> dict(G, DM, cost, tree, DP, order)"""
    Xw = np.asarray(Xw)
    nelem = int(np.prod(Xw.shape[2:]))
    top_k = nelem if top_k is None else top_k
    G = energy_map_tf(Xw, y)
    DM = discriminant_measure(G, kind, p)
    cost = ldb_costs_topk(DM, top_k)
    if Xw.ndim == 3:
        tree = tree_select(cost.copy(), Xw.shape[2], None, "max")
    else:
        tree = tree_select(cost.copy(), Xw.shape[3], Xw.shape[2], "max")
    if dp == "basis":
        DP, order = discriminant_power_basis(DM, tree)
    else:
        Xc = np.stack([getbasiscoef(Xw[i], tree) for i in range(Xw.shape[0])])
        DP, order = discriminant_power_fisher(Xc, y)
    return dict(G=G, DM=DM, cost=cost, tree=tree, DP=DP, order=order)
