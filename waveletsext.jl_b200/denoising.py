"""Denoising of a batch of decomposed signals -- row f-3 of SURVEY.md section 8: the step between ``getbasiscoefall`` and the
inverse transform in the reference's pipeline (Denoising.jl).

The order statistics (``noisest``, ``surethreshold``, ``relerrorthreshold``) and the thresholding run on the GPU for all
signals of the batch at once (csrc/wx_denoise.cu); the reconstruction is the batched inverse of the path (``idwtall``,
``iwptall``, ``isdwtall``, ``iswpdall``, ``iacdwtall``, ``iacwpdall``).  The threshold types and ``VisuShrink`` come from
Wavelets.jl in the reference; their definitions are restated here (HardTH, SoftTH, SemiSoftTH, SteinTH -- the types
``denoise`` can be called with; ``BiggestTH`` / ``PosTH`` / ``NegTH`` take no Float64 threshold and fail there too).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _dev as D
from .utils import (maxtransformlevels, maketree, getleaf, isdyadic, nodelength, finestdetailrange, coarsestscalingrange)
from . import dwt as _dwt, swt as _swt, acwt as _acwt, dist as _dist

__all__ = ["HardTH", "SoftTH", "SemiSoftTH", "SteinTH", "VisuShrink", "RelErrorShrink", "SureShrink", "noisest", "surethreshold",
           "relerrorthreshold", "threshold", "threshold_", "denoise", "denoiseall"]


# ------------------------------------------------------------------ threshold types and DNFT objects
class THType:
    code = -1

    def __repr__(self):
        return f"{type(self).__name__}()"


class HardTH(THType):
    code = 0


class SoftTH(THType):
    code = 1


class SemiSoftTH(THType):
    code = 2


class SteinTH(THType):
    code = 3


class DNFT:
    th: THType
    t: float


class VisuShrink(DNFT):
    """``VisuShrink(th, t)``, ``VisuShrink(n)`` (HardTH, t = sqrt(2 log n), Wavelets.jl) and ``VisuShrink(n, th)`` Denoising.jl:117-119"""

    def __init__(self, a, b=None):
        if isinstance(a, THType):
            self.th, self.t = a, float(b)
        else:
            self.th = HardTH() if b is None else b
            self.t = math.sqrt(2 * math.log(int(a)))
        assert isinstance(self.th, THType)


class RelErrorShrink(DNFT):
    """Denoising.jl:42-48"""

    def __init__(self, th=None, t=1.0):
        self.th = HardTH() if th is None else th
        self.t = float(t)
        assert isinstance(self.th, THType)


class SureShrink(DNFT):
    """``SureShrink(th, t)`` Denoising.jl:61-67 and ``SureShrink(xw[, redundant, tree, th])`` :94-100 (t = surethreshold of ONE
    decomposed signal)"""

    def __init__(self, a, b=None, tree=None, th=None):
        if isinstance(a, THType):
            self.th, self.t = a, float(b)
        else:
            self.th = HardTH() if th is None else th
            self.t = float(surethreshold(a, bool(b) if b is not None else False, tree))
        assert isinstance(self.th, THType)


# ------------------------------------------------------------------ batched order statistics (device vectors of N Float64)
def _as_batch(x, redundant):
    """single signal -> batch of one.  Vectors are (n,), redundant tables (K, n)"""
    x = D.dev(x, "x")
    want = 2 if redundant else 1
    return (x.unsqueeze(0), True) if x.dim() == want else (x, False)


def _slab(X):
    """(N, n) -> n, K = 1 ; (N, K, n) -> n, K"""
    if X.dim() == 2:
        return int(X.shape[1]), 1
    assert X.dim() == 3, "denoising covers 1-D signals"
    return int(X.shape[2]), int(X.shape[1])


def _leafmask(tree, K, strict=True):
    """leaf columns of a redundant table.  strict: the reference indexes with the BitVector itself (``coef[:, leaves]``,
    Denoising.jl:155,299), which needs one entry per column; otherwise it goes through ``findall`` (:541,:561) and only the set
    entries have to exist"""
    leaves = np.asarray(getleaf(tree, "binary"), dtype=np.uint8)
    if strict:
        if len(leaves) != K:
            raise IndexError(f"BoundsError: attempt to access {K} columns with a {len(leaves)}-element leaf mask")
        return np.ascontiguousarray(leaves)
    if len(leaves) > K and leaves[K:].any():
        raise IndexError(f"BoundsError: the tree has leaves beyond the {K} nodes of the table")
    out = np.zeros(K, np.uint8)
    m = min(K, len(leaves))
    out[:m] = leaves[:m]
    return out


def _noisest_all(X, redundant, tree):
    n, K = _slab(X)
    assert isdyadic(n), "AssertionError: isdyadic(size(x,1))"
    if not redundant and tree is None:
        off, ln = n // 2, n - n // 2
    elif not redundant:
        r = finestdetailrange(n, tree, False); off, ln = r.start, len(r)
    elif tree is None:
        off, ln = (K - 1) * n, n
    else:
        _, i = finestdetailrange(n, tree, True); off, ln = (i - 1) * n, n
        if i > K:
            raise IndexError(f"BoundsError: node {i} of a table with {K} nodes")
    sigma = torch.empty(X.shape[0], dtype=torch.float64, device=X.device)
    D.call("noisest", X, D.ptr(sigma), D.ptr(X), n * K, off, ln, X.shape[0], D.stream(X))
    return sigma


def _colmask(X, redundant, tree):
    n, K = _slab(X)
    if not redundant or tree is None:
        return None
    return _leafmask(tree, K)


def _mask_ptr(mask):
    return mask.ctypes.data_as(C.c_void_p) if mask is not None else None


def _sure_all(X, redundant, tree):
    n, K = _slab(X)
    mask = _colmask(X, redundant, tree)
    t = torch.empty(X.shape[0], dtype=torch.float64, device=X.device)
    D.call("surethreshold", X, D.ptr(t), D.ptr(X), n, K, _mask_ptr(mask), X.shape[0], D.stream(X))
    return t


def _relerr_all(X, redundant, tree, elbows):
    assert elbows >= 1, "AssertionError: elbows >= 1"
    n, K = _slab(X)
    mask = _colmask(X, redundant, tree)
    t = torch.empty(X.shape[0], dtype=torch.float64, device=X.device)
    D.call("relerrorthreshold", X, D.ptr(t), D.ptr(X), n, K, _mask_ptr(mask), int(elbows), X.shape[0], D.stream(X))
    return t


def _scalar_or_vec(v, single):
    return float(v[0].item()) if single else v


def noisest(x, redundant: bool, tree=None):
    """``noisest(x, redundant[, tree])`` Denoising.jl:214-232: MAD of the finest detail coefficients / 0.6745.  A single decomposed
    signal returns a float; a batch (one more leading dimension) returns a device vector with one estimate per signal."""
    X, single = _as_batch(x, redundant)
    return _scalar_or_vec(_noisest_all(X, redundant, tree), single)


def surethreshold(coef, redundant: bool, tree=None):
    """``surethreshold(coef, redundant[, tree])`` Denoising.jl:142-166"""
    X, single = _as_batch(coef, redundant)
    return _scalar_or_vec(_sure_all(X, redundant, tree), single)


def relerrorthreshold(coef, redundant: bool = False, tree=None, elbows: int = 2):
    """``relerrorthreshold(coef[, redundant, tree, elbows])`` Denoising.jl:285-328 (``makeplot`` is not built)"""
    X, single = _as_batch(coef, redundant)
    return _scalar_or_vec(_relerr_all(X, redundant, tree, elbows), single)


# ------------------------------------------------------------------ thresholding
def _threshold_into(Y, X, th, t, colmask=None, keep=(0, 0)):
    """Y <- threshold of the slabs of X; t: float or a device vector of per-signal thresholds"""
    assert isinstance(th, THType) and th.code >= 0, "threshold type must be HardTH / SoftTH / SemiSoftTH / SteinTH"
    n, K = _slab(X)
    if isinstance(t, torch.Tensor):
        sig = t.to(device=X.device, dtype=torch.float64).contiguous()
        assert sig.numel() == X.shape[0]
        sptr, tmul = D.ptr(sig), 1.0
    else:
        assert t >= 0, "AssertionError: t >= 0"
        sptr, tmul = None, float(t)
    D.call("threshold", X, D.ptr(Y), D.ptr(X), n, K, _mask_ptr(colmask), int(keep[0]), int(keep[1]), th.code, sptr, tmul, X.shape[0], D.stream(X))
    return Y


def threshold_(x, th, t):
    """``threshold!(x, th, t)`` (Wavelets.jl) on every coefficient of a device array, in place"""
    xo = D.out(x, "x")
    flat = xo.t.view(-1, xo.t.shape[-1]) if xo.t.dim() > 1 else xo.t.view(1, -1)
    _threshold_into(flat, flat, th, float(t))
    return xo.commit()


def threshold(x, th, t):
    x = D.dev(x, "x")
    return threshold_(x.clone(), th, t)


# ------------------------------------------------------------------ denoise / denoiseall
_TYPES = ("sig", "dwt", "wpt", "sdwt", "swpd", "acdwt", "acwpd")


def _estimate(X, estnoise, redundant, tree):
    """per-signal sigma: the batched kernels for this module's estimators, a host loop for any other callable"""
    if estnoise is noisest:
        return _noisest_all(X, redundant, tree)
    if estnoise is relerrorthreshold:
        return _relerr_all(X, redundant, tree, 2)
    if estnoise is surethreshold:
        return _sure_all(X, redundant, tree)
    vals = [float(estnoise(X[i], redundant, tree)) for i in range(X.shape[0])]
    return torch.tensor(vals, dtype=torch.float64, device=X.device)


def denoiseall(x, inputtype: str, wt, L=None, tree=None, dnt=None, estnoise=noisest, bestTH=None, smooth: str = "regular", group=None):
    """``denoiseall(x, inputtype, wt; L, tree, dnt, estnoise, bestTH, smooth)`` Denoising.jl:651-713 over ``denoise`` :483-600.

    x: (N, n) signals / dwt / wpt coefficients or (N, K, n) redundant tables (Julia (n, N) / (n, K, N)).  ``estnoise``: one of
    this module's estimators (run for the whole batch on the GPU), any callable ``(x_i, redundant, tree) -> float``, or a vector
    of N noise levels.  ``bestTH``: None or a function of the vector of noise levels (``np.mean``, ``np.median``).  ``wt = None``
    returns the thresholded coefficients instead of reconstructing.  With an initialised process group x is this rank's shard of
    the batch: nothing is exchanged unless ``bestTH`` is given, in which case the N noise levels are gathered (rank order) so that
    the summary threshold is that of the whole batch."""
    inputtype = str(inputtype).lstrip(":")
    smooth = str(smooth).lstrip(":")
    assert inputtype in _TYPES, "AssertionError: inputtype in [:sig, :dwt, :wpt, :sdwt, :swpd, :acdwt, :acwpd]"
    assert smooth in ("regular", "undersmooth"), "AssertionError: smooth in [:regular, :undersmooth]"
    X = D.dev(x, "x")
    assert X.dim() > 1, "AssertionError: ndims(x) > 1"
    n = int(X.shape[-1])
    L = maxtransformlevels(n) if L is None else int(L)
    if tree is None:
        tree = maketree(n, L, "dwt")
    if dnt is None:
        dnt = VisuShrink(n)
    if inputtype == "sig":
        if wt is None:
            raise RuntimeError("inputtype=:sig not supported with wt=nothing")
        X = _dwt.dwtall(X, wt, L)
        inputtype = "dwt"
    redundant = inputtype not in ("dwt", "wpt")
    if redundant:
        assert X.dim() == 3, "AssertionError: ndims(x) > 1"
    # the tree argument each branch of `denoise` hands to estnoise (:483-600; the bestTH branch of denoiseall :692-700 passes the
    # tree for :acdwt too, which only matters for user-supplied estimators)
    etree = tree if inputtype in ("wpt", "swpd", "acwpd") else None
    if bestTH is not None and inputtype == "acdwt":
        etree = tree
    N = X.shape[0]
    if callable(estnoise):
        sigma = _estimate(X, estnoise, redundant, etree)
    else:
        sigma = torch.as_tensor(np.asarray(estnoise, dtype=np.float64)).to(X.device)
        assert sigma.numel() == N, "one noise level per signal"
    if bestTH is not None:
        s = float(bestTH(_dist.allgather_host_vector(sigma.cpu().numpy(), group)))
        t = s * dnt.t
    else:
        t = sigma * dnt.t
    # which coefficients are thresholded
    K = _slab(X)[1]
    colmask, keep = None, (0, 0)
    if inputtype == "dwt":
        if smooth == "undersmooth":
            keep = (0, nodelength(n, L))
    elif inputtype == "wpt":
        if smooth == "undersmooth":
            r = coarsestscalingrange(n, tree, False); keep = (r.start, r.stop)
    elif inputtype in ("sdwt", "acdwt"):
        if smooth == "undersmooth":
            keep = (0, n)
    else:
        colmask = _leafmask(tree, K, strict=False)
        if smooth == "undersmooth":
            _, node = coarsestscalingrange(n, tree, True)
            colmask[node - 1] = 0
    Xt = _threshold_into(torch.empty_like(X), X, dnt.th, t, colmask, keep)
    if inputtype == "dwt":
        return Xt if wt is None else _dwt.idwtall(Xt, wt, L)
    if inputtype == "wpt":
        return Xt if wt is None else _dwt.iwptall(Xt, wt, tree)
    if inputtype == "sdwt":
        return Xt if wt is None else _swt.isdwtall(Xt, wt)
    if inputtype == "swpd":
        return Xt if wt is None else _swt.iswpdall(Xt, wt, tree)
    if inputtype == "acdwt":
        return _acwt.iacdwtall(Xt)
    return _acwt.iacwpdall(Xt, tree)


def denoise(x, inputtype: str, wt, L=None, tree=None, dnt=None, estnoise=noisest, smooth: str = "regular"):
    """``denoise(x, inputtype, wt; ...)`` Denoising.jl:483-600 for one signal: the batch path with N = 1.  ``estnoise`` may be a number."""
    x = D.dev(x, "x")
    it = str(inputtype).lstrip(":")
    if it in ("sdwt", "swpd", "acdwt", "acwpd"):
        assert x.dim() > 1, "AssertionError: ndims(x) > 1"
    if not callable(estnoise):
        estnoise = [float(estnoise)]
    return denoiseall(x.unsqueeze(0), it, wt, L=L, tree=tree, dnt=dnt, estnoise=estnoise, smooth=smooth)[0]
