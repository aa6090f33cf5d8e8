"""Orthogonal wavelet filters and the filter pairs the transforms consume.

Mirrors the pieces of Wavelets.jl (third party, NOT under /root/reference; Project.toml:30
compat "0.9, 0.10") that the reference calls on this path:

* ``wavelet(WT.db4)`` -> ``OrthoFilter`` holding ``qmf`` (Float64) and a name;
* ``WT.makereverseqmfpair(wt, true)`` -> ``g = reverse(qmf)`` (scaling), ``h = qmf .* (-1)^(0:F-1)``
  (detail) -- call sites DWT.jl:141,174,365,510,672, SWT.jl:119,...; bound as ``g, h = ...``;
* ``make_acreverseqmfpair`` (reference, acwt/acwt_utils.jl:7-72).

No tap table exists offline (no Wavelets.jl / PyWavelets in the image), so:
  haar/dbN   : spectral factorisation of the Daubechies half-band polynomial (minimum phase),
               same construction Wavelets.jl uses; validated by orthonormality + vanishing moments.
  symN       : same polynomial, root subset = the classical least-asymmetric choice (selected by
               minimising phase non-linearity; sym4/sym8 cross-checked against recalled tables).
  coif4      : (12 taps) Gauss-Newton solve of the coiflet moment equations from a recalled start.
Taps always cross the C ABI as data, exactly as in the Julia shim, so parity never depends on how a
table was obtained.
"""
from __future__ import annotations

import itertools
import math
from dataclasses import dataclass
from functools import lru_cache

import numpy as np

__all__ = ["OrthoFilter", "wavelet", "WT", "makereverseqmfpair", "makeqmfpair",
           "autocorr", "pfilter", "qfilter", "make_acqmfpair", "make_acreverseqmfpair",
           "check_orthonormal"]


@dataclass(frozen=True)
class OrthoFilter:
    """Wavelets.jl ``OrthoFilter``: ``qmf`` scaling filter (sum = sqrt 2, unit norm) + name."""
    qmf: tuple
    name: str

    @property
    def taps(self) -> np.ndarray:
        return np.asarray(self.qmf, dtype=np.float64)

    def __len__(self) -> int:
        return len(self.qmf)


def _halfband_roots(N: int) -> np.ndarray:
    """roots y_k of P(y) = sum_{k<N} C(N-1+k,k) y^k  (Daubechies)"""
    c = [math.comb(N - 1 + k, k) for k in range(N)]
    return np.roots(c[::-1]) if N > 1 else np.zeros(0)


def _z_pair(y: complex):
    """z + 1/z = 2 - 4y  -> (z_inside, z_outside)"""
    b = 2 - 4 * y
    r = np.sqrt(b * b - 4 + 0j)
    z1, z2 = (b + r) / 2, (b - r) / 2
    return (z1, z2) if abs(z1) < abs(z2) else (z2, z1)


def _poly_from(N: int, zroots) -> np.ndarray:
    p = np.array([1.0 + 0j])
    for _ in range(N):
        p = np.convolve(p, [1.0, 1.0])
    for z in zroots:
        p = np.convolve(p, [1.0, -z])
    q = np.real(p)
    q = q * (math.sqrt(2.0) / q.sum())
    return q


@lru_cache(maxsize=None)
def _daubechies(N: int) -> tuple:
    if N == 1:
        return (1 / math.sqrt(2.0), 1 / math.sqrt(2.0))
    zs = [_z_pair(y)[0] for y in _halfband_roots(N)]
    q = _poly_from(N, zs)
    # minimum-phase roots give Wavelets.jl's orientation: 0.2304, 0.7148, 0.6309, ... for db4
    return tuple(float(v) for v in q)


# recalled low-precision anchors (rec_lo orientation) used ONLY to pick the root subset / start Newton
_SYM_ANCHOR = {
    4: [0.0322231006, -0.0126039673, -0.0992195436, 0.2978577956, 0.8037387518, 0.4976186676,
        -0.0296355276, -0.0757657148],
    8: [0.0018899503, -0.0003029205, -0.0149522583, 0.0038087520, 0.0491371797, -0.0272190299,
        -0.0519458381, 0.3644418948, 0.7771857517, 0.4813596513, -0.0612733591, -0.1432942384,
        0.0076074873, 0.0316950878, -0.0005421323, -0.0033824160],
}
_COIF4_ANCHOR = [0.016387336463522112, -0.04146493678175915, -0.06737255472196302, 0.3861100668211622,
                 0.8127236354455423, 0.41700518442169254, -0.0764885990783064, -0.0594344186464569,
                 0.023680171946334084, 0.0056114348193944995, -0.0018232088707029932,
                 -0.0007205494453645122]


@lru_cache(maxsize=None)
def _symlet(N: int) -> tuple:
    ys = _halfband_roots(N)
    groups = []       # one entry per real root / conjugate pair: ([inside...], [outside...])
    used = np.zeros(len(ys), bool)
    for i, y in enumerate(ys):
        if used[i]:
            continue
        used[i] = True
        zi, zo = _z_pair(y)
        if abs(y.imag) < 1e-9:
            groups.append(([zi.real + 0j], [zo.real + 0j]))
        else:
            j = int(np.argmin([abs(yy - np.conj(y)) + (1e9 if used[k] else 0) for k, yy in enumerate(ys)]))
            used[j] = True
            groups.append(([zi, np.conj(zi)], [zo, np.conj(zo)]))
    cands = []
    for pick in itertools.product((0, 1), repeat=len(groups)):
        zs = [z for g, p in zip(groups, pick) for z in g[p]]
        cands.append(_poly_from(N, zs))
    anchor = _SYM_ANCHOR.get(N)
    if anchor is not None:
        a = np.asarray(anchor)
        best = min(cands, key=lambda q: min(np.abs(q - a).max(), np.abs(q[::-1] - a).max()))
        if np.abs(best[::-1] - a).max() < np.abs(best - a).max():
            best = best[::-1]
        if np.abs(best - a).max() > 1e-6:
            raise RuntimeError(f"sym{N}: no root subset matches the recalled table")
    else:       # unreachable through wavelet(): only anchored symlets are advertised
        raise ValueError(f"sym{N}: no verified table to select the root subset and orientation against")
    return tuple(float(v) for v in np.array(best))


@lru_cache(maxsize=None)
def _coif4() -> tuple:
    """12-tap coiflet (2K = 4 vanishing moments for both functions). Gauss-Newton on
    sum, orthonormality, wavelet moments p<4, scaling moments 1<=p<4 about the integer centre."""
    q = np.array(_COIF4_ANCHOR)
    F = len(q)
    k = np.arange(F, dtype=np.float64)
    c = round(float(np.dot(k, q) / q.sum()))
    sgn = (-1.0) ** np.arange(F)
    for _ in range(20):
        res, J = [q.sum() - math.sqrt(2.0)], [np.ones(F)]
        for m in range(F // 2):
            sh = 2 * m
            res.append(float(np.dot(q[: F - sh], q[sh:])) - (1.0 if m == 0 else 0.0))
            g = np.zeros(F); g[: F - sh] += q[sh:]; g[sh:] += q[: F - sh]
            J.append(g)
        for p in range(4):
            row = sgn * (k - c) ** p
            res.append(float(np.dot(row, q))); J.append(row)
        for p in range(1, 4):
            row = (k - c) ** p
            res.append(float(np.dot(row, q))); J.append(row)
        dq, *_ = np.linalg.lstsq(np.array(J), -np.array(res), rcond=None)
        q = q + dq
        if np.abs(dq).max() < 1e-17:
            break
    if np.abs(q - np.array(_COIF4_ANCHOR)).max() > 1e-6:
        raise RuntimeError("coif4 solve drifted from the recalled table")
    return tuple(float(v) for v in q)


class _WT:
    """``WT.db4``-style names (Wavelets.jl ``WT`` module constants)."""
    def __getattr__(self, name: str) -> str:
        if name.startswith("_"):
            raise AttributeError(name)
        return name


WT = _WT()


def wavelet(name: str) -> OrthoFilter:
    """``wavelet(WT.db4)`` -> OrthoFilter.  Supported: haar, db1..db16 (spectral factorisation, the construction Wavelets.jl
    itself uses for dbN), sym4, sym8 and coif4 (each checked against a recalled copy of the tabulated filter).  Wavelets.jl's
    other table filters (sym5-7, sym9-10, coif2/6/8, batt*, beyl, vaid) have no verified table in this offline image: they
    raise instead of returning taps that may differ from the reference's by a root choice or a time reversal -- pass the taps
    as data (``OrthoFilter(qmf, name)``), which is what crosses the C ABI anyway."""
    name = str(name).lower()
    if name in ("haar", "db1"):
        return OrthoFilter(_daubechies(1), "haar")
    if name.startswith("db") and name[2:].isdigit() and 1 <= int(name[2:]) <= 16:
        return OrthoFilter(_daubechies(int(name[2:])), name)
    if name in ("sym4", "sym8"):
        return OrthoFilter(_symlet(int(name[3:])), name)
    if name == "coif4":
        return OrthoFilter(_coif4(), name)
    raise ValueError(f"unknown or unverified wavelet class {name!r}: supported names are haar, db1..db16, sym4, sym8, coif4; "
                     f"for any other filter pass its qmf taps as data: OrthoFilter(taps, name)")


def _qmf(wt) -> np.ndarray:
    if isinstance(wt, OrthoFilter):
        return wt.taps
    return np.ascontiguousarray(wt, dtype=np.float64)


def makereverseqmfpair(wt, fw: bool = True):
    """Wavelets.jl ``WT.makereverseqmfpair(f, fw)`` -> (scfilter, dcfilter); the reference binds it
    as ``g, h`` (g = scaling, h = detail)."""
    q = _qmf(wt)
    mirror = q * (-1.0) ** np.arange(len(q))
    if fw:
        return q[::-1].copy(), mirror
    return q.copy(), mirror[::-1].copy()


def makeqmfpair(wt, fw: bool = True):
    sc, dc = makereverseqmfpair(wt, fw)
    return sc[::-1].copy(), dc[::-1].copy()


def autocorr(wt) -> np.ndarray:
    """acwt/acwt_utils.jl:7-18"""
    H = _qmf(wt)
    l = len(H)
    result = np.zeros(l - 1)
    for k in range(1, l):
        acc = 0.0
        for i in range(1, l - k + 1):
            acc += H[i - 1] * H[i + k - 1]
        result[k - 1] = acc * 2
    return result


def pfilter(wt) -> np.ndarray:
    """acwt/acwt_utils.jl:27-33"""
    a = autocorr(wt)
    c1 = 1 / math.sqrt(2.0)
    c2 = c1 / 2
    b = c2 * a
    return np.concatenate([b[::-1], [c1], b])


def qfilter(wt) -> np.ndarray:
    """acwt/acwt_utils.jl:42-48"""
    a = autocorr(wt)
    c1 = 1 / math.sqrt(2.0)
    c2 = c1 / 2
    b = -c2 * a
    return np.concatenate([b[::-1], [c1], b])


def make_acqmfpair(wt):
    """acwt/acwt_utils.jl:57-60"""
    return pfilter(wt), qfilter(wt)


def make_acreverseqmfpair(wt):
    """acwt/acwt_utils.jl:69-72 -> (reverse(P), reverse(Q))"""
    p, q = make_acqmfpair(wt)
    return p[::-1].copy(), q[::-1].copy()


def check_orthonormal(q) -> float:
    """max violation of sum = sqrt2 and even-shift orthonormality (used by tests)."""
    q = np.asarray(q, dtype=np.float64)
    F = len(q)
    err = abs(q.sum() - math.sqrt(2.0))
    for m in range(F // 2):
        v = float(np.dot(q[: F - 2 * m], q[2 * m:])) - (1.0 if m == 0 else 0.0)
        err = max(err, abs(v))
    return err
