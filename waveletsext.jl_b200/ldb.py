"""Local discriminant basis, first half of ``fitdec!`` (LDB.jl:186-245): time-frequency energy maps per class, discriminant
measure, node costs and the :max tree selection -- row f-2 of SURVEY.md section 8.

The per-class energy sums are accumulated on the GPU over the LOCAL shard of the batch and all-reduced like the JBB moments;
everything after that is per-position work on a (sz, K) map.  Feature ordering / ``transform`` of the LDB object stay with the
reference (they consume the tree and the expansion coefficients this module returns).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _dev as D
from . import _lib
from . import dist
from .bestbasis import _geom, _ncosts, bestbasis_treeselection

__all__ = ["TimeFrequency", "AsymmetricRelativeEntropy", "SymmetricRelativeEntropy", "LpDistance", "HellingerDistance",
           "energy_map", "discriminant_measure", "ldb_tree", "BasisDiscriminantMeasure", "FishersClassSeparability",
           "RobustFishersClassSeparability", "discriminant_power", "LocalDiscriminantBasis", "fit_", "fitdec_", "transform",
           "fit_transform", "inverse_transform", "change_nfeatures"]


@dataclass(frozen=True)
class TimeFrequency:
    """ldb/ldb_energymap.jl:21"""


@dataclass(frozen=True)
class AsymmetricRelativeEntropy:
    """ldb/ldb_measures.jl:43"""


@dataclass(frozen=True)
class SymmetricRelativeEntropy:
    """ldb/ldb_measures.jl:62"""


@dataclass(frozen=True)
class LpDistance:
    """ldb/ldb_measures.jl:74-76"""
    p: float = 2


@dataclass(frozen=True)
class HellingerDistance:
    """ldb/ldb_measures.jl:88"""


def _labels(y, device, classes=None):
    """class index of every signal.  classes = None: Julia's unique(y), i.e. order of first appearance in THIS label list; when the
    batch is sharded over ranks pass the global class list so that every rank numbers the classes identically."""
    ylist = y.tolist() if isinstance(y, (np.ndarray, torch.Tensor)) else list(y)
    if classes is None:
        classes = list(dict.fromkeys(ylist))
    index = {v: i for i, v in enumerate(classes)}
    lab = np.array([index[v] for v in ylist], dtype=np.int32)
    return torch.from_numpy(lab).to(device), list(classes)


def _energy_sums(Xw, y, group=None, classes=None):
    Xw = D.dev(Xw, "Xw")
    m, n, K, Nloc, szK = _geom(Xw)
    assert len(y) == Nloc, "AssertionError: Nx == Ny"
    lab, classes = _labels(y, Xw.device, classes)
    nc = len(classes)
    assert nc > 1, "AssertionError: nc > 1"
    esum = torch.empty((nc, szK), dtype=torch.float64, device=Xw.device)
    D.call("energy_map_tf", Xw, D.ptr(esum), D.ptr(Xw), D.ptr(lab), nc, szK, Nloc, D.stream(Xw))
    dist.allreduce_sum(esum, group)
    sz0 = szK // K
    norm = esum[:, :sz0].sum(dim=1)                          # sum over the class of ||x_k||^2 (level 0 = the signals themselves)
    return esum, norm, (m, n, K, szK), classes


def energy_map(Xw, y, method=None, group=None, classes=None):
    """``energy_map(Xw, y, TimeFrequency())`` ldb/ldb_energymap.jl:109-141.  Xw: local shard (N, K, n[, m]) on the device, y: its
    labels.  Returns Γ as a device tensor (nc, K, n[, m]) (the reference's (sz..., L, nc) in reversed-axes convention); with an
    initialised process group the maps are those of the concatenated batch (pass the global ``classes`` list then, so that
    every rank numbers the classes identically)."""
    method = TimeFrequency() if method is None else method
    if not isinstance(method, TimeFrequency):
        raise TypeError("only the TimeFrequency energy map runs on the B200 path")
    esum, norm, (m, n, K, szK), _ = _energy_sums(Xw, y, group, classes)
    G = (esum / norm[:, None]).to(Xw.dtype)
    return G.reshape((esum.shape[0],) + tuple(Xw.shape[1:]))


_KIND = {AsymmetricRelativeEntropy: 0, SymmetricRelativeEntropy: 1, LpDistance: 2, HellingerDistance: 3}


def discriminant_measure(G, dm=None):
    """``discriminant_measure(Γ, dm)`` ldb/ldb_measures.jl:139-183 for time-frequency maps: Γ (nc, K, n[, m]) -> D (K, n[, m])"""
    dm = AsymmetricRelativeEntropy() if dm is None else dm
    if type(dm) not in _KIND:
        raise TypeError(f"unsupported discriminant measure {type(dm).__name__}")
    G = D.dev(G, "Γ")
    nc = G.shape[0]
    assert nc > 1, "AssertionError: nc > 1"
    szK = int(np.prod(G.shape[1:]))
    g64 = G.reshape(nc, szK).to(torch.float64).contiguous()
    ones = torch.ones(nc, dtype=torch.float64, device=G.device)
    out = torch.empty(szK, dtype=torch.float64, device=G.device)
    with torch.cuda.device(G.device):
        _lib.call("wx_ldb_discriminant", D.ptr(out), D.ptr(g64), D.ptr(ones), nc, szK, _KIND[type(dm)],
                  C.c_double(float(getattr(dm, "p", 0.0))), G.element_size(), D.stream(G))
    return out.to(G.dtype).reshape(tuple(G.shape[1:]))


def ldb_tree(Xw, y, dm=None, en=None, group=None, classes=None):
    """The tree search of ``fitdec!`` (LDB.jl:209-240) with ``top_k`` = all coefficients: energy maps -> discriminant measure ->
    node costs (sum over the node) -> ``bestbasis_treeselection(cost, sz..., :max)``.  Returns (Γ, DM, cost, tree)."""
    dm = AsymmetricRelativeEntropy() if dm is None else dm
    G = energy_map(Xw, y, en, group, classes)
    DM = discriminant_measure(G, dm)
    m, n, K, _, szK = _geom(Xw)
    costs = np.empty(_ncosts(m, K, False), np.float64)
    d64 = DM.reshape(-1).to(torch.float64).contiguous()
    with torch.cuda.device(Xw.device):
        _lib.call("wx_node_costs", costs.ctypes.data, D.ptr(d64), m, n, K, 0, C.c_double(1.0), Xw.element_size(), D.stream(Xw))
    if Xw.dim() == 3:
        tree = bestbasis_treeselection(costs.copy(), Xw.shape[2], "max")
    else:
        tree = bestbasis_treeselection(costs.copy(), Xw.shape[3], Xw.shape[2], "max")
    return G, DM, costs, tree


# ------------------------------------------------------------------ the LDB object (LDB.jl:89-470)
@dataclass(frozen=True)
class BasisDiscriminantMeasure:
    """ldb/ldb_measures.jl:360"""


@dataclass(frozen=True)
class FishersClassSeparability:
    """ldb/ldb_measures.jl:374"""


@dataclass(frozen=True)
class RobustFishersClassSeparability:
    """ldb/ldb_measures.jl:388 (medians / MADs per class: not on the B200 path)"""


def _node_costs_topk(DM, top_k):
    """LDB.jl:217-237 with top_k < node size: per node, the sum of the top_k largest discriminant measures.  The (sz, K) map is a
    per-position summary, not batch data: like the tree selection it is finished on the host."""
    from .utils import getrowrange, getcolrange, getdepth
    Dh = DM.to(torch.float64).cpu().numpy()
    cost = lambda v: np.sort(v.reshape(-1))[::-1][:top_k].sum() if top_k < v.size else v.sum()
    if Dh.ndim == 2:
        K, n = Dh.shape
        return np.array([cost(Dh[d, j * (n >> d):(j + 1) * (n >> d)]) for d in range(K) for j in range(1 << d)])
    K, ncol, nrow = Dh.shape
    out = []
    for i in range(1, (4 ** K - 1) // 3 + 1):
        d = getdepth(i, "quad")
        rr, cc = getrowrange(nrow, i), getcolrange(ncol, i)
        out.append(cost(Dh[d, cc.start:cc.stop, rr.start:rr.stop]))
    return np.array(out)


def discriminant_power(a, b, dp=None):
    """``discriminant_power(D, tree, BasisDiscriminantMeasure())`` ldb/ldb_measures.jl:427-439 (a = DM map on the device, b = tree) /
    ``discriminant_power(coefs, y, FishersClassSeparability())`` :441-479 (a = best-basis coefficients (N, sz...), b = labels).
    Returns (power as a host array of the signal's shape, order as 0-based indices into the flattened signal)."""
    from .dwt import getbasiscoef
    dp = BasisDiscriminantMeasure() if dp is None else dp
    if isinstance(dp, BasisDiscriminantMeasure):
        DM = D.dev(a, "D")
        power = getbasiscoef(DM, b).to(torch.float64).cpu().numpy()
    elif isinstance(dp, FishersClassSeparability):
        Xc = D.dev(a, "coefs")
        N = Xc.shape[0]
        nelem = int(np.prod(Xc.shape[1:]))
        lab, classes = _labels(b, "cpu")
        lab = lab.numpy()
        nc = len(classes)
        sig = np.argsort(lab, kind="stable").astype(np.int32)
        off = np.concatenate([[0], np.cumsum(np.bincount(lab, minlength=nc))]).astype(np.int32)
        E = torch.empty((nc, nelem), dtype=torch.float64, device=Xc.device); V = torch.empty_like(E)
        sig_d, off_d = torch.from_numpy(sig).to(Xc.device), torch.from_numpy(off).to(Xc.device)      # named: they must outlive the launch
        D.call("class_moments", Xc, D.ptr(E), D.ptr(V), D.ptr(Xc), D.ptr(sig_d), D.ptr(off_d), nc, nelem, D.stream(Xc))
        E, V = E.cpu().numpy().T, V.cpu().numpy().T           # (nelem, nc) like the reference's Eαᵢ, Varαᵢ
        if Xc.dtype == torch.float32:
            E, V = E.astype(np.float32), V.astype(np.float32)
        Ni = np.diff(off).astype(E.dtype)
        Ea = E.mean(axis=1, keepdims=True)
        p = Ni / Ni.sum()
        with np.errstate(divide="ignore", invalid="ignore"):
            power = ((((E - Ea * E) ** 2) @ p) / (V @ p)).reshape(tuple(Xc.shape[1:]))
    else:
        raise TypeError(f"{type(dp).__name__} is not available on the B200 path (BasisDiscriminantMeasure, FishersClassSeparability are)")
    order = np.argsort(-power.reshape(-1), kind="stable")
    return power, order


class LocalDiscriminantBasis:
    """``LocalDiscriminantBasis(wt=, max_dec_level=, dm=, en=, dp=, top_k=, n_features=)`` LDB.jl:89-110 with the time-frequency
    energy map.  ``fit_``, ``fitdec_``, ``transform``, ``fit_transform``, ``inverse_transform`` and ``change_nfeatures`` below take the
    object first, like the reference's functions; signals and features are device arrays (N, sz...) / (N, n_features)."""

    def __init__(self, wt=None, max_dec_level=None, dm=None, en=None, dp=None, top_k=None, n_features=None):
        from .filters import wavelet
        self.wt = wavelet("haar") if wt is None else wt
        self.max_dec_level = max_dec_level
        self.dm = AsymmetricRelativeEntropy() if dm is None else dm
        self.en = TimeFrequency() if en is None else en
        self.dp = BasisDiscriminantMeasure() if dp is None else dp
        self.top_k, self.n_features = top_k, n_features
        self.sz = self.Γ = self.DM = self.cost = self.tree = self.DP = self.order = None


def _sz_of(X):
    return tuple(reversed(tuple(X.shape[1:])))


def fit_(f: LocalDiscriminantBasis, X, y, group=None, classes=None):
    """``fit!(f, X, y)`` LDB.jl:139-156"""
    from .dwt import wpdall
    from .utils import maxtransformlevels
    X = D.dev(X, "X")
    assert 2 <= X.dim() <= 3, "AssertionError: 2 <= ndims(X) <= 3"
    L = maxtransformlevels(min(_sz_of(X)))
    f.max_dec_level = L if f.max_dec_level is None else f.max_dec_level
    assert 1 <= f.max_dec_level <= L, "AssertionError: 1 <= f.max_dec_level <= L"
    fitdec_(f, wpdall(X, f.wt, f.max_dec_level), y, group, classes)


def fitdec_(f: LocalDiscriminantBasis, Xw, y, group=None, classes=None):
    """``fitdec!(f, Xw, y)`` LDB.jl:186-245.  With an initialised process group Xw / y are this rank's shard (only the energy sums
    are all-reduced; FishersClassSeparability then needs the whole batch on one rank)."""
    from .dwt import getbasiscoefall
    from .utils import maxtransformlevels
    Xw = D.dev(Xw, "Xw")
    assert 3 <= Xw.dim() <= 4, "AssertionError: 3 <= ndims(Xw) <= 4"
    f.sz = tuple(reversed(tuple(Xw.shape[2:])))
    L = Xw.shape[1]
    nelem = int(np.prod(f.sz))
    f.top_k = nelem if f.top_k is None else f.top_k
    f.n_features = nelem if f.n_features is None else f.n_features
    f.max_dec_level = L - 1 if f.max_dec_level is None else f.max_dec_level
    assert Xw.shape[0] == len(y), "AssertionError: Nx == Ny"
    assert 1 <= f.top_k <= nelem, "AssertionError: 1 <= f.top_k <= nelem"
    assert 1 <= f.n_features <= nelem, "AssertionError: 1 <= f.n_features <= nelem"
    assert f.max_dec_level + 1 == L, "AssertionError: f.max_dec_level+1 == L"
    assert 1 <= f.max_dec_level <= maxtransformlevels(min(f.sz)), "AssertionError: 1 <= f.max_dec_level <= maxtransformlevels"
    f.Γ = energy_map(Xw, y, f.en, group, classes)
    f.DM = discriminant_measure(f.Γ, f.dm)
    m, n, K, _, szK = _geom(Xw)
    if f.top_k >= nelem:
        costs = np.empty(_ncosts(m, K, False), np.float64)
        d64 = f.DM.reshape(-1).to(torch.float64).contiguous()
        with torch.cuda.device(Xw.device):
            _lib.call("wx_node_costs", costs.ctypes.data, D.ptr(d64), m, n, K, 0, C.c_double(1.0), Xw.element_size(), D.stream(Xw))
    else:
        costs = _node_costs_topk(f.DM, f.top_k)
    f.cost = costs
    f.tree = bestbasis_treeselection(costs.copy(), *f.sz, "max")
    if isinstance(f.dp, BasisDiscriminantMeasure):
        f.DP, f.order = discriminant_power(f.DM, f.tree, f.dp)
    else:
        f.DP, f.order = discriminant_power(getbasiscoefall(Xw, f.tree), y, f.dp)


def _check_fitted(f):
    for name in ("max_dec_level", "top_k", "n_features", "sz", "Γ", "DM", "cost", "tree", "DP", "order"):
        assert getattr(f, name) is not None, f"AssertionError: !isnothing(f.{name})"


def _select(f, Xb):
    """(N, sz...) best-basis coefficients -> (N, n_features) in f.order"""
    N = Xb.shape[0]
    nelem = int(np.prod(f.sz))
    order = torch.from_numpy(np.ascontiguousarray(f.order[:f.n_features]).astype(np.int32)).to(Xb.device)
    out = Xb.new_empty((N, f.n_features))
    D.call("select_features", Xb, D.ptr(out), D.ptr(Xb), D.ptr(order), f.n_features, nelem, N, D.stream(Xb))
    return out


def transform(f: LocalDiscriminantBasis, X):
    """``transform(f, X)`` LDB.jl:270-296: ``wptall`` along the fitted tree, then the n_features most discriminant coefficients"""
    from .dwt import wptall
    from .utils import maxtransformlevels
    X = D.dev(X, "X")
    assert 2 <= X.dim() <= 3, "AssertionError: 2 <= ndims(X) <= 3"
    _check_fitted(f)
    assert _sz_of(X) == tuple(f.sz), "AssertionError: sz == f.sz"
    return _select(f, wptall(X, f.wt, f.tree))


def fit_transform(f: LocalDiscriminantBasis, X, y, group=None, classes=None):
    """``fit_transform(f, X, y)`` LDB.jl:333-353"""
    from .dwt import wpdall, getbasiscoefall
    from .utils import maxtransformlevels
    X = D.dev(X, "X")
    assert 2 <= X.dim() <= 3, "AssertionError: 2 <= ndims(X) <= 3"
    L = maxtransformlevels(min(_sz_of(X)))
    f.max_dec_level = L if f.max_dec_level is None else f.max_dec_level
    assert 1 <= f.max_dec_level <= L, "AssertionError: 1 <= f.max_dec_level <= L"
    Xw = wpdall(X, f.wt, f.max_dec_level)
    fitdec_(f, Xw, y, group, classes)
    return _select(f, getbasiscoefall(Xw, f.tree))


def inverse_transform(f: LocalDiscriminantBasis, Xf):
    """``inverse_transform(f, X)`` LDB.jl:366-381: features back into a zero coefficient array, then ``iwptall``"""
    from .dwt import iwptall
    Xf = D.dev(Xf, "X")
    assert Xf.dim() == 2 and Xf.shape[1] == f.n_features, "AssertionError: size(X,1) == f.n_features"
    N = Xf.shape[0]
    nelem = int(np.prod(f.sz))
    order = torch.from_numpy(np.ascontiguousarray(f.order[:f.n_features]).astype(np.int32)).to(Xf.device)
    Xc = Xf.new_empty((N,) + tuple(reversed(f.sz)))
    D.call("scatter_features", Xf, D.ptr(Xc), D.ptr(Xf), D.ptr(order), f.n_features, nelem, N, D.stream(Xf))
    return iwptall(Xc, f.wt, f.tree)


def change_nfeatures(f: LocalDiscriminantBasis, x, n_features: int):
    """``change_nfeatures(f, x, n_features)`` LDB.jl:431-452"""
    import warnings
    x = D.dev(x, "x")
    assert f.n_features is not None, "AssertionError: !isnothing(f.n_features)"
    if x.shape[1] != f.n_features:
        raise ValueError("f.n_features and number of rows of x do not match!")
    assert 1 <= n_features <= int(np.prod(f.sz)), "AssertionError: 1 <= n_features <= prod(f.sz)"
    if f.n_features >= n_features:
        f.n_features = n_features
        return x[:, :n_features].contiguous()
    warnings.warn("Proposed n_features larger than currently saved n_features. Results will be less accurate since inverse_transform "
                  "and transform is involved.")
    X = inverse_transform(f, x)
    f.n_features = n_features
    return transform(f, X)
