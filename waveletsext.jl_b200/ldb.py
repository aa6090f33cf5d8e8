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
           "energy_map", "discriminant_measure", "ldb_tree"]


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
