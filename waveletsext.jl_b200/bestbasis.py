"""Joint best basis (JBB) and least statistically dependent basis (LSDB) on the B200: host mirror of BestBasis.jl,
bestbasis/bestbasis_tree.jl and bestbasis/bestbasis_costs.jl (JBB / LSDB rows of the hot path).

``tree_costs`` / ``bestbasistree`` are single calls into libwx_b200 (``wx_tree_costs_*`` / ``wx_bestbasistree_*``): the
per-position reductions on the local shard of the batch, the NCCL exchange of the small per-position state over the
library's own communicator when a process group is initialised (``dist.comm``), the per-node costs and the O(n) host
selection pass of the reference all run inside the library.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import ctypes as C
import numpy as np
import torch

from . import _dev as D
from . import _lib
from . import dist
from .utils import gettreelength

__all__ = ["LoglpCost", "NormCost", "DifferentialEntropyCost", "ShannonEntropyCost", "LogEnergyEntropyCost", "JBB", "LSDB", "BB",
           "tree_costs", "bestbasis_treeselection", "bestbasistree", "bestbasistreeall", "delete_subtree_"]


@dataclass(frozen=True)
class LoglpCost:
    """bestbasis/bestbasis_costs.jl:44-46"""
    p: float = 2


@dataclass(frozen=True)
class NormCost:
    """bestbasis/bestbasis_costs.jl:55-57"""
    p: float = 1


@dataclass(frozen=True)
class DifferentialEntropyCost:
    """bestbasis/bestbasis_costs.jl:66"""


@dataclass(frozen=True)
class ShannonEntropyCost:
    """bestbasis/bestbasis_costs.jl:75"""


@dataclass(frozen=True)
class LogEnergyEntropyCost:
    """bestbasis/bestbasis_costs.jl:85"""


@dataclass(frozen=True)
class BB:
    """bestbasis/bestbasis_tree.jl:61-64 : standard (per-signal) best basis"""
    cost: object = field(default_factory=ShannonEntropyCost)
    redundant: bool = False


@dataclass(frozen=True)
class JBB:
    """bestbasis/bestbasis_tree.jl:43-46"""
    cost: object = field(default_factory=LoglpCost)
    redundant: bool = False


@dataclass(frozen=True)
class LSDB:
    """bestbasis/bestbasis_tree.jl:25-28"""
    cost: object = field(default_factory=DifferentialEntropyCost)
    redundant: bool = False


def _geom(X):
    """X (N, K, n) or (N, K, n_cols, m_rows) -> (m, n, K, Nlocal, szK)"""
    assert 3 <= X.dim() <= 4, "AssertionError: 3 <= ndims(X) <= 4"
    N, K = X.shape[0], X.shape[1]
    if X.dim() == 3:
        m, n = 0, X.shape[2]
        sz = n
    else:
        n, m = X.shape[2], X.shape[3]
        sz = n * m
    return m, n, K, N, sz * K


def _ncosts(m, K, redundant):
    if redundant:
        return K
    return (4 ** K - 1) // 3 if m > 0 else (1 << K) - 1


def _bb_kind(method):
    if isinstance(method.cost, ShannonEntropyCost):
        return 0
    if isinstance(method.cost, LogEnergyEntropyCost):
        return 1
    raise TypeError("BB cost must be ShannonEntropyCost or LogEnergyEntropyCost")


def _bb_costs_batch(X, method):
    """X (N, K, ...) -> device costs (N, nnodes) Float64: tree_costs(Xi, ::BB) of every signal (bestbasis_tree.jl:210-256)"""
    m, n, K, N, _ = _geom(X)
    costs = torch.empty((N, _ncosts(m, K, method.redundant)), dtype=torch.float64, device=X.device)
    D.call("bb_costs", X, D.ptr(costs), D.ptr(X), m, n, K, N, int(method.redundant), _bb_kind(method), D.stream(X))
    return costs, m, n


def bestbasistreeall(X, method=None):
    """``bestbasistreeall(X, ::BB)`` BestBasis.jl:253-262: one best-basis tree per signal.  X (N, K, n[, m]) on the device ->
    boolean device tensor (N, ntree) (the reference's BitMatrix (ntree, k) in this package's reversed-axes convention).
    Costs and the bottom-up selection both run on the GPU; nothing depends on other signals, so shards need no exchange."""
    method = BB() if method is None else method
    assert isinstance(method, BB), "bestbasistreeall: method must be BB()"
    X = D.dev(X, "X")
    assert 3 <= X.dim() <= 4, "AssertionError: 3 <= ndims(X) <= 4"
    costs, m, n = _bb_costs_batch(X, method)
    N, nn = costs.shape
    ntree = gettreelength(m, n) if m > 0 else n - 1
    trees = torch.zeros((N, max(ntree, 0)), dtype=torch.uint8, device=X.device)
    with torch.cuda.device(X.device):
        _lib.call("wx_bb_select", D.ptr(trees), D.ptr(costs), nn, m, n, N, X.element_size(), D.stream(X))
    return trees.bool()


def _jbb_kind(method):
    if isinstance(method.cost, LoglpCost):
        return 0
    if isinstance(method.cost, NormCost):
        return 1
    raise TypeError("JBB cost must be LoglpCost or NormCost")


def tree_costs(X, method, group=None):
    """``tree_costs(X, method)`` bestbasis/bestbasis_tree.jl:104-256.  JBB / LSDB: X is the LOCAL shard (N_local, K, ...) of the
    packet table; with an initialised process group the costs are those of the concatenated batch -- the exchange of the
    per-position state runs inside libwx_b200 over its own NCCL communicator (``dist.comm``; ``wx_tree_costs_jbb`` /
    ``wx_tree_costs_lsdb``), not in Python.  BB: X is ONE decomposed signal (K, n[, m]) like in the reference."""
    X = D.dev(X, "X")
    if isinstance(method, BB):
        assert 2 <= X.dim() <= 3, "AssertionError: 2 <= ndims(X) <= 3"
        costs, _, _ = _bb_costs_batch(X.unsqueeze(0), method)
        c = costs[0].cpu().numpy()
        return c.astype(np.float32).astype(np.float64) if X.dtype == torch.float32 else c
    m, n, K, Nloc, szK = _geom(X)
    costs = np.empty(_ncosts(m, K, method.redundant), np.float64)
    cm = dist.comm(X.device, group)
    if isinstance(method, JBB):
        D.call("tree_costs_jbb", X, cm, costs.ctypes.data, D.ptr(X), m, n, K, Nloc, int(method.redundant), _jbb_kind(method),
               C.c_double(float(method.cost.p)), D.stream(X))
    elif isinstance(method, LSDB):
        D.call("tree_costs_lsdb", X, cm, costs.ctypes.data, D.ptr(X), m, n, K, Nloc, int(method.redundant), D.stream(X))
    else:
        raise TypeError(f"unsupported best-basis method {type(method).__name__} on the B200 path (JBB and LSDB)")
    return costs


def bestbasis_treeselection(costs, n, m=None, type="min"):
    """``bestbasis_treeselection(costs, n[, m], type)`` BestBasis.jl:59-110.  ``costs`` is modified in place like the
    reference's.  Returns the tree as a boolean vector."""
    if isinstance(m, str):
        m, type = None, m
    if type not in ("min", "max"):
        raise ValueError(f"ArgumentError: Unsupported type {type}.")
    costs = np.ascontiguousarray(costs, dtype=np.float64)
    if m is None:
        assert len(costs) <= gettreelength(2 * n), "AssertionError: k <= gettreelength(2*n)"
        tree = np.zeros(max(n - 1, 0), np.uint8)
        _lib.call("wx_tree_select", tree.ctypes.data, costs.ctypes.data, len(costs), 0, n, 0 if type == "min" else 1)
    else:
        assert len(costs) <= gettreelength(2 * n, 2 * m), "AssertionError: k <= gettreelength(2*n,2*m)"
        tree = np.zeros(gettreelength(n, m), np.uint8)
        _lib.call("wx_tree_select", tree.ctypes.data, costs.ctypes.data, len(costs), n, m, 0 if type == "min" else 1)
    return tree.astype(bool)


def delete_subtree_(bt, i, tree_type):
    """``delete_subtree!(bt, i, tree_type)`` BestBasis.jl:128-140"""
    assert 1 <= i <= len(bt), "AssertionError: 1 <= i <= length(bt)"
    assert tree_type in ("binary", "quad"), "AssertionError: tree_type in [:binary, :quad]"
    bt[i - 1] = False
    kids = (2 * i, 2 * i + 1) if tree_type == "binary" else (4 * i - 2, 4 * i - 1, 4 * i, 4 * i + 1)
    for c in kids:
        if c <= len(bt) and bt[c - 1]:
            delete_subtree_(bt, c, tree_type)
    return bt


def bestbasistree(X, method=None, group=None):
    """``bestbasistree(X, method)`` BestBasis.jl:185-217 for JBB / LSDB.  X: local shard (N_local, K, ...).  One call into
    the library: reduction kernels, NCCL exchange, node costs and ``bestbasis_treeselection`` (``wx_bestbasistree``)."""
    method = JBB() if method is None else method
    X = D.dev(X, "X")
    if isinstance(method, BB):                       # one signal (K, n[, m])   BestBasis.jl:206-213
        assert 2 <= X.dim() <= 3, "AssertionError: 2 <= ndims(X) <= 3"
        return bestbasistreeall(X.unsqueeze(0), method)[0].cpu().numpy()
    if not isinstance(method, (JBB, LSDB)):
        raise TypeError(f"unsupported best-basis method {type(method).__name__} on the B200 path (JBB, LSDB, BB)")
    m, n, K, Nloc, _ = _geom(X)
    # Julia sz = (rows, cols) = (shape[3], shape[2]); gettreelength is symmetric in them
    tree = np.zeros(gettreelength(m, n) if m > 0 else max(n - 1, 0), np.uint8)
    kind = _jbb_kind(method) if isinstance(method, JBB) else 0
    p = float(method.cost.p) if isinstance(method, JBB) else 0.0
    D.call("bestbasistree", X, dist.comm(X.device, group), 0 if isinstance(method, JBB) else 1, tree.ctypes.data, len(tree), None,
           D.ptr(X), m, n, K, Nloc, int(method.redundant), kind, C.c_double(p), D.stream(X))
    return tree.astype(bool)
