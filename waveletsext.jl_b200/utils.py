"""Host-side tree / index algebra (mirrors Utils.jl and utils/utils_tree.jl of the reference, plus the few
Wavelets.jl helpers they rely on).  Pure integer work; nothing here touches the GPU.

Conventions kept from the reference: node indices are 1-based heap indices (root = 1, binary children
2i / 2i+1, quad children 4i-2 .. 4i+1); a tree is a boolean vector whose entry i-1 says "node i is split".
Ranges are returned as Python ``range`` objects (0-based, half-open): Julia ``a:b`` == ``range(a-1, b)``.
"""
from __future__ import annotations

import numpy as np

__all__ = ["maxtransformlevels", "isdyadic", "ndyadicscales", "nodelength", "getchildindex", "getparentindex",
           "getdepth", "gettreelength", "maketree", "isvalidtree", "getleaf", "main2depthshift",
           "getrowrange", "getcolrange", "coarsestscalingrange", "finestdetailrange"]


def _sig_shape(x):
    """signal shape in Julia order for a single signal given in memory order (reversed dims)"""
    if isinstance(x, int):
        return (x,)
    if isinstance(x, (tuple, list)):
        return tuple(int(v) for v in x)
    return tuple(reversed(tuple(x.shape)))


def maxtransformlevels(x, dim: int | None = None) -> int:
    """Wavelets.jl ``maxtransformlevels``: largest k with 2^k dividing the (smallest) signal extent.
    ``x`` may be an int, a shape tuple (Julia order) or a single-signal array in memory order."""
    if isinstance(x, (int, np.integer)):
        n = int(x)
        if n <= 0:
            return 0
        k = 0
        while n % 2 == 0:
            n //= 2
            k += 1
        return k
    shp = _sig_shape(x)
    if dim is not None:
        assert 1 <= dim <= len(shp), "AssertionError: 1 <= dim <= ndims(x)"
        return maxtransformlevels(shp[dim - 1])
    return min(maxtransformlevels(s) for s in shp)


def isdyadic(n: int) -> bool:
    return n > 0 and (n & (n - 1)) == 0


def ndyadicscales(n: int) -> int:
    return int(n).bit_length() - 1


def nodelength(N: int, L: int) -> int:
    """Utils.jl:242"""
    return N >> L


_CHILD = {"left": None, "right": None, "topleft": -2, "topright": -1, "bottomleft": 0, "bottomright": 1}


def getchildindex(idx: int, child: str) -> int:
    """utils/utils_tree.jl:57-75"""
    assert child in _CHILD, "AssertionError: child in [:left, :right, :topleft, :topright, :bottomleft, :bottomright]"
    if child == "left":
        return idx << 1
    if child == "right":
        return (idx << 1) + 1
    return 4 * idx + _CHILD[child]


def getparentindex(idx: int, tree_type: str) -> int:
    """utils/utils_tree.jl:89-98"""
    assert tree_type in ("binary", "quad"), "AssertionError: tree_type in [:binary, :quad]"
    return idx >> 1 if tree_type == "binary" else (idx + 2) // 4


def getdepth(idx: int, tree_type: str) -> int:
    """utils/utils_tree.jl:252-262 (integer arithmetic instead of floating log)"""
    assert idx > 0, "AssertionError: idx > 0"
    assert tree_type in ("binary", "quad"), "AssertionError: tree_type in [:binary, :quad]"
    if tree_type == "binary":
        return int(idx).bit_length() - 1
    d, last, w = 0, 1, 1
    while idx > last:
        w *= 4
        last += w
        d += 1
    return d


def gettreelength(n: int, m: int | None = None) -> int:
    """utils/utils_tree.jl:285-293"""
    if m is None:
        return (1 << maxtransformlevels(n)) - 1
    L = maxtransformlevels(min(n, m))
    return ((1 << (2 * L)) - 1) // 3


def maketree(*args):
    """``maketree(n, L, s)`` (1-D, Wavelets.jl: n-1 entries) / ``maketree(n, m, L, s)`` (2-D,
    utils/utils_tree.jl:197-221) / ``maketree(x[, s])`` for a single signal array.  s in {"full","dwt"}."""
    s = "full"
    a = list(args)
    if a and isinstance(a[-1], str):
        s = a.pop()
    assert s in ("full", "dwt"), "AssertionError: s in [:full, :dwt]"
    if len(a) == 1 and not isinstance(a[0], (int, np.integer)):
        shp = _sig_shape(a[0])
        a = list(shp) + [maxtransformlevels(shp)]
    if len(a) == 2:
        n, L = int(a[0]), int(a[1])
        assert 0 <= L <= maxtransformlevels(n), "AssertionError: 0 <= L <= maxtransformlevels(n)"
        tree = np.zeros(max(n - 1, 0), dtype=bool)
        if s == "full":
            tree[: (1 << L) - 1] = True
        else:
            for i in range(L):
                tree[(1 << i) - 1] = True
        return tree
    n, m, L = int(a[0]), int(a[1]), int(a[2])
    assert 0 <= L <= maxtransformlevels(min(n, m)), "AssertionError: 0 <= L <= L0"
    tree = np.zeros(gettreelength(n, m), dtype=bool)
    if s == "full":
        tree[: (4 ** L - 1) // 3] = True
    else:
        if L > 0:
            tree[0] = True
        for i in range(0, L - 1):
            tree[((1 << (2 * i + 2)) + 2) // 3 - 1] = True
    return tree


def isvalidtree(x, tree) -> bool:
    """1-D: Wavelets.jl (length n-1, every node's parent exists); 2-D: utils/utils_tree.jl:13-29.
    ``x`` is a single signal (array in memory order), an int length or a Julia-order shape tuple."""
    shp = _sig_shape(x)
    tree = np.asarray(tree, dtype=bool)
    nb = len(tree)
    if len(shp) == 1:
        if nb != shp[0] - 1:
            return False
        ar = 2
    else:
        if gettreelength(shp[0], shp[1]) != nb:
            return False
        ar = 4
    # a split node needs a split parent (children of i: 2i, 2i+1 / 4i-2 .. 4i+1)
    on = np.flatnonzero(tree) + 1
    on = on[on > 1]
    par = on // 2 if ar == 2 else (on + 2) // 4
    return bool(np.all(tree[par - 1]))


def getleaf(tree, tree_type: str):
    """utils/utils_tree.jl:122-157"""
    assert tree_type in ("binary", "quad"), "AssertionError: tree_type in [:binary, :quad]"
    tree = np.asarray(tree, dtype=bool)
    nt = len(tree)
    assert nt >= 1, "AssertionError: tree is empty"
    L0 = getdepth(nt, tree_type)
    expected = (1 << (L0 + 1)) - 1 if tree_type == "binary" else ((1 << (2 * L0 + 2)) - 1) // 3
    assert expected == nt, "AssertionError: tree length does not match a full tree"
    n = 1 << (L0 + 1) if tree_type == "binary" else 1 << (2 * L0 + 2)
    ns = 1 << (L0 + 1)
    assert isvalidtree((ns,) if tree_type == "binary" else (ns, ns), tree), "AssertionError: isvalidtree(x, tree)"
    result = np.zeros(n + nt, dtype=bool)
    result[0] = True
    for i in range(1, nt + 1):
        if not tree[i - 1]:
            continue
        result[i - 1] = False
        if tree_type == "binary":
            result[2 * i - 1] = True
            result[2 * i] = True
        else:
            result[4 * i - 3: 4 * i + 1] = True
    return result


def main2depthshift(sm: int, L: int):
    """Utils.jl:297-305 -> list of L+1 cumulative shifts"""
    assert sm < (1 << L), "AssertionError: sm < 1<<L"
    sd, acc = [0], 0
    for d in range(L):
        acc += ((sm >> d) & 1) << d
        sd.append(acc)
    return sd


def _quadspan(n: int, idx: int, rows: bool):
    L0 = maxtransformlevels(n)
    k = ((1 << (2 * L0 + 2)) - 1) // 3
    assert 0 < idx <= k, "AssertionError: 0 < idx <= k"
    if idx == 1:
        return 0, n
    parent = (idx + 2) // 4
    p0, p1 = _quadspan(n, parent, rows)
    mid = (p0 + p1) // 2
    first = (idx < 4 * parent) if rows else (idx % 2 == 0)
    return (p0, mid) if first else (mid, p1)


def getrowrange(n: int, idx: int) -> range:
    """Utils.jl:465-491 (0-based half-open)"""
    return range(*_quadspan(n, idx, True))


def getcolrange(n: int, idx: int) -> range:
    """Utils.jl:516-542 (0-based half-open)"""
    return range(*_quadspan(n, idx, False))


def coarsestscalingrange(x, tree, redundant: bool = False):
    """Utils.jl:345-370 : follow the left (scaling) chain.  Returns ``range`` or ``(range, node index)``."""
    n = int(x) if isinstance(x, (int, np.integer)) else int(_sig_shape(x)[0])
    tree = np.asarray(tree, dtype=bool)
    L = getdepth(len(tree), "binary")
    assert L + 1 == maxtransformlevels(n), "AssertionError: L+1 == maxtransformlevels(n)"
    i, j = 1, 0
    while i < len(tree) and tree[i - 1]:
        i = getchildindex(i, "left")
        j += 1
    return (range(0, n), i) if redundant else range(0, n >> j)


def finestdetailrange(x, tree, redundant: bool = False):
    """Utils.jl:410-436 : follow the right (detail) chain."""
    n = int(x) if isinstance(x, (int, np.integer)) else int(_sig_shape(x)[0])
    tree = np.asarray(tree, dtype=bool)
    L = getdepth(len(tree), "binary")
    assert L + 1 == maxtransformlevels(n), "AssertionError: L+1 == maxtransformlevels(n)"
    i, j = 1, 0
    while i <= len(tree) and tree[i - 1]:
        i = getchildindex(i, "right")
        j += 1
    if redundant:
        return range(0, n), i
    n0 = nodelength(n, j)
    return range(n - n0, n)
