"""Best-basis values pinned from OUTSIDE the oracle: hand-derived numbers, written out below, that the CPU oracle (and, under
`-m gpu`, the CUDA path) must reproduce.  The reference's own tests only check `isvalidtree` (test/bestbasis.jl:13-41), so
without these the JBB / LSDB rows would rest on the restatement alone.

1. JBB (bestbasis/bestbasis_tree.jl:150-180, coefcost(::LoglpCost) bestbasis_costs.jl:127-130), haar, n = 8, two signals
   x1 = d, x2 = 0.  Per position E[x^2] - E[x]^2 = c^2/2 - c^2/4 = c^2/4, so sigma = |c|/2 where c is d's packet coefficient;
   with haar every coefficient of level l is an INTEGER combination of d divided by sqrt(2)^l (sums and differences of
   pairs), so  cost(node) = 2 * sum_i log(|c_i| / (2 * 2^(l/2)))  with the integer tables written out below.
2. ASH / DifferentialEntropyCost (bestbasis_costs.jl:135-155; AverageShiftedHistograms.jl, compat "0.8, 0.9" in
   /root/reference/Project.toml:22: `ash(x; rng, m, kernel)` bins with k = floor((x - a)/delta + 1.5), spreads every count with
   the triangular kernel 1 - |i-k|/m over |i-k| < m, normalises by 1/(sum(y) * delta); `pdf` interpolates linearly between
   grid points) for the three-point sample [2, 0, 13], every step in exact rational arithmetic.
"""
import math
from fractions import Fraction as Fr

import numpy as np
import pytest

# ---- 1. JBB, by hand ---------------------------------------------------------------------------------------------------------
D = [6, 8, 4, 8, 9, 1, 5, 4]
# level l, node j occupies entries j*(8>>l) ... ; first half of a node's children = pair sums, second half = pair differences
INT_TABLE = [
    [6, 8, 4, 8, 9, 1, 5, 4],                 # level 0: d
    [14, 12, 10, 9, 2, 4, -8, -1],            # level 1: (6+8, 4+8, 9+1, 5+4 | 8-6, 8-4, 1-9, 4-5)              / sqrt(2)
    [26, 19, -2, -1, 6, -9, 2, 7],            # level 2: (14+12, 10+9 | 12-14, 9-10 || 2+4, -8-1 | 4-2, -1+8)   / 2
    [45, -7, -3, 1, -3, -15, 9, 5],           # level 3: (26+19 | 19-26 || -2-1 | -1+2 || 6-9 | -9-6 || 2+7 | 7-2) / (2 sqrt(2))
]
HAND_TREE = [1, 1, 0, 1, 1, 0, 0]
# selection (BestBasis.jl:59-110), costs c1..c15 in heap order, rounded:
#   c1 13.9694 | c2 10.9298 c3 0.0 | c4 6.8599 c5 -4.1589 c6 2.4328 c7 -0.2671 | c8 4.1476 c9 0.4261 c10 -1.2685 c11 -3.4657
#   c12 -1.2685 c13 1.9504 c14 0.9287 c15 -0.2469
#   i=7: c14+c15 =  0.6818 >= c7            -> node 7 is a leaf
#   i=6: c12+c13 =  0.6819 <  c6            -> split, c6 := 0.6819
#   i=5: c10+c11 = -4.7342 <  c5            -> split, c5 := -4.7342
#   i=4: c8+c9   =  4.5737 <  c4            -> split, c4 := 4.5737
#   i=3: c6+c7   =  0.4148 >= c3 = 0        -> node 3 is a leaf, its subtree (node 6) is deleted
#   i=2: c4+c5   = -0.1605 <  c2            -> split
#   i=1: c2+c3   = -0.1605 <  c1            -> split


def hand_jbb_costs():
    costs = []
    for l, row in enumerate(INT_TABLE):
        p = 8 >> l
        for j in range(1 << l):
            costs.append(2.0 * sum(math.log(abs(c) / (2.0 * 2.0 ** (l / 2))) for c in row[j * p:(j + 1) * p]))
    return np.array(costs)


def test_hand_tables_are_consistent():
    """the integer tables above really are the pair sums / differences of the level before (a typo guard)"""
    for l in range(3):
        p = 8 >> l
        for j in range(1 << l):
            v = INT_TABLE[l][j * p:(j + 1) * p]
            kids = INT_TABLE[l + 1][j * p:(j + 1) * p]
            assert kids[:p // 2] == [v[2 * i] + v[2 * i + 1] for i in range(p // 2)]
            assert kids[p // 2:] == [v[2 * i + 1] - v[2 * i] for i in range(p // 2)]
    c = hand_jbb_costs()
    assert abs(c[2]) < 1e-14                          # node 3: |2*4*8*1| = 64 = (2 sqrt 2)^4, cost exactly 0
    assert np.allclose(np.round(c, 4), [13.9694, 10.9298, 0.0, 6.8599, -4.1589, 2.4328, -0.2671, 4.1476, 0.4261, -1.2685, -3.4657, -1.2685,
                                        1.9504, 0.9287, -0.2469], atol=5e-5)


def _haar_table(O, batch):
    g, h = O.makereverseqmfpair(np.array([1.0, 1.0]) / math.sqrt(2.0))
    return np.stack([O.wpd(np.asarray(x, np.float64), h, g, 3) for x in batch])


def test_oracle_jbb_reproduces_the_hand_computed_costs_and_tree(O):
    hand = hand_jbb_costs()
    for batch in ([D, [0] * 8], [D, [0] * 8, D, [0] * 8], [[0] * 8, D]):         # 2 signals, the pair twice (same sigma), swapped
        X = _haar_table(O, batch)
        c = O.tree_costs_jbb(X)
        assert np.abs(c - hand).max() <= 1e-13
        assert O.tree_select(c.copy(), 8).astype(int).tolist() == HAND_TREE
    # the detail sign convention does not matter for sigma; the packet coefficients themselves equal the integer tables / sqrt(2)^l
    X = _haar_table(O, [D])
    for l in range(4):
        assert np.allclose(np.abs(X[0, l]), np.abs(np.array(INT_TABLE[l])) / math.sqrt(2.0) ** l, rtol=0, atol=1e-13)
    # NormCost(1): sum |sigma| = sum |c| / (2 * 2^(l/2)), e.g. root = (6+8+4+8+9+1+5+4)/2 = 22.5
    cn = O.tree_costs_jbb(_haar_table(O, [D, [0] * 8]), False, cost="norm", p=1.0)
    assert abs(cn[0] - 22.5) <= 1e-13 and abs(cn[1] - (14 + 12 + 10 + 9) / (2 * math.sqrt(2))) <= 1e-13


# ---- 2. ASH differential entropy, by hand ----------------------------------------------------------------------------------------
def hand_ash_entropy():
    """x = [2, 0, 13]: mean 5, deviations (-3, -5, 8), corrected variance (9 + 25 + 64)/2 = 49, sigma = 7.
    N = 3: nbins = ceil(90^0.2) = ceil(2.4596) = 3, mbins = ceil(50/3) = 17, grid of (3+1)*17 = 68 points,
    delta = (13 - 0 + 7)/67 = 20/67, a = 0 - 7/2 = -7/2."""
    delta, a, m, npts = Fr(20, 67), Fr(-7, 2), 17, 68
    assert math.ceil(90 ** 0.2) == 3 and math.ceil(50 / 3) == 17
    xs = [Fr(2), Fr(0), Fr(13)]
    bins = [math.floor((x - a) / delta + Fr(3, 2)) for x in xs]
    assert bins == [19, 13, 56]                       # (5.5 * 3.35 + 1.5 = 19.925, 3.5 * 3.35 + 1.5 = 13.225, 16.5 * 3.35 + 1.5 = 56.775)
    y = [Fr(0)] * (npts + 1)                          # 1-based
    for i in bins:
        for k in range(max(1, i - m + 1), min(npts, i + m - 1) + 1):
            y[k] += 1 - Fr(abs(k - i), m)
    # a full triangle sums to 17; bin 13 loses offsets -16..-13 on the left edge (weights 1/17..4/17), bin 56 loses +13..+16
    assert sum(y) == 3 * 17 - 2 * Fr(10, 17) == Fr(847, 17)
    den = 1 / (sum(y) * delta)
    assert den == Fr(1139, 16940)
    pdfs = []
    for x in xs:
        t = (x - a) / delta
        i = math.floor(t) + 1                         # searchsortedlast(rng, x), 1-based
        pdfs.append(den * (y[i] + (y[i + 1] - y[i]) * (t - (i - 1))))
    # x = 2:  between grid points 19 (y = 28/17) and 20 (26/17), fraction 0.425  -> (28 - 0.85)/17 = 27.15/17
    # x = 0:  between 12 (26/17) and 13 (28/17), fraction 0.725               -> 27.45/17
    # x = 13: between 56 (1) and 57 (16/17), fraction 0.275                   -> 16.725/17
    assert pdfs == [den * Fr(2715, 1700), den * Fr(2745, 1700), den * Fr(16725, 17000)]
    return -sum(math.log(float(p)) for p in pdfs) / 3


def test_oracle_ash_reproduces_the_hand_computed_entropy(O):
    hand = hand_ash_entropy()
    assert abs(hand - 2.389191075365154) < 1e-14
    assert abs(O.diffentropy(np.array([2.0, 0.0, 13.0])) - hand) <= 1e-14
    assert abs(O.diffentropy(np.array([2.0, 0.0, 13.0], np.float32)) - hand) <= 1e-5
    # tree_costs(LSDB) sums the entropies of the positions of a node (bestbasis_tree.jl:113-122): two positions, same sample
    X = np.zeros((3, 1, 2))
    X[:, 0, 0] = [2.0, 0.0, 13.0]
    X[:, 0, 1] = [13.0, 2.0, 0.0]                     # the order of the samples does not matter
    assert abs(O.tree_costs_lsdb(X)[0] - 2 * hand) <= 1e-13


def test_oracle_ash_edge_cases(O):
    """a sample exactly on a grid point, and a constant column.  [0, 1, 2]: sigma = 1, delta = 3/67, a = -1/2, the middle
    sample sits at (1 + 1/2)/delta + 3/2 = 35 exactly -> bin 35 (ties go up: floor of an integer); bins 12 / 35 / 57, the outer
    triangles lose 15/17 each at the edges: sum(y) = 51 - 30/17 = 837/17, den = 1139/2511, pdf = den * (101/102, 33/34, 101/102)."""
    den = 1139 / 2511
    hand = -(2 * math.log(den * 101 / 102) + math.log(den * 33 / 34)) / 3
    assert abs(O.diffentropy(np.array([0.0, 1.0, 2.0])) - hand) <= 1e-13
    with pytest.raises(ValueError):                   # reference: rng = a:0.0:a throws ArgumentError (bestbasis_costs.jl:146)
        O.diffentropy(np.array([1.0, 1.0, 1.0]))
    X = np.zeros((3, 1, 2))
    X[:, 0, 0] = [2.0, 0.0, 13.0]
    X[:, 0, 1] = 4.0
    with pytest.raises(ValueError):
        O.tree_costs_lsdb(X)


# ---- the same numbers through the CUDA path ---------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_jbb_reproduces_the_hand_computed_costs_and_tree(wx, cuda):
    import torch
    hand = hand_jbb_costs()
    wt = wx.wavelet("haar")
    for batch in ([D, [0] * 8], [D, [0] * 8, D, [0] * 8]):
        Xw = wx.wpdall(torch.tensor(batch, dtype=torch.float64, device=cuda), wt, 3)
        c = wx.tree_costs(Xw, wx.JBB())
        assert np.abs(c - hand).max() <= 1e-12
        assert wx.bestbasistree(Xw, wx.JBB()).astype(int).tolist() == HAND_TREE
        assert wx.bestbasis_treeselection(c.copy(), 8).astype(int).tolist() == HAND_TREE
    cn = wx.tree_costs(wx.wpdall(torch.tensor([D, [0] * 8], dtype=torch.float64, device=cuda), wt, 3), wx.JBB(cost=wx.NormCost(1)))
    assert abs(cn[0] - 22.5) <= 1e-12


@pytest.mark.gpu
def test_gpu_lsdb_reproduces_the_hand_computed_entropy_and_edge_cases(wx, cuda):
    import torch
    hand = hand_ash_entropy()
    X = np.zeros((3, 1, 2))
    X[:, 0, 0] = [2.0, 0.0, 13.0]
    X[:, 0, 1] = [13.0, 2.0, 0.0]
    c = wx.tree_costs(torch.from_numpy(X).to(cuda), wx.LSDB())
    assert abs(c[0] - 2 * hand) <= 1e-12
    # a sample on a grid point
    den = 1139 / 2511
    tie = -(2 * math.log(den * 101 / 102) + math.log(den * 33 / 34)) / 3
    X[:, 0, 1] = [0.0, 1.0, 2.0]
    c = wx.tree_costs(torch.from_numpy(X).to(cuda), wx.LSDB())
    assert abs(c[0] - (hand + tie)) <= 1e-12
    # a position that is constant over the batch: the reference's range construction throws ArgumentError
    X[:, 0, 1] = 4.0
    with pytest.raises(ValueError):
        wx.tree_costs(torch.from_numpy(X).to(cuda), wx.LSDB())
    with pytest.raises(ValueError):
        wx.bestbasistree(torch.from_numpy(np.repeat(X, 2, axis=2)).to(cuda), wx.LSDB())
