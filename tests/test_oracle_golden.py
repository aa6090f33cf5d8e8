"""Pins the CPU oracle (oracle/wx_oracle.c) to every known-answer vector and structural identity the reference's own
tests hold for the hot path.  CPU only."""
import numpy as np
import pytest

DB4 = [0.23037781330889648, 0.7148465705529157, 0.6308807679298589, -0.027983769416859854,
       -0.18703481171909309, 0.030841381835560764, 0.0328830116668852, -0.010597401785069032]   # recalled literal, 16 digits


@pytest.fixture(scope="module")
def F():
    import importlib.util, os, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("wx_filters_for_tests", os.path.join(root, "waveletsext.jl_b200", "filters.py"))
    m = importlib.util.module_from_spec(spec); sys.modules[spec.name] = m; spec.loader.exec_module(m)
    return m


@pytest.fixture(scope="module")
def hg(O):
    g, h = O.makereverseqmfpair(np.array(DB4))
    return h, g


def test_filter_pair_convention(O, F):
    """WT.makereverseqmfpair(wt, true): g = reverse(qmf), h = qmf .* (-1)^(0:F-1); host module == oracle == literal"""
    q = F.wavelet("db4").taps
    assert np.abs(q - np.array(DB4)).max() < 5e-15
    g, h = F.makereverseqmfpair(q, True)
    g2, h2 = O.makereverseqmfpair(q)
    assert np.array_equal(g, g2) and np.array_equal(h, h2)
    assert np.array_equal(g, q[::-1]) and np.array_equal(h, q * np.array([1, -1] * 4))
    P, Q = F.make_acreverseqmfpair(q)
    P2, Q2 = O.make_acreverseqmfpair(q)
    assert np.array_equal(P, P2) and np.array_equal(Q, Q2)
    assert len(P) == 15 and np.array_equal(P, P[::-1]) and abs(P[7] - 1 / np.sqrt(2)) < 1e-16


def test_dwt_golden(O, hg):
    """test/transforms.jl:3-22"""
    h, g = hg
    x = np.array([2, 3, -4, 5.0])
    w1, w2 = O.dwt_step(x, h, g)
    assert np.round(np.concatenate([w1, w2]), 3).tolist() == [-0.524, 4.767, 1.803, 5.268]
    assert np.round(O.idwt_step(w1, w2, h, g), 3).tolist() == x.tolist()
    X = np.array([[2, 3], [-4, 5.0]])
    ws = O.dwt_step2(O.jl(X), h, g)
    assert [round(float(w[0, 0]), 3) for w in ws] == [3, 5, -2, 4]
    assert np.round(O.unjl(O.idwt_step2(*ws, h, g)), 3).tolist() == X.tolist()


def test_haar_golden(O):
    """test/wavemult.jl:26-30: ns_dwt([1,2,-3,4]) level 1 = dwt_step with haar"""
    g, h = O.makereverseqmfpair(np.array([1, 1]) / np.sqrt(2))
    w1, w2 = O.dwt_step(np.array([1, 2, -3, 4.0]), h, g)
    assert np.round(np.concatenate([w1, w2]), 4).tolist() == [2.1213, 0.7071, 0.7071, 4.9497]


def test_swt_golden(O, hg):
    """test/transforms.jl:54-87"""
    h, g = hg
    x = np.array([2, 3, -4, 5.0])
    w1, w2 = O.sdwt_step(x, 0, h, g)
    assert np.round(np.stack([w1, w2], 1), 3).tolist() == [[3.854, -6.181], [-0.524, 1.803], [0.389, -0.89], [4.767, 5.268]]
    for args in ((), (0, 0), (0, 1)):
        assert np.round(O.isdwt_step(w1, w2, 0, h, g, *args), 3).tolist() == x.tolist()
    for bad in ((-1, 0), (1, 0), (0, 2)):
        with pytest.raises(AssertionError):
            O.isdwt_step(w1, w2, 0, h, g, *bad)
    X = np.array([[2, 3], [-4, 5.0]])
    ws = O.sdwt_step2(O.jl(X), 0, h, g)
    assert [np.round(O.unjl(w), 3).tolist() for w in ws] == [[[3, 3], [3, 3]], [[-5, 5], [-5, 5]], [[2, 2], [-2, -2]], [[4, -4], [-4, 4]]]
    for args in ((), (0, 0), (0, 1)):
        assert np.round(O.unjl(O.isdwt_step2(*ws, 0, h, g, *args)), 3).tolist() == X.tolist()
    for bad in ((0, 2), (-1, 1), (1, 0)):
        with pytest.raises(AssertionError):
            O.isdwt_step2(*ws, 0, h, g, *bad)


def test_acwt_golden(O):
    """test/transforms.jl:124-145"""
    P, Q = O.make_acreverseqmfpair(np.array(DB4))
    g, h = P, Q
    x = np.array([2, 3, -4, 5.0])
    w1, w2 = O.acdwt_step(x, 0, h, g)
    assert (np.round(np.stack([w1, w2], 1), 3) + 0.0).tolist() == [[4.243, -1.414], [1.414, 2.828], [0, -5.657], [2.828, 4.243]]
    assert np.round(O.iacdwt_step(w1, w2), 3).tolist() == x.tolist()
    X = np.array([[2, 3], [-4, 5.0]])
    ws = O.acdwt_step2(O.jl(X), 0, h, g)
    assert [np.round(O.unjl(w), 3).tolist() for w in ws] == [[[3, 3], [3, 3]], [[-5, 5], [-5, 5]], [[2, 2], [-2, -2]], [[4, -4], [-4, 4]]]
    assert np.round(O.unjl(O.iacdwt_step2(*ws)), 3).tolist() == X.tolist()


def test_index_algebra_golden(O):
    """test/utils.jl:6-121"""
    assert O.main2depthshift(10, 4).tolist() == [0, 0, 2, 2, 10]
    assert O.main2depthshift(5, 5).tolist() == [0, 1, 1, 5, 5, 5]
    for bad in ((8, 3), (8, 2)):
        with pytest.raises(AssertionError):
            O.main2depthshift(*bad)
    Xw = O.jl(np.arange(1, 13, dtype=np.float64).reshape(3, 4).T)       # reshape(1:12, 4, 3)
    assert O.getbasiscoef(Xw, O.maketree1(4, 2, "dwt")).tolist() == [9, 10, 7, 8]
    assert O.getbasiscoef(Xw, O.maketree1(4, 2, "full")).tolist() == [9, 10, 11, 12]
    assert O.getleaf(O.maketree1(4, 2, "dwt"), "binary").astype(int).tolist() == [0, 0, 1, 1, 1, 0, 0]
    ql = np.zeros(21, int); ql[[2, 3, 4, 5, 6, 7, 8]] = 1
    assert O.getleaf(O.maketree2(4, 4, 2, "dwt"), "quad").astype(int).tolist() == ql.tolist()
    assert O.maketree2(4, 4, 2, "full").tolist() == [True] * 5 and O.maketree2(4, 4, 2, "dwt").astype(int).tolist() == [1, 1, 0, 0, 0]
    assert O.getdepth(5, "binary") == 2 and O.getdepth(5, "quad") == 1
    assert O.treelength2(8, 8) == 21 and O.treelength2(8, 16) == 21
    # getrowrange/getcolrange(8, idx): 2 -> rows 1:4 cols 1:4 ; 3 -> 1:4, 5:8 ; 4 -> 5:8, 1:4 ; 5 -> 5:8, 5:8
    assert [O.quadrange(8, 8, i)[:2] for i in (2, 3, 4, 5)] == [(0, 0), (0, 4), (4, 0), (4, 4)]
    assert O.isvalidtree(O.maketree2(4, 4, 2, "dwt"), 4) and not O.isvalidtree(np.array([0, 1, 0, 0, 0]), 4)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_structural_identities_1d(O, hg, dt):
    """test/transforms.jl:25-33 (wpd columns == wpt per level, iwpd round trips), :93-119, :151-174"""
    h, g = hg
    tol = 1e-12 if dt == np.float64 else 2e-5
    rng = np.random.default_rng(0)
    x = rng.standard_normal(8).astype(dt)
    y = O.wpd(x, h, g, 3)
    assert np.array_equal(y[0], x)
    for L in (1, 2, 3):
        assert np.abs(y[L] - O.wpt(x, O.maketree1(8, L, "full"), h, g)).max() <= tol
    for tree in (O.maketree1(8, 3, "full"), O.maketree1(8, 2, "full"), O.maketree1(8, 3, "dwt")):
        assert np.abs(O.iwpd(y, tree, h, g) - x).max() <= tol * 10
        assert np.abs(O.iwpt(O.wpt(x, tree, h, g), tree, h, g) - x).max() <= tol * 10
    # SWT
    xw = O.swpd(x, 3, h, g)
    assert np.array_equal(O.swpt(x, 3, h, g), xw[7:15])
    sm = 3
    assert np.abs(O.isdwt(O.sdwt(x, 3, h, g), h, g) - x).max() <= tol * 10
    assert np.abs(O.isdwt(O.sdwt(x, 3, h, g), h, g, sm) - x).max() <= tol * 10
    assert np.abs(O.iswpt(O.swpt(x, 3, h, g), h, g) - x).max() <= tol * 10
    assert np.abs(O.iswpt(O.swpt(x, 3, h, g), h, g, sm) - x).max() <= tol * 10
    for tree in (O.maketree1(8, 3, "full"), O.maketree1(8, 2, "full"), O.maketree1(8, 3, "dwt")):
        assert np.abs(O.iswpd(xw, tree, h, g) - x).max() <= tol * 10
        assert np.abs(O.iswpd(xw, tree, h, g, sm) - x).max() <= tol * 10
    # ACWT
    P, Q = O.make_acreverseqmfpair(np.array(DB4))
    aw = O.acwpd(x, 3, P, Q)
    assert np.array_equal(O.acwpt(x, 3, P, Q), aw[7:15])
    assert np.array_equal(O.acwpt(x, 2, P, Q), O.acwpd(x, 2, P, Q)[3:7])
    assert np.abs(O.iacdwt(O.acdwt(x, 3, P, Q)) - x).max() <= tol * 10
    assert np.abs(O.iacwpt(O.acwpt(x, 3, P, Q)) - x).max() <= tol * 10
    for tree in (O.maketree1(8, 3, "full"), O.maketree1(8, 2, "full"), O.maketree1(8, 3, "dwt")):
        assert np.abs(O.iacwpd(aw, tree) - x).max() <= tol * 10


def test_structural_identities_2d(O, hg):
    """test/transforms.jl:36-49, 104-119, 164-174"""
    h, g = hg
    rng = np.random.default_rng(1)
    x = rng.standard_normal((8, 8))
    y = O.wpd(x, h, g, 3)
    for L in (1, 2, 3):
        assert np.abs(y[L] - O.wpt(x, O.maketree2(8, 8, L, "full"), h, g)).max() <= 1e-12
    for tree in (O.maketree2(8, 8, 3, "full"), O.maketree2(8, 8, 2, "full"), O.maketree2(8, 8, 3, "dwt")):
        assert np.abs(O.iwpd(y, tree, h, g) - x).max() <= 1e-11
        assert np.abs(O.iwpt(O.wpt(x, tree, h, g), tree, h, g) - x).max() <= 1e-11
    xw = O.swpd(x, 3, h, g)
    assert np.array_equal(O.swpt(x, 3, h, g), xw[21:85])
    for sm in (None, 3):
        assert np.abs(O.isdwt(O.sdwt(x, 3, h, g), h, g, sm) - x).max() <= 1e-11
        assert np.abs(O.iswpt(O.swpt(x, 3, h, g), h, g, sm) - x).max() <= 1e-11
        for tree in (O.maketree2(8, 8, 3, "full"), O.maketree2(8, 8, 2, "full"), O.maketree2(8, 8, 3, "dwt")):
            assert np.abs(O.iswpd(xw, tree, h, g, sm) - x).max() <= 1e-11
    P, Q = O.make_acreverseqmfpair(np.array(DB4))
    aw = O.acwpd(x, 3, P, Q)
    assert np.array_equal(O.acwpt(x, 3, P, Q), aw[21:85])
    assert np.abs(O.iacdwt(O.acdwt(x, 3, P, Q)) - x).max() <= 1e-11
    assert np.abs(O.iacwpt(O.acwpt(x, 3, P, Q)) - x).max() <= 1e-11
    assert np.abs(O.iacwpd(aw, O.maketree2(8, 8, 3, "dwt")) - x).max() <= 1e-11


def test_batch_equals_singles(O, F):
    """test/transforms.jl:270-364 "Transform All": batch == cat of singles (also with OpenMP threads)"""
    q = F.wavelet("db4").taps
    g, h = O.makereverseqmfpair(q)
    x = np.random.default_rng(2).standard_normal((5, 64))
    for th in (1, 3):
        y = O.wpdall(x, q, 6, th)
        for k in range(5):
            assert np.array_equal(y[k], O.wpd(x[k], h, g, 6))
    w = np.random.default_rng(3).standard_normal((3, 8, 8))
    y2 = O.wpdall(w, q, 3)
    for k in range(3):
        assert np.array_equal(y2[k], O.wpd(w[k], h, g, 3))
    xw = O.rwpdall(0, x, 4, h, g, 2)
    for k in range(5):
        assert np.array_equal(xw[k], O.swpd(x[k], 4, h, g))


def test_bestbasis_oracle_sanity(O):
    """tree selection on hand-made costs (BestBasis.jl:59-83) and JBB cost of a constant-variance table"""
    c = np.array([5., 1, 1, 1, 1, 1, 1])
    assert O.tree_select(c, 4).tolist() == [True, False, False]
    c = np.array([5., 3, 3, 1, 1, 1, 1])
    assert O.tree_select(c, 4).tolist() == [True, True, True]
    assert O.tree_select(c, 4, minmax="max").tolist() == [True, False, False]
    rng = np.random.default_rng(4)
    X = rng.standard_normal((2000, 3, 8))
    costs = O.tree_costs_jbb(X)
    sig = X.std(axis=0)            # population std == sqrt(E[x^2] - E[x]^2)
    assert abs(costs[0] - 2 * np.log(sig[0]).sum()) < 1e-9
    assert abs(costs[1] - 2 * np.log(sig[1, :4]).sum()) < 1e-9
    assert abs(costs[6] - 2 * np.log(sig[2, 6:8]).sum()) < 1e-9
    # differential entropy of a wide sample is close to the Gaussian entropy
    e = O.diffentropy(rng.standard_normal(20000))
    assert abs(e - 0.5 * np.log(2 * np.pi * np.e)) < 0.05
