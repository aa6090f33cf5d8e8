"""CPU checks of the oracle's restatement of Denoising.jl (row f-3) against hand-computed values and the properties the
reference's own test/denoising.jl asserts.  No GPU."""
import numpy as np
import pytest


def heavisine(n):
    t = np.arange(n) / n
    return 4 * np.sin(4 * np.pi * t) - np.sign(t - 0.3) - np.sign(0.72 - t)


def test_mad_and_noisest_known_answers(O):
    y = np.array([1, 2, 3, 4, 100, 6, 7, 8.0])
    assert O.mad(y) == 2.5                               # median 5; |y-5| sorted = 1 1 2 2 3 3 4 95
    v = np.zeros(16); v[8:] = y
    assert O.noisest(v, False) == pytest.approx(2.5 / 0.6745, rel=1e-15)
    tab = np.zeros((5, 8)); tab[-1] = y                  # sdwt table: the last column is the finest detail
    assert O.noisest(tab, True) == pytest.approx(2.5 / 0.6745, rel=1e-15)
    full = O.maketree1(8, 3, "full")
    assert O.finestdetailrange(8, full) == (7, 8) and O.finestdetailrange(8, full, True) == 14
    assert O.coarsestscalingrange(8, full) == (0, 1) and O.coarsestscalingrange(8, full, True) == 7
    dtree = O.maketree1(8, 3, "dwt")
    assert O.finestdetailrange(8, dtree) == (4, 8) and O.finestdetailrange(8, dtree, True) == 2
    assert O.coarsestscalingrange(8, dtree) == (0, 1)


def test_threshold_known_answers(O):
    x = np.array([0.5, -1.0, 1.5, 2.0, 2.5, -3.0])
    assert np.array_equal(O.threshold(x, O.TH_HARD, 1.0), [0, 0, 1.5, 2.0, 2.5, -3.0])
    assert np.array_equal(O.threshold(x, O.TH_SOFT, 1.0), [0, -0.0, 0.5, 1.0, 1.5, -2.0])
    assert np.array_equal(O.threshold(x, O.TH_STEIN, 1.0), [0, -0.0, 1.5 * (1 - 1 / 2.25), 1.5, 2.5 * (1 - 1 / 6.25), -3.0 * (1 - 1 / 9)])
    # SemiSoftTH tests `x <= 2t` without abs (Wavelets.jl): negative values always enter the branch
    assert np.array_equal(O.threshold(x, O.TH_SEMISOFT, 1.0), [0, -0.0, 1.0, 2.0, 2.5, -3.0])
    with pytest.raises(AssertionError):
        O.threshold(x, O.TH_HARD, -0.1)
    assert O.threshold(x.astype(np.float32), O.TH_SOFT, 0.3).dtype == np.float32


def test_sure_and_relerror_small(O):
    # a = [1,4,9], b = [1,5,14], s = b + [2,1,0]*a = [3,9,14], risk = (3 - [2,4,6] + s)/3 = [4/3, 8/3, 11/3]
    assert O.surethreshold(np.array([3.0, -1.0, 2.0]), False) == 1.0
    c = np.array([4.0, 0.1, -0.2, 3.0, 0.05, 0.3, -0.15, 0.02])
    r = O.orth2relerror(c)
    assert r[0] == pytest.approx(np.sqrt((c ** 2).sum() - 16) / np.sqrt((c ** 2).sum()), rel=1e-14)
    assert np.all(np.diff(r) <= 1e-12)
    t1 = O.relerrorthreshold(c, False, None, 1)
    assert t1 in np.abs(c) or t1 == 0.0                  # the elbow sits on one of the magnitudes
    assert 0.3 <= t1 <= 3.0                              # between the noise floor and the two large coefficients
    tab = np.arange(24, dtype=float).reshape(3, 8)
    tree = O.maketree1(8, 1, "full")                     # leaves = nodes 2, 3 of a 15-node table
    with pytest.raises(IndexError):
        O.surethreshold(tab, True, tree)                 # 3 columns against a 15-entry leaf mask (BoundsError in the reference)


@pytest.mark.parametrize("kind", ["sig", "dwt", "wpt", "sdwt", "swpd", "acdwt", "acwpd"])
def test_denoise_brings_heavisine_closer(O, kind):
    """test/denoising.jl:14-46"""
    n = 256
    x0 = heavisine(n)
    x = x0 + 0.5 * np.random.default_rng(0).standard_normal(n)
    q = np.array([1.0, 1.0]) / np.sqrt(2)
    g, h = O.makereverseqmfpair(q); P, Q = O.make_acreverseqmfpair(q)
    relnorm = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
    err = relnorm(x, x0)
    kw = {}
    if kind == "sig": X = x
    elif kind == "dwt": X = O.wpt(x, O.maketree1(n, 8, "dwt"), h, g)
    elif kind == "wpt": X = O.wpt(x, O.maketree1(n, 8, "full"), h, g); kw["tree"] = O.maketree1(n, 8, "full")
    elif kind == "sdwt": X = O.sdwt(x, 8, h, g)
    elif kind == "swpd": X = O.swpd(x, 8, h, g)
    elif kind == "acdwt": X = O.acdwt(x, 8, P, Q)
    else: X = O.acwpd(x, 8, P, Q)
    y = O.denoise(X, kind, q, th=O.TH_HARD, t=O.visushrink_t(2) if kind != "swpd" else None, smooth="undersmooth", **kw)
    assert y.shape == (n,)
    assert relnorm(y, x0) <= 2 * err
    if kind in ("sig", "sdwt", "swpd", "acdwt", "acwpd"):
        assert relnorm(O.denoise(X, kind, q, **kw), x0) <= err
    # a zero noise level reconstructs the input
    assert relnorm(O.denoise(X, kind, q, estnoise=0.0, **kw), x) <= 1e-12


def test_denoiseall_matches_single_and_besttTH(O):
    n, N = 64, 4
    rng = np.random.default_rng(3)
    x = np.stack([np.roll(heavisine(n), 3 * k) for k in range(N)]) + 0.4 * rng.standard_normal((N, n))
    q = np.array([1.0, 1.0]) / np.sqrt(2)
    Y = O.denoiseall(x, "sig", q)
    for i in range(N):
        assert np.array_equal(Y[i], O.denoise(x[i], "sig", q))
    g, h = O.makereverseqmfpair(q)
    X = np.stack([O.wpt(x[i], O.maketree1(n, 6, "dwt"), h, g) for i in range(N)])
    s = np.mean([O.noisest(X[i], False) for i in range(N)])
    Y = O.denoiseall(X, "dwt", q, bestTH=np.mean)
    for i in range(N):
        assert np.allclose(Y[i], O.denoise(X[i], "dwt", q, estnoise=s), rtol=0, atol=1e-14)
