"""GPU parity tests for the LocalDiscriminantBasis object (LDB.jl:89-470) with the time-frequency energy map: fit!/fitdec!,
transform, fit_transform, inverse_transform, change_nfeatures against the oracle's numpy restatement, plus the shape checks of
test/ldb.jl."""
import warnings

import numpy as np
import pytest
import torch

from test_gpu_dwt import dev, pair
from test_gpu_bestbasis import rel

pytestmark = pytest.mark.gpu


def class_signals(n, per, seed, dt=np.float64):
    """three classes of noisy shapes (a stand-in for generateclassdata(ClassData(:tri, ...)), utils_dataset.jl)"""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / n
    base = [np.sin(6 * np.pi * t), np.sign(np.sin(10 * np.pi * t)), 8 * t * (1 - t)]
    X, y = [], []
    for c in range(3):
        for _ in range(per):
            X.append(np.roll(base[c], int(rng.integers(0, 3))) + 0.3 * rng.standard_normal(n)); y.append(c + 1)
    return np.stack(X).astype(dt), y


@pytest.mark.parametrize("dm,kind,p", [("are", "are", 0), ("sre", "sre", 0), ("lp", "lp", 2), ("hd", "hd", 0)])
@pytest.mark.parametrize("dp", ["basis", "fisher"])
@pytest.mark.parametrize("top_k", [None, 5])
def test_ldb_1d_fit_transform(wx, O, cuda, dm, kind, p, dp, top_k):
    n, per = 32, 5
    X, y = class_signals(n, per, 3)
    wt = wx.wavelet("haar")
    dmo = {"are": wx.AsymmetricRelativeEntropy(), "sre": wx.SymmetricRelativeEntropy(), "lp": wx.LpDistance(2), "hd": wx.HellingerDistance()}[dm]
    dpo = wx.BasisDiscriminantMeasure() if dp == "basis" else wx.FishersClassSeparability()
    f = wx.LocalDiscriminantBasis(wt=wt, max_dec_level=4, dm=dmo, dp=dpo, top_k=top_k, n_features=5)
    Xd = dev(X, cuda)
    Xc = wx.fit_transform(f, Xd, y)
    assert tuple(Xc.shape) == (15, 5)                                    # test/ldb.jl: size(Xc) == (5, 15)
    h, g = pair(wx, wt)
    Xw = np.stack([O.wpd(x, h, g, 4) for x in X])
    ref = O.ldb_fitdec(Xw, y, kind, p, top_k, dp)
    assert rel(f.cost, ref["cost"]) <= 1e-10
    assert np.array_equal(f.tree, ref["tree"])
    assert rel(np.nan_to_num(f.DP, posinf=0, neginf=0), np.nan_to_num(ref["DP"], posinf=0, neginf=0)) <= 1e-9
    # the order is a sort of DP: compare the sorted powers (ties may permute) and, where DP is tie-free, the indices
    top = np.sort(np.nan_to_num(ref["DP"].reshape(-1), nan=-np.inf))[::-1][:5]
    got = np.nan_to_num(ref["DP"].reshape(-1), nan=-np.inf)[f.order[:5]]
    assert np.allclose(got, top, rtol=1e-9, atol=0)
    coef = np.stack([O.wpt(x, ref["tree"], h, g) for x in X])
    assert rel(Xc.cpu().numpy(), coef[:, f.order[:5]]) <= 1e-12
    wx.fit_(f, Xd, y)
    Xt = wx.transform(f, Xd)
    assert torch.equal(Xt, Xc)
    Xr = wx.inverse_transform(f, Xc)
    assert tuple(Xr.shape) == (15, n)                                    # size(X̂) == (32, 15)
    z = np.zeros_like(coef); z[:, f.order[:5]] = coef[:, f.order[:5]]
    refr = np.stack([O.iwpt(v, ref["tree"], h, g) for v in z])
    assert rel(Xr.cpu().numpy(), refr) <= 1e-12


def test_ldb_change_nfeatures_and_errors(wx, cuda):
    X, y = class_signals(32, 5, 4)
    f = wx.LocalDiscriminantBasis(wt=wx.wavelet("haar"), max_dec_level=4, top_k=5, n_features=8)
    Xc = wx.fit_transform(f, dev(X, cuda), y)
    x5 = wx.change_nfeatures(f, Xc, 5)
    assert tuple(x5.shape) == (15, 5) and f.n_features == 5 and torch.equal(x5, Xc[:, :5])
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        x10 = wx.change_nfeatures(f, x5, 10)
    assert tuple(x10.shape) == (15, 10) and any("less accurate" in str(i.message) for i in w)
    assert torch.allclose(x10[:, :5], x5, rtol=0, atol=1e-12) and float(x10[:, 5:].abs().max()) <= 1e-12     # "additional features tend to be zeros"
    with pytest.raises(ValueError):
        wx.change_nfeatures(f, x5, 10)                                    # ArgumentError: rows of x do not match f.n_features
    with pytest.raises(TypeError):
        g = wx.LocalDiscriminantBasis(dp=wx.RobustFishersClassSeparability(), max_dec_level=3)
        wx.fit_(g, dev(X, cuda), y)
    with pytest.raises(AssertionError):
        wx.transform(wx.LocalDiscriminantBasis(), dev(X, cuda))          # not fitted
    with pytest.raises(AssertionError):
        wx.fit_(wx.LocalDiscriminantBasis(max_dec_level=9), dev(X, cuda), y)


@pytest.mark.parametrize("dt,tol", [(np.float64, 1e-11), (np.float32, 3e-4)])
def test_ldb_2d_object(wx, O, cuda, dt, tol):
    """test/ldb.jl "2D LDB": three classes of 8 x 8 images with different means"""
    rng = np.random.default_rng(11)
    X = np.concatenate([rng.normal(c, 1, (5, 8, 8)) for c in range(3)]).astype(dt)
    y = [1] * 5 + [2] * 5 + [3] * 5
    for dpo, dpn in ((wx.BasisDiscriminantMeasure(), "basis"), (wx.FishersClassSeparability(), "fisher")):
        f = wx.LocalDiscriminantBasis(max_dec_level=2, top_k=5, n_features=5, dp=dpo)
        Xc = wx.fit_transform(f, dev(X, cuda), y)
        assert tuple(Xc.shape) == (15, 5)
        h, g = pair(wx, f.wt)
        Xw = np.stack([O.wpd(x, h, g, 2) for x in X])
        ref = O.ldb_fitdec(Xw.astype(np.float64), y, "are", 0, 5, dpn)
        assert rel(f.cost, ref["cost"]) <= tol * 10
        if dt == np.float64:
            assert np.array_equal(f.tree, ref["tree"])
            coef = np.stack([O.wpt(x, ref["tree"], h, g) for x in X])
            assert rel(Xc.cpu().numpy(), coef.reshape(15, -1)[:, f.order[:5]]) <= 1e-12
        Xr = wx.inverse_transform(f, Xc)
        assert tuple(Xr.shape) == (15, 8, 8)
        assert torch.equal(wx.transform(f, dev(X, cuda)), Xc)


def test_ldb_large_batch_features(wx, O, cuda):
    """the feature gather / scatter and the class moments on a batch that spans many CTAs"""
    n, per = 256, 700
    X, y = class_signals(n, per, 9)
    f = wx.LocalDiscriminantBasis(wt=wx.wavelet("db2"), max_dec_level=5, dp=wx.FishersClassSeparability(), n_features=40)
    Xd = dev(X, cuda)
    Xc = wx.fit_transform(f, Xd, y)
    h, g = pair(wx, f.wt)
    coef = wx.wptall(Xd, f.wt, f.tree).cpu().numpy()
    DPref, _ = O.discriminant_power_fisher(coef, y)
    assert rel(f.DP, DPref) <= 1e-10
    assert rel(Xc.cpu().numpy(), coef[:, f.order[:40]]) == 0.0
    Xr = wx.inverse_transform(f, Xc).cpu().numpy()
    z = np.zeros_like(coef); z[:, f.order[:40]] = coef[:, f.order[:40]]
    k = 1234
    assert rel(Xr[k], O.iwpt(z[k], f.tree, h, g)) <= 1e-12
