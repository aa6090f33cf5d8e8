"""GPU parity tests for row f-4: the SIWT steps (siwt/siwt_one_level.jl) and the nonstandard-form transform
(wavemult/transforms.jl), against the oracle's literal restatement and the reference's own test vectors."""
import numpy as np
import pytest
import torch

from test_gpu_dwt import dev, pair

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("wname", ["haar", "db2", "db4", "coif4", "sym8"])
@pytest.mark.parametrize("dt,tol", [(np.float64, 1e-13), (np.float32, 2e-6)])
def test_sidwt_steps_match_oracle(wx, O, cuda, wname, dt, tol):
    wt = wx.wavelet(wname)
    h, g = pair(wx, wt)
    rng = np.random.default_rng(len(wname))
    for n in (2, 4, 8, 50, 256):
        v = rng.standard_normal(n).astype(dt)
        for s in (False, True):
            w1 = torch.empty(n // 2, dtype=torch.from_numpy(v).dtype, device=cuda); w2 = torch.empty_like(w1)
            wx.sidwt_step_(w1, w2, dev(v, cuda), h, g, s)
            r1, r2 = O.sidwt_step(v.astype(np.float64), h, g, s)
            assert np.abs(w1.cpu().numpy() - r1).max() <= tol * max(1, np.abs(r1).max())
            assert np.abs(w2.cpu().numpy() - r2).max() <= tol * max(1, np.abs(r2).max())
            # s = false is dwt_step!, s = true is dwt_step! of the signal delayed by one sample
            d1, d2 = wx.dwt_step(dev(np.roll(v, 1) if s else v, cuda), h, g)
            assert torch.equal(d1, w1) and torch.equal(d2, w2)
            # inverse: parity with the oracle and perfect reconstruction
            back = torch.empty(n, dtype=w1.dtype, device=cuda)
            wx.isidwt_step_(back, w1, w2, h, g, s)
            rb = O.isidwt_step(r1, r2, h, g, s)
            assert np.abs(back.cpu().numpy() - rb).max() <= tol * max(1, np.abs(rb).max())
            assert np.abs(back.cpu().numpy() - v).max() <= 20 * tol * max(1, np.abs(v).max())
    with pytest.raises(AssertionError):
        wx.sidwt_step_(torch.empty(3, device=cuda, dtype=torch.float64), torch.empty(4, device=cuda, dtype=torch.float64),
                       torch.empty(8, device=cuda, dtype=torch.float64), h, g, False)


def test_ns_dwt_reference_vectors(wx, cuda):
    """test/wavemult.jl:24-36"""
    wt = wx.wavelet("haar")
    x = dev(np.array([1, 2, -3, 4.0]), cuda)
    y = np.array([2, 0, 2, -1, 2.1213, 0.7071, 0.7071, 4.9497])
    z = np.array([3.5, 4.5, -1.5, 5.5])
    assert np.array_equal(np.round(wx.ns_dwt(x, wt).cpu().numpy(), 4) + 0.0, y)
    assert np.array_equal(np.round(wx.ns_idwt(dev(y, cuda), wt).cpu().numpy(), 4), z)
    for L in (3, 0):
        with pytest.raises(AssertionError):
            wx.ns_dwt(x, wt, L)
        with pytest.raises(AssertionError):
            wx.ns_idwt(dev(y, cuda), wt, L)
    assert wx.ndyad(1, 4, False) == range(16, 24) and wx.ndyad(1, 4, True) == range(24, 32)      # 17:24, 25:32 one-based
    for bad in ((5, 4, True), (5, 4, False), (0, 4, False)):
        with pytest.raises(AssertionError):
            wx.ndyad(*bad)


@pytest.mark.parametrize("wname", ["haar", "db4", "coif4"])
@pytest.mark.parametrize("dt,tol", [(np.float64, 1e-12), (np.float32, 1e-5)])
def test_ns_dwt_matches_oracle(wx, O, cuda, wname, dt, tol):
    wt = wx.wavelet(wname)
    rng = np.random.default_rng(2)
    for n, N in ((8, 3), (64, 17), (1024, 5)):
        x = rng.standard_normal((N, n)).astype(dt)
        Lmax = int(np.log2(n))
        for L in sorted({1, max(1, Lmax // 2), Lmax}):
            Y = wx.ns_dwt(dev(x, cuda), wt, L)
            assert tuple(Y.shape) == (N, 2 * n)
            ref = np.stack([O.ns_dwt(x[i].astype(np.float64), wt.taps, L) for i in range(N)])
            assert np.abs(Y.cpu().numpy() - ref).max() <= tol * np.abs(ref).max()
            Z = wx.ns_idwt(Y, wt, L)
            refz = np.stack([O.ns_idwt(ref[i], wt.taps, L) for i in range(N)])
            assert np.abs(Z.cpu().numpy() - refz).max() <= tol * np.abs(refz).max()
            # single-vector form
            assert torch.equal(wx.ns_dwt(dev(x[0], cuda), wt, L), Y[0])
            assert torch.equal(wx.ns_idwt(Y[0].contiguous(), wt, L), Z[0])
