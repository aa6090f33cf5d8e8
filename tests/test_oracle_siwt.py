"""CPU checks of the oracle's restatement of row f-4 (SIWT steps, nonstandard-form transform) against the reference's own test
vectors (test/wavemult.jl:24-36) and the structural identities of the shifted steps.  No GPU."""
import numpy as np
import pytest


def test_ns_dwt_reference_vectors(O):
    q = np.array([1.0, 1.0]) / np.sqrt(2)
    x = np.array([1, 2, -3, 4.0])
    y = np.array([2, 0, 2, -1, 2.1213, 0.7071, 0.7071, 4.9497])
    z = np.array([3.5, 4.5, -1.5, 5.5])
    assert np.array_equal(np.round(O.ns_dwt(x, q), 4) + 0.0, y)
    assert np.array_equal(np.round(O.ns_idwt(y, q), 4), z)
    for L in (3, 0):
        with pytest.raises(AssertionError):
            O.ns_dwt(x, q, L)
        with pytest.raises(AssertionError):
            O.ns_idwt(y, q, L)
    assert O.ndyad(1, 4, False) == (16, 24) and O.ndyad(1, 4, True) == (24, 32)        # 17:24 and 25:32 one-based


def test_sidwt_step_is_the_delayed_dwt_step(O):
    rng = np.random.default_rng(0)
    import waveletsext_b200 as wx
    for name in ("haar", "db3", "coif4"):
        g, h = O.makereverseqmfpair(wx.wavelet(name).taps)
        for n in (2, 8, 30):
            v = rng.standard_normal(n)
            a1, a2 = O.sidwt_step(v, h, g, False)
            d1, d2 = O.dwt_step(v, h, g)
            assert np.allclose(a1, d1, rtol=0, atol=1e-14) and np.allclose(a2, d2, rtol=0, atol=1e-14)
            s1, s2 = O.sidwt_step(v, h, g, True)
            d1, d2 = O.dwt_step(np.roll(v, 1), h, g)
            assert np.allclose(s1, d1, rtol=0, atol=1e-14) and np.allclose(s2, d2, rtol=0, atol=1e-14)
            for s, (w1, w2) in ((False, (a1, a2)), (True, (s1, s2))):
                assert np.allclose(O.isidwt_step(w1, w2, h, g, s), v, rtol=0, atol=1e-12)
