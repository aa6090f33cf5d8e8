"""GPU parity tests for the denoising row (f-3): noisest / surethreshold / relerrorthreshold / threshold / denoiseall against the
oracle's restatement of Denoising.jl, plus the property checks test/denoising.jl makes (the denoised signal is closer to the
clean one)."""
import numpy as np
import pytest
import torch

from test_gpu_dwt import dev
from test_gpu_bestbasis import signals, rel

pytestmark = pytest.mark.gpu

TH = {"hard": 0, "soft": 1, "semisoft": 2, "stein": 3}


def thobj(wx, name):
    return {"hard": wx.HardTH, "soft": wx.SoftTH, "semisoft": wx.SemiSoftTH, "stein": wx.SteinTH}[name]()


def heavisine(n):
    t = np.arange(n) / n
    return 4 * np.sin(4 * np.pi * t) - np.sign(t - 0.3) - np.sign(0.72 - t)


def tables(wx, cuda, x, wt, kind, L=None):
    f = {"dwt": wx.dwtall, "wpt": wx.wptall, "sdwt": wx.sdwtall, "swpd": wx.swpdall, "acdwt": wx.acdwtall, "acwpd": wx.acwpdall}[kind]
    return f(dev(x, cuda), wt) if L is None else f(dev(x, cuda), wt, L)


@pytest.mark.parametrize("dt,tol", [(np.float64, 1e-13), (np.float32, 1e-6)])
def test_noisest_all_shapes(wx, O, cuda, dt, tol):
    wt = wx.wavelet("db2")
    n, N = 256, 37
    x = signals(n, N, 3, dt)
    full = wx.maketree(n, 8, "full"); part = wx.maketree(n, 3, "full"); dtree = wx.maketree(n, 8, "dwt")
    for kind, red, trees in (("dwt", False, [None, dtree, part]), ("wpt", False, [full]), ("sdwt", True, [None]),
                             ("swpd", True, [dtree, part, full]), ("acwpd", True, [part])):
        X = tables(wx, cuda, x, wt, kind)
        Xh = X.cpu().numpy()
        for tr in trees:
            s = wx.noisest(X, red, tr).cpu().numpy()
            ref = np.array([O.noisest(Xh[i], red, tr) for i in range(N)])
            assert np.abs(s - ref).max() <= tol * max(1.0, np.abs(ref).max()), (kind, tr is None)
            assert wx.noisest(X[0], red, tr) == pytest.approx(ref[0], rel=tol, abs=tol)     # single-signal form


def test_noisest_known_answer(wx, cuda):
    # hand-computed: y = [1, 2, 3, 4, 100, 6, 7, 8] as the finest detail half of a 16-vector; median = 5, |y-5| sorted =
    # [1,1,2,2,3,3,4,95] -> median 2.5 -> sigma = 2.5/0.6745
    v = np.zeros(16); v[8:] = [1, 2, 3, 4, 100, 6, 7, 8]
    assert wx.noisest(dev(v, cuda), False) == pytest.approx(2.5 / 0.6745, rel=1e-15)


def test_noisest_long_range_uses_global_scratch(wx, O, cuda):
    n, N = 1 << 17, 3                      # finest detail range 65536 doubles = 512 KB > shared memory
    rng = np.random.default_rng(9)
    x = rng.standard_normal((N, n))
    s = wx.noisest(dev(x, cuda), False).cpu().numpy()
    ref = np.array([O.noisest(x[i], False) for i in range(N)])
    assert np.abs(s - ref).max() <= 1e-14


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_sure_and_relerror_thresholds(wx, O, cuda, dt):
    wt = wx.wavelet("haar")
    n, N = 128, 21
    x = signals(n, N, 11, dt)
    part = wx.maketree(n, 4, "full"); dtree = wx.maketree(n, 7, "dwt")
    tol = 1e-12 if dt == np.float64 else 1e-6
    for kind, red, trees in (("dwt", False, [None]), ("wpt", False, [None]), ("sdwt", True, [None]), ("swpd", True, [dtree, part]), ("acwpd", True, [dtree])):
        X = tables(wx, cuda, x, wt, kind)
        Xh = X.cpu().numpy()
        for tr in trees:
            t = wx.surethreshold(X, red, tr).cpu().numpy()
            ref = np.array([O.surethreshold(Xh[i], red, tr) for i in range(N)])
            assert np.abs(t - ref).max() <= tol * np.abs(ref).max(), ("sure", kind)
            for elbows in (1, 2, 3):
                t = wx.relerrorthreshold(X, red, tr, elbows).cpu().numpy()
                ref = np.array([O.relerrorthreshold(Xh[i], red, tr, elbows) for i in range(N)])
                # the elbow is an argmax over a curve with cancellation noise ~ sqrt(eps): allow a rare neighbouring pick
                bad = np.abs(t - ref) > tol * np.abs(ref).max()
                assert bad.sum() <= (0 if dt == np.float64 else 2), ("relerr", kind, elbows, t[bad], ref[bad])
    assert wx.surethreshold(X[0], True, dtree) == pytest.approx(O.surethreshold(Xh[0], True, dtree), rel=tol)
    s = wx.SureShrink(X[0], True, dtree, wx.SoftTH())
    assert isinstance(s.th, wx.SoftTH) and s.t == pytest.approx(O.surethreshold(Xh[0], True, dtree), rel=tol)


def test_thresholds_cta_sort_path(wx, O, cuda):
    """more than 1024 selected coefficients per signal: the per-CTA shared-memory sort instead of the warp-resident one"""
    wt = wx.wavelet("db2")
    n, N = 512, 6
    x = signals(n, N, 8)
    X = wx.sdwtall(dev(x, cuda), wt)                     # (N, 10, 512): 5120 coefficients per signal
    Xh = X.cpu().numpy()
    t = wx.surethreshold(X, True).cpu().numpy()
    assert np.abs(t - [O.surethreshold(Xh[i], True) for i in range(N)]).max() <= 1e-12
    t = wx.relerrorthreshold(X, True).cpu().numpy()
    assert np.abs(t - [O.relerrorthreshold(Xh[i], True) for i in range(N)]).max() <= 1e-12
    xl = signals(4096, 3, 9)
    s = wx.noisest(dev(xl, cuda), False).cpu().numpy()   # 2048-element range
    assert np.abs(s - [O.noisest(xl[i], False) for i in range(3)]).max() <= 1e-13


def test_sure_large_selection_global_scratch(wx, O, cuda):
    wt = wx.wavelet("haar")
    n, N, L = 4096, 2, 3                   # 8 leaves x 4096 = 32768 doubles = 256 KB > shared memory
    x = signals(n, N, 5)
    X = wx.swpdall(dev(x, cuda), wt, L)
    tree = wx.maketree(n, L, "full")
    assert len(tree) == n - 1
    # tables with fewer levels than the tree length implies: select the leaves by hand through the C ABI mask
    leaves = np.zeros(X.shape[1], np.uint8); leaves[7:15] = 1
    from waveletsext_b200 import denoising as dn
    t = torch.empty(N, dtype=torch.float64, device=cuda)
    dn.D.call("surethreshold", X, dn.D.ptr(t), dn.D.ptr(X), n, X.shape[1], dn._mask_ptr(leaves), N, dn.D.stream(X))
    Xh = X.cpu().numpy()
    ref = [O.surethreshold(Xh[i][7:15], False) for i in range(N)]
    assert np.abs(t.cpu().numpy() - ref).max() <= 1e-12
    dn.D.call("relerrorthreshold", X, dn.D.ptr(t), dn.D.ptr(X), n, X.shape[1], dn._mask_ptr(leaves), 2, N, dn.D.stream(X))
    ref = [O.relerrorthreshold(Xh[i][7:15], False) for i in range(N)]
    assert np.abs(t.cpu().numpy() - ref).max() <= 1e-12


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("name", ["hard", "soft", "semisoft", "stein"])
def test_threshold_types(wx, O, cuda, dt, name):
    rng = np.random.default_rng(4)
    x = rng.standard_normal((7, 333)).astype(dt)
    x[0, :5] = [0.0, -0.0, 0.5, -0.5, 1.0]
    for t in (0.0, 0.5, 1.3):
        y = wx.threshold(dev(x, cuda), thobj(wx, name), t).cpu().numpy()
        ref = O.threshold(x, TH[name], t)
        assert np.array_equal(np.isnan(y), np.isnan(ref))
        assert np.array_equal(np.nan_to_num(y), np.nan_to_num(ref)), (name, t)      # bit-exact: same Float64 expression, one rounding
    xd = dev(x, cuda)
    assert wx.threshold_(xd, thobj(wx, name), 0.7) is xd
    with pytest.raises(Exception):
        wx.threshold(dev(x, cuda), thobj(wx, name), -1.0)


CASES = [("sig", None), ("dwt", None), ("wpt", "full"), ("wpt", "part"), ("sdwt", None), ("swpd", None), ("swpd", "part"),
         ("acdwt", None), ("acwpd", None), ("acwpd", "part")]


@pytest.mark.parametrize("kind,treekind", CASES)
@pytest.mark.parametrize("smooth", ["regular", "undersmooth"])
def test_denoiseall_matches_oracle(wx, O, cuda, kind, treekind, smooth):
    wt = wx.wavelet("db2")
    q = wt.taps
    n, N, L = 128, 9, 7
    x = signals(n, N, 21)
    tree = None if treekind is None else (wx.maketree(n, L, "full") if treekind == "full" else wx.maketree(n, 3, "full"))
    X = dev(x, cuda) if kind == "sig" else tables(wx, cuda, x, wt, kind)
    Xh = X.cpu().numpy()
    kw = {} if tree is None else {"tree": tree}
    for th in ("hard", "soft"):
        dnt = wx.VisuShrink(n, thobj(wx, th))
        y = wx.denoiseall(X, kind, wt, dnt=dnt, smooth=smooth, **kw).cpu().numpy()
        ref = O.denoiseall(Xh, kind, q, th=TH[th], t=dnt.t, smooth=smooth, **kw)
        assert y.shape == (N, n)
        assert rel(y, ref) <= 1e-12, (kind, th)
    # one summary threshold for the batch, thresholds from the relative-error curve
    dnt = wx.RelErrorShrink(wx.HardTH(), 0.3)
    if kind == "acdwt":
        # Denoising.jl:692-700 hands the (default :dwt) tree to the estimator for :acdwt: relerrorthreshold then indexes the (n, L+1)
        # table with a 2n-1 leaf mask and throws; noisest reads column 3
        with pytest.raises(IndexError):
            wx.denoiseall(X, kind, wt, dnt=dnt, estnoise=wx.relerrorthreshold, bestTH=np.mean, smooth=smooth)
        with pytest.raises(IndexError):
            O.denoiseall(Xh, kind, q, th=0, t=0.3, estnoise=O.relerrorthreshold, bestTH=np.mean, smooth=smooth)
        y = wx.denoiseall(X, kind, wt, bestTH=np.median, smooth=smooth).cpu().numpy()
        ref = O.denoiseall(Xh, kind, q, bestTH=np.median, smooth=smooth)
    else:
        y = wx.denoiseall(X, kind, wt, dnt=dnt, estnoise=wx.relerrorthreshold, bestTH=np.mean, smooth=smooth, **kw).cpu().numpy()
        ref = O.denoiseall(Xh, kind, q, th=0, t=0.3, estnoise=O.relerrorthreshold, bestTH=np.mean, smooth=smooth, **kw)
    assert np.abs(y - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())       # column 3 taken as "noise" can zero a whole :acdwt table
    # precomputed noise levels
    sig = np.linspace(0.2, 0.6, N)
    y = wx.denoiseall(X, kind, wt, estnoise=sig, smooth=smooth, **kw).cpu().numpy()
    ref = O.denoiseall(Xh, kind, q, estnoise=sig, smooth=smooth, **kw)
    assert rel(y, ref) <= 1e-12


def test_denoiseall_without_reconstruction_and_single(wx, O, cuda):
    wt = wx.wavelet("haar")
    n, N = 64, 5
    x = signals(n, N, 2)
    X = wx.wptall(dev(x, cuda), wt, wx.maketree(n, 2, "full"))
    tree = wx.maketree(n, 2, "full")
    y = wx.denoiseall(X, "wpt", None, tree=tree, dnt=wx.VisuShrink(wx.SteinTH(), 1.5)).cpu().numpy()
    ref = O.denoiseall(X.cpu().numpy(), "wpt", None, tree=tree, th=3, t=1.5)
    assert rel(y, ref) <= 1e-13
    S = wx.swpdall(dev(x, cuda), wt)
    one = wx.denoise(S[1], "swpd", wt, smooth="undersmooth").cpu().numpy()
    ref = O.denoise(S[1].cpu().numpy(), "swpd", wt.taps, smooth="undersmooth")
    assert rel(one, ref) <= 1e-12
    one = wx.denoise(dev(x[2], cuda), "sig", wt, estnoise=0.4).cpu().numpy()
    assert rel(one, O.denoise(x[2], "sig", wt.taps, estnoise=0.4)) <= 1e-12
    with pytest.raises(AssertionError):
        wx.denoiseall(X, "nope", wt)
    with pytest.raises(AssertionError):
        wx.denoiseall(X, "wpt", wt, smooth="bumpy")
    with pytest.raises(RuntimeError):
        wx.denoiseall(dev(x, cuda), "sig", None)
    # a table with fewer levels than the tree allows: thresholding goes through findall (fine), the BitVector-indexed estimators throw
    S3 = wx.swpdall(dev(x, cuda), wt, 3)
    t3 = wx.maketree(n, 3, "full")
    y = wx.denoiseall(S3, "swpd", wt, tree=t3, estnoise=np.full(N, 0.3)).cpu().numpy()
    h, g = O.makereverseqmfpair(wt.taps)[1], O.makereverseqmfpair(wt.taps)[0]
    S3h = S3.cpu().numpy()
    lv = np.flatnonzero(np.asarray(O.getleaf(t3, "binary")))
    for i in range(N):
        xt = S3h[i].copy(); xt[lv] = O.threshold(xt[lv], 0, 0.3 * np.sqrt(2 * np.log(n)))
        assert rel(y[i], O.iswpd(xt, t3, h, g)) <= 1e-12
    with pytest.raises(IndexError):
        wx.relerrorthreshold(S3, True, t3)


def test_denoise_reduces_the_error(wx, cuda):
    """test/denoising.jl:14-85: every input type brings the noisy heavisine closer to the clean one"""
    wt = wx.wavelet("haar")
    n, N = 256, 5
    x0 = np.stack([np.roll(heavisine(n), 2 * k) for k in range(N)])
    rng = np.random.default_rng(0)
    x = x0 + 0.5 * rng.standard_normal((N, n))
    relnorm = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
    max_err = max(relnorm(x[i], x0[i]) for i in range(N))
    dnt = wx.VisuShrink(2, wx.HardTH())
    runs = [("sig", dict(dnt=dnt, bestTH=np.mean)), ("dwt", dict(dnt=dnt)), ("sdwt", {}), ("swpd", dict(tree=wx.maketree(n, 7, "full"), dnt=wx.RelErrorShrink(wx.HardTH(), 0.3), estnoise=wx.relerrorthreshold)),
            ("acdwt", {}), ("acwpd", dict(tree=wx.maketree(n, 7, "full"), dnt=wx.RelErrorShrink(wx.HardTH(), 0.3), estnoise=wx.relerrorthreshold, bestTH=np.mean))]
    for kind, kw in runs:
        X = dev(x, cuda) if kind == "sig" else tables(wx, cuda, x, wt, kind)
        y = wx.denoiseall(X, kind, wt, **kw).cpu().numpy()
        assert np.mean([relnorm(y[i], x0[i]) for i in range(N)]) <= max_err, kind


def test_denoiseall_full_size_properties(wx, cuda):
    """size-independent checks on a larger batch: zero threshold = perfect reconstruction, a huge threshold with :undersmooth keeps
    only the coarsest scaling coefficient (the signal mean for haar)"""
    wt = wx.wavelet("haar")
    n, N = 1024, 4096
    g = torch.Generator(device=cuda).manual_seed(1)
    x = torch.randn(N, n, dtype=torch.float64, device=cuda, generator=g)
    X = wx.dwtall(x, wt)
    y = wx.denoiseall(X, "dwt", wt, estnoise=np.zeros(N))
    assert float((y - x).abs().max()) <= 1e-12
    y = wx.denoiseall(X, "dwt", wt, estnoise=np.full(N, 1e6), smooth="undersmooth")
    assert float((y - x.mean(dim=1, keepdim=True)).abs().max()) <= 1e-12
    s = wx.noisest(X, False)
    assert s.shape == (N,) and float((s - 1.0).abs().max()) < 0.25          # white noise of unit variance
