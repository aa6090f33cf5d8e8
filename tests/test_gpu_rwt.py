"""GPU parity tests for the redundant families (stationary SWT.jl, autocorrelation ACWT.jl): C ABI vs the CPU oracle."""
import numpy as np
import pytest
import torch

from test_gpu_dwt import TOL, dev, pair, relerr

pytestmark = pytest.mark.gpu


def acpair(wx, wt):
    P, Q = wx.make_acreverseqmfpair(wt)
    return P, Q


def test_golden_steps_swt_acwt(wx, cuda):
    """test/transforms.jl:54-87 (SWT) and :124-145 (ACWT) known answers through the CUDA kernels"""
    wt = wx.wavelet(wx.WT.db4)
    h, g = pair(wx, wt)
    x = dev(np.array([2, 3, -4, 5.0]), cuda)
    w1, w2 = wx.sdwt_step(x, 0, h, g)
    assert np.round(torch.stack([w1, w2], 1).cpu().numpy(), 3).tolist() == [[3.854, -6.181], [-0.524, 1.803], [0.389, -0.89], [4.767, 5.268]]
    assert np.round(wx.isdwt_step(w1, w2, 0, h, g).cpu().numpy(), 3).tolist() == [2, 3, -4, 5]
    assert np.round(wx.isdwt_step(w1, w2, 0, 0, 0, h, g).cpu().numpy(), 3).tolist() == [2, 3, -4, 5]
    assert np.round(wx.isdwt_step(w1, w2, 0, 0, 1, h, g).cpu().numpy(), 3).tolist() == [2, 3, -4, 5]
    for sv, sw in ((-1, 0), (1, 0), (0, 2)):                 # @test_throws AssertionError (:63-65)
        with pytest.raises(AssertionError):
            wx.isdwt_step(w1, w2, 0, sv, sw, h, g)
    X = np.array([[2, 3], [-4, 5.0]])
    exp = [[[3, 3], [3, 3]], [[-5, 5], [-5, 5]], [[2, 2], [-2, -2]], [[4, -4], [-4, 4]]]
    ws = wx.sdwt_step(dev(X.T, cuda), 0, h, g)
    assert [np.round(w.cpu().numpy().T, 3).tolist() for w in ws] == exp
    assert np.round(wx.isdwt_step(*ws, 0, h, g).cpu().numpy().T, 3).tolist() == X.tolist()
    assert np.round(wx.isdwt_step(*ws, 0, 0, 0, h, g).cpu().numpy().T, 3).tolist() == X.tolist()
    assert np.round(wx.isdwt_step(*ws, 0, 0, 1, h, g).cpu().numpy().T, 3).tolist() == X.tolist()
    with pytest.raises(AssertionError):
        wx.isdwt_step(*ws, 0, 0, 2, h, g)
    # ACWT: g, h = make_acreverseqmfpair(wt); acdwt_step(x, 0, h, g)
    gg, hh = acpair(wx, wt)
    w1, w2 = wx.acdwt_step(x, 0, hh, gg)
    got = np.round(torch.stack([w1, w2], 1).cpu().numpy(), 3) + 0.0
    assert got.tolist() == [[4.243, -1.414], [1.414, 2.828], [0, -5.657], [2.828, 4.243]]
    assert np.round(wx.iacdwt_step(w1, w2).cpu().numpy(), 3).tolist() == [2, 3, -4, 5]
    ws = wx.acdwt_step(dev(X.T, cuda), 0, hh, gg)
    assert [np.round(w.cpu().numpy().T, 3).tolist() for w in ws] == exp
    assert np.round(wx.iacdwt_step(*ws).cpu().numpy().T, 3).tolist() == X.tolist()


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("name", ["haar", "db4", "coif4", "db7", "db9", "db12"])
@pytest.mark.parametrize("n,L", [(8, 3), (64, 4), (2048, 8), (96, 3)])
def test_swt_family_1d(wx, O, cuda, dt, name, n, L):
    wt = wx.wavelet(name)
    h, g = pair(wx, wt)
    N = 3
    x = np.random.default_rng(n + L).standard_normal((N, n)).astype(dt)
    xd = dev(x, cuda)
    refs = {"dwt": O.sdwt, "wpt": O.swpt, "wpd": O.swpd}
    outs = {"dwt": wx.sdwtall(xd, wt, L), "wpt": wx.swptall(xd, wt, L), "wpd": wx.swpdall(xd, wt, L)}
    ref = {k: np.stack([f(x[i], L, h, g) for i in range(N)]) for k, f in refs.items()}
    for k in outs:
        assert outs[k].shape == ref[k].shape
        assert relerr(outs[k].cpu().numpy(), ref[k]) <= TOL[dt], k
    # swpt == swpd leaves (test/transforms.jl:95-96)
    assert torch.equal(outs["wpt"], outs["wpd"][:, (1 << L) - 1:])
    rt = 1e-10 if dt == np.float64 else 3e-4
    sm = 3 % (1 << L)
    # inverses: average and shift based, vs oracle and as round trips
    for nm, fwd, inv, oinv in (("dwt", outs["dwt"], wx.isdwtall, O.isdwt), ("wpt", outs["wpt"], wx.iswptall, O.iswpt)):
        for s in (None, sm):
            if nm == "dwt" and s is not None and not (0 <= np.log2(s) < L):
                continue
            got = inv(fwd, wt, s).cpu().numpy()
            want = np.stack([oinv(ref[nm][i], h, g, s) for i in range(N)])
            assert relerr(got, want) <= TOL[dt] * 20, (nm, s)
            assert relerr(got, x) <= rt, (nm, s)
    tree_full = wx.maketree(n, L, "full")
    tree_dwt = wx.maketree(n, L, "dwt")
    for tree in (tree_full, tree_dwt, min(2, L)):
        for s in (None, sm):
            got = wx.iswpdall(outs["wpd"], wt, tree, s).cpu().numpy()
            t = wx.maketree(n, tree, "full") if isinstance(tree, int) else tree
            want = np.stack([O.iswpd(ref["wpd"][i], t, h, g, s) for i in range(N)])
            assert relerr(got, want) <= TOL[dt] * 20
            assert relerr(got, x) <= rt
    # single-signal API == batch
    assert torch.equal(wx.swpd(xd[0], wt, L), outs["wpd"][0])
    assert torch.equal(wx.sdwt(xd[1], wt, L), outs["dwt"][1])
    assert torch.equal(wx.swpt(xd[2], wt, L), outs["wpt"][2])


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("name", ["haar", "db4", "db7", "db9"])
@pytest.mark.parametrize("n,L", [(8, 3), (64, 4), (2048, 8)])
def test_acwt_family_1d(wx, O, cuda, dt, name, n, L):
    wt = wx.wavelet(name)
    P, Q = acpair(wx, wt)
    N = 3
    x = np.random.default_rng(n + L + 1).standard_normal((N, n)).astype(dt)
    xd = dev(x, cuda)
    outs = {"dwt": wx.acdwtall(xd, wt, L), "wpt": wx.acwptall(xd, wt, L), "wpd": wx.acwpdall(xd, wt, L)}
    refs = {"dwt": O.acdwt, "wpt": O.acwpt, "wpd": O.acwpd}
    ref = {k: np.stack([f(x[i], L, P, Q) for i in range(N)]) for k, f in refs.items()}
    for k in outs:
        assert relerr(outs[k].cpu().numpy(), ref[k]) <= TOL[dt], k
    assert torch.equal(outs["wpt"], outs["wpd"][:, (1 << L) - 1:])          # test/transforms.jl:153-154
    rt = 1e-10 if dt == np.float64 else 3e-4
    assert relerr(wx.iacdwtall(outs["dwt"]).cpu().numpy(), x) <= rt
    assert relerr(wx.iacdwtall(outs["dwt"], wt).cpu().numpy(), np.stack([O.iacdwt(ref["dwt"][i]) for i in range(N)])) <= TOL[dt] * 20
    assert relerr(wx.iacwptall(outs["wpt"]).cpu().numpy(), x) <= rt
    assert relerr(wx.iacwptall(outs["wpt"]).cpu().numpy(), np.stack([O.iacwpt(ref["wpt"][i]) for i in range(N)])) <= TOL[dt] * 20
    # the default L is maxtransformlevels(n) (acwt_all.jl:302), only valid for a full-depth table
    first = () if L == wx.maxtransformlevels(n) else (L,)
    for args in (first, (min(2, L),), (wt, min(2, L)), (wx.maketree(n, L, "dwt"),), (wt, wx.maketree(n, L, "dwt"))):
        got = wx.iacwpdall(outs["wpd"], *args).cpu().numpy()
        assert relerr(got, x) <= rt
    t = wx.maketree(n, L, "dwt")
    want = np.stack([O.iacwpd(ref["wpd"][i], t) for i in range(N)])
    assert relerr(wx.iacwpdall(outs["wpd"], t).cpu().numpy(), want) <= TOL[dt] * 20
    with pytest.raises(AssertionError):                                      # test/transforms.jl:162
        wx.iacwpd_(torch.zeros(n // 2, dtype=xd.dtype, device=cuda), outs["wpd"][0], wt, t)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("ac", [False, True])
def test_redundant_2d(wx, O, cuda, dt, ac):
    wt = wx.wavelet("db4")
    nr, nc, L, N = 8, 8, 3, 2
    x = np.random.default_rng(31 + ac).standard_normal((N, nc, nr)).astype(dt)
    xd = dev(x, cuda)
    rt = 1e-10 if dt == np.float64 else 3e-4
    if ac:
        P, Q = acpair(wx, wt)
        fw = {"dwt": (wx.acdwtall, lambda a: O.acdwt(a, L, P, Q)), "wpt": (wx.acwptall, lambda a: O.acwpt(a, L, P, Q)),
              "wpd": (wx.acwpdall, lambda a: O.acwpd(a, L, P, Q))}
    else:
        h, g = pair(wx, wt)
        fw = {"dwt": (wx.sdwtall, lambda a: O.sdwt(a, L, h, g)), "wpt": (wx.swptall, lambda a: O.swpt(a, L, h, g)),
              "wpd": (wx.swpdall, lambda a: O.swpd(a, L, h, g))}
    outs, ref = {}, {}
    for k, (f, of) in fw.items():
        outs[k] = f(xd, wt, L)
        ref[k] = np.stack([of(x[i]) for i in range(N)])
        assert outs[k].shape == ref[k].shape, k
        assert relerr(outs[k].cpu().numpy(), ref[k]) <= TOL[dt], k
    assert torch.equal(outs["wpt"], outs["wpd"][:, 21:])                      # [:,:,22:85]  test/transforms.jl:111-112
    tree = wx.maketree(nr, nc, L, "dwt")
    if ac:
        assert relerr(wx.iacdwtall(outs["dwt"]).cpu().numpy(), x) <= rt
        assert relerr(wx.iacwptall(outs["wpt"], wt).cpu().numpy(), x) <= rt
        assert relerr(wx.iacwpdall(outs["wpd"]).cpu().numpy(), x) <= rt
        assert relerr(wx.iacwpdall(outs["wpd"], wt, tree).cpu().numpy(), x) <= rt
        want = np.stack([O.iacwpd(ref["wpd"][i], tree) for i in range(N)])
        assert relerr(wx.iacwpdall(outs["wpd"], tree).cpu().numpy(), want) <= TOL[dt] * 20
        # complete trees take the one-pass pairwise sum over the slices (quad tree of depth L = binary tree of depth 2L): against the oracle,
        # on the reference's own tables, for the wpt layout, the full wpd tree and a complete tree of depth 2 inside the depth-3 table
        want = np.stack([O.iacwpt(ref["wpt"][i]) for i in range(N)])
        assert relerr(wx.iacwptall(dev(ref["wpt"], cuda)).cpu().numpy(), want) <= TOL[dt] * 20
        for tr in (wx.maketree(nr, nc, L, "full"), wx.maketree(nr, nc, 2, "full")):
            want = np.stack([O.iacwpd(ref["wpd"][i], tr) for i in range(N)])
            assert relerr(wx.iacwpdall(dev(ref["wpd"], cuda), tr).cpu().numpy(), want) <= TOL[dt] * 20
    else:
        h, g = pair(wx, wt)
        for s in (None, 3):
            assert relerr(wx.isdwtall(outs["dwt"], wt, s).cpu().numpy(), x) <= rt
            assert relerr(wx.iswptall(outs["wpt"], wt, s).cpu().numpy(), x) <= rt
            assert relerr(wx.iswpdall(outs["wpd"], wt, None, s).cpu().numpy(), x) <= rt
            assert relerr(wx.iswpdall(outs["wpd"], wt, 2, s).cpu().numpy(), x) <= rt
            got = wx.iswpdall(outs["wpd"], wt, tree, s).cpu().numpy()
            assert relerr(got, x) <= rt
            want = np.stack([O.iswpd(ref["wpd"][i], tree, h, g, s) for i in range(N)])
            assert relerr(got, want) <= TOL[dt] * 20
            want = np.stack([O.isdwt(ref["dwt"][i], h, g, s) for i in range(N)])
            assert relerr(wx.isdwtall(outs["dwt"], wt, s).cpu().numpy(), want) <= TOL[dt] * 20


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("ac", [False, True])
@pytest.mark.parametrize("name,nr,nc,L", [("haar", 64, 32, 3), ("db4", 96, 64, 2), ("db2", 128, 128, 3), ("db4", 32, 160, 1)])
def test_redundant_2d_tiles(wx, O, cuda, dt, ac, name, nr, nc, L):
    """images larger than a tile: the fused 2-D a-trous step (row halo patches, one column coset per tile, sliding windows, shifted detail
    outputs, periodic wrap at the image border, parent copies for the in-place swpt / sdwt layouts) at depths 0 .. 2"""
    wt = wx.wavelet(name)
    x = np.random.default_rng(nr + nc + L).standard_normal((2, nc, nr)).astype(dt)
    xd = dev(x, cuda)
    if ac:
        P, Q = acpair(wx, wt)
        fw = {"dwt": (wx.acdwtall, lambda a: O.acdwt(a, L, P, Q)), "wpt": (wx.acwptall, lambda a: O.acwpt(a, L, P, Q)),
              "wpd": (wx.acwpdall, lambda a: O.acwpd(a, L, P, Q))}
    else:
        h, g = pair(wx, wt)
        fw = {"dwt": (wx.sdwtall, lambda a: O.sdwt(a, L, h, g)), "wpt": (wx.swptall, lambda a: O.swpt(a, L, h, g)),
              "wpd": (wx.swpdall, lambda a: O.swpd(a, L, h, g))}
    for k, (f, of) in fw.items():
        out = f(xd, wt, L)
        ref = np.stack([of(x[i]) for i in range(2)])
        assert out.shape == ref.shape, k
        assert relerr(out.cpu().numpy(), ref) <= TOL[dt], k
    rt = 1e-10 if dt == np.float64 else 3e-4
    if ac:
        assert relerr(wx.iacwptall(wx.acwptall(xd, wt, L), wt).cpu().numpy(), x) <= rt
    else:
        assert relerr(wx.iswptall(wx.swptall(xd, wt, L), wt).cpu().numpy(), x) <= rt


def test_swt_argument_errors(wx, cuda):
    """test/transforms.jl:340-341,352-353: isdwtall / iswptall with sm = 12 on L = 3 tables -> AssertionError"""
    wt = wx.wavelet("db4")
    x = torch.randn((3, 8), dtype=torch.float64, device=cuda)
    with pytest.raises(AssertionError):
        wx.isdwtall(wx.sdwtall(x, wt), wt, 12)
    with pytest.raises(AssertionError):
        wx.iswptall(wx.swptall(x, wt), wt, 12)
    with pytest.raises(ValueError):                      # ArgumentError: too many levels
        wx.swpdall(x, wt, 4)
    with pytest.raises(ValueError):                      # ArgumentError: L >= 1
        wx.sdwtall(x, wt, 0)
    w = torch.randn((3, 8, 8), dtype=torch.float64, device=cuda)
    with pytest.raises(AssertionError):
        wx.isdwtall(wx.sdwtall(w, wt), wt, 12)
    with pytest.raises(AssertionError):
        wx.iswptall(wx.swptall(w, wt), wt, 12)


def test_swpd_fullsize_properties(wx, cuda):
    """config 3 shape per GPU slab (2048 samples, L = 8): swpd energy identity -- an undecimated orthogonal
    filter-bank step doubles the energy: sum over the 2^d nodes of depth d == 2^d * energy(x) -- and batch slicing."""
    wt = wx.wavelet("db4")
    n, L, N = 2048, 8, 64
    gen = torch.Generator(device=cuda).manual_seed(20243)
    x = torch.randn((N, n), dtype=torch.float64, device=cuda, generator=gen)
    xw = wx.swpdall(x, wt, L)
    e0 = (x * x).sum(dim=1)
    for d in range(L + 1):
        lo, hi = (1 << d) - 1, (1 << (d + 1)) - 1
        ed = (xw[:, lo:hi] ** 2).sum(dim=(1, 2))
        assert torch.allclose(ed, e0 * (1 << d), rtol=1e-10)
    assert torch.equal(wx.swpdall(x[5:7].contiguous(), wt, L), xw[5:7])
    xr = wx.iswpdall(xw, wt, L)
    assert (xr - x).abs().max().item() <= 1e-10 * x.abs().max().item()
    xa = wx.acwpdall(x, wt, L)
    xr = wx.iacwpdall(xa, L)
    assert (xr - x).abs().max().item() <= 1e-10 * x.abs().max().item()


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("name", ["db4", "sym8"])
@pytest.mark.parametrize("n,L", [(4096, 6), (1024, 10), (256, 8), (8192, 3), (16384, 2)])
def test_redundant_fused_stage_plans(wx, O, cuda, dt, name, n, L):
    """shapes that exercise every plan of the fused redundant kernels: several stages (long stacks), a fused prefix followed
    by the per-depth path (levels whose cosets are shorter than a thread's run), one and two CTAs per SM"""
    wt = wx.wavelet(name)
    h, g = pair(wx, wt)
    P, Q = acpair(wx, wt)
    x = np.random.default_rng(n + L).standard_normal((2, n)).astype(dt)
    xd = dev(x, cuda)
    rt = 1e-10 if dt == np.float64 else 5e-4
    for k, fn, ref in (("swpd", wx.swpdall, lambda v: O.swpd(v, L, h, g)), ("sdwt", wx.sdwtall, lambda v: O.sdwt(v, L, h, g)),
                       ("swpt", wx.swptall, lambda v: O.swpt(v, L, h, g)), ("acwpd", wx.acwpdall, lambda v: O.acwpd(v, L, P, Q)),
                       ("acdwt", wx.acdwtall, lambda v: O.acdwt(v, L, P, Q)), ("acwpt", wx.acwptall, lambda v: O.acwpt(v, L, P, Q))):
        out = fn(xd, wt, L)
        want = np.stack([ref(x[i]) for i in range(2)])
        assert out.shape == want.shape, k
        assert relerr(out.cpu().numpy(), want) <= TOL[dt], k
    # inverses (average based = the fused tree / chain reductions; shift based; autocorrelation sums) as round trips
    assert relerr(wx.iswptall(wx.swptall(xd, wt, L), wt).cpu().numpy(), x) <= rt
    assert relerr(wx.iswptall(wx.swptall(xd, wt, L), wt, 3 % (1 << L)).cpu().numpy(), x) <= rt
    assert relerr(wx.isdwtall(wx.sdwtall(xd, wt, L), wt).cpu().numpy(), x) <= rt
    assert relerr(wx.iswpdall(wx.swpdall(xd, wt, L), wt, L).cpu().numpy(), x) <= rt
    assert relerr(wx.iswpdall(wx.swpdall(xd, wt, L), wt, min(2, L)).cpu().numpy(), x) <= rt
    assert relerr(wx.iacwptall(wx.acwptall(xd, wt, L)).cpu().numpy(), x) <= rt
    assert relerr(wx.iacdwtall(wx.acdwtall(xd, wt, L)).cpu().numpy(), x) <= rt
    assert relerr(wx.iacwpdall(wx.acwpdall(xd, wt, L), L).cpu().numpy(), x) <= rt
    # average-based inverses against the oracle (not only as round trips)
    sw = wx.swptall(xd, wt, L)
    want = np.stack([O.iswpt(sw[i].cpu().numpy(), h, g) for i in range(2)])
    assert relerr(wx.iswptall(sw, wt).cpu().numpy(), want) <= TOL[dt] * 20
