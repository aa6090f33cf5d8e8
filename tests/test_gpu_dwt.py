"""GPU parity tests for the decimated family: every result goes through the C ABI of libwx_b200.so and is compared
with the CPU oracle on the same seeded inputs.  Tolerances are the north-star ones: Float64 <= 1e-12, Float32 <= 1e-5,
relative to the max-norm of each signal's reference output."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = {np.float64: 1e-12, np.float32: 1e-5}


def relerr(got, ref, batch_axes=(0,)):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    red = tuple(i for i in range(ref.ndim) if i not in batch_axes)
    den = np.abs(ref).max(axis=red) if red else np.abs(ref)
    den = np.where(den == 0, 1.0, den)
    return float((np.abs(got - ref).max(axis=red) / den).max()) if red else float((np.abs(got - ref) / den).max())


def dev(a, cuda):
    return torch.from_numpy(np.ascontiguousarray(a)).to(cuda)


def pair(wx, wt):
    g, h = wx.makereverseqmfpair(wt, True)
    return h, g


# ------------------------------------------------------------------ golden vectors through the GPU path
def test_golden_steps_gpu(wx, cuda):
    """test/transforms.jl:3-22 known answers, computed by the CUDA kernels"""
    wt = wx.wavelet(wx.WT.db4)
    h, g = pair(wx, wt)
    x = dev(np.array([2, 3, -4, 5.0]), cuda)
    w1, w2 = wx.dwt_step(x, h, g)
    assert np.round(torch.cat([w1, w2]).cpu().numpy(), 3).tolist() == [-0.524, 4.767, 1.803, 5.268]
    assert np.round(wx.idwt_step(w1, w2, h, g).cpu().numpy(), 3).tolist() == [2, 3, -4, 5]
    X = np.array([[2, 3], [-4, 5.0]])
    ws = wx.dwt_step(dev(X.T, cuda), h, g)          # Julia memory order = transpose
    assert [round(float(w.item()), 3) for w in ws] == [3, 5, -2, 4]
    back = wx.idwt_step(*ws, h, g).cpu().numpy().T
    assert np.round(back, 3).tolist() == X.tolist()


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("name", ["haar", "db2", "db4", "coif4", "sym8", "db10", "db3", "db7", "db9", "db12", "db11"])
@pytest.mark.parametrize("n,L", [(4096, 12), (1024, 10), (1024, 4), (64, 6), (8, 3), (16, 0), (96, 5), (24, 3)])
def test_wpdall_parity(wx, O, cuda, dt, name, n, L):
    wt = wx.wavelet(name)
    rng = np.random.default_rng(20240 + n + L)
    N = 5 if n >= 1024 else 37
    x = rng.standard_normal((N, n)).astype(dt)
    y = wx.wpdall(dev(x, cuda), wt, L)
    ref = O.wpdall(x, wt.taps, L)
    assert y.shape == ref.shape
    assert relerr(y.cpu().numpy(), ref) <= TOL[dt]


def test_wpdall_edge_cases(wx, O, cuda):
    wt = wx.wavelet("db4")
    # empty batch
    y = wx.wpdall(torch.empty((0, 64), dtype=torch.float64, device=cuda), wt, 3)
    assert tuple(y.shape) == (0, 4, 64)
    # too many levels -> AssertionError like the reference (dwt_all.jl:266)
    with pytest.raises(AssertionError):
        wx.wpdall(torch.zeros((2, 24), dtype=torch.float64, device=cuda), wt, 4)
    with pytest.raises(AssertionError):
        wx.wpdall(torch.zeros(24, dtype=torch.float64, device=cuda), wt)      # ndims(x) > 1
    # host tensors are rejected (no CPU fallback)
    with pytest.raises(RuntimeError):
        wx.wpdall(torch.zeros((2, 64), dtype=torch.float64), wt, 2)
    # batch of one == single-signal wpd (test/transforms.jl:300-305 "batch == cat of singles")
    x = np.random.default_rng(1).standard_normal((3, 256))
    ya = wx.wpdall(dev(x, cuda), wt, 8).cpu().numpy()
    for k in range(3):
        yk = wx.wpd(dev(x[k], cuda), wt).cpu().numpy()
        assert np.array_equal(ya[k], yk)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_wpd_long_signal(wx, O, cuda, dt):
    """signals longer than the shared-memory ping-pong: leading levels run through the per-level kernel"""
    wt = wx.wavelet("db4")
    n = 1 << 16
    x = np.random.default_rng(7).standard_normal((3, n)).astype(dt)
    y = wx.wpdall(dev(x, cuda), wt, 16)
    ref = O.wpdall(x, wt.taps, 16, nthreads=4)
    assert relerr(y.cpu().numpy(), ref) <= TOL[dt]


def test_wpd_structural_identity(wx, cuda):
    """test/transforms.jl:25-30: wpd(x) columns == wpt(x, L) for each L"""
    wt = wx.wavelet("db4")
    x = dev(np.random.default_rng(3).standard_normal(8), cuda)
    y = wx.wpd(x, wt)
    for L in (1, 2, 3):
        assert torch.allclose(y[L], wx.wpt(x, wt, L), rtol=1e-12, atol=1e-13)
    assert torch.equal(y[0], x)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("n", [8, 256, 4096])
def test_wpt_iwpt_trees(wx, O, cuda, dt, n):
    wt = wx.wavelet("db4")
    h, g = pair(wx, wt)
    rng = np.random.default_rng(n)
    x = rng.standard_normal((6, n)).astype(dt)
    L = wx.maxtransformlevels(n)
    trees = [wx.maketree(n, L, "full"), wx.maketree(n, L, "dwt"), wx.maketree(n, min(2, L), "full")]
    # a random valid tree
    t = np.zeros(n - 1, bool)
    for i in range(1, n):
        if (i == 1 or t[i // 2 - 1]) and rng.random() < 0.7:
            t[i - 1] = True
    trees.append(t)
    for tree in trees:
        yw = wx.wptall(dev(x, cuda), wt, tree)
        ref = np.stack([O.wpt(x[k], tree, h, g) for k in range(x.shape[0])])
        assert relerr(yw.cpu().numpy(), ref) <= TOL[dt]
        xr = wx.iwptall(yw, wt, tree)
        refi = np.stack([O.iwpt(ref[k], tree, h, g) for k in range(x.shape[0])])
        assert relerr(xr.cpu().numpy(), refi) <= TOL[dt] * 10
        assert relerr(xr.cpu().numpy(), x) <= (1e-10 if dt == np.float64 else 2e-4)


def random_tree(n, rng, pr=0.7, maxdepth=None):
    t = np.zeros(n - 1, bool)
    for i in range(1, n):
        if maxdepth is not None and int(np.log2(i)) >= maxdepth:
            break
        if (i == 1 or t[i // 2 - 1]) and rng.random() < pr:
            t[i - 1] = True
    return t


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("name", ["haar", "db2", "sym8", "db10", "db7", "db9", "db12"])
@pytest.mark.parametrize("n", [16, 96, 48, 1024])
def test_wpt_iwpt_trees_filters_and_lengths(wx, O, cuda, dt, name, n):
    """fused all-level tree kernels: every supported filter length, non power-of-two lengths, random trees"""
    wt = wx.wavelet(name)
    h, g = pair(wx, wt)
    rng = np.random.default_rng(n + len(name))
    x = rng.standard_normal((5, n)).astype(dt)
    L = wx.maxtransformlevels(n)
    for tree in (wx.maketree(n, L, "full"), random_tree(n, rng, maxdepth=L), random_tree(n, rng, 0.9, maxdepth=L)):
        yw = wx.wptall(dev(x, cuda), wt, tree)
        ref = np.stack([O.wpt(x[k], tree, h, g) for k in range(x.shape[0])])
        assert relerr(yw.cpu().numpy(), ref) <= TOL[dt]
        xr = wx.iwptall(yw, wt, tree)
        refi = np.stack([O.iwpt(ref[k], tree, h, g) for k in range(x.shape[0])])
        assert relerr(xr.cpu().numpy(), refi) <= TOL[dt] * 10
        assert relerr(xr.cpu().numpy(), x) <= (1e-10 if dt == np.float64 else 3e-4)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_tree_long_signal(wx, O, cuda, dt):
    """a signal too long for shared memory: coarse levels per-level, deep levels fused (forward and inverse)"""
    n = 1 << 16
    wt = wx.wavelet("db4")
    h, g = pair(wx, wt)
    rng = np.random.default_rng(77)
    x = rng.standard_normal((2, n)).astype(dt)
    for tree in (wx.maketree(n, 9, "full"), random_tree(n, rng, 0.8, maxdepth=10)):
        yw = wx.wptall(dev(x, cuda), wt, tree)
        ref = np.stack([O.wpt(x[k], tree, h, g) for k in range(2)])
        assert relerr(yw.cpu().numpy(), ref) <= TOL[dt]
        xr = wx.iwptall(yw, wt, tree)
        assert relerr(xr.cpu().numpy(), np.stack([O.iwpt(ref[k], tree, h, g) for k in range(2)])) <= TOL[dt] * 10
        assert relerr(xr.cpu().numpy(), x) <= (1e-10 if dt == np.float64 else 3e-4)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("n", [64, 1024, 96])
def test_iwpd_random_trees(wx, O, cuda, dt, n):
    """iwpd by tree (DWT.jl:337-351): the getbasiscoef gather is fused into the inverse kernel's staging loads"""
    wt = wx.wavelet("coif4")
    h, g = pair(wx, wt)
    rng = np.random.default_rng(n)
    x = rng.standard_normal((7, n)).astype(dt)
    L = wx.maxtransformlevels(n)
    y = wx.wpdall(dev(x, cuda), wt, L)
    yh = y.cpu().numpy()
    for tree in (random_tree(n, rng, maxdepth=L), random_tree(n, rng, 0.95, maxdepth=L), wx.maketree(n, L, "full"), wx.maketree(n, min(3, L), "full")):
        got = wx.iwpdall(y, wt, tree).cpu().numpy()
        ref = np.stack([O.iwpd(yh[k], tree, h, g) for k in range(x.shape[0])])
        assert relerr(got, ref) <= TOL[dt] * 10
        assert relerr(got, x) <= (1e-10 if dt == np.float64 else 3e-4)
        # == getbasiscoefall + iwptall
        two = wx.iwptall(wx.getbasiscoefall(y, tree), wt, tree).cpu().numpy()
        assert relerr(got, two) <= TOL[dt]


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_getbasiscoef_and_iwpd(wx, O, cuda, dt):
    """test/utils.jl:6-25 index vectors + iwpd round trips (test/transforms.jl:31-33)"""
    Xw = np.arange(1, 13, dtype=dt).reshape(3, 4)          # Julia reshape(1:12,4,3) in memory order
    d = dev(Xw, cuda)
    assert wx.getbasiscoef(d, wx.maketree(4, 2, "dwt")).cpu().numpy().tolist() == [9, 10, 7, 8]
    assert wx.getbasiscoef(d, wx.maketree(4, 2, "full")).cpu().numpy().tolist() == [9, 10, 11, 12]
    Xw2 = np.arange(1, 25, dtype=dt).reshape(2, 3, 4)      # reshape(1:24,4,3,2)
    t1, t2 = wx.maketree(4, 2, "dwt"), wx.maketree(4, 2, "full")
    assert wx.getbasiscoefall(dev(Xw2, cuda), t1).cpu().numpy().T.tolist() == [[9, 21], [10, 22], [7, 19], [8, 20]]
    assert wx.getbasiscoefall(dev(Xw2, cuda), t2).cpu().numpy().T.tolist() == [[9, 21], [10, 22], [11, 23], [12, 24]]
    assert wx.getbasiscoefall(dev(Xw2, cuda), np.stack([t1, t2], 1)).cpu().numpy().T.tolist() == [[9, 21], [10, 22], [7, 23], [8, 24]]
    with pytest.raises(AssertionError):
        wx.getbasiscoef(d, np.array([0, 1, 0], bool))
    with pytest.raises(ValueError):                          # ArgumentError: not enough levels
        wx.getbasiscoef(torch.zeros((2, 4), dtype=torch.float64, device=cuda), wx.maketree(4, 2, "dwt"))
    wt = wx.wavelet("db4")
    h, g = pair(wx, wt)
    x = np.random.default_rng(5).standard_normal((4, 64)).astype(dt)
    y = wx.wpdall(dev(x, cuda), wt)
    for arg in (None, 2, wx.maketree(64, 6, "dwt")):
        xr = wx.iwpdall(y, wt, arg)
        assert relerr(xr.cpu().numpy(), x) <= (1e-10 if dt == np.float64 else 2e-4)
    tree = wx.maketree(64, 6, "dwt")
    ref = np.stack([O.iwpd(y[k].cpu().numpy(), tree, h, g) for k in range(4)])
    assert relerr(wx.iwpdall(y, wt, tree).cpu().numpy(), ref) <= TOL[dt] * 10


# ------------------------------------------------------------------ 2-D
@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("name", ["haar", "db4"])
@pytest.mark.parametrize("m,n,L", [(8, 8, 3), (32, 16, 3), (64, 64, 5), (24, 40, 3)])
def test_wpd2d_parity(wx, O, cuda, dt, name, m, n, L):
    wt = wx.wavelet(name)
    h, g = pair(wx, wt)
    x = np.random.default_rng(m * n).standard_normal((3, n, m)).astype(dt)     # (N, cols, rows)
    y = wx.wpdall(dev(x, cuda), wt, L)
    ref = np.stack([O.wpd(x[k], h, g, L) for k in range(3)])
    assert relerr(y.cpu().numpy(), ref) <= TOL[dt]
    xr = wx.iwpdall(y, wt, L)
    assert relerr(xr.cpu().numpy(), x) <= (1e-10 if dt == np.float64 else 2e-4)
    tree = wx.maketree(m, n, L, "dwt")
    refi = np.stack([O.iwpd(ref[k], tree, h, g) for k in range(3)])
    assert relerr(wx.iwpdall(y, wt, tree).cpu().numpy(), refi) <= TOL[dt] * 10


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("m,n,L", [(384, 96, 5), (160, 224, 4), (1024, 64, 6), (128, 32, 1), (256, 256, 5), (64, 64, 5), (48, 80, 3)])
def test_wpd2d_haar_all_levels_kernel(wx, O, cuda, dt, m, n, L):
    """two-tap filters: one launch keeps an image tile in shared memory for every level (no halo); tile shapes shrink to divide
    the image, shapes that do not fit go through the level-by-level kernels -- all bit-identical to each other"""
    import os
    wt = wx.wavelet("haar")
    h, g = pair(wx, wt)
    x = np.random.default_rng(m * 7 + n + L).standard_normal((3, n, m)).astype(dt)
    y = wx.wpdall(dev(x, cuda), wt, L)
    ref = np.stack([O.wpd(x[k], h, g, L) for k in range(3)])
    assert relerr(y.cpu().numpy(), ref) <= TOL[dt]
    assert torch.equal(y[:, 0], dev(x, cuda))
    assert relerr(wx.iwpdall(y, wt, L).cpu().numpy(), x) <= (1e-10 if dt == np.float64 else 3e-4)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("name", ["haar", "db4", "sym8", "db10"])
@pytest.mark.parametrize("m,n,L", [(256, 128, 4), (96, 160, 3), (512, 512, 5), (64, 1024, 2), (128, 128, 7)])
def test_wpd2d_large_images(wx, O, cuda, dt, name, m, n, L):
    """images larger than shared memory: halo-tile kernel for the coarse levels, whole-node kernel below"""
    wt = wx.wavelet(name)
    h, g = pair(wx, wt)
    x = np.random.default_rng(m + n + L).standard_normal((2, n, m)).astype(dt)
    y = wx.wpdall(dev(x, cuda), wt, L)
    ref = np.stack([O.wpd(x[k], h, g, L) for k in range(2)])
    assert relerr(y.cpu().numpy(), ref) <= TOL[dt]
    assert torch.equal(y[:, 0], dev(x, cuda))                                      # level 0 is a bit copy of x
    # inverses through the fused 2-D kernels (whole-node kernel for the deep levels, halo tiles above): by level, by the
    # dwt tree and by a random quad tree, against the oracle and as round trips; forward by tree == table + leaf gather
    rt = 1e-10 if dt == np.float64 else 3e-4
    assert relerr(wx.iwpdall(y, wt, L).cpu().numpy(), x) <= rt
    rng = np.random.default_rng(L)
    nt = wx.gettreelength(m, n)
    rtree = np.zeros(nt, bool)
    for i in range(1, nt + 1):
        if wx.getdepth(i, "quad") >= L:
            break
        if (i == 1 or rtree[(i + 2) // 4 - 1]) and rng.random() < 0.8:
            rtree[i - 1] = True
    for tree in (wx.maketree(m, n, L, "dwt"), rtree, np.zeros(nt, bool)):
        got = wx.iwpdall(y, wt, tree).cpu().numpy()
        refi = np.stack([O.iwpd(ref[k], tree, h, g) for k in range(2)])
        assert relerr(got, refi) <= TOL[dt] * 10
        assert relerr(got, x) <= rt
        yt = wx.wptall(dev(x, cuda), wt, tree)
        assert relerr(yt.cpu().numpy(), np.stack([O.wpt(x[k], tree, h, g) for k in range(2)])) <= TOL[dt]
        assert relerr(wx.iwptall(yt, wt, tree).cpu().numpy(), x) <= rt


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("name", ["db3", "db5", "coif4"])
def test_wpd2d_more_filter_lengths(wx, O, cuda, dt, name):
    """6, 10 and 12 taps through the halo-tile / whole-node kernels (10 and 12 taps take the 8-pair windows of the whole-node kernel),
    table and by-tree forms, against the oracle"""
    wt = wx.wavelet(name)
    h, g = pair(wx, wt)
    m, n, L = 256, 128, 4
    x = np.random.default_rng(77).standard_normal((2, n, m)).astype(dt)
    y = wx.wpdall(dev(x, cuda), wt, L)
    ref = np.stack([O.wpd(x[k], h, g, L) for k in range(2)])
    assert relerr(y.cpu().numpy(), ref) <= TOL[dt]
    rt = 1e-10 if dt == np.float64 else 3e-4
    for tree in (wx.maketree(m, n, L, "full"), wx.maketree(m, n, L, "dwt")):
        yt = wx.wptall(dev(x, cuda), wt, tree)
        assert relerr(yt.cpu().numpy(), np.stack([O.wpt(x[k], tree, h, g) for k in range(2)])) <= TOL[dt]
        assert relerr(wx.iwptall(yt, wt, tree).cpu().numpy(), x) <= rt


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_wpt2d_trees(wx, O, cuda, dt):
    wt = wx.wavelet("db4")
    h, g = pair(wx, wt)
    m = n = 16
    x = np.random.default_rng(11).standard_normal((2, n, m)).astype(dt)
    rng = np.random.default_rng(12)
    t = np.zeros(wx.gettreelength(m, n), bool)
    for i in range(1, len(t) + 1):
        if (i == 1 or t[(i + 2) // 4 - 1]) and rng.random() < 0.6:
            t[i - 1] = True
    for tree in (wx.maketree(m, n, 4, "full"), wx.maketree(m, n, 3, "dwt"), wx.maketree(m, n, 2, "full"), t):
        yw = wx.wptall(dev(x, cuda), wt, tree)
        ref = np.stack([O.wpt(x[k], tree, h, g) for k in range(2)])
        assert relerr(yw.cpu().numpy(), ref) <= TOL[dt]
        xr = wx.iwptall(yw, wt, tree)
        assert relerr(xr.cpu().numpy(), np.stack([O.iwpt(ref[k], tree, h, g) for k in range(2)])) <= TOL[dt] * 10
        assert relerr(xr.cpu().numpy(), x) <= (1e-10 if dt == np.float64 else 2e-4)
    # single-image API + structural identity wpd[:,:,L] == wpt(x, L)   (test/transforms.jl:36-43)
    xs = dev(x[0], cuda)
    y = wx.wpd(xs, wt)
    for L in (1, 2, 3):
        assert relerr(y[L].cpu().numpy()[None], wx.wpt(xs, wt, L).cpu().numpy()[None]) <= TOL[dt]


# ------------------------------------------------------------------ size-independent properties at full bench size
def test_wpdall_fullsize_properties(wx, cuda):
    """config 2 shape (4096 samples, L = 12) on a slab of 2048 signals: energy conservation per level (orthonormal
    filter bank), linearity, and exact agreement with a small-batch run of the same signals."""
    wt = wx.wavelet("db4")
    n, L, N = 4096, 12, 2048
    gen = torch.Generator(device=cuda).manual_seed(20242)
    x = torch.randn((N, n), dtype=torch.float64, device=cuda, generator=gen)
    y = wx.wpdall(x, wt, L)
    e0 = (x * x).sum(dim=1)
    for lvl in range(L + 1):
        el = (y[:, lvl] * y[:, lvl]).sum(dim=1)
        assert torch.allclose(el, e0, rtol=1e-11)
    x2 = torch.randn((N, n), dtype=torch.float64, device=cuda, generator=gen)
    y2 = wx.wpdall(x2, wt, L)
    ys = wx.wpdall(2.0 * x - 0.5 * x2, wt, L)
    assert (ys - (2.0 * y - 0.5 * y2)).abs().max().item() <= 1e-11 * y.abs().max().item()
    sub = wx.wpdall(x[100:103].contiguous(), wt, L)
    assert torch.equal(sub, y[100:103])
    xr = wx.iwptall(y[:, L].contiguous(), wt, L)
    assert (xr - x).abs().max().item() <= 1e-10 * x.abs().max().item()


def test_wpdall_host_pipeline(wx, O, cuda):
    """host-buffer entry point == device path, including ragged last chunk"""
    wt = wx.wavelet("db4")
    x = np.random.default_rng(9).standard_normal((1000, 512))
    y = wx.host.wpdall_host(x, wt, 9, chunk=300)
    yd = wx.wpdall(dev(x, cuda), wt, 9).cpu().numpy()
    assert np.array_equal(y, yd)
    ref = O.wpdall(x[:8], wt.taps, 9)
    assert relerr(y[:8], ref) <= 1e-12
    # the slots come from the stream-ordered pool and stay cached; trimming returns them and the next call still works
    wx.host.trim_scratch(0)
    assert np.array_equal(wx.host.wpdall_host(x, wt, 9, chunk=300), yd)
    wx.host.trim_scratch(1 << 20)
    # level 0 is filled on the host from x while levels 1..L cross PCIe: L = 0 (nothing to bring back), Float32, and a batch large
    # enough for the two-thread fill (>= 64 MB of x)
    assert np.array_equal(wx.host.wpdall_host(x[:10], wt, 0), x[:10, None, :])
    xf = x.astype(np.float32)
    assert np.array_equal(wx.host.wpdall_host(xf, wt, 4, chunk=333), wx.wpdall(dev(xf, cuda), wt, 4).cpu().numpy())
    xb = np.random.default_rng(10).standard_normal((8200, 1024))
    yb = wx.host.wpdall_host(xb, wt, 2)
    assert np.array_equal(yb[:, 0], xb)
    assert np.array_equal(yb, wx.wpdall(dev(xb, cuda), wt, 2).cpu().numpy())


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_dwtall_idwtall(wx, O, cuda, dt):
    """dwtall / idwtall (dwt/dwt_all.jl:39-110): the :dwt tree of the packet transform; batch == singles and the inverse
    round trip are what the reference tests (test/transforms.jl:278-283); haar level 1 is pinned by test/wavemult.jl:26-30"""
    haar = wx.wavelet("haar")
    y = wx.dwtall(dev(np.array([[1, 2, -3, 4.0]], dtype=dt), cuda), haar, 1)
    assert np.abs(y.cpu().numpy()[0].astype(np.float64) - np.array([2.1213, 0.7071, 0.7071, 4.9497])).max() < 6e-5
    wt = wx.wavelet("db4")
    h, g = pair(wx, wt)
    x = np.random.default_rng(31).standard_normal((5, 256)).astype(dt)
    for L in (None, 3, 0):
        Lx = 8 if L is None else L
        y = wx.dwtall(dev(x, cuda), wt) if L is None else wx.dwtall(dev(x, cuda), wt, L)
        ref = np.stack([O.wpt(x[k], O.maketree1(256, Lx, "dwt"), h, g) for k in range(5)])
        assert relerr(y.cpu().numpy(), ref) <= TOL[dt]
        xr = wx.idwtall(y, wt) if L is None else wx.idwtall(y, wt, L)
        assert relerr(xr.cpu().numpy(), x) <= (1e-10 if dt == np.float64 else 3e-4)
    img = np.random.default_rng(32).standard_normal((3, 32, 16)).astype(dt)
    y2 = wx.dwtall(dev(img, cuda), wt, 2)
    ref2 = np.stack([O.wpt(img[k], O.maketree2(16, 32, 2, "dwt"), h, g) for k in range(3)])
    assert relerr(y2.cpu().numpy(), ref2) <= TOL[dt]
    assert relerr(wx.idwtall(y2, wt, 2).cpu().numpy(), img) <= (1e-10 if dt == np.float64 else 3e-4)
    with pytest.raises(AssertionError):
        wx.dwtall(dev(x, cuda), wt, 9)


# ------------------------------------------------------------------ strided views as OUTPUT arguments
def test_in_place_wrappers_write_through_strided_views(wx, O, cuda):
    """The reference's own wpd! hands sub-block views of y to dwt_step! (DWT.jl:145-156).  An output argument that is a strided view
    must receive the result (it used to be silently replaced by a temporary contiguous copy)."""
    wt = wx.wavelet("db4")
    h, g = pair(wx, wt)
    rng = np.random.default_rng(5)
    n, L = 64, 3
    x = rng.standard_normal(n)
    # Julia column-major y(n, L+1) with level columns as views: here a (n, L+1) tensor whose COLUMNS are strided views
    y = torch.full((n, L + 1), float("nan"), dtype=torch.float64, device=cuda)
    y[:, 0] = dev(x, cuda)
    for d in range(L):                                    # the loop of wpd! (DWT.jl:145-156)
        n0 = n >> d
        for j in range(1 << d):
            v = y[j * n0:(j + 1) * n0, d]
            w1 = y[j * n0:j * n0 + n0 // 2, d + 1]
            w2 = y[j * n0 + n0 // 2:(j + 1) * n0, d + 1]
            assert not w1.is_contiguous() or n0 // 2 == 1
            wx.dwt_step_(w1, w2, v, h, g)
    ref = O.wpd(x, h, g, L)
    assert relerr(y.cpu().numpy().T[None], ref[None]) <= 1e-12
    # inverse into a strided view, batched kernels into strided outputs, redundant and SIWT steps
    back = torch.zeros((n, 2), dtype=torch.float64, device=cuda)
    wx.idwt_step_(back[:, 1], y[:n // 2, 1].contiguous(), y[n // 2:, 1].contiguous(), h, g)
    assert relerr(back[:, 1].cpu().numpy()[None], x[None]) <= 1e-12 and float(back[:, 0].abs().max()) == 0.0
    xb = dev(rng.standard_normal((6, n)), cuda)
    big = torch.zeros((6, L + 1, 2 * n), dtype=torch.float64, device=cuda)
    yv = big[:, :, ::2]
    wx.dwt._wpd_batch(xb, wt, L, yv)
    assert torch.equal(yv, wx.wpdall(xb, wt, L)) and float(big[:, :, 1::2].abs().max()) == 0.0
    holder = torch.zeros((n, 4), dtype=torch.float64, device=cuda)
    w1s, w2s = wx.sdwt_step_(holder[:, 0], holder[:, 2], dev(x, cuda), 1, h, g)
    r1, r2 = wx.sdwt_step(dev(x, cuda), 1, h, g)
    assert torch.equal(holder[:, 0], r1) and torch.equal(holder[:, 2], r2) and w1s.data_ptr() == holder[:, 0].data_ptr()
    # accumulate-into-output kernel: the packed copy must start with the view's contents
    acc = torch.ones((n, 2), dtype=torch.float64, device=cuda)
    ref_acc = torch.ones(n, dtype=torch.float64, device=cuda)
    wx.isdwt_step_(acc[:, 0], r1, r2, 1, 0, 0, h, g, add2out=True)
    wx.isdwt_step_(ref_acc, r1, r2, 1, 0, 0, h, g, add2out=True)
    assert torch.equal(acc[:, 0], ref_acc) and float((acc[:, 1] - 1).abs().max()) == 0.0
    t = torch.randn((8, 2 * n), dtype=torch.float64, device=cuda)
    keep = t.clone()
    wx.threshold_(t[:, ::2], wx.HardTH(), 0.5)
    exp = keep[:, ::2].clone(); exp[exp.abs() <= 0.5] = 0
    assert torch.equal(t[:, ::2], exp) and torch.equal(t[:, 1::2], keep[:, 1::2])
    with pytest.raises(RuntimeError):
        wx.dwt_step_(torch.zeros(2), torch.zeros(2), dev(x[:4], cuda), h, g)       # host outputs are rejected, not copied
