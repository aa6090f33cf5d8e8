"""Host-side logic of the product package (no GPU needed): index algebra mirrored from Utils.jl / utils_tree.jl with the
reference's own test vectors (test/utils.jl), filter construction, and the rule that the product never touches oracle/."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_index_algebra_vectors(wx):
    """test/utils.jl:36-39, 58-68, 71-119"""
    assert wx.main2depthshift(10, 4) == [0, 0, 2, 2, 10]
    assert wx.main2depthshift(5, 5) == [0, 1, 1, 5, 5, 5]
    for bad in ((8, 3), (8, 2)):
        with pytest.raises(AssertionError):
            wx.main2depthshift(*bad)
    assert wx.nodelength(8, 2) == 2
    assert wx.getrowrange(8, 2) == range(0, 4) and wx.getrowrange(8, 3) == range(0, 4)
    assert wx.getrowrange(8, 4) == range(4, 8) and wx.getrowrange(8, 5) == range(4, 8)
    assert wx.getcolrange(8, 2) == range(0, 4) and wx.getcolrange(8, 3) == range(4, 8)
    assert wx.getcolrange(8, 4) == range(0, 4) and wx.getcolrange(8, 5) == range(4, 8)
    for f in (wx.getrowrange, wx.getcolrange):
        with pytest.raises(AssertionError):
            f(8, 86)
    assert [wx.getchildindex(3, c) for c in ("left", "right", "topleft", "topright", "bottomleft", "bottomright")] == [6, 7, 10, 11, 12, 13]
    with pytest.raises(AssertionError):
        wx.getchildindex(3, "fail")
    assert wx.getparentindex(4, "binary") == 2 and wx.getparentindex(5, "binary") == 2
    assert [wx.getparentindex(i, "quad") for i in (10, 11, 12, 13)] == [3, 3, 3, 3]
    with pytest.raises(AssertionError):
        wx.getparentindex(15, "fail")
    assert wx.getdepth(5, "binary") == 2 and wx.getdepth(5, "quad") == 1
    with pytest.raises(AssertionError):
        wx.getdepth(5, "fail")
    with pytest.raises(AssertionError):
        wx.getdepth(0, "binary")
    assert wx.gettreelength(8) == 7 and wx.gettreelength(8, 8) == 21 and wx.gettreelength(8, 16) == 21
    # getdepth(:quad) agrees with floor(log(4, 3i-2)) wherever the float formula is safe
    for i in range(1, 5000):
        assert wx.getdepth(i, "quad") == int(np.floor(np.log(3 * i - 2) / np.log(4) + 1e-12))


def test_trees(wx):
    """test/utils.jl:71-74, 94-112"""
    z44 = np.zeros((4, 4))
    assert wx.isvalidtree(z44, wx.maketree(4, 4, 2, "full")) and wx.isvalidtree(z44, wx.maketree(4, 4, 2, "dwt"))
    assert not wx.isvalidtree(z44, np.array([0, 1, 0, 0, 0], bool)) and not wx.isvalidtree(z44, np.ones(4, bool))
    assert wx.getleaf(wx.maketree(4, 2, "dwt"), "binary").astype(int).tolist() == [0, 0, 1, 1, 1, 0, 0]
    ql = np.zeros(21, int); ql[[2, 3, 4, 5, 6, 7, 8]] = 1
    assert wx.getleaf(wx.maketree(4, 4, 2, "dwt"), "quad").astype(int).tolist() == ql.tolist()
    for tree, kind in ((wx.maketree(4, 2, "dwt"), "fail"), (wx.maketree(4, 4, 2, "dwt"), "binary"), (np.array([0, 1, 0], bool), "binary"),
                       (np.array([0, 1], bool), "binary"), (wx.maketree(4, 2, "dwt"), "quad"), (np.array([0, 1, 1, 1, 1], bool), "quad"),
                       (np.array([0, 1], bool), "quad")):
        with pytest.raises(AssertionError):
            wx.getleaf(tree, kind)
    assert wx.maketree(z44).tolist() == [True] * 5
    assert wx.maketree(z44, "dwt").astype(int).tolist() == [1, 1, 0, 0, 0]
    assert wx.maketree(4, 4, 2).tolist() == [True] * 5 and wx.maketree(4, 4, 2, "dwt").astype(int).tolist() == [1, 1, 0, 0, 0]
    with pytest.raises(AssertionError):
        wx.maketree(4, 4, 3, "dwt")
    with pytest.raises(AssertionError):
        wx.maketree(4, 4, 2, "fail")
    assert wx.maxtransformlevels((4, 2), 1) == 2 and wx.maxtransformlevels((4, 2), 2) == 1
    with pytest.raises(AssertionError):
        wx.maxtransformlevels((4, 2), 3)
    tree = wx.maketree(4, 1, "dwt")
    assert wx.coarsestscalingrange(4, tree) == range(0, 2) and wx.coarsestscalingrange(4, tree, True) == (range(0, 4), 2)
    assert wx.finestdetailrange(4, tree) == range(2, 4) and wx.finestdetailrange(4, tree, True) == (range(0, 4), 3)
    for f in (wx.coarsestscalingrange, wx.finestdetailrange):
        with pytest.raises(AssertionError):
            f(5, tree, True)
    bt = np.ones(7, bool)
    assert wx.delete_subtree_(bt, 2, "binary").astype(int).tolist() == [1, 0, 1, 0, 0, 1, 1]


def test_tree_selection_host(wx, O):
    """wx_tree_select (host part of libwx_b200, no GPU needed) == oracle restatement of BestBasis.jl:59-110"""
    rng = np.random.default_rng(0)
    for n in (4, 16, 64):
        for K in range(2, int(np.log2(n)) + 2):
            c = rng.standard_normal((1 << K) - 1)
            for mm in ("min", "max"):
                assert np.array_equal(wx.bestbasis_treeselection(c.copy(), n, mm), O.tree_select(c, n, minmax=mm))
    for n, K in ((8, 3), (16, 4), (16, 2)):
        c = rng.standard_normal((4 ** K - 1) // 3)
        t = wx.bestbasis_treeselection(c.copy(), n, n)
        assert np.array_equal(t, O.tree_select(c, n, n))
        assert wx.isvalidtree((n, n), t)
    with pytest.raises(ValueError):
        wx.bestbasis_treeselection(rng.standard_normal(15), 8, "fail")      # test/bestbasis.jl:44
    with pytest.raises(AssertionError):
        wx.bestbasis_treeselection(rng.standard_normal(32), 8)              # test/bestbasis.jl:43


def test_filters(wx):
    for name, F in (("haar", 2), ("db2", 4), ("db4", 8), ("db6", 12), ("db8", 16), ("db10", 20), ("sym4", 8), ("sym8", 16), ("coif4", 12)):
        wt = wx.wavelet(name)
        assert len(wt) == F
        assert wx.filters.check_orthonormal(wt.taps) < 5e-15, name
    q = wx.wavelet(wx.WT.db4).taps
    assert abs(q[0] - 0.2303778133088965) < 1e-15 and abs(q[-1] + 0.010597401785069032) < 1e-15
    # db4: four vanishing moments of the detail filter
    _, h = wx.makereverseqmfpair(wx.wavelet("db4"))
    k = np.arange(8.0)
    for p in range(4):
        assert abs(np.dot(h, k ** p)) < 1e-10
    a = wx.autocorr(wx.wavelet("db4"))
    assert np.abs(a[1::2]).max() < 1e-15            # even lags vanish for an orthonormal filter
    P, Q = wx.make_acqmfpair(wx.wavelet("db4"))
    assert np.allclose(P + Q, np.eye(15)[7] * np.sqrt(2))
    with pytest.raises(ValueError):
        wx.wavelet("nope")


def test_product_never_touches_the_oracle():
    """the oracle is test infrastructure: nothing under the product package may import, link or execute it"""
    pkg = os.path.join(ROOT, "waveletsext.jl_b200")
    pat = re.compile(r"oracle|wx_oracle|libwx_oracle", re.I)
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert not pat.search(txt), f"{os.path.join(dp, f)} mentions the oracle"
    assert not pat.search(open(os.path.join(ROOT, "include", "wx_b200.h")).read())
    assert not pat.search(open(os.path.join(ROOT, "waveletsext_b200.py")).read())


def test_no_cpu_fallback(wx):
    """device arrays only: host tensors are rejected instead of being transformed on the CPU"""
    import torch
    wt = wx.wavelet("haar")
    with pytest.raises(RuntimeError):
        wx.wpdall(torch.zeros((2, 8), dtype=torch.float64), wt)
    with pytest.raises(RuntimeError):
        wx.swpdall(torch.zeros((2, 8), dtype=torch.float64), wt, 2)
    with pytest.raises(TypeError):
        wx.wpdall(np.zeros((2, 8)), wt)


def test_isvalidtree_matches_the_definition(wx):
    """vectorised isvalidtree == the definition (no split node below an unsplit one), binary and quad, valid and invalid"""
    import numpy as np
    rng = np.random.default_rng(0)

    def slow(tree, ar):
        nb = len(tree)
        for i in range(1, nb + 1):
            if tree[i - 1]:
                continue
            for c in range(ar):
                ch = 2 * i + c if ar == 2 else 4 * i - 2 + c
                if ch <= nb and tree[ch - 1]:
                    return False
        return True

    for _ in range(200):
        t = rng.random(63) < 0.6
        assert wx.isvalidtree(64, t) == slow(t, 2)
        q = rng.random(21) < 0.6
        assert wx.isvalidtree((8, 8), q) == slow(q, 4)
    assert wx.isvalidtree(64, wx.maketree(64, 6, "full")) and wx.isvalidtree((8, 8), wx.maketree(8, 8, 3, "dwt"))
    assert not wx.isvalidtree(64, np.ones(62, bool))                      # wrong length


def test_ldb_class_numbering_and_bb_types(wx):
    """host pieces of the f-1 / f-2 rows: class labels are numbered in order of first appearance (Julia unique) unless a global
    class list is given (needed when the batch is sharded); BB option types mirror the reference's defaults"""
    import numpy as np
    import torch
    lab, classes = wx.ldb._labels(["b", "a", "b", 3, "a"], torch.device("cpu"))
    assert classes == ["b", "a", 3] and lab.tolist() == [0, 1, 0, 2, 1]
    lab, classes = wx.ldb._labels(["a", 3], torch.device("cpu"), classes=["b", "a", 3])
    assert lab.tolist() == [1, 2] and classes == ["b", "a", 3]
    assert isinstance(wx.BB().cost, wx.ShannonEntropyCost) and wx.BB().redundant is False
    assert wx.LpDistance().p == 2 and wx.JBB().cost.p == 2 and wx.NormCost().p == 1


def test_denoising_host_objects(wx):
    """DNFT constructors (Denoising.jl:42-119, Wavelets.jl VisuShrink), the leaf masks the thresholding and the BitVector-indexed
    estimators use, ndyad (wavemult/utils.jl:146-155) -- host logic only, no device work"""
    import math
    from waveletsext_b200 import denoising as dn
    v = wx.VisuShrink(256)
    assert isinstance(v.th, wx.HardTH) and v.t == math.sqrt(2 * math.log(256))
    v = wx.VisuShrink(2, wx.SoftTH())
    assert isinstance(v.th, wx.SoftTH) and v.t == math.sqrt(2 * math.log(2))
    v = wx.VisuShrink(wx.HardTH(), 8)                       # test/denoising.jl:2
    assert v.t == 8.0
    r = wx.RelErrorShrink()
    assert isinstance(r.th, wx.HardTH) and r.t == 1.0
    assert wx.RelErrorShrink(wx.HardTH(), 1).t == 1.0 and isinstance(wx.RelErrorShrink(wx.SteinTH()).th, wx.SteinTH)
    s = wx.SureShrink(wx.HardTH(), 1)
    assert s.t == 1.0
    assert [c().code for c in (wx.HardTH, wx.SoftTH, wx.SemiSoftTH, wx.SteinTH)] == [0, 1, 2, 3]
    # leaves of maketree(8, 1, :full) in a 15-node table: nodes 2 and 3
    tree = wx.maketree(8, 1, "full")
    m = dn._leafmask(tree, 15)
    assert m.tolist() == [0, 1, 1] + [0] * 12
    assert dn._leafmask(tree, 3, strict=False).tolist() == [0, 1, 1]            # findall-style: a shallower table is fine
    with pytest.raises(IndexError):
        dn._leafmask(tree, 3)                                                    # BitVector-style: BoundsError
    with pytest.raises(IndexError):
        dn._leafmask(wx.maketree(8, 3, "full"), 7, strict=False)                 # leaves beyond the table
    assert wx.ndyad(1, 4, False) == range(16, 24) and wx.ndyad(4, 4, True) == range(3, 4)
    assert wx.finestdetailrange(8, wx.maketree(8, 3, "dwt")) == range(4, 8)
    assert wx.coarsestscalingrange(8, wx.maketree(8, 3, "dwt")) == range(0, 1)
    # the LDB object keeps the reference's defaults (LDB.jl:89-110)
    f = wx.LocalDiscriminantBasis()
    assert f.wt.name == "haar" and isinstance(f.dm, wx.AsymmetricRelativeEntropy) and isinstance(f.en, wx.TimeFrequency)
    assert isinstance(f.dp, wx.BasisDiscriminantMeasure) and f.top_k is None and f.n_features is None and f.tree is None


def test_wavelet_only_returns_verified_filters(wx):
    """Wavelets.jl's table filters without a verified copy here raise instead of returning a heuristic root choice"""
    import numpy as np
    for name in ("haar", "db2", "db4", "db7", "db10", "db12", "db16", "sym4", "sym8", "coif4"):
        q = wx.wavelet(name).taps
        assert wx.filters.check_orthonormal(q) < 1e-12, name
    for name in ("sym5", "sym6", "sym7", "sym9", "sym10", "coif2", "coif6", "beyl", "vaid", "batt2", "db17"):
        with pytest.raises(ValueError):
            wx.wavelet(name)
    # any other filter crosses as data
    q = wx.wavelet("db3").taps
    f = wx.filters.OrthoFilter(tuple(q), "custom")
    assert np.array_equal(wx.makereverseqmfpair(f, True)[0], q[::-1])
