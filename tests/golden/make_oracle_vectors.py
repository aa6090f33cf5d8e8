#!/usr/bin/env python
"""Generates tests/golden/oracle_vectors.npz: seeded inputs, the taps used, and the outputs of the PINNED CPU oracle
(oracle/wx_oracle.c, checked against the reference's known answers by tests/test_oracle_golden.py) for every function of
the hot path.  The GPU parity tests compare the CUDA path with these committed vectors; the CPU tests re-run the oracle
against them so that a change of the oracle cannot go unnoticed.

    python tests/golden/make_oracle_vectors.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402
import importlib.util  # noqa: E402

spec = importlib.util.spec_from_file_location("wx_filters_gold", os.path.join(ROOT, "waveletsext.jl_b200", "filters.py"))
F = importlib.util.module_from_spec(spec); sys.modules[spec.name] = F; spec.loader.exec_module(F)


def build():
    out = {}
    rng = np.random.default_rng(20241017)
    for name in ("haar", "db4", "coif4", "sym8"):
        q = F.wavelet(name).taps
        g, h = O.makereverseqmfpair(q)
        P, Q = O.make_acreverseqmfpair(q)
        out[f"{name}/qmf"] = q
        x = rng.standard_normal((3, 64))
        out[f"{name}/x"] = x
        out[f"{name}/wpd"] = np.stack([O.wpd(x[k], h, g, 6) for k in range(3)])
        tree = np.zeros(63, bool)
        for i in range(1, 64):
            if (i == 1 or tree[i // 2 - 1]) and rng.random() < 0.75:
                tree[i - 1] = True
        out[f"{name}/tree"] = tree
        wpt = np.stack([O.wpt(x[k], tree, h, g) for k in range(3)])
        out[f"{name}/wpt"] = wpt
        out[f"{name}/iwpt"] = np.stack([O.iwpt(wpt[k], tree, h, g) for k in range(3)])
        out[f"{name}/swpd"] = np.stack([O.swpd(x[k], 4, h, g) for k in range(3)])
        out[f"{name}/sdwt"] = np.stack([O.sdwt(x[k], 4, h, g) for k in range(3)])
        out[f"{name}/acwpd"] = np.stack([O.acwpd(x[k], 4, P, Q) for k in range(3)])
        sw = out[f"{name}/swpd"]
        out[f"{name}/iswpd_avg"] = np.stack([O.iswpd(sw[k], O.maketree1(64, 4, "full"), h, g) for k in range(3)])
        out[f"{name}/iswpd_sm5"] = np.stack([O.iswpd(sw[k], O.maketree1(64, 4, "full"), h, g, 5) for k in range(3)])
        img = rng.standard_normal((2, 32, 16))                    # (N, cols, rows): images of 16 rows x 32 columns
        out[f"{name}/img"] = img
        out[f"{name}/wpd2d"] = np.stack([O.wpd(img[k], h, g, 3) for k in range(2)])
    # best basis (db4 packet table of a noisy shifted heavisine batch)
    q = F.wavelet("db4").taps
    g, h = O.makereverseqmfpair(q)
    n, N = 64, 200
    t = np.arange(n) / n
    hs = 4 * np.sin(4 * np.pi * t) - np.sign(t - 0.3) - np.sign(0.72 - t)
    X = np.stack([np.roll(hs, 2 * (k % n)) for k in range(N)]) + 0.5 * rng.standard_normal((N, n))
    Xw = np.stack([O.wpd(X[k], h, g, 6) for k in range(N)])
    out["bb/X"] = X
    out["bb/jbb_costs"] = O.tree_costs_jbb(Xw)
    out["bb/jbb_tree"] = O.tree_select(out["bb/jbb_costs"].copy(), n)
    out["bb/lsdb_costs"] = O.tree_costs_lsdb(Xw)
    out["bb/lsdb_tree"] = O.tree_select(out["bb/lsdb_costs"].copy(), n)
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **build())
    print("wrote", os.path.join(HERE, "oracle_vectors.npz"))
