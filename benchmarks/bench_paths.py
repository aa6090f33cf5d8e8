#!/usr/bin/env python
"""Secondary measurements for the other hot-path rows of SURVEY.md section 8 (configs 2-5 of BASELINE.json): device-timed,
inputs resident, algorithmic bytes / CUDA-event time against the measured HBM peak.  One JSON line per path.

    python benchmarks/bench_paths.py [--only name,...] [--scale S]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import waveletsext_b200 as wx  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def timeit(fn, steps=5, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    l0 = wx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, (wx.launch_count() - l0) // steps


def report(name, ms, launches, alg_bytes, units, unit_name, extra=None):
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    out = {"path": name, "ms": round(ms, 4), "launches_per_call": launches, "algorithmic_GB": round(alg_bytes / 1e9, 3), "achieved_GBps": round(gbs, 1),
           "frac_of_measured_hbm_peak": round(gbs / peak(), 4), unit_name: round(units / (ms * 1e-3) / 1e9, 3)}
    if extra:
        out.update(extra)
    print(json.dumps(out), flush=True)


def fp64_bound(fmas):
    """time of `fmas` double-precision FMAs at the B200's FP64 pipe rate (64 DFMA per clock per SM, 148 SMs, max SM clock of
    MEASURED_PEAKS.json): the second roofline of the long filters in Float64"""
    try:
        mhz = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["sm_max_mhz"])
    except Exception:
        mhz = 1965.0
    return fmas / (64.0 * 148.0 * mhz * 1e6) * 1e3


def cpu_rate(fn, units, budget_s=2.0):
    """units per second of the CPU port: repeat fn() (which processes `units` samples / pixels) for about budget_s seconds"""
    import time
    fn()
    t0 = time.perf_counter(); reps = 0
    while True:
        fn(); reps += 1
        el = time.perf_counter() - t0
        if el >= budget_s or reps >= 50:
            break
    return reps * units / el / 1e9


def cpu_baselines():
    """CPU port (oracle/, test infrastructure) timed on the host, one thread, bounded samples; one JSON line per path"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    rng = np.random.default_rng(1)
    q = wx.wavelet("db4").taps
    g, h = O.makereverseqmfpair(q)
    P, Q = O.make_acreverseqmfpair(q)
    out = {}
    x = rng.standard_normal((64, 4096))
    out["wpdall_f64_db4"] = (cpu_rate(lambda: O.wpdall(x, q, 12, 1), x.size), "GSamples_per_s", "64 signals x 4096, L=12")
    leaves = O.wpdall(x, q, 12, 1)[:, 12].copy()
    tree = O.maketree1(4096, 12, "full")
    out["iwptall_f64_db4"] = (cpu_rate(lambda: O.iwptall(leaves, q, tree, 1), x.size), "GSamples_per_s", "64 signals x 4096, full tree")
    xs = rng.standard_normal((16, 2048))
    out["swpdall_f64"] = (cpu_rate(lambda: O.rwpdall(0, xs, 8, h, g, 1), xs.size), "GSamples_per_s", "16 signals x 2048, L=8")
    out["acwpdall_f64"] = (cpu_rate(lambda: O.rwpdall(1, xs, 8, Q, P, 1), xs.size), "GSamples_per_s", "16 signals x 2048, L=8")
    img = rng.standard_normal((4, 512, 512))
    out["wpd2d_f64_db4"] = (cpu_rate(lambda: O.wpdall(img, q, 5, 1), img.size), "GPixels_per_s", "4 images 512 x 512, L=5")
    Xw = O.wpdall(rng.standard_normal((2048, 1024)), q, 10, 1)
    out["jbb_tree_costs_f64"] = (cpu_rate(lambda: O.tree_costs_jbb(Xw), 2048 * 1024), "GSamples_per_s", "2048 signals x 1024 x 11 levels")
    out["lsdb_tree_costs_f64"] = (cpu_rate(lambda: O.tree_costs_lsdb(Xw), 2048 * 1024), "GSamples_per_s", "2048 signals x 1024 x 11 levels")
    for k, (v, unit, sample) in out.items():
        print(json.dumps({"path": k, "cpu_port_1thread": {unit: round(v, 5), "sample": sample, "kind": "port (C restatement of the reference loops, "
                                                                                                    "Julia unavailable)"}}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--scale", type=float, default=1.0, help="scale the batch sizes (1.0 = the per-GPU sizes of BASELINE.json)")
    ap.add_argument("--cpu", action="store_true", help="also time the CPU port of the reference loops (oracle, 1 thread like the "
                    "single-threaded reference) on a bounded sample of each workload")
    a = ap.parse_args()
    only = set(filter(None, a.only.split(",")))
    if a.cpu:
        cpu_baselines()
        if a.only == "cpu":
            return
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(7)
    want = lambda nm: (not only) or nm in only

    for dt, es in ((torch.float64, 8), (torch.float32, 4)):
        tag = "f64" if es == 8 else "f32"
        # config 2: wpdall + iwptall round trip, 65536 x 4096, L = 12
        n, N, L = 4096, int(65536 * a.scale), 12
        for wname in ("db4", "coif4", "sym8"):
            if want(f"iwptall_{tag}_{wname}") or want(f"wpdall_{tag}_{wname}"):
                wt = wx.wavelet(wname)
                x = torch.randn((N, n), dtype=dt, device=dev, generator=gen)
                y = torch.empty((N, L + 1, n), dtype=dt, device=dev)
                if want(f"wpdall_{tag}_{wname}"):
                    ms, nl = timeit(lambda: wx.dwt._wpd_batch(x, wt, L, y))
                    extra = None
                    if es == 8:          # F FMAs per sample and level for each of the two filters = F*L*n*N in total... per OUTPUT pair 2F, i.e. F per sample
                        fb = fp64_bound(float(len(wt.taps)) * L * n * N)
                        extra = {"fp64_pipe_bound_ms": round(fb, 3), "frac_of_fp64_bound": round(fb / ms, 4)}
                    report(f"wpdall_{tag}_{wname}", ms, nl, es * n * N * (L + 2), n * N, "GSamples_per_s", extra)
                if want(f"iwptall_{tag}_{wname}"):
                    wx.dwt._wpd_batch(x, wt, L, y)
                    leaves = y[:, L].contiguous()
                    del y
                    out = torch.empty_like(leaves)
                    tree = wx.maketree(n, L, "full")
                    ms, nl = timeit(lambda: wx.dwt._tree_batch("iwpt", leaves, wt, tree, out))
                    err = float((out - x).abs().max() / x.abs().max())
                    extra = {"roundtrip_relerr": err}
                    if es == 8:
                        fb = fp64_bound(float(len(wt.taps)) * L * n * N)
                        extra.update({"fp64_pipe_bound_ms": round(fb, 3), "frac_of_fp64_bound": round(fb / ms, 4)})
                    report(f"iwptall_{tag}_{wname}", ms, nl, 2 * es * n * N, n * N, "GSamples_per_s", extra)
                    del leaves, out
                del x
                torch.cuda.empty_cache()
        # config 3: swpd / acwpd 2048 signals x 2048 samples per GPU (16384 over 8 GPUs), L = 8
        n, N, L = 2048, int(2048 * a.scale), 8
        for nm, fn in (("swpdall", wx.swt), ("acwpdall", wx.acwt)):
            if want(f"{nm}_{tag}"):
                wt = wx.wavelet("db4")
                x = torch.randn((N, n), dtype=dt, device=dev, generator=gen)
                xw = torch.empty((N, (1 << (L + 1)) - 1, n), dtype=dt, device=dev)
                ac = nm.startswith("ac")
                ms, nl = timeit(lambda: wx._rwt.forward(ac, "wpd", x, wt, L, xw), steps=3, warmup=1)
                report(f"{nm}_{tag}", ms, nl, es * n * N * (1 << (L + 1)), n * N, "GSamples_per_s")
                del x, xw
                torch.cuda.empty_cache()
        # config 4: 2-D wpd, 4096 images 512 x 512, L = 5
        m = n2 = 512
        N, L = int(4096 * a.scale), 5
        for wname in ("haar", "db4", "sym8"):
            if want(f"wpd2d_{tag}_{wname}"):
                wt = wx.wavelet(wname)
                x = torch.randn((N, n2, m), dtype=dt, device=dev, generator=gen)
                y = torch.empty((N, L + 1, n2, m), dtype=dt, device=dev)
                ms, nl = timeit(lambda: wx.dwt._wpd_batch(x, wt, L, y), steps=3, warmup=1)
                report(f"wpd2d_{tag}_{wname}", ms, nl, es * m * n2 * N * (L + 2), m * n2 * N, "GPixels_per_s")
                del x, y
                torch.cuda.empty_cache()
    # general path (one launch per depth): the other output shapes and the inverses of the redundant families, 2-D by tree
    if any(want(k) for k in ("general", "sdwt", "iswt", "tree2d", "swt2d")):
        dt, es = torch.float64, 8
        n, N, L = 2048, int(2048 * a.scale), 8
        wt = wx.wavelet("db4")
        x = torch.randn((N, n), dtype=dt, device=dev, generator=gen)
        if want("general") or want("sdwt"):
            for nm, fn in (("sdwtall", wx.sdwtall), ("acdwtall", wx.acdwtall)):
                ms, nl = timeit(lambda: fn(x, wt, L), steps=3, warmup=1)
                report(f"{nm}_f64", ms, nl, es * n * N * (L + 2), n * N, "GSamples_per_s")
        if want("general") or want("iswt"):
            xd = wx.sdwtall(x, wt, L)
            for sm in (None, 5):
                ms, nl = timeit(lambda: wx.isdwtall(xd, wt, sm), steps=3, warmup=1)
                report(f"isdwtall_f64_{'avg' if sm is None else 'shift'}", ms, nl, es * n * N * (L + 2), n * N, "GSamples_per_s")
            ms, nl = timeit(lambda: wx.iacdwtall(wx.acdwtall(x, wt, L)), steps=3, warmup=1)
            del xd
            xt = wx.swptall(x, wt, L)
            for sm in (None, 5):
                ms, nl = timeit(lambda: wx.iswptall(xt, wt, sm), steps=3, warmup=1)
                err = float((wx.iswptall(xt, wt, sm) - x).abs().max() / x.abs().max())
                report(f"iswptall_f64_{'avg' if sm is None else 'shift'}", ms, nl, es * n * N * ((1 << L) + 1), n * N, "GSamples_per_s", {"roundtrip_relerr": err})
            xa = wx.acwptall(x, wt, L)
            ms, nl = timeit(lambda: wx.iacwptall(xa), steps=3, warmup=1)
            report("iacwptall_f64", ms, nl, es * n * N * ((1 << L) + 1), n * N, "GSamples_per_s")
            del xt, xa
            torch.cuda.empty_cache()
        del x
        if want("general") or want("tree2d"):
            m = n2 = 512
            N2, L2 = int(1024 * a.scale), 5
            xi = torch.randn((N2, n2, m), dtype=dt, device=dev, generator=gen)
            tree = wx.maketree(m, n2, L2, "full")
            ms, nl = timeit(lambda: wx.wptall(xi, wt, tree), steps=3, warmup=1)
            report("wptall2d_f64", ms, nl, 2 * es * m * n2 * N2, m * n2 * N2, "GPixels_per_s")
            yi = wx.wptall(xi, wt, tree)
            ms, nl = timeit(lambda: wx.iwptall(yi, wt, tree), steps=3, warmup=1)
            err = float((wx.iwptall(yi, wt, tree) - xi).abs().max() / xi.abs().max())
            report("iwptall2d_f64", ms, nl, 2 * es * m * n2 * N2, m * n2 * N2, "GPixels_per_s", {"roundtrip_relerr": err})
            del xi, yi
            torch.cuda.empty_cache()
        if want("general") or want("swt2d"):
            m = n2 = 256
            N2, L2 = int(64 * a.scale) or 1, 3
            xi = torch.randn((N2, n2, m), dtype=dt, device=dev, generator=gen)
            nsl = (4 ** (L2 + 1) - 1) // 3
            ms, nl = timeit(lambda: wx.swpdall(xi, wt, L2), steps=3, warmup=1)
            report("swpdall2d_f64", ms, nl, es * m * n2 * N2 * (nsl + 1), m * n2 * N2, "GPixels_per_s")
            ms, nl = timeit(lambda: wx.acwpdall(xi, wt, L2), steps=3, warmup=1)
            report("acwpdall2d_f64", ms, nl, es * m * n2 * N2 * (nsl + 1), m * n2 * N2, "GPixels_per_s")
            del xi
            torch.cuda.empty_cache()
    # the paper's pipeline end to end from HOST buffers (paper/paper.md:100-118): x (pinned host) -> H2D -> wpdall -> bestbasistree(JBB)
    # -> getbasiscoefall -> D2H of the best-basis coefficients.  The 13x larger packet table never crosses PCIe.
    if want("pipeline_host"):
        import time
        n, N, L = 4096, int(65536 * a.scale), 12
        wt = wx.wavelet("db4")
        xh = torch.randn((N, n), dtype=torch.float64).pin_memory()
        ch = torch.empty((N, n), dtype=torch.float64).pin_memory()
        y = torch.empty((N, L + 1, n), dtype=torch.float64, device=dev)

        def run():
            xd = xh.to(dev, non_blocking=True)
            wx.dwt._wpd_batch(xd, wt, L, y)
            tree = wx.bestbasistree(y, wx.JBB())
            ch.copy_(wx.getbasiscoefall(y, tree), non_blocking=True)
            torch.cuda.synchronize()
            return tree
        run()
        t0 = time.perf_counter()
        for _ in range(3):
            tree = run()
        el = (time.perf_counter() - t0) / 3
        print(json.dumps({"path": "pipeline_host_wpd_jbb_basiscoef_f64", "ms": round(el * 1e3, 3), "GSamples_per_s": round(n * N / el / 1e9, 3),
                          "h2d_bytes": 8 * n * N, "d2h_bytes": 8 * n * N, "tree_nodes": int(tree.sum()),
                          "timer": "host wall clock, pinned host buffers, copies inside the timed region"}), flush=True)
        del xh, ch, y
        torch.cuda.empty_cache()
    # next row f-3: denoiseall on the 131072 x 1024 batch of config 5 (dwt coefficients and swpd tables of a smaller batch)
    if want("denoise"):
        n, N, L = 1024, int(131072 * a.scale), 10
        wt = wx.wavelet("db4")
        x = torch.randn((N, n), dtype=torch.float64, device=dev, generator=gen)
        X = wx.dwtall(x, wt)
        ms, nl = timeit(lambda: wx.noisest(X, False), steps=3, warmup=1)
        report("noisest_dwt_f64", ms, nl, 8 * (n // 2) * N, n * N, "GSamples_per_s")
        sig = wx.noisest(X, False)
        Y = torch.empty_like(X)
        ms, nl = timeit(lambda: wx.denoising._threshold_into(Y, X, wx.HardTH(), sig * 3.0), steps=3, warmup=1)
        report("threshold_hard_f64", ms, nl, 2 * 8 * n * N, n * N, "GSamples_per_s")
        ms, nl = timeit(lambda: wx.denoising._threshold_into(Y, X, wx.SoftTH(), sig * 3.0), steps=3, warmup=1)
        report("threshold_soft_f64", ms, nl, 2 * 8 * n * N, n * N, "GSamples_per_s")
        ms, nl = timeit(lambda: wx.denoiseall(X, "dwt", wt), steps=3, warmup=1)
        report("denoiseall_dwt_visushrink_f64", ms, nl, (8 * (n // 2) + 4 * 8 * n) * N, n * N, "GSamples_per_s")
        ms, nl = timeit(lambda: wx.surethreshold(X, False), steps=2, warmup=1)
        report("surethreshold_dwt_f64", ms, nl, 8 * n * N, n * N, "GSamples_per_s")
        ms, nl = timeit(lambda: wx.relerrorthreshold(X, False), steps=2, warmup=1)
        report("relerrorthreshold_dwt_f64", ms, nl, 8 * n * N, n * N, "GSamples_per_s")
        del x, X, Y
        n, N, L = 2048, int(2048 * a.scale), 8
        x = torch.randn((N, n), dtype=torch.float64, device=dev, generator=gen)
        S = wx.swpdall(x, wt, L)
        tree = wx.maketree(n, L, "full")
        K = S.shape[1]
        ms, nl = timeit(lambda: wx.denoiseall(S, "swpd", wt, tree=tree), steps=3, warmup=1)
        report("denoiseall_swpd_fulltree_f64", ms, nl, 8 * n * N * (2 * K + (1 << L) + 2), n * N, "GSamples_per_s")
        del x, S
        torch.cuda.empty_cache()
    # config 5: JBB / LSDB best basis + getbasiscoefall + iwptall on 131072 signals x 1024 per GPU (1M over 8 GPUs)
    n, N, L = 1024, int(131072 * a.scale), 10
    if want("jbb") or want("lsdb") or want("basis_iwpt") or want("bb"):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        wt = wx.wavelet("db4")
        t = torch.arange(n, device=dev, dtype=torch.float64) / n
        hs = 4 * torch.sin(4 * np.pi * t) - torch.sign(t - 0.3) - torch.sign(0.72 - t)
        x = hs[None, :].repeat(N, 1)
        idx = (torch.arange(n, device=dev)[None, :] - 2 * (torch.arange(N, device=dev)[:, None] % n)) % n
        x = torch.gather(x, 1, idx) + 0.5 * torch.randn((N, n), dtype=torch.float64, device=dev, generator=gen)
        Xw = wx.wpdall(x, wt, L)
        K = L + 1
        tree = None
        if want("jbb"):
            ms, nl = timeit(lambda: wx.tree_costs(Xw, wx.JBB()), steps=3, warmup=1)
            report("jbb_tree_costs_f64", ms, nl, 8 * n * K * N, n * N, "GSamples_per_s")
            tree = wx.bestbasistree(Xw, wx.JBB())
        if want("lsdb"):
            ms, nl = timeit(lambda: wx.tree_costs(Xw, wx.LSDB()), steps=2, warmup=1)
            report("lsdb_tree_costs_f64", ms, nl, 3 * 8 * n * K * N, n * N, "GSamples_per_s")
        if want("bb"):
            ms, nl = timeit(lambda: wx.bestbasistreeall(Xw, wx.BB()), steps=3, warmup=1)
            report("bestbasistreeall_bb_f64", ms, nl, 8 * n * K * N, n * N, "GSamples_per_s")
            trees = wx.bestbasistreeall(Xw, wx.BB())
            ms, nl = timeit(lambda: wx.getbasiscoefall(Xw, trees), steps=3, warmup=1)
            report("getbasiscoefall_per_signal_trees_f64", ms, nl, 2 * 8 * n * N, n * N, "GSamples_per_s", {"mean_tree_nodes": float(trees.sum()) / N})
            del trees
        if want("basis_iwpt"):
            if tree is None:
                tree = wx.bestbasistree(Xw, wx.JBB())
            ms, nl = timeit(lambda: wx.iwptall(wx.getbasiscoefall(Xw, tree), wt, tree), steps=3, warmup=1)
            xr = wx.iwptall(wx.getbasiscoefall(Xw, tree), wt, tree)
            report("getbasiscoefall+iwptall_f64", ms, nl, 3 * 8 * n * N, n * N, "GSamples_per_s",
                   {"tree_nodes": int(tree.sum()), "roundtrip_relerr": float((xr - x).abs().max() / x.abs().max())})
            # the same reconstruction in one launch: iwpdall gathers the best-basis coefficients from the table inside the kernel
            ms, nl = timeit(lambda: wx.iwpdall(Xw, wt, tree), steps=3, warmup=1)
            xr = wx.iwpdall(Xw, wt, tree)
            report("iwpdall_fused_gather_f64", ms, nl, 2 * 8 * n * N, n * N, "GSamples_per_s",
                   {"roundtrip_relerr": float((xr - x).abs().max() / x.abs().max())})


if __name__ == "__main__":
    main()
