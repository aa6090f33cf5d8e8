"""timings of the 2-D redundant inverses (64 x 256^2, L = 3, db4, F64): isdwtall / iswptall / iswpdall, average and shift based, iacwptall"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import waveletsext_b200 as wx
dev = torch.device("cuda:0")
m = n = 256; N = 64; L = 3
wt = wx.wavelet("db4")
x = torch.randn((N, n, m), dtype=torch.float64, device=dev)
def t(f, reps=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    l0 = wx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / reps, 4), (wx.launch_count() - l0) // reps
es = 8
img = m * n * N * es
for name, fwd, inv, nsl in (("isdwtall2d", lambda: wx.sdwtall(x, wt, L), lambda y: wx.isdwtall(y, wt), 3 * L + 1),
                            ("isdwtall2d_shift", lambda: wx.sdwtall(x, wt, L), lambda y: wx.isdwtall(y, wt, 5), 3 * L + 1),
                            ("iswptall2d", lambda: wx.swptall(x, wt, L), lambda y: wx.iswptall(y, wt), 4 ** L),
                            ("iswptall2d_shift", lambda: wx.swptall(x, wt, L), lambda y: wx.iswptall(y, wt, 5), 4 ** L),
                            ("iswpdall2d", lambda: wx.swpdall(x, wt, L), lambda y: wx.iswpdall(y, wt, L), 4 ** L),
                            ("iacwptall2d", lambda: wx.acwptall(x, wt, L), lambda y: wx.iacwptall(y, wt), 4 ** L)):
    try:
        y = fwd()
        ms, nl = t(lambda: inv(y))
        err = float((inv(y) - x).abs().max())
        print(json.dumps({"path": name, "ms": ms, "launches": nl, "algorithmic_GB": round((nsl + 1) * img / 1e9, 3),
                          "frac_of_hbm_peak": round((nsl + 1) * img / (ms * 1e-3) / 1e9 / 6552.0, 4), "roundtrip_abs_err": err}), flush=True)
    except Exception as e:
        print(json.dumps({"path": name, "error": repr(e)[:200]}), flush=True)
