"""iwptall / wptall timings (complete tree, 65536 x 4096 and 131072 x 1024) for A-B builds of the fused tree kernel (WX_B200_LIB)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import waveletsext_b200 as wx
dev = torch.device("cuda:0")
for dt in (torch.float64, torch.float32):
    for n, N, L in ((4096, 65536, 12), (1024, 131072, 10)):
        x = torch.randn((N, n), dtype=dt, device=dev)
        tree = wx.maketree(n, L, "full")
        out = torch.empty_like(x); back = torch.empty_like(x)
        for w in ("db4", "coif4", "sym8"):
            wt = wx.wavelet(w)
            wx.dwt._tree_batch("wpt", x, wt, tree, out)
            res = {}
            for nm, src, dst in (("iwpt", out, back), ("wpt", x, out)):
                best = 1e9
                for rep in range(3):
                    wx.dwt._tree_batch(nm, src, wt, tree, dst)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(5): wx.dwt._tree_batch(nm, src, wt, tree, dst)
                    e1.record(); torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1) / 5)
                res[nm] = round(best, 4)
            err = float((back - x).abs().max() / x.abs().max())
            print(json.dumps({"lib": os.path.basename(os.environ.get("WX_B200_LIB", "default")), "dtype": str(dt)[-7:], "n": n, "wavelet": w, **res, "roundtrip": err}), flush=True)
        del x, out, back
