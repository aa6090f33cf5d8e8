"""2-D wpdall (1024 x 512^2, L = 5) for long filters: whole-node kernel with the default / WIDE windows (WX_B200_WPD2D_BLKWIDE=0|1)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import waveletsext_b200 as wx
dev = torch.device("cuda:0")
m = n = 512; N = 1024; L = 5
for dt in (torch.float64, torch.float32):
    es = 8 if dt == torch.float64 else 4
    x = torch.randn((N, n, m), dtype=dt, device=dev); y = torch.empty((N, L + 1, n, m), dtype=dt, device=dev)
    for w in ("db4", "db5", "coif4", "sym8", "db10"):
        wt = wx.wavelet(w)
        best = 1e9
        for rep in range(3):
            for _ in range(2): wx.dwt._wpd_batch(x, wt, L, y)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): wx.dwt._wpd_batch(x, wt, L, y)
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 5)
        print(json.dumps({"blkwide": os.environ.get("WX_B200_WPD2D_BLKWIDE", "rule"), "dtype": "f64" if es == 8 else "f32", "wavelet": w, "ms": round(best, 4),
                          "frac": round(es * m * n * N * (L + 2) / (best * 1e-3) / 1e9 / 6552.0, 4)}), flush=True)
    del x, y
