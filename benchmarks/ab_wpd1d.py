#!/usr/bin/env python
"""A/B of the fused 1-D wpdall kernel with the tuned residency: L2 evict_first hints on / off (WX_B200_WPD1D_L2HINT), interleaved
repeats so that clock / power drift shows up as spread instead of bias.  One JSON line per (dtype, n, wavelet, hint)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import waveletsext_b200 as wx  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6552.0
    for dt in (torch.float64, torch.float32):
        es = 8 if dt == torch.float64 else 4
        for n, N, L in ((4096, 65536, 12), (1024, 131072, 10), (2048, 65536, 11)):
            x = torch.randn((N, n), dtype=dt, device=dev)
            y = torch.empty((N, L + 1, n), dtype=dt, device=dev)
            for wname in ("haar", "db2", "db4", "coif4", "sym8", "db7"):
                wt = wx.wavelet(wname)
                res = {0: [], 1: []}
                for rep in range(3):
                    for hint in (0, 1):
                        os.environ["WX_B200_WPD1D_L2HINT"] = str(hint)
                        for _ in range(2):
                            wx.dwt._wpd_batch(x, wt, L, y)          # first call of a (shape, hint) tunes the residency
                        torch.cuda.synchronize()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        for _ in range(5):
                            wx.dwt._wpd_batch(x, wt, L, y)
                        e1.record()
                        torch.cuda.synchronize()
                        res[hint].append(e0.elapsed_time(e1) / 5)
                b = es * n * N * (L + 2)
                for hint in (0, 1):
                    ms = min(res[hint])
                    print(json.dumps({"dtype": "f64" if es == 8 else "f32", "n": n, "N": N, "L": L, "wavelet": wname, "taps": len(wt.taps), "l2hint": hint,
                                      "ms_min": round(ms, 4), "ms_all": [round(v, 4) for v in res[hint]], "frac": round(b / (ms * 1e-3) / 1e9 / peak, 4)}), flush=True)
            del x, y
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
