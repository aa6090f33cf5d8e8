#!/usr/bin/env python
"""Residency sweep of the fused 1-D wpdall kernel (wpd1d_tma_k): resident CTAs per SM (WX_B200_WPD1D_OCC) x threads per CTA
(WX_B200_WPD1D_THREADS) for every filter length / element type / signal length of the BASELINE configs.  One JSON line per point;
the launcher's residency table (csrc/wx_wpd1d.cu) is derived from this output (profiles/r2_wpd1d_residency_sweep.jsonl)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import waveletsext_b200 as wx  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    peak = 6552.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    shapes = [(4096, 65536, 12), (1024, 131072, 10)]
    for dt in (torch.float64, torch.float32):
        es = 8 if dt == torch.float64 else 4
        for n, N, L in shapes:
            x = torch.randn((N, n), dtype=dt, device=dev)
            y = torch.empty((N, L + 1, n), dtype=dt, device=dev)
            for wname in ("haar", "db2", "db3", "db4", "db5", "coif4", "sym8", "db10"):
                wt = wx.wavelet(wname)
                for thr in (256, 512):
                    for occ in (0, 1, 2, 3, 4, 5, 6, 8):
                        os.environ["WX_B200_WPD1D_THREADS"] = str(thr)
                        if occ:
                            os.environ["WX_B200_WPD1D_OCC"] = str(occ)
                        else:
                            os.environ.pop("WX_B200_WPD1D_OCC", None)
                        for _ in range(2):
                            wx.dwt._wpd_batch(x, wt, L, y)
                        torch.cuda.synchronize()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        for _ in range(4):
                            wx.dwt._wpd_batch(x, wt, L, y)
                        e1.record()
                        torch.cuda.synchronize()
                        ms = e0.elapsed_time(e1) / 4
                        b = es * n * N * (L + 2)
                        print(json.dumps({"dtype": "f64" if es == 8 else "f32", "n": n, "N": N, "L": L, "wavelet": wname, "taps": len(wt.taps),
                                          "threads": thr, "occ": occ or "default", "ms": round(ms, 4), "frac": round(b / (ms * 1e-3) / 1e9 / peak, 4)}),
                              flush=True)
            del x, y
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
