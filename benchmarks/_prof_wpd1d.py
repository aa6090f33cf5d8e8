"""three launches of the fused 1-D wpdall at the headline shape (65536 x 4096, L = 12) for ncu: python benchmarks/_prof_wpd1d.py [wavelet] [f64|f32] [n] [N]
Residency comes from WX_B200_WPD1D_OCC (set it, together with WX_B200_AUTOTUNE=0, so that the profiled launch is the shipped one)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import waveletsext_b200 as wx
dev = torch.device("cuda:0")
wname = sys.argv[1] if len(sys.argv) > 1 else "db4"
dt = torch.float32 if (len(sys.argv) > 2 and sys.argv[2] == "f32") else torch.float64
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
N = int(sys.argv[4]) if len(sys.argv) > 4 else 65536
L = n.bit_length() - 1
wt = wx.wavelet(wname)
x = torch.randn((N, n), dtype=dt, device=dev)
y = torch.empty((N, L + 1, n), dtype=dt, device=dev)
for _ in range(3):
    wx.dwt._wpd_batch(x, wt, L, y)
torch.cuda.synchronize()
