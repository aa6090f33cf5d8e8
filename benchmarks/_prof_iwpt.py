"""three launches of the fused inverse (iwptall, complete tree) at the config-2 shape for ncu: python benchmarks/_prof_iwpt.py [wavelet] [f64|f32]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import waveletsext_b200 as wx
dev = torch.device("cuda:0")
wname = sys.argv[1] if len(sys.argv) > 1 else "db4"
dt = torch.float32 if (len(sys.argv) > 2 and sys.argv[2] == "f32") else torch.float64
n, N, L = 4096, 65536, 12
wt = wx.wavelet(wname)
x = torch.randn((N, n), dtype=dt, device=dev)
out = torch.empty_like(x)
tree = wx.maketree(n, L, "full")
for _ in range(3):
    wx.dwt._tree_batch("iwpt", x, wt, tree, out)
torch.cuda.synchronize()
