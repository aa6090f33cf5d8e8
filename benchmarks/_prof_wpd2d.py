import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import waveletsext_b200 as wx
dev = torch.device("cuda:0")
m = n = 512; N = 1024; L = 5
wt = wx.wavelet(sys.argv[1] if len(sys.argv) > 1 else "db4")
x = torch.randn((N, n, m), dtype=torch.float64, device=dev)
y = torch.empty((N, L + 1, n, m), dtype=torch.float64, device=dev)
for _ in range(2):
    wx.dwt._wpd_batch(x, wt, L, y)
torch.cuda.synchronize()
