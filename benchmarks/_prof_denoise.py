import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import waveletsext_b200 as wx
dev = torch.device("cuda:0")
n, N = 1024, 131072
wt = wx.wavelet("db4")
x = torch.randn((N, n), dtype=torch.float64, device=dev)
X = wx.dwtall(x, wt)
for _ in range(2):
    s = wx.noisest(X, False)
    Y = wx.denoising._threshold_into(torch.empty_like(X), X, wx.HardTH(), s * 3.0)
    t = wx.relerrorthreshold(X, False)
torch.cuda.synchronize()
