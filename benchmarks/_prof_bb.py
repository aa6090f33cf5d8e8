"""two calls of bestbasistreeall(BB) on the config-5 per-GPU table for ncu"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import waveletsext_b200 as wx
dev = torch.device("cuda:0")
n, N, L = 1024, 131072, 10
x = torch.randn((N, n), dtype=torch.float64, device=dev)
Xw = wx.wpdall(x, wx.wavelet("db4"), L)
for _ in range(2):
    wx.bestbasistreeall(Xw, wx.BB())
torch.cuda.synchronize()
