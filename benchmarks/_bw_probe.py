"""write-only / copy bandwidth of this box (context for the write-dominated wpdall roofline): cudaMemset of the 27.9 GB table vs a
device-to-device copy of half of it"""
import json
import torch
dev = torch.device("cuda:0")
n = 27_917_287_424 // 8
y = torch.empty(n, dtype=torch.float64, device=dev)
def t(fn, steps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / steps
ms = t(lambda: y.zero_())
out = {"memset_GBps": round(n * 8 / ms / 1e6, 1), "memset_ms": round(ms, 3)}
ms = t(lambda: y.fill_(1.5))
out["fill_kernel_GBps"] = round(n * 8 / ms / 1e6, 1)
h = n // 2
ms = t(lambda: y[:h].copy_(y[h:2 * h]))
out["copy_GBps_read_plus_write"] = round(2 * h * 8 / ms / 1e6, 1)
print(json.dumps(out))
