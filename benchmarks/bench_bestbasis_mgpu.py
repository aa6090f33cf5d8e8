#!/usr/bin/env python
"""BASELINE.json configs[4]: JBB and LSDB bestbasistree + getbasiscoefall + iwptall on 2^20 signals x 1024 samples sharded over the
GPUs of one box (131072 signals per GPU at 8 GPUs), one process per GPU, NCCL all-reduce / all-gather of the cost-tree state.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        benchmarks/bench_bestbasis_mgpu.py [--per-gpu 131072] [--steps 5]

Prints one JSON line per stage (rank 0): device time (CUDA events, max over ranks), including the collectives and the host
selection.  The shards never move: only the per-position state (JBB 180 KB, LSDB < 7 MB) crosses NVLink.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import waveletsext_b200 as wx  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--per-gpu", type=int, default=131072)
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, N, L = a.n, a.per_gpu, int(np.log2(a.n))
    wt = wx.wavelet("db4")
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    t = torch.arange(n, device=dev, dtype=torch.float64) / n
    hs = 4 * torch.sin(4 * np.pi * t) - torch.sign(t - 0.3) - torch.sign(0.72 - t)
    idx = (torch.arange(n, device=dev)[None, :] - 2 * ((torch.arange(N, device=dev)[:, None] + rank * N) % n)) % n
    x = torch.gather(hs[None, :].repeat(N, 1), 1, idx) + 0.5 * torch.randn((N, n), dtype=torch.float64, device=dev, generator=gen)
    del idx
    Xw = wx.wpdall(x, wt, L)

    def timed(fn, steps):
        fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    res = {}
    res["wpdall"] = timed(lambda: wx.dwt._wpd_batch(x, wt, L, Xw), a.steps)[0]
    res["bestbasistree_JBB"], tj = timed(lambda: wx.bestbasistree(Xw, wx.JBB()), a.steps)
    res["bestbasistree_LSDB"], tl = timed(lambda: wx.bestbasistree(Xw, wx.LSDB()), max(a.steps // 2, 1))
    res["getbasiscoefall+iwptall"], xr = timed(lambda: wx.iwptall(wx.getbasiscoefall(Xw, tj), wt, tj), a.steps)
    err = float((xr - x).abs().max() / x.abs().max())
    # every rank must hold the same trees
    for tr in (tj, tl):
        tt = torch.from_numpy(tr.astype(np.int64)).to(dev)
        lo, hi = tt.clone(), tt.clone()
        if world > 1:
            dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), "ranks disagree on the best-basis tree"
    if rank == 0:
        tot = N * world
        for k, ms in res.items():
            print(json.dumps({"stage": k, "n_gpus": world, "signals_total": tot, "n": n, "ms": round(ms, 4),
                              "GSamples_per_s": round(tot * n / (ms * 1e-3) / 1e9, 2)}), flush=True)
        print(json.dumps({"check": "roundtrip_relerr", "value": err, "jbb_tree_nodes": int(tj.sum()), "lsdb_tree_nodes": int(tl.sum())}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
