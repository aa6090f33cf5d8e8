#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 filter-bank path.

Metric (BASELINE.json): wpdall GSamples/s, Float64, db4, full depth.  Workload = BASELINE.json configs[1]:
65536 signals x 4096 samples, L = 12, on each GPU (weak scaling: every rank owns a full 65536-signal shard,
no data-path collective).  One "step" = one wpdall pass over the resident batch = ONE launch of the fused kernel.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  `value` is device-timed with inputs resident in HBM; `e2e` is the same metric through
the host-buffer C-ABI call (pinned host arrays, H2D + kernel + D2H inside the timed region); `roofline` compares the
algorithmic bytes (L+2)*n*N*sizeof(T) per launch with the measured HBM peak; `cpu_baseline` times the oracle port of
the reference loops (1 thread, like the single-threaded reference) on a bounded sample.
`--impl reference` times that CPU port with all host threads instead (Julia is not available in this image).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(n=4096, N=65536, L=12, wavelet="db4", dtype="f64")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=WORKLOAD["n"])
    ap.add_argument("--N", type=int, default=WORKLOAD["N"], help="signals per GPU")
    ap.add_argument("--L", type=int, default=WORKLOAD["L"])
    ap.add_argument("--wavelet", default=WORKLOAD["wavelet"])
    ap.add_argument("--dtype", default=WORKLOAD["dtype"], choices=["f64", "f32"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    return ap.parse_args()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.th = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_rate(n, L, q, dtype, threads, target_s):
    """oracle (CPU port of the reference loops) throughput in GSamples/s on a bounded sample"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import oracle as O
    dt = np.float64 if dtype == "f64" else np.float32
    chunk = 1024
    x = np.random.default_rng(20242).standard_normal((chunk, n)).astype(dt)
    y = np.empty((chunk, L + 1, n), dt)
    O.wpdall_into(y, x, q, L, threads)                      # warm-up + page-in
    t0 = time.perf_counter()
    reps = 0
    while True:
        O.wpdall_into(y, x, q, L, threads)
        reps += 1
        el = time.perf_counter() - t0
        if el >= target_s or reps >= 64:
            break
    return reps * chunk * n / el / 1e9, reps * chunk, el


def run_reference(a):
    """reference arm: the CPU port of the reference's wpdall on all host threads (rank 0 only)"""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import oracle as O
    from importlib import util
    spec = util.spec_from_file_location("wx_filters", os.path.join(ROOT, "waveletsext.jl_b200", "filters.py"))
    F = util.module_from_spec(spec); sys.modules["wx_filters"] = F; spec.loader.exec_module(F)
    q = F.wavelet(a.wavelet).taps
    # every host core this process may run on -- NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 to its workers, which
    # would silently turn the all-core reference arm into a single-thread run at N > 1
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    dt = np.float64 if a.dtype == "f64" else np.float32
    ns = 4096                                               # signals per step (bounded sample of the 65536-signal workload)
    x = np.random.default_rng(20242).standard_normal((ns, a.n)).astype(dt)
    y = np.empty((ns, a.L + 1, a.n), dt)
    for _ in range(max(a.warmup, 1)):
        O.wpdall_into(y, x, q, a.L, threads)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        O.wpdall_into(y, x, q, a.L, threads)
    el = time.perf_counter() - t0
    val = a.steps * ns * a.n / el / 1e9
    out = {"impl": "reference", "metric": "wpdall_gsamples_per_s", "value": val, "unit": "GSamples/s", "n_gpus": a.gpus, "steps": a.steps,
           "warmup": a.warmup, "ms_per_step": el / a.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": a.dtype, "data": "synthetic",
           "config": {"workload": f"wpdall {a.N} signals x {a.n} samples, {a.wavelet}, L={a.L}, {a.dtype}",
                      "sample": f"{ns} signals per step on the host CPU"},
           "cpu_baseline": {"value": val, "unit": "GSamples/s", "cores": threads, "kind": "port",
                            "sample": f"{ns} signals x {a.n} samples per step, {a.steps} steps; C restatement of the reference loops "
                                      f"(Julia unavailable), OpenMP over signals"},
           "e2e": {"value": val, "unit": "GSamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)

    import numpy as np
    import torch
    import torch.distributed as dist
    import waveletsext_b200 as wx

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    tdt = torch.float64 if a.dtype == "f64" else torch.float32
    esz = 8 if a.dtype == "f64" else 4
    wt = wx.wavelet(a.wavelet)
    n, N, L = a.n, a.N, a.L
    gen = torch.Generator(device=dev).manual_seed(20242 + rank)
    x = torch.randn((N, n), dtype=tdt, device=dev, generator=gen)
    y = torch.empty((N, L + 1, n), dtype=tdt, device=dev)

    def step():
        wx.dwt._wpd_batch(x, wt, L, y)

    for _ in range(max(a.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = wx.launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    evs[0].record()
    for i in range(a.steps):
        step()
        evs[i + 1].record()
    torch.cuda.synchronize()
    launches = wx.launch_count() - l0
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None

    total_ms = evs[0].elapsed_time(evs[-1])
    per_launch = [evs[i].elapsed_time(evs[i + 1]) for i in range(a.steps)]
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / a.steps
    value = n * N * world / (ms_per_step * 1e-3) / 1e9

    # ---- end-to-end through the host-buffer C-ABI call (pinned host arrays, copies inside the timed region)
    e2e = None
    if not a.no_e2e:
        Ne = N
        per_sig = (L + 2) * n * esz
        while Ne > 1024 and Ne * per_sig * world > 96e9:        # keep the pinned host footprint of all ranks under ~96 GB
            Ne //= 2
        xh = yh = None
        while Ne >= 1024:
            try:
                xh_np, xh = wx.host.pinned_empty((Ne, n), np.float64 if a.dtype == "f64" else np.float32)
                yh_np, yh = wx.host.pinned_empty((Ne, L + 1, n), np.float64 if a.dtype == "f64" else np.float32)
                break
            except Exception:
                xh = yh = None
                Ne //= 2
        if xh is not None:
            xh.copy_(x[:Ne])
            wx.host.wpdall_host(xh_np, wt, L, out=yh_np, device=local)              # warm-up
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(a.e2e_steps):
                wx.host.wpdall_host(xh_np, wt, L, out=yh_np, device=local)          # synchronous call
            el = time.perf_counter() - t0
            te = torch.tensor([el], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            el = float(te.item())
            ok = bool(torch.equal(yh[:4].to(dev), y[:4])) if Ne == N else True
            e2e = {"value": a.e2e_steps * Ne * n * world / el / 1e9, "unit": "GSamples/s", "h2d_bytes_per_step": Ne * n * esz,
                   "d2h_bytes_per_step": Ne * (L + 1) * n * esz, "steps": a.e2e_steps, "signals_per_step": Ne,
                   "matches_device_path": ok, "timer": "host wall clock around the synchronous C-ABI call, max over ranks"}
            del xh, yh

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    alg_bytes = esz * n * N * (L + 2)
    avg_launch_ms = sum(per_launch) / len(per_launch)
    achieved = alg_bytes / (avg_launch_ms * 1e-3) / 1e9
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        key = f"wpd1d_{a.dtype}_n{n}_N{N}_L{L}_{a.wavelet}"
        traffic = tj.get(key, {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": "wpd1d_tma_k", "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": avg_launch_ms, "peak_source": peak_src}

    cpu = None
    if not a.no_cpu and world == 1:
        rate, nsig, el = cpu_port_rate(n, L, wt.taps, a.dtype, 1, 12.0)
        cpu = {"value": rate, "unit": "GSamples/s", "cores": 1, "kind": "port",
               "sample": f"{nsig} signals x {n} samples ({el:.1f} s), C restatement of the reference loops, 1 thread like the "
                         f"single-threaded reference (Julia unavailable)"}

    out = {"metric": "wpdall_gsamples_per_s", "value": value, "unit": "GSamples/s", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
           "config": {"workload": f"wpdall {N} signals x {n} samples per GPU, {a.wavelet}, L={L}, {a.dtype} (BASELINE.json configs[1])",
                      "l2": "inputs+outputs (>= 30 GB per step) exceed the 126 MB L2, no flush needed", "sharding": "batch dimension, no collective"},
           "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
