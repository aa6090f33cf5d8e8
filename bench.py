#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 filter-bank path.

Metric (BASELINE.json): wpdall GSamples/s, Float64, db4, full depth.  Workload = BASELINE.json configs[1]:
65536 signals x 4096 samples, L = 12, on each GPU (weak scaling: every rank owns a full 65536-signal shard,
no data-path collective).  One "step" = one wpdall pass over the resident batch = ONE launch of the fused kernel.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--blocks a,b,...]

Prints ONE JSON line (rank 0).  `value` is device-timed with inputs resident in HBM; `e2e` is the same metric through
the host-buffer C-ABI call (pinned host arrays, H2D + kernel + D2H inside the timed region); `roofline` compares the
algorithmic bytes (L+2)*n*N*sizeof(T) per launch with the measured HBM peak; `cpu_baseline` times the oracle port of
the reference loops (1 thread like the single-threaded reference, plus all cores) on a bounded sample.

Further blocks of the same line (the other BASELINE.json configs, each on its per-GPU share so that N GPUs run the named total):
  pipeline      configs[4]: 131072 signals x 1024 per GPU (2^20 over 8): wpdall -> bestbasistree(JBB) and (LSDB), whose cost-tree
                state crosses ranks through libwx_b200's own NCCL communicator -> getbasiscoefall + iwptall; per-stage device
                ms (max over ranks), bytes exchanged, the all-reduce latency on its own, tree hashes of every rank.
  config3       configs[2]: swpdall / acwpdall on 2048 signals x 2048 per GPU (16384 over 8), L = 8.
  config1       configs[0]: wpdall 1024 x 1024, L = 10 (the reference's CPU-runnable shape): GPU ms (L2 flushed between launches)
                and the CPU port at that shape, 1 thread and all cores.
  e2e_pipeline  x (pinned host) -> wpdall -> bestbasistree(JBB) -> getbasiscoefall -> best-basis coefficients (pinned host) through
                ONE C-ABI call (wx_wpd_bestbasis_host); the 13x larger packet table never leaves HBM.
`--impl reference` times the CPU port with all host threads instead (Julia is not available in this image).
"""
from __future__ import annotations

import argparse
import ctypes
import hashlib
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
ALL_BLOCKS = "config1,pipeline,config3,e2e,e2e_pipeline,cpu"


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
    ap.add_argument("--blocks", default=ALL_BLOCKS, help="comma list of the extra blocks to run (see the docstring); '' = headline only")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    return ap.parse_args()


def workload_config(a):
    """identical in both arms"""
    return {"workload": f"wpdall {a.N} signals x {a.n} samples per GPU, {a.wavelet}, L={a.L}, {a.dtype} (BASELINE.json configs[1])",
            "l2": "inputs+outputs (>= 30 GB per step) exceed the 126 MB L2, no flush needed", "sharding": "batch dimension, no collective"}


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


# ---- CPU legs (the oracle port: the only place bench.py executes oracle/) --------------------------------------------
def host_threads():
    # every host core this process may run on -- NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 to its workers
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def cpu_port_rate(n, L, q, dtype, threads, target_s, chunk=1024, max_reps=64):
    """oracle (CPU port of the reference loops) throughput in GSamples/s on a bounded sample"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import oracle as O
    dt = np.float64 if dtype == "f64" else np.float32
    x = np.random.default_rng(20242).standard_normal((chunk, n)).astype(dt)
    y = np.empty((chunk, L + 1, n), dt)
    O.wpdall_into(y, x, q, L, threads)                      # warm-up + page-in
    t0 = time.perf_counter()
    reps = 0
    while True:
        O.wpdall_into(y, x, q, L, threads)
        reps += 1
        el = time.perf_counter() - t0
        if el >= target_s or reps >= max_reps:
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
    threads = host_threads()
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
           "dtype": a.dtype, "data": "synthetic", "config": workload_config(a),
           "cpu_baseline": {"value": val, "unit": "GSamples/s", "cores": threads, "kind": "port",
                            "sample": f"{ns} of the {a.N} signals x {a.n} samples per step, {a.steps} steps (a rate over an O(N) loop of independent "
                                      f"signals); C restatement of the reference loops (Julia unavailable), OpenMP over signals"},
           "e2e": {"value": val, "unit": "GSamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


# ---- NUMA placement of pinned host buffers ----------------------------------------------------------------------------
def gpu_numa_node(props):
    try:
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        v = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        return v if v >= 0 else None
    except Exception:
        return None


def set_mempolicy_preferred(node):
    """set_mempolicy(MPOL_PREFERRED, {node}) for this thread: pinned allocations that follow land on the GPU's NUMA node
    (works even when the cgroup's CPU set does not include that node's cores).  node None: back to the default policy."""
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        SYS_set_mempolicy = 238                              # x86_64
        if node is None:
            return libc.syscall(SYS_set_mempolicy, 0, None, 0) == 0
        mask = (ctypes.c_ulong * 16)()
        mask[node // 64] = 1 << (node % 64)
        return libc.syscall(SYS_set_mempolicy, 1, mask, 16 * 64 + 1) == 0     # MPOL_PREFERRED = 1
    except Exception:
        return False


def node_cpus(node):
    try:
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        return cpus
    except Exception:
        return set()


def host_mem_available():
    avail = None
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        pass
    for p in ("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory/memory.limit_in_bytes"):
        try:
            v = open(p).read().strip()
            if v != "max":
                lim = int(v)
                used = 0
                try:
                    used = int(open(p.replace("memory.max", "memory.current").replace("limit_in_bytes", "usage_in_bytes")).read())
                except Exception:
                    pass
                avail = min(avail, lim - used) if avail is not None else lim - used
        except Exception:
            pass
    return avail


def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)

    import numpy as np
    import torch
    import torch.distributed as dist
    import waveletsext_b200 as wx
    from waveletsext_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    blocks = set(b for b in a.blocks.split(",") if b)
    if a.no_e2e:
        blocks -= {"e2e", "e2e_pipeline"}
    if a.no_cpu:
        blocks.discard("cpu")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def minr(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return float(t.item())

    def timed(fn, steps, warm=1):
        """device ms per call (CUDA events on the launching stream, max over ranks) and the last result"""
        out = None
        for _ in range(warm):
            out = fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return maxr(e0.elapsed_time(e1) / steps), out

    tdt = torch.float64 if a.dtype == "f64" else torch.float32
    npdt = np.float64 if a.dtype == "f64" else np.float32
    esz = 8 if a.dtype == "f64" else 4
    wt = wx.wavelet(a.wavelet)
    peak, peak_src = measured_peak()
    n, N, L = a.n, a.N, a.L
    gen = torch.Generator(device=dev).manual_seed(20242 + rank)
    x = torch.randn((N, n), dtype=tdt, device=dev, generator=gen)
    y = torch.empty((N, L + 1, n), dtype=tdt, device=dev)

    def step():
        wx.dwt._wpd_batch(x, wt, L, y)

    # ---- headline: device-timed wpdall on configs[1] ---------------------------------------------------------------------
    for _ in range(max(a.warmup, 3)):
        step()
    barrier()
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
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = evs[0].elapsed_time(evs[-1])
    per_launch = [evs[i].elapsed_time(evs[i + 1]) for i in range(a.steps)]
    ms_per_step = maxr(total_ms) / a.steps
    value = n * N * world / (ms_per_step * 1e-3) / 1e9
    y_head = y[:4].clone()
    y_tail = y[N - 4:N].clone()

    # ---- e2e: wpdall through the host-buffer C-ABI call (pinned host arrays, copies inside the timed region) ----------------
    orig_aff = os.sched_getaffinity(0)
    numa = {"gpu_node": gpu_numa_node(torch.cuda.get_device_properties(local)), "policy": None}
    if numa["gpu_node"] is not None:
        cpus = node_cpus(numa["gpu_node"]) & set(os.sched_getaffinity(0))
        if cpus:
            try:
                os.sched_setaffinity(0, cpus)                 # the copy-issuing thread and first touch on the GPU's node
                numa["cpus_bound"] = len(cpus)
            except Exception:
                pass
        numa["policy"] = "preferred" if set_mempolicy_preferred(numa["gpu_node"]) else None

    probe = None
    if blocks & {"e2e", "e2e_pipeline"}:
        # measured PCIe ceiling of this rank while every rank copies at once (1 GiB pinned, both directions)
        pb = torch.empty(1 << 27, dtype=torch.float64, pin_memory=True)
        db = torch.empty(1 << 27, dtype=torch.float64, device=dev)
        db.copy_(pb, non_blocking=True); barrier()
        res = {}
        for name, dst, src in (("d2h", pb, db), ("h2d", db, pb)):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record()
            for _ in range(4):
                dst.copy_(src, non_blocking=True)
            e1.record(); torch.cuda.synchronize()
            gbs = 4 * pb.numel() * 8 / (e0.elapsed_time(e1) * 1e-3) / 1e9
            res[name] = {"min_gbs": round(minr(gbs), 1), "max_gbs": round(maxr(gbs), 1)}
        probe = {"what": "1 GiB pinned cudaMemcpyAsync per rank, all ranks at once, min / max over ranks", **res}
        del pb, db

    e2e = None
    if "e2e" in blocks:
        Ne = N
        per_sig = (L + 2) * n * esz
        avail = host_mem_available()
        while Ne > 1024 and avail is not None and Ne * per_sig * world > 0.6 * avail:      # keep the pinned footprint of all ranks safe
            Ne //= 2
        Ne = int(minr(Ne))
        xh = yh = None
        while Ne >= 1024:
            try:
                xh_np, xh = wx.host.pinned_empty((Ne, n), npdt)
                yh_np, yh = wx.host.pinned_empty((Ne, L + 1, n), npdt)
                ok_alloc = 1.0
            except Exception:
                xh = yh = None
                ok_alloc = 0.0
            if minr(ok_alloc) >= 1.0:
                break
            xh = yh = None
            Ne //= 2
        if xh is not None:
            xh.copy_(x[:Ne])
            wx.host.wpdall_host(xh_np, wt, L, out=yh_np, device=local)              # warm-up
            barrier()
            t0 = time.perf_counter()
            for _ in range(a.e2e_steps):
                wx.host.wpdall_host(xh_np, wt, L, out=yh_np, device=local)          # synchronous call
            el = maxr(time.perf_counter() - t0)
            ok = bool(torch.equal(yh[:4].to(dev), y_head)) and (Ne != N or bool(torch.equal(yh[Ne - 4:Ne].to(dev), y_tail)))
            host_l0 = os.environ.get("WX_B200_HOST_LEVEL0", "1")[:1] != "0"         # wx_host.cu: level 0 (= x) is filled from the host copy of x
            e2e = {"value": a.e2e_steps * Ne * n * world / el / 1e9, "unit": "GSamples/s", "h2d_bytes_per_step": Ne * n * esz,
                   "d2h_bytes_per_step": Ne * (L if host_l0 else L + 1) * n * esz, "steps": a.e2e_steps, "signals_per_step": Ne,
                   "matches_device_path": ok, "timer": "host wall clock around the synchronous C-ABI call (wx_wpdall_host), max over ranks",
                   "numa": numa, "pcie_probe": probe,
                   "level0": "rows y[:,0,:] = x are written by host threads from the caller's x while levels 1..L cross PCIe" if host_l0 else "whole table over PCIe",
                   "bound": "PCIe / host DRAM: the packet table is 13x the input, so every step pulls L*n*N*8 bytes back to the host"}
            del xh, yh, xh_np, yh_np

    # ---- e2e_pipeline: x (host) -> wpdall -> bestbasistree(JBB) -> getbasiscoefall -> coefficients (host), one C-ABI call ---
    e2e_pipeline = None
    if "e2e_pipeline" in blocks:
        xh_np, xh = wx.host.pinned_empty((N, n), npdt)
        ch_np, ch = wx.host.pinned_empty((N, n), npdt)
        xh.copy_(x)
        del y
        torch.cuda.empty_cache()
        coef, tree = wx.host.wpd_bestbasis_host(xh_np, wt, L, wx.JBB(), out=ch_np, device=local)       # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.e2e_steps):
            coef, tree = wx.host.wpd_bestbasis_host(xh_np, wt, L, wx.JBB(), out=ch_np, device=local)
        el = maxr(time.perf_counter() - t0)
        # check against the device-resident path on a slab
        yy = wx.wpdall(x[:256], wt, L)
        ok = bool(torch.equal(wx.getbasiscoefall(yy, tree), ch[:256].to(dev)))
        e2e_pipeline = {"value": a.e2e_steps * N * n * world / el / 1e9, "unit": "GSamples/s", "h2d_bytes_per_step": N * n * esz,
                        "d2h_bytes_per_step": N * n * esz, "steps": a.e2e_steps, "signals_per_step": N, "tree_nodes": int(tree.sum()),
                        "matches_device_path": ok, "ms_per_step": el / a.e2e_steps * 1e3,
                        "what": "wx_wpd_bestbasis_host: x (pinned host) -> wpdall -> bestbasistree(JBB, NCCL exchange across ranks) -> "
                                "getbasiscoefall -> coefficients (pinned host); the packet table stays in HBM",
                        "timer": "host wall clock around the synchronous C-ABI call, max over ranks"}
        del xh, ch, xh_np, ch_np, yy
    else:
        del y
    del x
    torch.cuda.empty_cache()
    set_mempolicy_preferred(None)
    try:
        os.sched_setaffinity(0, orig_aff)
    except Exception:
        pass

    # ---- config1: the reference's CPU-runnable shape ------------------------------------------------------------------------
    config1 = None
    if "config1" in blocks and rank == 0:
        n1, N1, L1 = 1024, 1024, 10
        x1 = torch.randn((N1, n1), dtype=torch.float64, device=dev, generator=gen)
        y1 = torch.empty((N1, L1 + 1, n1), dtype=torch.float64, device=dev)
        flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)                 # 256 MB > 126 MB L2
        w1 = wx.wavelet("db4")
        for _ in range(3):
            wx.dwt._wpd_batch(x1, w1, L1, y1)
        ts = []
        for _ in range(20):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); wx.dwt._wpd_batch(x1, w1, L1, y1); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms1 = statistics.median(ts)
        b1 = 8 * n1 * N1 * (L1 + 2)
        config1 = {"workload": "wpdall 1024 signals x 1024 samples, db4, L=10, f64 (BASELINE.json configs[0])", "ms": ms1,
                   "GSamples_per_s": n1 * N1 / (ms1 * 1e-3) / 1e9, "achieved_gbs": b1 / (ms1 * 1e-3) / 1e9, "frac_of_hbm_peak": b1 / (ms1 * 1e-3) / 1e9 / peak,
                   "l2": "256 MB written between launches (the 100 MB working set would otherwise sit in L2)",
                   "note": "12.6 us of HBM traffic: launch-latency sized; one launch of the fused kernel"}
        del x1, y1, flush

    # ---- pipeline: configs[4], the cost-tree exchange over libwx_b200's own NCCL communicator -------------------------------
    pipeline = None
    if "pipeline" in blocks:
        np_, Np, Lp = 1024, 131072, 10
        w5 = wx.wavelet("db4")
        g5 = torch.Generator(device=dev).manual_seed(100 + rank)
        t = torch.arange(np_, device=dev, dtype=torch.float64) / np_
        hs = 4 * torch.sin(4 * np.pi * t) - torch.sign(t - 0.3) - torch.sign(0.72 - t)          # heavisine, utils_dataset.jl:135-137
        idx = (torch.arange(np_, device=dev)[None, :] - 2 * ((torch.arange(Np, device=dev)[:, None] + rank * Np) % np_)) % np_
        xp = torch.gather(hs[None, :].repeat(Np, 1), 1, idx) + 0.5 * torch.randn((Np, np_), dtype=torch.float64, device=dev, generator=g5)
        del idx
        Xw = torch.empty((Np, Lp + 1, np_), dtype=torch.float64, device=dev)
        K = Lp + 1
        st = {}
        lp0 = wx.launch_count()
        st["wpdall"], _ = timed(lambda: wx.dwt._wpd_batch(xp, w5, Lp, Xw), 5)
        st["bestbasistree_JBB"], tj = timed(lambda: wx.bestbasistree(Xw, wx.JBB()), 5)
        st["bestbasistree_LSDB"], tl = timed(lambda: wx.bestbasistree(Xw, wx.LSDB()), 3)
        st["getbasiscoefall+iwptall"], xr = timed(lambda: wx.iwptall(wx.getbasiscoefall(Xw, tj), w5, tj), 5)
        st["iwpdall_fused_gather"], xr2 = timed(lambda: wx.iwpdall(Xw, w5, tj), 5)
        lp1 = wx.launch_count()
        err = maxr(float((xr - xp).abs().max() / xp.abs().max()))
        err2 = maxr(float((xr2 - xp).abs().max() / xp.abs().max()))
        hj, hl = hashlib.sha1(np.packbits(tj).tobytes()).hexdigest()[:16], hashlib.sha1(np.packbits(tl).tobytes()).hexdigest()[:16]
        hashes = [(hj, hl)]
        if world > 1:
            hashes = [None] * world
            dist.all_gather_object(hashes, (hj, hl))
        # the collective on its own: in-place all-reduce of the JBB moment buffer (2*n*K+1 doubles) over the library's communicator
        ar_us = None
        nccl_ver = None
        cm = wx.dist.comm(dev)
        if cm is not None:
            buf = torch.zeros(2 * np_ * K + 1, dtype=torch.float64, device=dev)
            s_ = int(torch.cuda.current_stream().cuda_stream)
            for _ in range(5):
                _lib.call("wx_allreduce", cm, buf.data_ptr(), buf.numel(), 0, 0, s_)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                _lib.call("wx_allreduce", cm, buf.data_ptr(), buf.numel(), 0, 0, s_)
            e1.record(); torch.cuda.synchronize()
            ar_us = maxr(e0.elapsed_time(e1) / 50 * 1e3)
            v = ctypes.c_int()
            _lib.call("wx_nccl_version", ctypes.byref(v))
            nccl_ver = v.value
        tot = Np * world
        nb = int(np.ceil((30.0 * tot) ** 0.2)); npts = (nb + 1) * int(np.ceil(50.0 / nb))
        szK = np_ * K
        pipeline = {"workload": f"{tot} signals x {np_} samples ({Np} per GPU), db4, L={Lp}, f64: wpdall -> bestbasistree(JBB), (LSDB) -> getbasiscoefall + iwptall "
                                f"(BASELINE.json configs[4] at 8 GPUs)",
                    "stage_ms": {k: round(v, 4) for k, v in st.items()},
                    "GSamples_per_s": {k: round(tot * np_ / (v * 1e-3) / 1e9, 2) for k, v in st.items()},
                    "frac_of_hbm_peak": {"wpdall": 8 * np_ * Np * (Lp + 2) / (st["wpdall"] * 1e-3) / 1e9 / peak,
                                         "bestbasistree_JBB": 8 * szK * Np / (st["bestbasistree_JBB"] * 1e-3) / 1e9 / peak,
                                         "bestbasistree_LSDB": 3 * 8 * szK * Np / (st["bestbasistree_LSDB"] * 1e-3) / 1e9 / peak},
                    "collective": {"library": "libwx_b200 wx_comm (NCCL resolved by dlopen)", "nccl_version": nccl_ver, "ranks": world,
                                   "jbb_bytes_allreduced": (2 * szK + 1) * 8 if world > 1 else 0,
                                   "lsdb_bytes_exchanged": ((szK + 2 * 2 * szK * world + 2 * szK + npts * szK + 2 * szK * world) * 8 + 8 * world) if world > 1 else 0,
                                   "jbb_allreduce_latency_us": ar_us},
                    "tree_hashes_jbb_lsdb_per_rank": hashes, "trees_identical_on_all_ranks": len(set(hashes)) == 1,
                    "tree_nodes": {"jbb": int(tj.sum()), "lsdb": int(tl.sum())},
                    "roundtrip_relerr": {"getbasiscoefall+iwptall": err, "iwpdall": err2},
                    "timer": "CUDA events on the launching stream around each stage incl. its NCCL exchange and host selection, max over ranks",
                    "launches": int(lp1 - lp0)}
        del xp, Xw, xr, xr2
        torch.cuda.empty_cache()

    # ---- config3: redundant tables, configs[2] on the per-GPU share -----------------------------------------------------------
    config3 = None
    if "config3" in blocks:
        n3, N3, L3 = 2048, 2048, 8
        w3 = wx.wavelet("db4")
        x3 = torch.randn((N3, n3), dtype=torch.float64, device=dev, generator=gen)
        b3 = 8 * n3 * N3 * (1 << (L3 + 1))
        out3 = {}
        tab = torch.empty((N3, (1 << (L3 + 1)) - 1, n3), dtype=torch.float64, device=dev)
        for name, ac in (("swpdall", False), ("acwpdall", True)):
            ms3, _ = timed(lambda: wx._rwt.forward(ac, "wpd", x3, w3, L3, tab), 5)
            out3[name] = {"ms": round(ms3, 4), "GSamples_per_s": round(n3 * N3 * world / (ms3 * 1e-3) / 1e9, 3),
                          "achieved_gbs": round(b3 / (ms3 * 1e-3) / 1e9, 1), "frac_of_hbm_peak": round(b3 / (ms3 * 1e-3) / 1e9 / peak, 4)}
        config3 = {"workload": f"swpdall / acwpdall {N3 * world} signals x {n3} samples ({N3} per GPU), db4, L={L3}, f64 (BASELINE.json configs[2] at 8 GPUs); "
                               f"shard-local, no collective", "algorithmic_bytes_per_gpu": b3, **out3,
                   "timer": "CUDA events on the launching stream, max over ranks; output table preallocated like the headline's"}
        del x3, tab
        torch.cuda.empty_cache()

    if world > 1:
        wx.dist.destroy_comms()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    alg_bytes = esz * n * N * (L + 2)
    avg_launch_ms = sum(per_launch) / len(per_launch)
    achieved = alg_bytes / (avg_launch_ms * 1e-3) / 1e9
    traffic = traffic_src = None
    try:
        tj_ = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        key = f"wpd1d_{a.dtype}_n{n}_N{N}_L{L}_{a.wavelet}"
        traffic = tj_.get(key, {}).get("dram_bytes_per_launch")
        traffic_src = "static: " + str(tj_.get(key, {}).get("source", "ncu --set full capture under profiles/")) if traffic else None
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": traffic_src, "kernel": "wpd1d_tma_k", "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": avg_launch_ms,
                "peak_source": peak_src,
                "frac_of_write_stream_ceiling": achieved / 7500.0,
                "write_stream_ceiling": "7.5 TB/s: a fill kernel / cudaMemset over the 27.9 GB table on these boxes (profiles/r1_bw_probe.json); "
                                        "the kernel is 93 % writes"}

    cpu = None
    if "cpu" in blocks and world == 1:
        rate, nsig, el = cpu_port_rate(n, L, wt.taps, a.dtype, 1, 10.0)
        threads = host_threads()
        rate_all, nsig_all, el_all = cpu_port_rate(n, L, wt.taps, a.dtype, threads, 6.0, chunk=4096, max_reps=32)
        cpu = {"value": rate, "unit": "GSamples/s", "cores": 1, "kind": "port",
               "sample": f"{nsig} signals x {n} samples ({el:.1f} s), C restatement of the reference loops, 1 thread like the "
                         f"single-threaded reference (Julia unavailable)",
               "all_cores": {"value": rate_all, "unit": "GSamples/s", "cores": threads,
                             "sample": f"{nsig_all} signals x {n} samples ({el_all:.1f} s), OpenMP over signals (what Threads.@threads over eachslice would give)"}}
        if config1 is not None:
            w1 = wx.wavelet("db4")
            r1, s1, t1 = cpu_port_rate(1024, 10, w1.taps, "f64", 1, 4.0, chunk=1024, max_reps=64)
            rA, sA, tA = cpu_port_rate(1024, 10, w1.taps, "f64", threads, 3.0, chunk=1024, max_reps=256)
            config1["cpu_baseline"] = {"value": r1, "unit": "GSamples/s", "cores": 1, "kind": "port", "sample": f"{s1} signals x 1024 samples ({t1:.1f} s)",
                                       "ms_per_pass_of_1024_signals": 1024 * 1024 / r1 / 1e6,
                                       "all_cores": {"value": rA, "unit": "GSamples/s", "cores": threads, "sample": f"{sA} signals x 1024 samples ({tA:.1f} s)"}}

    out = {"metric": "wpdall_gsamples_per_s", "value": value, "unit": "GSamples/s", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
           "config": workload_config(a),
           "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_pipeline": e2e_pipeline, "pipeline": pipeline, "config3": config3,
           "config1": config1, "gpu_launches": int(launches), "clocks": clocks}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
