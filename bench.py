#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the ternary sparse-GEMM hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2] [--algo auto]
    python bench.py --impl reference ...      # the reference's own CPU implementation, same workload
    torchrun --nproc-per-node N bench.py --gpus N ...   # one rank per GPU, N-column sharding

A "step" is one pass of the hot path  Y = X·W + b  over one batch X of the named BASELINE.json
workload (default c2 = configs[1]: M=1 K=4096 N=4096 s=3, the GEMV-style decode shape).
At N GPUs the workload is weak-scaled along the sharded axis: every rank owns an N-column
slice of W with the workload's own column count (global W is K × N·G), X is replicated
(broadcast once from rank 0 over NCCL), each rank writes its own Y slice, no reduction.

One JSON line is printed by rank 0:
  value        whole-job effective GFLOP/s (flops = M·N_total·(1+K/s), readme.md:84-85) with all
               inputs resident in HBM; exactly K launches captured in one CUDA graph, timed with
               CUDA events on the launching stream, max over ranks.  The matrix is rotated over
               enough distinct HBM copies that consecutive launches never find it in L2.
  e2e          same metric through the reference-facing C ABI call with HOST buffers
               (tsg_spmm: H2D X,b -> kernel -> D2H Y inside the timed region; at N>1 additionally
               the NCCL broadcast of X).
  roofline     dominant kernel vs the measured HBM peak: algorithmic bytes are the reference's
               own "Total Input Size" (main.cpp:267) = 4(MK+MN+N)+4(2(N+1)+nnz).
  cpu_baseline the reference's fastest registered function (DoubleUnrolledTCSC_K4_M4,
               main.cpp:125-130) built in place (oracle/_ref) on this box's host, 1 core.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

_OUT = None


def emit(line: dict):
    """The JSON line goes to the process's original stdout; everything else libraries print on
    fd 1 (e.g. NCCL's version banner at N > 1) has been pointed at stderr by main()."""
    out = _OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


METRIC = "ternary spGEMM effective GFLOP/s (flops = M*N*(1+K/s))"
UNIT = "GFLOP/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--algo", default="auto")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the informational other-workload timings")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default): every rank owns the workload's N columns; strong: the "
                         "workload's N columns are split across the ranks (BASELINE config 4)")
    ap.add_argument("--seed", type=int, default=1234)
    return ap.parse_args()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_tensor_peak():
    """dense bf16 TFLOP/s (burst) from MEASURED_PEAKS.json, else the profiling guide's fallback."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p)).get("bf16_tflops", 1590.0))
    return 1590.0


def recorded_traffic(workload: str, algo: str):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        return json.load(open(p)).get(f"{workload}:{algo}")
    return None


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """NVML clocks / throttle reasons sampled while the GPU is busy (the recipe's clocks line)."""

    def __init__(self, index: int):
        self.samples, self.reasons, self._stop, self._thr = [], set(), threading.Event(), None
        self.max_mhz = None
        self.interval = 0.002
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def sample(self):
        if not self.nv:
            return
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            names = {
                "hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown,
                "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown,
                "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap,
            }
            for k, bit in names.items():
                if r & bit:
                    self.reasons.add(k)
        except Exception:
            pass

    def _loop(self):
        while not self._stop.is_set():
            self.sample()
            time.sleep(self.interval)

    def __enter__(self):
        self._stop.clear()
        self._thr = threading.Thread(target=self._loop, daemon=True)
        self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thr.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "samples": len(s), "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------
def cpu_reference_leg(cfg, seed, steps, warmup, budget_s=12.0):
    """Time the reference's fastest registered function on this host (1 core, as written)."""
    import ctypes as C

    import numpy as np
    from oracle import pyoracle

    M, K, N, s = cfg["M"], cfg["K"], cfg["N"], cfg["s"]
    pyoracle.build()
    orc = pyoracle.Oracle()
    kind = "reference" if pyoracle.have_reference() else "port"
    # bounded sample: whole workload when one call fits the budget, else a row subset (x4: the
    # 4-row unroll of DoubleUnrolledTCSC) and, for very wide W, a column subset
    from_flops = lambda m, n: m * n * (1.0 + K / s)
    est_rate = 1.2e9
    Ms, Ns = M, N
    while from_flops(Ms, Ns) / est_rate > budget_s / max(1, (steps or 3)) and Ms > 4:
        Ms = max(4, (Ms // 2) // 4 * 4)
    while K * Ns > (1 << 27) and Ns > 1024:
        Ns //= 2
    W = orc.generate_sparse_matrix(K, Ns, s, seed)
    X = orc.init_x(Ms, K, seed + 1)
    b = np.full(Ns, 2.0, np.float32)
    Y = np.zeros((Ms, Ns), np.float32)
    if kind == "reference":
        ref = pyoracle.Reference()
        h = ref.tcsc_handle(W)
        fn = lambda: ref.lib.ref_double_unrolled_tcsc_k4_m4(h.h, X, b, Y, Ms, Ns, K)
        name = "DoubleUnrolledTCSC<float,4,4> (reference, built in place)"
    else:
        t = orc.tcsc(W)
        fn = lambda: orc.lib.orc_double_unrolled_tcsc_k4_m4(X, *t.arrays, b, Y, Ms, Ns, K)
        name = "DoubleUnrolledTCSC<float,4,4> order (oracle port)"
    for _ in range(max(1, warmup or 1)):
        fn()
    reps = steps
    if reps is None:
        t0 = time.perf_counter()
        fn()
        one = time.perf_counter() - t0
        reps = int(min(200, max(3, budget_s / max(one, 1e-6))))
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    dt = (time.perf_counter() - t0) / reps
    try:
        model = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        model = "unknown"
    return {"value": from_flops(Ms, Ns) / dt / 1e9, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": f"{name}; M={Ms} of {M} rows, N={Ns} of {N} cols, K={K}, s={s}; "
                      f"{reps} calls of {dt * 1e3:.2f} ms; cpu: {model}",
            "ms_per_step": dt * 1e3, "steps": reps}


def run_reference(args, cfg, rank, world):
    if rank != 0:
        return
    steps, warmup = args.steps or 5, args.warmup if args.warmup is not None else 1
    leg = cpu_reference_leg(cfg, args.seed, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": leg["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": leg["steps"], "warmup": warmup,
        "ms_per_step": leg["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, cfg, 1), "M": cfg["M"], "K": cfg["K"],
                   "N": cfg["N"], "s": cfg["s"], "note": "CPU, single thread as written; at N>1 "
                   "the reference has no multi-device path: rank 0 runs one instance"},
        "cpu_baseline": {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": leg["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_name(key, cfg, world):
    fmt = "fp32 packed-value CSC" if cfg.get("fmt") == "pcsc" else "fp32 TCSC"
    if "N_full" in cfg:   # strong scaling: the workload's own N, split across the ranks
        return (f"{key}: M={cfg['M']} K={cfg['K']} N={cfg['N_full']} (N-sharded over {world} GPUs, "
                f"~{cfg['N']} cols/GPU) s={cfg['s']} {fmt}" + (" +bias+PReLU" if cfg.get("prelu") else " +bias"))
    return (f"{key}: M={cfg['M']} K={cfg['K']} N={cfg['N']}"
            + (f"x{world} (N-sharded, {cfg['N']} cols/GPU)" if world > 1 else "")
            + f" s={cfg['s']} {fmt}" + (" +bias+PReLU" if cfg.get("prelu") else " +bias"))


# ------------------------------------------------------------------------------------------------
class Workload:
    """One BASELINE.json shape resident on this rank's GPU: the rank's W shard (rotated over
    enough HBM copies to defeat L2), X (replicated), b, alpha, output buffers."""

    def __init__(self, tsg, synth, torch, cfg, seed, rank, dev, algo):
        self.tsg, self.torch, self.cfg, self.algo = tsg, torch, cfg, algo
        M, K, N, s = cfg["M"], cfg["K"], cfg["N"], cfg["s"]
        self.M, self.K, self.N, self.s, self.prelu = M, K, N, s, bool(cfg.get("prelu"))
        info = tsg.device_info(dev.index)
        Wd = synth.device_ternary(K, N, s, seed + 7919 * rank, device=dev)
        self.fmt = cfg.get("fmt", "tcsc")
        io_bytes = 4 * (M * K + M * N + N + (N if self.prelu else 0))
        if self.fmt == "pcsc":
            # BASELINE config 4 names the packed-value CSC format: the handle is built from W by the
            # device-side packed builder and the call goes through tsg_pcsc_spmm_dev; the algorithmic
            # bytes are the packed structure's: 4(N+1) + 4 nnz + ceil(nnz/5)
            base = tsg.PackedCSC.from_device_dense(Wd, K, N, elem_bytes=1)
            self.nnz = base.sizes[0]
            ds = base.getDataStructureSize()
            self.bytes_per_launch = io_bytes + ds
            self.bytes_model = "4(MK+MN+N) + 4(N+1) + 4 nnz + ceil(nnz/5)  (packed-value CSC, DESIGN.md §6)"
        else:
            base = tsg.TCSC.from_device_dense(Wd, K, N, elem_bytes=1)
            self.nnz = sum(base.nnz)
            self.bytes_per_launch = base.spmm_bytes(M, self.prelu)
            ds = base.getDataStructureSize()
            self.bytes_model = "4(MK+MN+N[+N alpha]) + 4(2(N+1)+nnz)  (main.cpp:267)"
        replicas = int(min(64, max(1, -(-2 * info["l2_bytes"] // max(ds, 1)) + 1)))
        if replicas * ds > 30e9:
            replicas = max(1, int(30e9 // ds))
        if self.fmt == "pcsc":
            self.mats = [base] + [tsg.PackedCSC.from_device_dense(Wd, K, N, elem_bytes=1) for _ in range(replicas - 1)]
        else:
            self.mats = [base] + [base.slice_cols(0, N) for _ in range(replicas - 1)]
        del Wd
        self.replicas = replicas
        self.l2_policy = (f"rotating {replicas} HBM copies of W ({replicas * ds / 1e6:.0f} MB > L2 "
                          f"{info['l2_bytes'] / 1e6:.0f} MB)") if replicas * ds > info["l2_bytes"] \
            else f"W ({ds / 1e6:.0f} MB) x {replicas} copies"
        self.X = synth.device_x(M, K, seed + 1, device=dev)
        self.b = torch.full((N,), 2.0, device=dev)
        self.alpha = torch.full((N,), 0.1, device=dev) if self.prelu else None
        self.Ys = [torch.empty(M, N, device=dev) for _ in range(min(replicas, 4))]
        if algo != tsg.ALGO_AUTO:
            self.resolved = algo
        else:
            self.resolved = base.pick(M)
        self.l2_warm = False

    def step(self, i, stream):
        if self.l2_warm:
            i = 0   # always the same copy of W: it stays in L2 when it fits
        self.mats[i % self.replicas].spmm_dev(self.X, self.b, self.Ys[i % len(self.Ys)], self.M,
                                              alpha=self.alpha, algo=self.algo,
                                              stream=stream.cuda_stream)

    def time_graph(self, steps, warmup, stream, barrier, sampler=None):
        """Exactly `steps` launches captured in ONE CUDA graph, timed with CUDA events on the
        launching stream.  Returns (ms_total, kernels launched per replay)."""
        torch, tsg = self.torch, self.tsg
        with torch.cuda.stream(stream):
            for i in range(max(warmup, self.replicas)):
                self.step(i, stream)
            stream.synchronize()
            graph = torch.cuda.CUDAGraph()
            l0 = tsg.launch_count()
            with torch.cuda.graph(graph, stream=stream):
                for i in range(steps):
                    self.step(i, stream)
            launches = tsg.launch_count() - l0
            graph.replay()                       # untimed replay (graph upload)
            stream.synchronize()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if sampler is not None:
                sampler.__enter__()
            e0.record(stream)
            graph.replay()
            e1.record(stream)
            while not e1.query():
                if sampler is not None:
                    sampler.sample()
            if sampler is not None:
                sampler.__exit__()
            barrier()
            return e0.elapsed_time(e1), launches


def run_ours(args, cfg, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge
    tsg = ge.load_package()
    from ternary_spgemm_b200 import shard, synth

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if args.scaling == "strong" and world > 1:   # this rank's share of the workload's columns
        lo, hi = tsg.shard_columns(cfg["N"], world, rank)
        cfg = dict(cfg, N=hi - lo, N_full=cfg["N"])
    M, K, N, s, prelu = cfg["M"], cfg["K"], cfg["N"], cfg["s"], bool(cfg.get("prelu"))
    n_total = cfg.get("N_full", N * world)
    algo = {v: k for k, v in tsg.ALGO_NAMES.items()}[args.algo]
    steps = args.steps if args.steps is not None else (2000 if M * N * K / s < 5e8 else 200)
    warmup = max(3, args.warmup if args.warmup is not None else 20)
    info = tsg.device_info(local_rank)

    # this rank's shard: columns [rank*N, (rank+1)*N) of the global K x (N*world) weight
    wl = Workload(tsg, synth, torch, cfg, args.seed, rank, dev, algo)
    # rank 0 draws X, broadcast once (NCCL) — the only collective on the path
    shard.broadcast_x(wl.X, src=0)
    stream = torch.cuda.Stream(device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    ms_total, launches_per_replay = wl.time_graph(steps, warmup, stream, barrier, sampler)
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / steps
    total_flops = synth.flops(M, n_total, K, s)
    value = total_flops / (ms_step * 1e-3) / 1e9
    # the same launches against ONE copy of W (L2-resident when it fits): reported, not the headline
    wl.l2_warm = True
    ms_warm, _ = wl.time_graph(steps, 3, stream, barrier)
    wl.l2_warm = False
    t = torch.tensor([ms_warm], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_warm = float(t.item()) / steps
    kernel_name = tsg.ALGO_NAMES[wl.resolved]
    run_meta = {"nnz_per_gpu": wl.nnz, "kernel": kernel_name, "l2": wl.l2_policy}
    bytes_model = wl.bytes_model
    run_roof = {"bytes_per_launch": wl.bytes_per_launch,
                "traffic": recorded_traffic(args.workload, kernel_name)}

    # ---- e2e: the reference-facing call with HOST (pinned) buffers ------------------------------
    X, b, alpha, Ys, mats, replicas = wl.X, wl.b, wl.alpha, wl.Ys, wl.mats, wl.replicas
    Xh = X.cpu().pin_memory()
    bh = b.cpu().pin_memory()
    ah = alpha.cpu().pin_memory() if prelu else None
    Yh = torch.empty(M, N).pin_memory()
    e2e_steps = min(steps, 2000 if 4 * (M * K + M * N + N) < (1 << 20) else 200)
    Xd2 = torch.empty_like(X)
    # N > 1: X is not broadcast — rank 0 owns it in symmetric memory and the other ranks' kernels
    # read it over NVLink in place (shard.PeerX); NCCL broadcast only if that is unavailable
    # N > 1, X lives on the HOST of rank 0 (the reference's comp_func contract).  Default: rank 0
    # publishes it through host shared memory and every rank's host-pointer call pulls it over its own
    # PCIe link (shard.HostSharedX: no inter-GPU traffic, no device-side barrier).  TSG_BENCH_X=peer:
    # rank 0 copies X into symmetric memory and the other ranks' kernels read it over NVLink in place
    # (shard.PeerX); TSG_BENCH_X=nccl: H2D on rank 0 + NCCL broadcast.
    peer, hostx, x_transport = None, None, "none (single GPU)"
    if world > 1:
        mode = os.environ.get("TSG_BENCH_X", "nccl" if os.environ.get("TSG_BENCH_NCCL_X") else "shm")
        try:
            if mode == "shm":
                hostx = shard.HostSharedX(M, K)
                for buf in hostx.buffers():
                    if rank == 0:
                        buf[...] = Xh.numpy()     # the producer's X, written in place (not part of a step)
                x_transport = ("rank 0 publishes X in host shared memory (registered with CUDA), every rank's "
                               "tsg_spmm pulls it over its own PCIe link (no collective, no NVLink)")
            elif mode == "peer":
                peer = shard.PeerX(M, K, dev)
                x_transport = "peer reads of rank 0's symmetric-memory X over NVLink inside the kernel (no collective)"
            else:
                x_transport = "NCCL broadcast from rank 0"
        except Exception as e:  # shared / symmetric memory not usable on this box
            peer, hostx = None, None
            x_transport = f"NCCL broadcast from rank 0 ({type(e).__name__}: {str(e)[:80]})"
        ok = torch.tensor([1 if (peer is not None or hostx is not None or mode == "nccl") else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)    # every rank must take the same path
        if int(ok.item()) == 0:
            peer, hostx, x_transport = None, None, "NCCL broadcast from rank 0 (preferred transport failed on a rank)"

    xp, bp, ap, yp = Xh.data_ptr(), bh.data_ptr(), (ah.data_ptr() if prelu else None), Yh.data_ptr()
    hx_ptr = [b_.ctypes.data for b_ in hostx.buffers()] if hostx is not None else None

    def e2e_step(i):
        m = mats[i % replicas]
        if world == 1:
            m.spmm_host_ptr(xp, bp, ap, yp, M, algo=algo)
        elif hostx is not None:
            x = hostx.next(None)                  # rank 0 publishes the step, the others wait for it
            m.spmm_host_ptr(hx_ptr[hostx.step & 1], bp, ap, yp, M, algo=algo)
            hostx.done()
        else:
            with torch.cuda.stream(stream):
                if peer is not None:
                    xin = peer.stage(Xh)
                else:
                    if rank == 0:
                        Xd2.copy_(Xh, non_blocking=True)
                    shard.broadcast_x(Xd2, src=0)
                    xin = Xd2
                # Y goes straight to the pinned host buffer (mapped under UVA): no separate D2H operation
                m.spmm_dev(xin, b, yp, M, alpha=alpha, algo=algo, stream=stream.cuda_stream)
            stream.synchronize()

    for i in range(max(warmup, replicas, 100 if e2e_steps > 200 else 5)):
        e2e_step(i)
    barrier()
    with sampler:
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            e2e_step(i)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
    barrier()
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item()) / e2e_steps * 1e3
    e2e_value = total_flops / (e2e_ms * 1e-3) / 1e9
    # sanity: the e2e result must equal the device-path result
    torch.cuda.synchronize(dev)
    with torch.cuda.stream(stream):
        mats[0].spmm_dev(X, b, Ys[0], M, alpha=alpha, algo=algo, stream=stream.cuda_stream)
    stream.synchronize()
    e2e_step(0)
    if not torch.equal(Ys[0].cpu(), Yh):
        raise RuntimeError("e2e result differs from device-path result")
    if hostx is not None:
        barrier()
        hostx.close()

    # ---- the other BASELINE shapes, device-timed the same way (N=1 only; informational) --------
    others = []
    if world == 1 and not args.no_others:
        peak_o, _ = measured_peak()
        for key in ("c1", "c3", "c4", "c5a", "c5b"):
            if key == args.workload:
                continue
            try:
                del wl
                torch.cuda.empty_cache()
                ocfg = synth.CONFIGS[key]
                wl = Workload(tsg, synth, torch, ocfg, args.seed, 0, dev, tsg.ALGO_AUTO)
                osteps = 200 if ocfg["M"] * ocfg["N"] * ocfg["K"] / ocfg["s"] < 5e8 else 20
                oms, _ = wl.time_graph(osteps, 3, stream, barrier)
                ous = oms / osteps * 1e3
                others.append({
                    "workload": workload_name(key, ocfg, 1), "kernel": tsg.ALGO_NAMES[wl.resolved],
                    "us_per_launch": round(ous, 3),
                    "gflops": round(synth.flops(ocfg["M"], ocfg["N"], ocfg["K"], ocfg["s"]) / ous / 1e3, 1),
                    "hbm_frac_tcsc_bytes": round(wl.bytes_per_launch / (ous * 1e-6) / 1e9 / peak_o, 4),
                    "dense_tflops": round(2.0 * ocfg["M"] * ocfg["N"] * ocfg["K"] / ous / 1e6, 1)
                    if tsg.ALGO_NAMES[wl.resolved] == "dense_tc" else None,
                    "tensor_frac_of_measured_bf16_peak": round(
                        2.0 * ocfg["M"] * ocfg["N"] * ocfg["K"] / ous / 1e6 / measured_tensor_peak(), 4)
                    if tsg.ALGO_NAMES[wl.resolved] == "dense_tc" else None,
                    "l2": wl.l2_policy})
            except Exception as e:  # informational only
                others.append({"workload": key, "error": f"{type(e).__name__}: {str(e)[:120]}"})
        wl = None
        torch.cuda.empty_cache()

    if rank != 0:
        return
    peak, peak_src = measured_peak()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(run_meta, **{"workload": workload_name(args.workload, cfg, world), "M": M, "K": K,
                   "N_per_gpu": N, "N_total": n_total, "s": s,
                   "timing": "one CUDA graph of `steps` launches, CUDA events on the launch stream, "
                             "max over ranks; X resident (broadcast once before the timed region)",
                   "parallelism": f"N-column sharding x{world}, no data-path collective"}),
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "steps": e2e_steps,
                "h2d_bytes_per_step": 4 * (M * K + N + (N if prelu else 0)),
                "d2h_bytes_per_step": 4 * M * N,
                "path": "tsg_spmm(host ptrs), synchronous: inputs -> one staging block -> ONE H2D copy (inline in the "
                        "command stream up to 64 KB), kernel stores Y to mapped host memory (calls < 1 MB); "
                        "cudaMemcpyAsync H2D/D2H otherwise"
                        + ("; N > 1: " + x_transport if world > 1 else "")},
        "l2_warm": {"us_per_launch": ms_warm * 1e3, "value": total_flops / (ms_warm * 1e-3) / 1e9, "unit": UNIT,
                    "note": "same launches, one copy of W (stays in L2 when it fits); informational"},
        "gpu_launches": int(launches_per_replay),
        "clocks": sampler.summary(),
        "device": info["name"],
    }
    line["roofline"] = dict(run_roof, **{"bound": "hbm", "peak": peak, "unit": "GB/s",
                            "us_per_launch": ms_step * 1e3, "peak_source": peak_src,
                            "bytes_model": bytes_model})
    line["roofline"]["achieved"] = line["roofline"]["bytes_per_launch"] / (ms_step * 1e-3) / 1e9
    line["roofline"]["frac"] = line["roofline"]["achieved"] / peak
    if kernel_name in ("code_gemv", "dense_tc"):
        # transparency: these kernels stream the 2-bit codes (K*N/4 bytes), not the index arrays the
        # algorithmic byte count above is defined on; what they actually move per launch is:
        stream = (K * N) // 4 + 4 * (M * K + N + (N if prelu else 0) + M * N)
        tensor_bound = kernel_name == "dense_tc" and M >= 128
        line["roofline"]["kernel_stream"] = {
            "bytes_per_launch": stream, "achieved": stream / (ms_step * 1e-3) / 1e9, "unit": "GB/s",
            "frac": stream / (ms_step * 1e-3) / 1e9 / peak,
            "note": "bytes the kernel itself streams (2-bit code stream + X + b + Y); "
                    + ("at this M the kernel is bound by the tensor pipe (see roofline_tensor)" if tensor_bound else
                       "at this size the kernel is bound by launch + first-HBM latency and "
                       + ("the FMA pipe" if kernel_name == "code_gemv" else "the tensor core's A-operand feed")
                       + ", not by HBM")}
        if tensor_bound:
            # dense-equivalent MMA work: 2*M*K*N per term of X (integer-valued X = one fp16 term);
            # a `steps`-long graph of 0.3-0.7 ms launches runs under the power cap: sustained peak
            pj = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
                os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
            long_run = ms_step * steps > 50.0
            tpeak = float(pj.get("bf16_tflops_sustained" if long_run else "bf16_tflops", 1414.0 if long_run else 1590.0))
            tfl = 2.0 * M * K * N / (ms_step * 1e-3) / 1e12
            line["roofline_tensor"] = {
                "bound": "tensor", "achieved": tfl, "peak": tpeak, "unit": "TFLOP/s", "frac": tfl / tpeak,
                "peak_source": ("measured (MEASURED_PEAKS.json " + ("bf16_tflops_sustained" if long_run else "bf16_tflops") + ")")
                if pj else "fallback (B200_PROFILING.md)",
                "flops_model": "2*M*K*N_per_gpu x 1 term (integer-valued X is exact in one fp16 term)",
                "traffic": None}
    if others:
        line["other_workloads"] = others
    if world == 1 and not args.no_cpu_baseline:
        try:
            leg = cpu_reference_leg(cfg, args.seed, None, 1)
            line["cpu_baseline"] = {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:  # the GPU numbers stand on their own
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "unavailable",
                                    "sample": f"{type(e).__name__}: {e}"}
    emit(line)


def main():
    global _OUT
    args = parse_args()
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    ge.load_package()
    from ternary_spgemm_b200 import synth
    keys = args.workload.split(",")   # several workloads: one JSON line each (developer use)
    for key in keys:
        if key not in synth.CONFIGS:
            raise SystemExit(f"unknown workload {key}; have {sorted(synth.CONFIGS)}")
    if args.impl == "reference":
        for key in keys:
            args.workload = key
            run_reference(args, synth.CONFIGS[key], rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        import torch
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        for key in keys:
            args.workload = key
            run_ours(args, synth.CONFIGS[key], rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
