#!/usr/bin/env python
"""tools/sweep.py — device-time every kernel on the BASELINE.json shapes (developer tool).

    python tools/sweep.py [--workloads c1,c2,...] [--algos gather,bitplane,dense_tc] [--steps 200]

Same timing method as bench.py (one CUDA graph of `steps` launches, CUDA events, matrix rotated
over > 2×L2 of HBM copies); prints one JSON line per (workload, algo) with GFLOP/s, µs/launch
and the fraction of the measured HBM roofline (reference byte model, main.cpp:267).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def time_algo(tsg, torch, mats, X, b, alpha, Ys, M, algo, steps, stream):
    def step(i):
        mats[i % len(mats)].spmm_dev(X, b, Ys[i % len(Ys)], M, alpha=alpha, algo=algo,
                                     stream=stream.cuda_stream)
    with torch.cuda.stream(stream):
        for i in range(max(3, len(mats))):
            step(i)
        stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            for i in range(steps):
                step(i)
        g.replay()
        stream.synchronize()
        best = 1e30
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            g.replay()
            e1.record(stream)
            stream.synchronize()
            best = min(best, e0.elapsed_time(e1) / steps)
    return best  # ms per launch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="c1,c2,c3,c5a")
    ap.add_argument("--algos", default="gather,dense_tc,code_gemv")
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--shape", default=None, help="M,K,N,s ad-hoc shape instead of --workloads")
    ap.add_argument("--x", default="int", help="int (initX regime), real (U(-1,1) fp32), bf16 (bf16-valued fp32); comma list")
    args = ap.parse_args()
    import torch
    import __graft_entry__ as ge
    tsg = ge.load_package()
    from ternary_spgemm_b200 import synth
    peak = 6536.7
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    info = tsg.device_info(0)
    names = {v: k for k, v in tsg.ALGO_NAMES.items()}
    if args.shape:
        M, K, N, s = (int(v) for v in args.shape.split(","))
        todo = {f"M{M}K{K}N{N}s{s}": dict(M=M, K=K, N=N, s=s, prelu=False)}
    else:
        todo = {k: synth.CONFIGS[k] for k in args.workloads.split(",")}
    stream = torch.cuda.Stream()
    for key, cfg in todo.items():
        M, K, N, s, prelu = cfg["M"], cfg["K"], cfg["N"], cfg["s"], bool(cfg.get("prelu"))
        Wd = synth.device_ternary(K, N, s, 1234)
        base = tsg.TCSC.from_device_dense(Wd, K, N, elem_bytes=1)
        del Wd
        ds = base.getDataStructureSize()
        reps = int(min(32, max(1, -(-2 * info["l2_bytes"] // ds) + 1)))
        mats = [base] + [base.slice_cols(0, N) for _ in range(reps - 1)]
        b = torch.full((N,), 2.0, device="cuda")
        alpha = torch.full((N,), 0.1, device="cuda") if prelu else None
        Ys = [torch.empty(M, N, device="cuda") for _ in range(2)]
        byts = base.spmm_bytes(M, prelu)
        ref = None
        for an, xk in ((a, x) for x in args.x.split(",") for a in args.algos.split(",")):
            algo = names[an]
            X = synth.device_x(M, K, 1, integer=(xk == "int"))
            if xk == "bf16":
                X = X.to(torch.bfloat16).to(torch.float32)
            steps = args.steps if M * N * K / s < 2e9 else max(10, args.steps // 10)
            try:
                ms = time_algo(tsg, torch, mats, X, b, alpha, Ys, M, algo, steps, stream)
            except tsg.TsgError as e:
                print(json.dumps({"workload": key, "algo": an, "error": str(e)[:120]}), flush=True)
                continue
            y = Ys[0].clone() if steps % 2 == 1 else Ys[(steps - 1) % 2].clone()
            if ref is None:
                ref = y
            print(json.dumps({
                "workload": key, "M": M, "K": K, "N": N, "s": s, "algo": an, "x": xk,
                "us": round(ms * 1e3, 3), "gflops": round(synth.flops(M, N, K, s) / ms / 1e6, 1),
                "tcsc_bytes": byts, "hbm_frac_tcsc_model": round(byts / (ms * 1e-3) / 1e9 / peak, 4),
                "replicas": reps, "agrees_with_first": bool(torch.equal(y, ref)) if xk == args.x.split(",")[0] else None}), flush=True)
        del mats, base
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
