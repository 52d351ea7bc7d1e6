#!/usr/bin/env python
"""tools/crossover.py — gather vs dense-expand (tcgen05) vs code_gemv crossover: device time per launch over
M x s for a fixed (K, N); writes profiles/crossover_<K>x<N>.json (north-star item 3: the dense
path is kept only where it is measured faster, and the crossover is recorded)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import __graft_entry__ as ge  # noqa: E402

tsg = ge.load_package()
from ternary_spgemm_b200 import synth  # noqa: E402
from tools.sweep import time_algo  # noqa: E402

K, N = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "4096,4096").split(","))
Ms = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "1,2,4,8,16,32,64").split(",")]
Ss = [int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "2,4,8,16").split(",")]
info = tsg.device_info(0)
stream = torch.cuda.Stream()
out = []
for s in Ss:
    Wd = synth.device_ternary(K, N, s, 1234)
    base = tsg.TCSC.from_device_dense(Wd, K, N, elem_bytes=1)
    del Wd
    ds = base.getDataStructureSize()
    reps = int(min(16, max(1, -(-2 * info["l2_bytes"] // ds) + 1)))
    mats = [base] + [base.slice_cols(0, N) for _ in range(reps - 1)]
    b = torch.full((N,), 2.0, device="cuda")
    for M in Ms:
        X = synth.device_x(M, K, 1)
        Ys = [torch.empty(M, N, device="cuda") for _ in range(2)]
        row = {"K": K, "N": N, "s": s, "M": M}
        cands = [("gather", tsg.ALGO_GATHER), ("dense_tc", tsg.ALGO_DENSE_TC)]
        if M <= 2:
            cands.append(("code_gemv", tsg.ALGO_CODE_GEMV))
        for name, algo in cands:
            steps = 100 if M * N * K / s < 2e9 else 20
            row[name + "_us"] = round(time_algo(tsg, torch, mats, X, b, None, Ys, M, algo, steps, stream) * 1e3, 2)
        row["winner"] = min((row[n + "_us"], n) for n, _ in cands)[1]
        row["auto_picks"] = tsg.ALGO_NAMES[base.pick(M)]
        out.append(row)
        print(json.dumps(row), flush=True)
    del mats, base
    torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"crossover_{K}x{N}.json"), "w"), indent=1)
