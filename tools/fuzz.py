#!/usr/bin/env python
"""tools/fuzz.py — random shapes through every kernel (developer tool, GPU).

    python tools/fuzz.py [cases] [seed]

For each random (M, K, N, s): W drawn on the device, integer-valued X; every kernel must be
bit-identical to the reference-order kernel (gather_seq) — with and without PReLU — and AUTO's
pick is timed against the alternatives; cases where AUTO is more than 1.5x off the best are listed.
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import __graft_entry__ as ge  # noqa: E402

tsg = ge.load_package()
from ternary_spgemm_b200 import synth  # noqa: E402
from tools.sweep import time_algo  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rnd = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
stream = torch.cuda.Stream()
bad, slow = [], []
for case in range(cases):
    M = rnd.choice([1, 1, 2, 3, 4, 7, 8, 15, 16, 17, 31, 32, 33, 60, 64, 65, 100, 128, 129, 200, 256, 300, 513])
    K = rnd.choice([37, 64, 100, 256, 500, 1024, 2000, 4096, 5000, 8192])
    N = rnd.choice([29, 128, 130, 500, 1024, 3000, 4096, 10000, 14336])
    s = rnd.choice([2, 3, 4, 8, 16, 32])
    if M * K * N > 4e10 or N // s < 2:
        continue
    Wd = synth.device_ternary(K, N, s, 1000 + case)
    t = tsg.TCSC.from_device_dense(Wd, K, N, elem_bytes=1)
    del Wd
    X = synth.device_x(M, K, 2000 + case)
    b = torch.randint(-8, 9, (N,), device="cuda").float()
    al = torch.full((N,), 0.25, device="cuda")
    Yref, Y = torch.empty(M, N, device="cuda"), torch.empty(M, N, device="cuda")
    row = {"M": M, "K": K, "N": N, "s": s, "auto": tsg.ALGO_NAMES[t.pick(M)]}
    for alpha in (None, al):
        t.spmm_dev(X, b, Yref, M, alpha=alpha, algo=tsg.ALGO_GATHER_SEQ)
        for name, algo in (("gather", tsg.ALGO_GATHER), ("dense_tc", tsg.ALGO_DENSE_TC),
                           ("code_gemv", tsg.ALGO_CODE_GEMV), ("auto", tsg.ALGO_AUTO)):
            if name == "code_gemv" and M > 8:
                continue
            try:
                Y.fill_(float("nan"))
                t.spmm_dev(X, b, Y, M, alpha=alpha, algo=algo)
                torch.cuda.synchronize()
            except tsg.TsgError as e:
                if e.status == -5:
                    continue
                raise
            if not torch.equal(Y, Yref):
                bad.append(dict(row, algo=name, prelu=alpha is not None,
                                maxdiff=float((Y - Yref).abs().nan_to_num(1e30).max())))
    times = {}
    for name, algo in (("gather", tsg.ALGO_GATHER), ("dense_tc", tsg.ALGO_DENSE_TC), ("code_gemv", tsg.ALGO_CODE_GEMV)):
        if name == "code_gemv" and M > 2:
            continue
        try:
            times[name] = round(time_algo(tsg, torch, [t], X, b, None, [Y], M, algo, 30, stream) * 1e3, 2)
        except tsg.TsgError:
            pass
    row["us"] = times
    best = min(times, key=times.get)
    if row["auto"] in times and times[row["auto"]] > 1.5 * times[best]:
        slow.append(row)
    print(json.dumps(row), flush=True)
    del t
    torch.cuda.empty_cache()
print(json.dumps({"mismatches": bad, "auto_more_than_1.5x_off": slow}, indent=1))
sys.exit(1 if bad else 0)
