#!/usr/bin/env python
"""tools/fuzz.py — random shapes through every kernel (developer tool, GPU).

    python tools/fuzz.py [cases] [seed]

For each random (M, K, N, s): W drawn on the device; integer-valued X: every kernel must be
bit-identical to the reference-order kernel (gather_seq), with and without PReLU; real-valued X
(random scale, sometimes mixed with integer tiles): within 4e-6 of the forward scale, exact and
fast split (the tensor core's truncating fp32 accumulator reaches 2.6e-6 when small-magnitude tiles
follow integer tiles of ±512 at K = 8192; uniform data stays below 4e-7); AUTO's pick is timed against the alternatives and cases where it is more than 1.5x
off the best are listed.
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
    # real-valued X (random scale): every kernel against the reference-order kernel, relative to the
    # forward scale sum|x||w| + |b|; the tensor path with the exact split and with the opt-in fast split
    Xr = synth.device_x(M, K, 3000 + case, integer=False) * (10.0 ** rnd.choice([-3, -1, 0, 0, 2, 4]))
    if rnd.random() < 0.3:
        Xr[:, : K // 2] = synth.device_x(M, K, 4000 + case)[:, : K // 2]      # integer tiles next to real ones
    t.spmm_dev(Xr, b, Yref, M, algo=tsg.ALGO_GATHER_SEQ)
    Wabs = (t_dense := torch.from_numpy(t.getVectorRepresentation()).cuda().abs().float())
    scale = Xr.abs() @ Wabs + b.abs()[None, :]
    floor = Wabs.sum(dim=0)[None, :] * 2.0 ** -25      # the fast split's absolute term: 2^-25 per non-zero
    del t_dense, Wabs
    for fast in (False, True):
        tsg.set_fast_split(fast)
        for name, algo in (("gather", tsg.ALGO_GATHER), ("dense_tc", tsg.ALGO_DENSE_TC),
                           ("code_gemv", tsg.ALGO_CODE_GEMV), ("auto", tsg.ALGO_AUTO)):
            if (name == "code_gemv" and M > 8) or (fast and name in ("gather", "code_gemv")):
                continue
            try:
                Y.fill_(float("nan"))
                t.spmm_dev(Xr, b, Y, M, algo=algo)
                torch.cuda.synchronize()
            except tsg.TsgError as e:
                if e.status == -5:
                    continue
                raise
            err = (Y - Yref).abs() - (floor if fast else 0.0)   # (contract of tsg_set_fast_split, include/tsg.h)
            rel = float((err.clamp_min(0.0) / scale.clamp_min(1e-30)).nan_to_num(1e30).max())   # (empty columns with b = 0: scale 0, error 0)
            if rel > 4e-6:
                bad.append(dict(row, algo=name, real=True, fast=fast, rel=rel))
    tsg.set_fast_split(False)
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
