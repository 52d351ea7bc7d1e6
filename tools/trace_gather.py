#!/usr/bin/env python
"""tools/trace_gather.py — per-CTA phase timeline of the gather kernel (developer tool).
Run with TSG_GATHER_TRACE=1; prints phase durations (ns, %globaltimer) averaged over CTAs."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["TSG_GATHER_TRACE"] = "1"
import torch  # noqa: E402
import __graft_entry__ as ge  # noqa: E402

tsg = ge.load_package()
from ternary_spgemm_b200 import synth  # noqa: E402

M, K, N, s = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1,4096,4096,3").split(","))
Wd = synth.device_ternary(K, N, s, 1234)
base = tsg.TCSC.from_device_dense(Wd, K, N, elem_bytes=1)
NREP = int(sys.argv[2]) if len(sys.argv) > 2 else 13
mats = [base] + [base.slice_cols(0, N) for _ in range(NREP - 1)]
X = synth.device_x(M, K, 1)
b = torch.full((N,), 2.0, device="cuda")
Y = torch.empty(M, N, device="cuda")
st = torch.cuda.Stream()
names = ["start->staged", "staged->sync1", "sync1->streamed", "streamed->sync2", "sync2->end"]
L = tsg.lib()
L.tsg_debug_gather_trace.argtypes = [C.c_void_p, C.c_int]
for it in range(30):
    mats[it % NREP].spmm_dev(X, b, Y, M, algo=tsg.ALGO_GATHER, stream=st.cuda_stream)
    st.synchronize()
buf = np.zeros(148 * 16 * 8, np.uint64)
rows = []
for it in range(13):
    mats[it % NREP].spmm_dev(X, b, Y, M, algo=tsg.ALGO_GATHER, stream=st.cuda_stream)
    st.synchronize()
    L.tsg_debug_gather_trace(buf.ctypes.data, buf.size)
    t = buf.reshape(-1, 8)[:148 * 16, :6].astype(np.int64)
    t0 = t[:, 0].min()
    rows.append((t - t0))
r = np.stack(rows)            # iters x ctas x 7
print("CTA start spread (ns): mean", r[:, :, 0].mean().round(0), "max", r[:, :, 0].max(axis=1).mean().round(0))
d = np.diff(r, axis=2)
for i, n in enumerate(names):
    print(f"{n:22s} mean {d[:, :, i].mean():8.0f} ns   max-over-CTAs {d[:, :, i].max(axis=1).mean():8.0f} ns")
print("kernel span (first start -> last end):", (r[:, :, 5].max(axis=1)).mean().round(0), "ns")
for slot in range(6):
    print("slot", slot, "abs time: mean", r[:, :, slot].mean().round(0), "min", r[:, :, slot].min(axis=1).mean().round(0), "max", r[:, :, slot].max(axis=1).mean().round(0))
per_warp = r[:, :, 3].reshape(r.shape[0], 148, 16).mean(axis=(0, 1))
print("streamed-done time by warp id (ns):", per_warp.round(0).tolist())
per_cta = r[:, :, 3].reshape(r.shape[0], 148, 16).max(axis=2).mean(axis=0)
print("streamed-done by CTA: min", per_cta.min().round(0), "median", np.median(per_cta).round(0), "max", per_cta.max().round(0))
