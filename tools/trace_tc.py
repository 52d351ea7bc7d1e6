#!/usr/bin/env python
"""tools/trace_tc.py — per-CTA phase timeline of the dense tensor-core kernel (developer tool).

    python tools/trace_tc.py M,K,N,s [replicas]

Sets TSG_TC_TRACE=1; prints, per trace slot, SM-clock cycles since the CTA's first stamp
(median / p10 / p90 over CTAs and repetitions).  Slots (tsg_dense_tc.cu, TC_TRACE):
  0 CTA start          1 prologue done (barriers, TMEM alloc, __syncthreads)
  2 first k-block expanded (group 0)      3 first full-barrier arrive (group 0)
  4 group 0 finished its k-blocks         5 MMA warp saw the first full barrier
  6 MMA warp issued the last commit       7 expanders saw tmem_full
  8 accumulators in registers             9 (overlay only) cluster barrier before the push
 10 cluster barrier after the push       11 outputs written      12 TMEM freed, CTA done
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["TSG_TC_TRACE"] = "1"
import torch  # noqa: E402
import __graft_entry__ as ge  # noqa: E402

tsg = ge.load_package()
from ternary_spgemm_b200 import synth  # noqa: E402

M, K, N, s = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1,4096,4096,3").split(","))
NREP = int(sys.argv[2]) if len(sys.argv) > 2 else 8
Wd = synth.device_ternary(K, N, s, 1234)
base = tsg.TCSC.from_device_dense(Wd, K, N, elem_bytes=1)
del Wd
mats = [base] + [base.slice_cols(0, N) for _ in range(NREP - 1)]
X = synth.device_x(M, K, 1)
b = torch.full((N,), 2.0, device="cuda")
Y = torch.empty(M, N, device="cuda")
st = torch.cuda.Stream()
L = tsg.lib()
L.tsg_debug_tc_trace.argtypes = [C.c_void_p, C.c_int]
L.tsg_debug_tc_trace.restype = C.c_int
for it in range(2 * NREP):
    mats[it % NREP].spmm_dev(X, b, Y, M, algo=tsg.ALGO_DENSE_TC, stream=st.cuda_stream)
st.synchronize()
MAXC = 1 << 16
buf = np.zeros(MAXC * 16, np.uint64)
rows = []
for it in range(NREP):
    mats[it % NREP].spmm_dev(X, b, Y, M, algo=tsg.ALGO_DENSE_TC, stream=st.cuda_stream)
    st.synchronize()
    n = L.tsg_debug_tc_trace(buf.ctypes.data, MAXC)
    t = buf[: n * 16].reshape(n, 16).astype(np.int64)
    rows.append(t[:, :13] - t[:, :1])
r = np.concatenate(rows)
print(f"M={M} K={K} N={N} s={s}: {r.shape[0] // NREP} CTAs x {NREP} launches; cycles since CTA start")
for slot in range(13):
    v = r[:, slot]
    v = v[v >= 0]
    if slot and not np.any(v):
        continue
    print(f"slot {slot:2d}: median {np.median(v):9.0f}  p10 {np.percentile(v, 10):9.0f}  p90 {np.percentile(v, 90):9.0f}")
