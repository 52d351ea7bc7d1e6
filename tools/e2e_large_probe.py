#!/usr/bin/env python
"""tools/e2e_large_probe.py — per-call wall time of the large (pipelined) host-pointer call (developer tool).

    python tools/e2e_large_probe.py c3 [real|int] [handles]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import __graft_entry__ as ge  # noqa: E402

tsg = ge.load_package()
from ternary_spgemm_b200 import synth  # noqa: E402

key = sys.argv[1] if len(sys.argv) > 1 else "c3"
regime = sys.argv[2] if len(sys.argv) > 2 else "real"
nh = int(sys.argv[3]) if len(sys.argv) > 3 else 1
cfg = synth.CONFIGS[key]
M, K, N, s = cfg["M"], cfg["K"], cfg["N"], cfg["s"]
Wd = synth.device_ternary(K, N, s, 1234)
mats = [tsg.TCSC.from_device_dense(Wd, K, N, elem_bytes=1) for _ in range(nh)]
Xd = synth.device_x(M, K, 1, integer=(regime == "int"))
Xh = Xd.cpu().pin_memory()
bh = torch.full((N,), 2.0).pin_memory()
ah = torch.full((N,), 0.1).pin_memory() if cfg.get("prelu") else None
Yh = torch.empty(M, N).pin_memory()
xp, bp, ap, yp = Xh.data_ptr(), bh.data_ptr(), (ah.data_ptr() if ah is not None else None), Yh.data_ptr()
# the device path first, like bench.py
bd = bh.cuda()
Yd = torch.empty(M, N, device="cuda")
st = torch.cuda.Stream()
for m in mats:
    m.spmm_dev(Xd, bd, Yd, M, stream=st.cuda_stream)
st.synchronize()
ts = []
for i in range(12):
    t0 = time.perf_counter()
    mats[i % nh].spmm_host_ptr(xp, bp, ap, yp, M)
    ts.append((time.perf_counter() - t0) * 1e3)
print(key, regime, f"handles={nh}", "ms per call:", " ".join(f"{t:.2f}" for t in ts), flush=True)
