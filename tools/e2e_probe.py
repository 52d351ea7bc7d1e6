#!/usr/bin/env python
"""tools/e2e_probe.py — where the host-pointer call's time goes at c2 (developer tool):
per-call wall time of tsg_spmm from Python with and without the NVML sampler thread running,
and (TSG_E2E_TRACE=1) the C-side phase split printed by libtsg on stderr."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import __graft_entry__ as ge  # noqa: E402
import bench  # noqa: E402

tsg = ge.load_package()
from ternary_spgemm_b200 import synth  # noqa: E402

cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
M, K, N, s = cfg["M"], cfg["K"], cfg["N"], cfg["s"]
Wd = synth.device_ternary(K, N, s, 1234)
m = tsg.TCSC.from_device_dense(Wd, K, N, elem_bytes=1)
mats = [m] + [m.slice_cols(0, N) for _ in range(12)]   # bench.py rotates 13 HBM copies of W at c2
Xh = synth.device_x(M, K, 1).cpu().pin_memory()
bh = torch.full((N,), 2.0).pin_memory()
Yh = torch.empty(M, N).pin_memory()
xp, bp, yp = Xh.data_ptr(), bh.data_ptr(), Yh.data_ptr()


def loop(n, rotate=False):
    t0 = time.perf_counter()
    for i in range(n):
        (mats[i % 13] if rotate else m).spmm_host_ptr(xp, bp, None, yp, M)
    return (time.perf_counter() - t0) / n * 1e6


loop(200)
for rep in range(3):
    print(f"one handle: {loop(2000):.2f} us/call   13 handles in rotation: {loop(2000, True):.2f} us/call", flush=True)
for interval in (0.002, 0.02):
    sm = bench.ClockSampler(0)
    sm.interval = interval
    with sm:
        r = [loop(2000) for _ in range(3)]
    print(f"sampler every {interval * 1e3:.0f} ms: " + " ".join(f"{v:.2f}" for v in r) + f" us/call ({len(sm.samples)} samples)", flush=True)
