#!/usr/bin/env python
"""tools/build_probe.py — run the device-side TCSC builder a few times on a BASELINE shape (developer
tool; wrap it in `ncu --metrics gpu__time_duration.sum -k regex:"encode|scan|emit|tile_codes"` for
the per-kernel times)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("TSG_BUILD_TIMING", "1")
import torch  # noqa: E402
import __graft_entry__ as ge  # noqa: E402

tsg = ge.load_package()
from ternary_spgemm_b200 import synth  # noqa: E402

key = sys.argv[1] if len(sys.argv) > 1 else "c4"
eb = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = synth.CONFIGS[key]
K, N, s = cfg["K"], cfg["N"], cfg["s"]
Wd = synth.device_ternary(K, N, s, 1234)
if eb == 4:
    Wd = Wd.to(torch.int32)
for i in range(3):
    t = tsg.TCSC.from_device_dense(Wd, K, N, elem_bytes=eb)
    torch.cuda.synchronize()
    print(key, f"int{8 * eb}", "device ms", round(float(tsg.lib().tsg_debug_last_build_device_ms()), 4), "nnz", sum(t.nnz), flush=True)
    t.close()
