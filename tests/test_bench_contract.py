"""CPU: bench.py's contract on a box without a GPU — the reference arm prints ONE JSON line with
the keys the driver reads (the reference's own CPU implementation, built in place, timed on a
bounded sample of the same workload), and our arm fails loudly instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True,
                          text=True, timeout=600, cwd=ROOT)


def test_reference_arm_line():
    r = run_bench("--impl", "reference", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                                   # stdout carries the JSON line and nothing else
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GFLOP/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("ternary spGEMM effective GFLOP/s")
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    # default workload: c4, the largest single-GPU configuration of BASELINE.json
    assert d["config"]["workload"].startswith("c4:") and (d["config"]["M"], d["config"]["K"]) == (2048, 8192)
    assert d["config"]["N"] == 28672 and d["config"]["s"] == 8
    cb = d["cpu_baseline"]
    # one instance of the (single-threaded) reference function per usable core, on disjoint row blocks
    ncpu = len(os.sched_getaffinity(0))
    assert cb["kind"] in ("reference", "port") and 1 <= cb["cores"] <= ncpu and cb["value"] == d["value"]
    assert cb["cores"] >= (min(ncpu, 2048 // 4) + 1) // 2 and cb["one_core_value"] > 0   # (row blocks are multiples of 4 rows)
    assert "DoubleUnrolledTCSC" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_our_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = run_bench("--steps", "1", "--warmup", "1", "--no-others", "--no-cpu-baseline", "--no-builder")
    assert r.returncode != 0                                 # no silent CPU path
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_roofline_args_for_the_reference_plot_script():
    """tools/roofline_args.py prints --beta / --pi in the units plots/plot_roofline.py:597-598 expects
    (bytes and flops per host TSC cycle) — sane magnitudes for a B200 next to a GHz-class host."""
    import re
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "roofline_args.py")], capture_output=True,
                       text=True, timeout=60)
    assert p.returncode == 0, p.stderr
    m = re.match(r"--beta ([0-9.]+) --pi ([0-9.]+)\s*$", p.stdout)
    assert m, p.stdout
    beta, pi = float(m.group(1)), float(m.group(2))
    assert 500 < beta < 20000 and 1e5 < pi < 5e6
