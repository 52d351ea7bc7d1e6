"""GPU: a short run of tools/fuzz.py — random (M, K, N, s) through every kernel.  Integer X must be
bit-identical to the reference-order kernel (with and without PReLU); real-valued X of random
scale, sometimes mixed with integer tiles, within 4e-6 of the forward scale for the exact and the
opt-in fast split; AUTO never more than 1.5x off the best kernel it could have picked is reported,
not asserted (timing noise).  The shapes cover partial waves, tail launches, ragged K and N."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [101, 202])
def test_fuzz_every_kernel(seed):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz.py"), "40", str(seed)],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    tail = p.stdout[p.stdout.rfind("{\n"):]
    assert p.returncode == 0, tail[-3000:] + p.stderr[-2000:]
    assert json.loads(tail)["mismatches"] == []
