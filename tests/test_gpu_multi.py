"""-m gpu, needs >= 2 GPUs (skipped otherwise): the N-sharded CUDA path on 2 (and 4, when present)
ranks against the oracle — see tests/mgpu_worker.py.  tests/test_shard_gloo.py covers the same host
logic on CPU with gloo."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_cuda_path_against_oracle(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, {torch.cuda.device_count()} visible")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "mgpu_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0 and f"mgpu ok {world}" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]
