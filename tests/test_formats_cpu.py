"""CPU: the packed-value CSC layout restatement used by the GPU tests is self-consistent
(pack -> unpack == W), and the oracle's TCSR restatement agrees with the compiled reference."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from pcsc_layout import pack_reference, pack_reference_csr  # noqa: E402


def unpack(cp, ri, vv, K, N):
    W = np.zeros((K, N), np.int32)
    d = (vv[:, None].astype(np.int64) // np.array([1, 3, 9, 27, 81])) % 3
    d = d.reshape(-1)[: ri.size] - 1
    for n in range(N):
        W[ri[cp[n]:cp[n + 1]], n] = d[cp[n]:cp[n + 1]]
    return W


def test_pack_round_trip(orc):
    for K, N, s, seed in [(3, 4, 2, 0), (37, 29, 2, 1), (100, 130, 4, 2)]:
        W = orc.generate_sparse_matrix(K, N, s, seed)
        cp, ri, vv = pack_reference(W)
        assert vv.size == (ri.size + 4) // 5 and int(vv.max(initial=0)) <= 242
        assert np.array_equal(unpack(cp, ri, vv, K, N), W)


def test_pack_csr_round_trip(orc):
    for K, N, s, seed in [(3, 4, 2, 0), (37, 29, 2, 1), (100, 130, 4, 2)]:
        W = orc.generate_sparse_matrix(K, N, s, seed)
        rp, ci, vv = pack_reference_csr(W)
        assert rp.shape == (K + 1,) and vv.size == (ci.size + 4) // 5
        assert np.array_equal(unpack(rp, ci, vv, N, K).T, W)
        t = orc.tcsr(W).arrays                      # TCSR.h:13-41: merged pointers = pos + neg
        assert np.array_equal(rp, t[0] + t[1])


def test_oracle_tcsr_vs_reference(orc, ref):
    W = orc.generate_sparse_matrix(64, 96, 4, 9)
    a, b = orc.tcsr(W), ref.tcsr(W)
    for x, y in zip(a.arrays, b.arrays):
        assert np.array_equal(x, y)
    X = orc.init_x(3, 64, 5)
    bias = np.full(96, 2.0, np.float32)
    assert np.array_equal(orc.base_tcsr(X, a, bias), ref.base_tcsr(W, X, bias))
