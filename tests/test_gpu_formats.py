"""GPU: the two 'next' formats — TCSR (bit-exact vs the reference constructor and
BaseTCSR) and packed-value CSC (layout defined here; pinned by round trip and Y parity)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from pcsc_layout import pack_reference, pack_reference_csr  # noqa: E402

pytestmark = pytest.mark.gpu

SHAPES = [(3, 4, 2, 0), (37, 29, 2, 1), (100, 130, 4, 2), (512, 2048, 16, 3), (1024, 256, 8, 4), (65, 257, 3, 5)]


@pytest.mark.parametrize("K,N,s,seed", SHAPES)
def test_tcsr_bit_exact_and_round_trip(tsg, orc, K, N, s, seed):
    W = orc.generate_sparse_matrix(K, N, s, seed)
    want = orc.tcsr(W)                       # restatement of TCSR.h:13-41, pinned vs the reference
    t = tsg.TCSR(W)
    for got, exp, name in zip(t.export(), want.arrays, ("rsp", "rsn", "cip", "cin")):
        assert got.dtype == np.int32 and np.array_equal(got, exp), name
    assert t.getDataStructureSize() == 4 * (2 * (K + 1) + want.arrays[2].size + want.arrays[3].size)   # TCSR.h:43-49
    assert np.array_equal(t.getVectorRepresentation(K, N), W)


@pytest.mark.parametrize("K,N,s,seed", SHAPES[:5])
@pytest.mark.parametrize("M", [1, 5])
def test_tcsr_spmm(tsg, orc, K, N, s, seed, M):
    W = orc.generate_sparse_matrix(K, N, s, seed)
    rng = np.random.default_rng(seed)
    Xr = rng.uniform(-1, 1, (M, K)).astype(np.float32)     # real-valued: order matters
    b = rng.uniform(-1, 1, N).astype(np.float32)
    al = rng.uniform(0.01, 0.3, N).astype(np.float32)
    want = orc.base_tcsr(Xr, orc.tcsr(W), b)               # BaseTCSR, comp.h:478-528
    t = tsg.TCSR(W)
    got = t.spmm(Xr, b, algo=tsg.ALGO_TCSR_SEQ)
    assert np.array_equal(got, want)                       # same order => bit-identical
    gotp = t.spmm(Xr, b, al, algo=tsg.ALGO_TCSR_SEQ)
    assert np.array_equal(gotp, np.where(want > 0, want, al * want).astype(np.float32))
    # the engine's kernels on the same W: integer X is exact in any order
    Xi = orc.init_x(M, K, seed + 7)
    wi = orc.base_tcsr(Xi, orc.tcsr(W), b * 0 + 2)
    for algo in (tsg.ALGO_AUTO, tsg.ALGO_DENSE_TC, tsg.ALGO_GATHER):
        assert np.array_equal(t.spmm(Xi, b * 0 + 2, algo=algo), wi)


@pytest.mark.parametrize("K,N,s,seed", SHAPES)
def test_pcsc_layout_and_round_trip(tsg, orc, K, N, s, seed):
    W = orc.generate_sparse_matrix(K, N, s, seed)
    cp, ri, vv = pack_reference(W)
    p = tsg.PackedCSC(W)
    gcp, gri, gvv = p.export()
    assert np.array_equal(gcp, cp) and np.array_equal(gri, ri) and np.array_equal(gvv, vv)
    nnz, nb = p.sizes
    assert nnz == ri.size and nb == (nnz + 4) // 5
    assert p.getDataStructureSize() == 4 * (N + 1) + 4 * nnz + nb
    assert np.array_equal(p.getVectorRepresentation(K, N), W)
    # adopt the arrays (interchange) and decode again
    p2 = tsg.PackedCSC.from_arrays(cp, ri, vv, K, N)
    assert np.array_equal(p2.getVectorRepresentation(K, N), W)


@pytest.mark.parametrize("K,N,s,seed", SHAPES[:5])
@pytest.mark.parametrize("M", [1, 4, 9])
def test_pcsc_spmm(tsg, orc, K, N, s, seed, M):
    W = orc.generate_sparse_matrix(K, N, s, seed)
    tref = orc.tcsc(W)
    Xi = orc.init_x(M, K, seed + 3)
    b = np.full(N, 2.0, np.float32)
    al = np.full(N, 0.1, np.float32)
    want, wantp = orc.base_tcsc(Xi, tref, b), orc.base_tcsc_prelu(Xi, tref, b, al)
    p = tsg.PackedCSC(W)
    for algo in (tsg.ALGO_PCSC_GATHER, tsg.ALGO_AUTO):
        assert np.array_equal(p.spmm(Xi, b, algo=algo), want)          # integer X: exact
        assert np.array_equal(p.spmm(Xi, b, al, algo=algo), wantp)
    rng = np.random.default_rng(seed)
    Xr = rng.uniform(-1, 1, (M, K)).astype(np.float32)
    wr = orc.base_tcsc(Xr, tref, b).astype(np.float64)
    gr = p.spmm(Xr, b, algo=tsg.ALGO_PCSC_GATHER).astype(np.float64)
    scale = np.abs(Xr).astype(np.float64) @ np.abs(W).astype(np.float64) + np.abs(b)
    assert np.max(np.abs(gr - wr) / scale) <= 1e-5                    # north star: max rel err 1e-5


@pytest.mark.parametrize("K,N,s,seed", SHAPES)
def test_pcsr_layout_and_round_trip(tsg, orc, K, N, s, seed):
    """Packed-value CSR: the packed layout along rows (== packed CSC of the transpose)."""
    W = orc.generate_sparse_matrix(K, N, s, seed)
    rp, ci, vv = pack_reference_csr(W)
    p = tsg.PackedCSR(W)
    grp, gci, gvv = p.export()
    assert grp.shape == (K + 1,) and np.array_equal(grp, rp)
    assert np.array_equal(gci, ci) and np.array_equal(gvv, vv)
    # the merged column list of a row is the sorted union of TCSR's two lists (TCSR.h:13-41)
    t = orc.tcsr(W).arrays
    assert np.array_equal(grp, t[0] + t[1])
    nnz, nb = p.sizes
    assert nnz == ci.size and nb == (nnz + 4) // 5
    assert p.getDataStructureSize() == 4 * (K + 1) + 4 * nnz + nb
    assert np.array_equal(p.getVectorRepresentation(K, N), W)
    p2 = tsg.PackedCSR.from_arrays(rp, ci, vv, K, N)
    assert np.array_equal(p2.getVectorRepresentation(K, N), W)
    assert all(np.array_equal(a, b) for a, b in zip(p2.export(), (rp, ci, vv)))


@pytest.mark.parametrize("K,N,s,seed", SHAPES[:5])
@pytest.mark.parametrize("M", [1, 5])
def test_pcsr_spmm(tsg, orc, K, N, s, seed, M):
    W = orc.generate_sparse_matrix(K, N, s, seed)
    rng = np.random.default_rng(seed)
    Xr = rng.uniform(-1, 1, (M, K)).astype(np.float32)     # real-valued: order matters
    b = rng.uniform(-1, 1, N).astype(np.float32)
    al = rng.uniform(0.01, 0.3, N).astype(np.float32)
    want = orc.base_tcsr(Xr, orc.tcsr(W), b)               # BaseTCSR, comp.h:478-528
    p = tsg.PackedCSR(W)
    got = p.spmm(Xr, b, algo=tsg.ALGO_PCSR_SEQ)
    assert np.array_equal(got, want)                       # same per-column order => bit-identical
    gotp = p.spmm(Xr, b, al, algo=tsg.ALGO_PCSR_SEQ)
    assert np.array_equal(gotp, np.where(want > 0, want, al * want).astype(np.float32))
    Xi = orc.init_x(M, K, seed + 7)
    wi = orc.base_tcsc(Xi, orc.tcsc(W), b * 0 + 2)
    for algo in (tsg.ALGO_AUTO, tsg.ALGO_DENSE_TC, tsg.ALGO_GATHER, tsg.ALGO_PCSR_SEQ):
        assert np.array_equal(p.spmm(Xi, b * 0 + 2, algo=algo), wi)


@pytest.mark.parametrize("K,N,s,seed,B", [(1024, 256, 4, 1, 512), (1100, 130, 2, 2, 512), (512, 2048, 16, 3, 512),
                                          (96, 37, 2, 4, 32), (300, 64, 4, 5, 64)])
def test_blocked_tcsc_bit_exact(tsg, orc, K, N, s, seed, B):
    """BlockedTCSC<B> (BlockedTCSC.h:15-43): device-built arrays == the reference constructor's,
    including the dropped tail rows when B does not divide K."""
    W = orc.generate_sparse_matrix(K, N, s, seed)
    want = orc.blocked(W, B)
    got = tsg.TCSC(W).blocked(B)
    for g, e, name in zip(got, want.arrays, ("csp", "csn", "rip", "rin")):
        assert g.shape == e.shape and np.array_equal(g, e), name
