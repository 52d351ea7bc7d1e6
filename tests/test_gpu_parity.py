"""GPU parity tests: the CUDA path (through the C ABI) against the oracle / golden fixtures.

Bars (BASELINE.json north_star):
  * format conversion: bit-exact against the reference TCSC arrays;
  * Y: bit-exact on the reference's integer-valued inputs for every kernel; on real-valued
    inputs bit-exact for TSG_ALGO_GATHER_SEQ (reference summation order) and within
    REL_TOL = 1e-5 for the re-ordered kernels, measured as max|Y-Yref| / max|Yref| (the
    max-norm relative error; element-wise relative error is unbounded where sums cancel) and
    backed by the rigorous forward bound  |err| <= n·eps·Σ|x|  per element.
"""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
EPS = np.finfo(np.float32).eps
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def algos(tsg):
    return [tsg.ALGO_GATHER, tsg.ALGO_GATHER_SEQ, tsg.ALGO_DENSE_TC, tsg.ALGO_CODE_GEMV, tsg.ALGO_AUTO]


def run_or_skip(tsg, fn):
    """Kernels may declare a shape unsupported (TSG_ERR_UNSUPPORTED = -5); anything else is a bug."""
    try:
        return fn()
    except tsg.TsgError as e:
        if e.status == -5:
            return None
        raise


def rel_err(Y, Yref):
    scale = max(float(np.abs(Yref).max()), 1e-30)
    return float(np.abs(Y.astype(np.float64) - Yref.astype(np.float64)).max()) / scale


def abs_sum_bound(X, tcsc, extra=1.0):
    """Per-element forward error bound of any fp32 summation order: (n+2)·eps·(Σ|x| + |b|)."""
    absX = np.abs(X).astype(np.float64)
    csp, csn, rip, rin = tcsc.arrays
    N = tcsc.cols
    tot = np.zeros((X.shape[0], N))
    cnt = np.zeros(N)
    for n in range(N):
        idx = np.concatenate([rip[csp[n]:csp[n + 1]], rin[csn[n]:csn[n + 1]]])
        tot[:, n] = absX[:, idx].sum(axis=1)
        cnt[n] = idx.size
    return (cnt[None, :] + 2) * EPS * (tot + extra)


# ------------------------------------------------------------------------------------------------
# a1: device-side TCSC builder
# ------------------------------------------------------------------------------------------------
BUILD_SHAPES = [(1, 1, 1, 0), (3, 4, 2, 0), (31, 33, 2, 1), (32, 32, 2, 2), (33, 31, 3, 3),
                (100, 130, 4, 7), (257, 1000, 8, 4), (1024, 1100, 16, 5), (512, 2048, 2, 0),
                (2048, 512, 4, 1), (4096, 1024, 16, 2), (1024, 4096, 4, 1234)]


@pytest.mark.parametrize("K,N,s,seed", BUILD_SHAPES)
def test_builder_bit_exact(tsg, orc, K, N, s, seed):
    W = orc.generate_sparse_matrix(K, N, s, seed)
    want = orc.tcsc(W)
    t = tsg.TCSC(W)
    assert (t.getNumRows(), t.getNumCols()) == (K, N)
    assert t.nnz == (want.row_index_pos.size, want.row_index_neg.size)
    for got, exp, name in zip(t.export(), want.arrays, ("csp", "csn", "rip", "rin")):
        assert got.dtype == np.int32 and got.shape == exp.shape and np.array_equal(got, exp), name
    assert t.getDataStructureSize() == orc.tcsc_size_bytes(want)
    assert np.array_equal(t.getVectorRepresentation(K, N), W)  # test_data_structure.cpp:62-73


def test_builder_edge_cases(tsg, orc):
    rng = np.random.default_rng(0)
    cases = {
        "all_zero": np.zeros((40, 50), np.int32),
        "all_plus": np.ones((70, 9), np.int32),
        "all_minus": -np.ones((33, 65), np.int32),
        "junk_values_ignored": rng.integers(-3, 4, (129, 77)).astype(np.int32),  # TCSC.h:26-35
        "single_row": rng.integers(-1, 2, (1, 300)).astype(np.int32),
        "single_col": rng.integers(-1, 2, (300, 1)).astype(np.int32),
        "empty_cols": np.concatenate([np.zeros((64, 40), np.int32),
                                      rng.integers(-1, 2, (64, 3)).astype(np.int32),
                                      np.zeros((64, 40), np.int32)], axis=1),
    }
    for name, W in cases.items():
        want = orc.tcsc(W)
        t = tsg.TCSC(W)
        for got, exp in zip(t.export(), want.arrays):
            assert np.array_equal(got, exp), name
        Wt = np.where(np.abs(W) == 1, W, 0)
        assert np.array_equal(t.getVectorRepresentation(), Wt), name


def test_builder_zero_sized(tsg):
    for K, N in ((0, 5), (5, 0), (0, 0)):
        t = tsg.TCSC(np.zeros((K, N), np.int32))
        csp, csn, rip, rin = t.export()
        assert csp.tolist() == [0] * (N + 1) and csn.tolist() == [0] * (N + 1)
        assert rip.size == 0 and rin.size == 0
        Y = t.spmm(np.zeros((3, K), np.float32), np.ones(N, np.float32))
        assert Y.shape == (3, N)


def test_builder_int8_device_input_and_shards(tsg, orc):
    import torch
    K, N, s = 512, 2048, 4
    W = orc.generate_sparse_matrix(K, N, s, 3)
    want = orc.tcsc(W)
    Wd8 = torch.from_numpy(W.astype(np.int8)).cuda()
    Wd32 = torch.from_numpy(W).cuda()
    for Wd, eb in ((Wd8, 1), (Wd32, 4)):
        t = tsg.TCSC.from_device_dense(Wd, K, N, elem_bytes=eb)
        for got, exp in zip(t.export(), want.arrays):
            assert np.array_equal(got, exp)
    # N-column shards three ways: from host cols, from device cols, device-side slice
    full = tsg.TCSC(W)
    for world in (2, 3, 8):
        for r in range(world):
            lo, hi = tsg.shard_columns(N, world, r)
            exp = orc.tcsc(W[:, lo:hi])
            for t in (tsg.TCSC(W, col_range=(lo, hi)),
                      tsg.TCSC.from_device_dense(Wd8, K, N, elem_bytes=1, col_range=(lo, hi)),
                      full.slice_cols(lo, hi)):
                for got, e in zip(t.export(), exp.arrays):
                    assert np.array_equal(got, e), (world, r)


def test_from_arrays_roundtrip(tsg, orc):
    W = orc.generate_sparse_matrix(300, 200, 4, 9)
    o = orc.tcsc(W)
    t = tsg.TCSC.from_arrays(*o.arrays, 300, 200)
    assert np.array_equal(t.getVectorRepresentation(), W)
    X = orc.init_x(3, 300, 1)
    b = np.full(200, 2.0, np.float32)
    for algo in algos(tsg):
        Y = run_or_skip(tsg, lambda: t.spmm(X, b, algo=algo))
        if Y is not None:
            assert np.array_equal(Y, orc.base_tcsc(X, o, b)), tsg.ALGO_NAMES[algo]


def test_from_arrays_rejects_malformed_input(tsg, orc):
    """The interchange entry points adopt caller-made arrays: anything the reference constructors
    could not have produced (TCSC.h:24-36: pointers from 0, non-decreasing, rows ascending inside
    [0, K)) must come back as TSG_ERR_INVALID, not as a write outside device buffers."""
    K, N = 64, 40
    W = orc.generate_sparse_matrix(K, N, 4, 3)
    csp, csn, rip, rin = (a.copy() for a in orc.tcsc(W).arrays)

    def bad(**kw):
        arrs = dict(csp=csp.copy(), csn=csn.copy(), rip=rip.copy(), rin=rin.copy())
        for k, f in kw.items():
            f(arrs[k])
        with pytest.raises(tsg.TsgError) as e:
            tsg.TCSC.from_arrays(arrs["csp"], arrs["csn"], arrs["rip"], arrs["rin"], K, N)
        assert e.value.status == -1, e.value
    bad(csp=lambda a: a.__setitem__(0, 1))                      # does not start at 0
    bad(csn=lambda a: a.__setitem__(5, a[4] - 1))               # decreasing pointer
    bad(rip=lambda a: a.__setitem__(3, K))                      # row out of range
    bad(rin=lambda a: a.__setitem__(0, -1))                     # negative row
    first_long = int(np.argmax(np.diff(csp) >= 2))
    bad(rip=lambda a: a.__setitem__(csp[first_long] + 1, a[csp[first_long]]))   # not strictly ascending
    with pytest.raises(tsg.TsgError) as e:                      # row 5 listed as +1 and as -1
        tsg.TCSC.from_arrays(np.array([0, 2], np.int32), np.array([0, 2], np.int32),
                             np.array([1, 5], np.int32), np.array([5, 9], np.int32), 16, 1)
    assert e.value.status == -1
    t = tsg.TCSC.from_arrays(csp, csn, rip, rin, K, N)          # the untouched arrays are accepted
    assert np.array_equal(t.getVectorRepresentation(), W)
    # packed formats
    p = tsg.PackedCSC(W)
    cp, ri, vv = p.export()
    for mutate in (lambda: (np.r_[1, cp[1:]], ri, vv), lambda: (cp, np.r_[K, ri[1:]], vv),
                   lambda: (cp, ri, np.r_[np.uint8(250), vv[1:]]),
                   lambda: (np.r_[cp[:3], cp[2] - 1, cp[4:]] if cp[2] > 0 else np.r_[cp[:-1], cp[-1] + 1], ri, vv)):
        with pytest.raises(tsg.TsgError) as e:
            tsg.PackedCSC.from_arrays(*mutate(), K, N)
        assert e.value.status == -1
    r = tsg.PackedCSR(W)
    rp, ci, vv = r.export()
    for mutate in (lambda: (rp, np.r_[N, ci[1:]], vv), lambda: (np.r_[rp[:-1], rp[-1] + 5], ci, vv)):
        with pytest.raises(tsg.TsgError) as e:
            tsg.PackedCSR.from_arrays(*mutate(), K, N)
        assert e.value.status == -1


# ------------------------------------------------------------------------------------------------
# golden fixtures (outputs of the unmodified reference)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_golden(tsg, orc, path):
    g = np.load(path)
    M, K, N, s, seed = (int(v) for v in g["shape"])
    W = g["W"].astype(np.int32)
    t = tsg.TCSC(W)
    for got, key in zip(t.export(), ("csp", "csn", "rip", "rin")):
        assert np.array_equal(got, g[key]), key
    assert t.getDataStructureSize() == int(g["ds_bytes"])
    b2, a01 = np.full(N, 2.0, np.float32), np.full(N, 0.1, np.float32)
    o = orc.tcsc(W)
    bound = abs_sum_bound(g["X_real"], o, extra=np.abs(g["b"]).max())
    ran = 0
    for algo in algos(tsg):
        Y = run_or_skip(tsg, lambda: t.spmm(g["X_int"], b2, algo=algo))
        if Y is None:
            continue
        ran += 1
        name = tsg.ALGO_NAMES[algo]
        assert np.array_equal(Y, g["Y_int"]), name                      # integer regime: exact
        assert np.array_equal(t.spmm(g["X_int"], b2, a01, algo=algo), g["Y_int_prelu"]), name
        assert orc.compare_results(Y, g["Y_int_dense"]), name           # reference -correctness
        Yr = t.spmm(g["X_real"], g["b"], algo=algo)
        Yp = t.spmm(g["X_real"], g["b"], g["alpha"], algo=algo)
        if algo == tsg.ALGO_GATHER_SEQ:
            assert np.array_equal(Yr, g["Y_real"]) and np.array_equal(Yp, g["Y_real_prelu"])
        else:
            assert rel_err(Yr, g["Y_real"]) <= REL_TOL, name
            assert rel_err(Yp, g["Y_real_prelu"]) <= REL_TOL, name
            assert np.all(np.abs(Yr.astype(np.float64) - g["Y_real"]) <= 2 * bound), name
    assert ran >= 4


# ------------------------------------------------------------------------------------------------
# a2/a3: kernels vs oracle on seeded inputs
# ------------------------------------------------------------------------------------------------
SPMM_SHAPES = [  # M, K, N, s
    (1, 64, 48, 2), (1, 1000, 333, 3), (2, 512, 512, 4), (3, 37, 29, 3), (4, 1024, 256, 8),
    (5, 2048, 300, 16), (7, 4096, 257, 2), (8, 1024, 4096, 4), (16, 512, 1024, 4),
    (33, 256, 512, 2), (64, 1024, 384, 8), (1, 8192, 1024, 8), (130, 128, 200, 2),
]


@pytest.mark.parametrize("M,K,N,s", SPMM_SHAPES)
def test_spmm_vs_oracle(tsg, orc, M, K, N, s):
    seed = M * 7 + K + N + s
    W = orc.generate_sparse_matrix(K, N, s, seed)
    o = orc.tcsc(W)
    t = tsg.TCSC(W)
    rng = np.random.default_rng(seed)
    Xi = orc.init_x(M, K, seed)
    Xr = rng.standard_normal((M, K)).astype(np.float32)
    b = rng.uniform(-2, 2, N).astype(np.float32)
    al = rng.uniform(0.0, 0.5, N).astype(np.float32)
    Yi, Yip = orc.base_tcsc(Xi, o, b), orc.base_tcsc_prelu(Xi, o, b, al)
    Yr, Yrp = orc.base_tcsc(Xr, o, b), orc.base_tcsc_prelu(Xr, o, b, al)
    ran = 0
    for algo in algos(tsg):
        name = tsg.ALGO_NAMES[algo]
        Y = run_or_skip(tsg, lambda: t.spmm(Xi, b, algo=algo))
        if Y is None:
            continue
        ran += 1
        # b is real-valued here, so the final "+ b" rounds once, identically in every kernel
        # only if the integer partial sums agree: still exact.
        assert np.array_equal(Y, Yi), name
        assert np.array_equal(t.spmm(Xi, b, al, algo=algo), Yip), name
        Y2, Y2p = t.spmm(Xr, b, algo=algo), t.spmm(Xr, b, al, algo=algo)
        if algo == tsg.ALGO_GATHER_SEQ:
            assert np.array_equal(Y2, Yr) and np.array_equal(Y2p, Yrp), name
        else:
            assert rel_err(Y2, Yr) <= REL_TOL, (name, rel_err(Y2, Yr))
            assert rel_err(Y2p, Yrp) <= REL_TOL, (name, rel_err(Y2p, Yrp))
    assert ran >= 3


def test_spmm_deterministic_and_overwrites(tsg, orc):
    """Y is overwritten, not accumulated (perf.cpp:63-66 re-uses Y across calls), and results
    are run-to-run identical (fixed reduction order, no float atomics)."""
    M, K, N, s = 4, 2048, 1024, 4
    W = orc.generate_sparse_matrix(K, N, s, 2)
    t = tsg.TCSC(W)
    X = np.random.default_rng(1).standard_normal((M, K)).astype(np.float32)
    b = np.ones(N, np.float32)
    for algo in algos(tsg):
        Y0 = run_or_skip(tsg, lambda: t.spmm(X, b, algo=algo))
        if Y0 is None:
            continue
        out = np.full((M, N), 1e30, np.float32)
        for _ in range(3):
            t.spmm(X, b, algo=algo, out=out)
            assert np.array_equal(out, Y0), tsg.ALGO_NAMES[algo]


def test_spmm_special_values(tsg, orc):
    """PReLU branch uses strict y>0 (comp_prelu.h:57); zeros, negative zero and exact
    cancellation must follow the reference."""
    K, N = 64, 96
    W = orc.generate_sparse_matrix(K, N, 2, 5)
    o = orc.tcsc(W)
    t = tsg.TCSC(W)
    X = np.zeros((2, K), np.float32)
    X[1] = 1.0
    b = np.zeros(N, np.float32)
    b[::3] = -0.0
    al = np.full(N, -0.25, np.float32)
    for algo in algos(tsg):
        Y = run_or_skip(tsg, lambda: t.spmm(X, b, al, algo=algo))
        if Y is not None:
            assert np.array_equal(Y, orc.base_tcsc_prelu(X, o, b, al)), tsg.ALGO_NAMES[algo]


@pytest.mark.parametrize("M", [1, 2, 3, 8, 16, 40, 64, 100, 300])
def test_non_finite_x_follows_reference(tsg, orc, M):
    """comp.h:44-61 never touches x where W is 0: an inf / NaN there must not reach Y, and one at a
    non-zero position must arrive exactly as the reference's sequential sum delivers it (inf, or NaN
    from inf - inf).  Values near FLT_MAX must not overflow where the reference does not.  Checked
    against the oracle for every kernel, AUTO included (the oracle itself is pinned against the
    unmodified reference on the same kind of input in tests/test_oracle_pinning.py)."""
    K, N, s = 512, 640, 4
    W = orc.generate_sparse_matrix(K, N, s, 21)
    o = orc.tcsc(W)
    t = tsg.TCSC(W)
    rng = np.random.default_rng(M)
    X = orc.init_x(M, K, 31)
    b = rng.uniform(-1, 1, N).astype(np.float32)
    al = rng.uniform(0.01, 0.3, N).astype(np.float32)
    specials = [np.inf, -np.inf, np.nan, 3.0e38, -3.4e38, 2.0 ** 100, np.float32(1e-40)]
    for i, v in enumerate(specials):
        X[rng.integers(0, M), rng.integers(0, K)] = v
    # a column of W that is all zero in some rows: make sure at least one special sits on a zero of W
    kz = int(np.argmin(np.abs(W).sum(axis=1)))
    X[0, kz] = np.inf
    want = orc.base_tcsc(X, o, b)
    wantp = orc.base_tcsc_prelu(X, o, b, al)
    for algo in algos(tsg):
        Y = run_or_skip(tsg, lambda: t.spmm(X, b, algo=algo))
        if Y is None:
            continue
        name = tsg.ALGO_NAMES[algo]
        if algo == tsg.ALGO_GATHER:
            # re-ordered adds: same inf/NaN pattern wherever the reference's sum is finite or inf
            fin = np.isfinite(want)
            assert np.array_equal(np.isfinite(Y)[fin], fin[fin]), name
            continue
        assert np.array_equal(Y, want, equal_nan=True), name
        assert np.array_equal(t.spmm(X, b, al, algo=algo), wantp, equal_nan=True), name


def test_two_streams_one_handle(tsg, orc):
    """The tensor-core path keeps ONE split scratch per handle: calls arriving on different streams
    must serialise on it instead of overwriting each other's operand tiles."""
    import torch
    M, K, N, s = 200, 2048, 1024, 4
    W = orc.generate_sparse_matrix(K, N, s, 4)
    o = orc.tcsc(W)
    t = tsg.TCSC(W)
    b = np.full(N, 2.0, np.float32)
    db = torch.from_numpy(b).cuda()
    Xs = [orc.init_x(M, K, 40 + i) for i in range(6)]
    dXs = [torch.from_numpy(x).cuda() for x in Xs]
    dYs = [torch.empty(M, N, device="cuda") for _ in Xs]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    for rep in range(3):
        for i, (dx, dy) in enumerate(zip(dXs, dYs)):
            st = streams[i & 1]
            t.spmm_dev(dx, db, dy, M, algo=tsg.ALGO_DENSE_TC, stream=st.cuda_stream)
        torch.cuda.synchronize()
        for x, dy in zip(Xs, dYs):
            assert np.array_equal(dy.cpu().numpy(), orc.base_tcsc(x, o, b))


def test_shape_mismatch_is_an_error(tsg, orc):
    t = tsg.TCSC(orc.generate_sparse_matrix(64, 32, 2, 0))
    with pytest.raises(tsg.TsgError):
        t.spmm(np.zeros((2, 63), np.float32), np.zeros(32, np.float32))
    with pytest.raises(tsg.TsgError):
        t.getVectorRepresentation(32, 64)


@pytest.mark.parametrize("M", [3, 100])
def test_device_pointer_entry(tsg, orc, M):
    import torch
    K, N, s = 1024, 2048, 4
    W = orc.generate_sparse_matrix(K, N, s, 8)
    o = orc.tcsc(W)
    t = tsg.TCSC(W)
    X = orc.init_x(M, K, 3)
    b = np.full(N, 2.0, np.float32)
    al = np.full(N, 0.1, np.float32)
    dX, db, da = (torch.from_numpy(a).cuda() for a in (X, b, al))
    # padded leading dimensions
    dXp = torch.zeros(M, K + 24, device="cuda")
    dXp[:, :K] = dX
    dY = torch.full((M, N + 8), -7.0, device="cuda")
    st = torch.cuda.Stream()
    for algo in algos(tsg):
        try:
            with torch.cuda.stream(st):
                t.spmm_dev(dXp, db, dY, M, alpha=da, algo=algo, ldx=K + 24, ldy=N + 8,
                           stream=st.cuda_stream)
            st.synchronize()
        except tsg.TsgError as e:
            if e.status == -5:
                continue
            raise
        Y = dY.cpu().numpy()
        assert np.array_equal(Y[:, :N], orc.base_tcsc_prelu(X, o, b, al)), tsg.ALGO_NAMES[algo]
        assert np.all(Y[:, N:] == -7.0)
        dY.fill_(-7.0)


# ------------------------------------------------------------------------------------------------
# BASELINE configs at full size: oracle where it finishes in seconds, properties elsewhere
# ------------------------------------------------------------------------------------------------
def test_config1_readme_example(tsg, orc):
    M, K, N, s = 32, 1024, 4096, 4
    W = orc.generate_sparse_matrix(K, N, s, 1234)
    o = orc.tcsc(W)
    t = tsg.TCSC(W)
    for got, exp in zip(t.export(), o.arrays):
        assert np.array_equal(got, exp)
    X = orc.init_x(M, K, 5)
    b, al = np.full(N, 2.0, np.float32), np.full(N, 0.1, np.float32)
    Yd = orc.gemm(X, W, b)
    for algo in algos(tsg):
        Y = run_or_skip(tsg, lambda: t.spmm(X, b, algo=algo))
        if Y is not None:
            assert np.array_equal(Y, orc.base_tcsc(X, o, b))
            assert orc.compare_results(Y, Yd)  # what `-correctness` checks
            assert np.array_equal(t.spmm(X, b, al, algo=algo), orc.base_tcsc_prelu(X, o, b, al))


def test_config2_decode_shape(tsg, orc):
    M, K, N, s = 1, 4096, 4096, 3
    W = orc.generate_sparse_matrix(K, N, s, 1234)
    o = orc.tcsc(W)
    assert o.nnz == 5_586_944  # N/s non-integer
    t = tsg.TCSC(W)
    for got, exp in zip(t.export(), o.arrays):
        assert np.array_equal(got, exp)
    X = orc.init_x(M, K, 6)
    Xr = np.random.default_rng(6).standard_normal((M, K)).astype(np.float32)
    b = np.full(N, 2.0, np.float32)
    for algo in algos(tsg):
        Y = run_or_skip(tsg, lambda: t.spmm(X, b, algo=algo))
        if Y is not None:
            assert np.array_equal(Y, orc.base_tcsc(X, o, b))
            assert rel_err(t.spmm(Xr, b, algo=algo), orc.base_tcsc(Xr, o, b)) <= REL_TOL


def _device_ternary(K, N, s, seed):
    """Synthetic W on the device (int8): exactly N//s non-zeros per row, random signs — the
    reference generator's marginal distribution without its O(KN) host cost."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    W = torch.zeros(K, N, dtype=torch.int8, device="cuda")
    rows_per = max(1, (1 << 24) // N)
    for k0 in range(0, K, rows_per):
        k1 = min(K, k0 + rows_per)
        cols = torch.rand(k1 - k0, N, device="cuda", generator=g).argsort(dim=1)[:, : N // s]
        sign = (torch.randint(0, 2, cols.shape, device="cuda", generator=g) * 2 - 1).to(torch.int8)
        W[k0:k1].scatter_(1, cols, sign)
    return W


@pytest.mark.parametrize("M,K,N,s,prelu", [(1, 4096, 4096, 3, False),      # config 2 (decode)
                                          (32, 1024, 4096, 4, True),      # config 1
                                          (256, 4096, 14336, 4, True),    # config 3
                                          (32, 8192, 28672, 8, False),    # config 4/5 family
                                          (512, 8192, 14336, 2, False)])  # config 5 family, M = 512
def test_full_size_properties(tsg, M, K, N, s, prelu):
    """Full BASELINE sizes, size-independent checks: (1) builder round trip dense->TCSC->dense,
    pointer monotonicity, ascending rows; (2) linearity Y(X1+X2)-b = (Y(X1)-b)+(Y(X2)-b) exactly on
    integer inputs; (3) column-sum checksum: Σ_n Y[m,n] = X[m,:]·rowsum(W) + Σb; (4) kernels agree
    with the reference-order kernel on a row subset."""
    import torch
    Wd = _device_ternary(K, N, s, 11)
    t = tsg.TCSC.from_device_dense(Wd, K, N, elem_bytes=1)
    csp, csn, rip, rin = t.export()
    assert csp[0] == 0 and csn[0] == 0 and np.all(np.diff(csp) >= 0) and np.all(np.diff(csn) >= 0)
    assert csp[-1] + csn[-1] == K * (N // s)          # _device_ternary: exactly N//s per row
    for ptr, idx in ((csp, rip), (csn, rin)):
        d = np.diff(idx.astype(np.int64))
        starts = ptr[1:-1][(ptr[1:-1] > 0) & (ptr[1:-1] < idx.size)]
        interior = np.ones(idx.size - 1, bool)
        interior[starts - 1] = False
        assert np.all(d[interior] > 0)  # strictly ascending inside every column
    Wh = Wd.cpu().numpy().astype(np.int32)
    assert np.array_equal(t.getVectorRepresentation(), Wh)
    rng = np.random.default_rng(3)
    X1 = rng.integers(-512, 513, (M, K)).astype(np.float32)
    X2 = rng.integers(-512, 513, (M, K)).astype(np.float32)
    b = np.full(N, 2.0, np.float32)
    al = np.full(N, 0.1, np.float32) if prelu else None
    rowsum = Wh.sum(axis=1).astype(np.float64)
    Mseq = min(M, 4)
    Yseq = t.spmm(X1[:Mseq], b, al, algo=tsg.ALGO_GATHER_SEQ)
    for algo in (tsg.ALGO_GATHER, tsg.ALGO_DENSE_TC, tsg.ALGO_AUTO) + ((tsg.ALGO_CODE_GEMV,) if M <= 32 else ()):
        Y1 = run_or_skip(tsg, lambda: t.spmm(X1, b, algo=algo))
        if Y1 is None:
            continue
        name = tsg.ALGO_NAMES[algo]
        Y2 = t.spmm(X2, b, algo=algo)
        Y12 = t.spmm(X1 + X2, b, algo=algo)
        assert np.array_equal(Y12 - 2.0, (Y1 - 2.0) + (Y2 - 2.0)), name           # linearity
        assert np.array_equal(Y1.astype(np.float64).sum(axis=1),
                              X1.astype(np.float64) @ rowsum + 2.0 * N), name     # checksum
        Ya = t.spmm(X1[:Mseq], b, al, algo=algo)
        assert np.array_equal(Ya, Yseq), name
    del Wd
    torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------------------
# Every tile height / occupancy variant of the tensor-core kernel, forced through its developer
# overrides (the heuristics reach some of them only at the largest shapes), against the
# reference-order kernel on the device — integer X: bit-identical; real X: the written tolerance.
VARIANTS = [  # M, K, N, s, env
    (8, 1024, 4000, 4, {"TSG_TC_EW": "8"}),        # in-kernel X conversion, two CTAs per SM
    (8, 1024, 4000, 4, {"TSG_TC_EW": "16"}),
    (16, 300, 260, 2, {"TSG_TC_EW": "8"}),
    (30, 4096, 4096, 8, {"TSG_TC_EW": "8"}),       # TMA path, 32-row tiles, two CTAs per SM
    (30, 4096, 4096, 8, {"TSG_TC_EW": "16"}),
    (50, 2048, 4096, 4, {}),                       # 64-row tiles
    (100, 1024, 700, 4, {"TSG_TC_NT": "64"}),
    (100, 1024, 700, 4, {"TSG_TC_NT": "128"}),
    (100, 1024, 700, 4, {"TSG_TC_NT": "256"}),
    (300, 512, 1200, 2, {"TSG_TC_NT": "256"}),
    (300, 512, 1200, 2, {"TSG_TC_NT": "208"}),      # run-time tile height (any multiple of 16)
    (300, 512, 1200, 2, {"TSG_TC_NT": "80"}),
    (300, 512, 1200, 2, {"TSG_TC_NT": "128", "TSG_TC_PDL": "0"}),
    (300, 512, 1200, 2, {"TSG_TC_NT": "256", "TSG_TC_FAST": "1"}),    # opt-in two-fp16-term tiles, term-major
    (300, 512, 1200, 2, {"TSG_TC_NT": "80", "TSG_TC_FAST": "1"}),
    (50, 2048, 4096, 4, {"TSG_TC_FAST": "1"}),                        # ... side by side (64-row tiles)
    (30, 4096, 4096, 8, {"TSG_TC_EW": "8", "TSG_TC_FAST": "1"}),
    # the partial last wave of a tall-tile grid, K-split through global memory (forced: the heuristic
    # wants longer K): 148 whole tiles + 2 split four ways; a grid that is all tail, rows beyond M;
    # an uneven three-way split of four stages at a run-time tile height
    (256, 1024, 19200, 4, {"TSG_TC_NT": "256", "TSG_TC_TAIL": "8"}),
    (256, 1024, 19200, 4, {"TSG_TC_NT": "256", "TSG_TC_TAIL": "8", "TSG_TC_FAST": "1"}),
    (300, 1024, 1200, 2, {"TSG_TC_NT": "256", "TSG_TC_TAIL": "8"}),
    (300, 1024, 1200, 2, {"TSG_TC_NT": "208", "TSG_TC_TAIL": "3"}),
    (300, 2048, 1200, 2, {"TSG_TC_NT": "96", "TSG_TC_TAIL": "2", "TSG_TC_PDL": "0"}),
]


@pytest.mark.parametrize("M,K,N,s,env", VARIANTS)
def test_dense_tc_variants(tsg, orc, M, K, N, s, env):
    import subprocess
    import sys
    import textwrap
    code = textwrap.dedent(f"""
        import sys, numpy as np
        sys.path.insert(0, {ROOT!r})
        import __graft_entry__ as ge
        from oracle.pyoracle import Oracle
        tsg, orc = ge.load_package(), Oracle()
        M, K, N, s = {M}, {K}, {N}, {s}
        W = orc.generate_sparse_matrix(K, N, s, 17)
        t = tsg.TCSC(W)
        rng = np.random.default_rng(5)
        b = rng.uniform(-1, 1, N).astype(np.float32)
        al = rng.uniform(0.01, 0.3, N).astype(np.float32)
        Xi = orc.init_x(M, K, 23)
        for alpha in (None, al):
            want = t.spmm(Xi, b, alpha, algo=tsg.ALGO_GATHER_SEQ)
            got = t.spmm(Xi, b, alpha, algo=tsg.ALGO_DENSE_TC)
            assert np.array_equal(got, want), "integer X must be bit-identical"
        # real-valued X: three bf16 terms, exact.  With TSG_TC_FAST=1 (some variants): scale 1 and 3e4
        # are inside fp16's range -> two fp16 terms (x carried to max(2^-24 |x|, 2^-25)); 1e-3 lies
        # below it and 1e5 above it -> three bf16 terms again
        for scale in (1.0, 1e-3, 3e4, 1e5):
            Xr = (rng.uniform(-1, 1, (M, K)) * scale).astype(np.float32)
            want = t.spmm(Xr, b, algo=tsg.ALGO_GATHER_SEQ).astype(np.float64)
            got = t.spmm(Xr, b, algo=tsg.ALGO_DENSE_TC).astype(np.float64)
            bound = np.abs(Xr).astype(np.float64) @ np.abs(W).astype(np.float64) + np.abs(b)
            rel = np.max(np.abs(got - want) / bound)
            assert rel <= 1e-6, (scale, rel)       # the bar is 1e-5; both splits sit at fp32 summation noise
            exact = (Xr.astype(np.float64) @ W.astype(np.float64)) + b
            assert np.max(np.abs(got - exact) / bound) <= 1e-6, scale
        # operand format is chosen per X tile (m-tile x 64 k): integer tiles (one fp16 term), tiles of
        # 17-bit integers (three bf16 terms), bf16-valued tiles (one bf16 term), all-zero tiles — all
        # integer-valued, every partial sum < 2^24, so any order is exact: bit-identical
        Xm = Xi.copy()
        for kb in range(0, K, 192):
            Xm[:, kb:kb + 64] = rng.integers(-60000, 60001, (M, min(64, K - kb)))
        for kb in range(64, K, 256):
            Xm[:, kb:kb + 64] = rng.integers(-3, 4, (M, min(64, K - kb))) * 4096.0
        Xm[:, 128:192] = 0.0
        Xm[M // 2:, :64] = Xi[M // 2:, :64]        # formats differ between m-tiles of the same k-block
        want = t.spmm(Xm, b, al, algo=tsg.ALGO_GATHER_SEQ)
        assert np.array_equal(t.spmm(Xm, b, al, algo=tsg.ALGO_DENSE_TC), want), "mixed tile formats"
        # real-valued tiles mixed with integer tiles: the written tolerance
        Xq = Xi.copy()
        Xq[:, K // 2:] = rng.uniform(-300, 300, (M, K - K // 2))
        want = t.spmm(Xq, b, algo=tsg.ALGO_GATHER_SEQ).astype(np.float64)
        got = t.spmm(Xq, b, algo=tsg.ALGO_DENSE_TC).astype(np.float64)
        bound = np.abs(Xq).astype(np.float64) @ np.abs(W).astype(np.float64) + np.abs(b)
        assert np.max(np.abs(got - want) / bound) <= 1e-5
        # inf / NaN / huge values: the tile falls back to the reference's own order (comp.h:44-61)
        Xn = Xi.copy()
        Xn[0, 3], Xn[M - 1, K - 1], Xn[M // 2, K // 2] = np.inf, -np.inf, np.nan
        Xn[M // 3, 7] = 3.0e38
        want = t.spmm(Xn, b, al, algo=tsg.ALGO_GATHER_SEQ)
        got = t.spmm(Xn, b, al, algo=tsg.ALGO_DENSE_TC)
        assert np.array_equal(got, want, equal_nan=True), "non-finite X"
        print("ok")
    """)
    e = dict(os.environ)
    e.update(env)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=e, timeout=600)
    assert p.returncode == 0 and "ok" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]


# ------------------------------------------------------------------------------------------------
# code_gemv beyond two CTAs per SM: one CTA walks several 32-column blocks (software-pipelined
# code loads, double-buffered partial sums, rotating reducer warp)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("K,N,s,seed", [(500, 12011, 4, 3), (1100, 20000, 2, 5), (64, 9473, 3, 7), (4100, 9600, 16, 9)])
@pytest.mark.parametrize("M", [1, 2, 5])
def test_code_gemv_many_column_blocks(tsg, orc, K, N, s, seed, M):
    W = orc.generate_sparse_matrix(K, N, s, seed)
    tref = orc.tcsc(W)
    t = tsg.TCSC(W)
    Xi = orc.init_x(M, K, seed + 1)
    rng = np.random.default_rng(seed)
    b = rng.integers(-8, 9, N).astype(np.float32)
    al = np.full(N, 0.25, np.float32)
    want, wantp = orc.base_tcsc(Xi, tref, b), orc.base_tcsc_prelu(Xi, tref, b, al)
    assert np.array_equal(t.spmm(Xi, b, algo=tsg.ALGO_CODE_GEMV), want)            # integer X: exact
    assert np.array_equal(t.spmm(Xi, b, al, algo=tsg.ALGO_CODE_GEMV), wantp)
    Xr = rng.uniform(-1, 1, (M, K)).astype(np.float32)
    wr = orc.base_tcsc(Xr, tref, b)
    gr = t.spmm(Xr, b, algo=tsg.ALGO_CODE_GEMV)
    assert rel_err(gr, wr) <= REL_TOL
    assert np.all(np.abs(gr.astype(np.float64) - wr.astype(np.float64)) <= 2 * abs_sum_bound(Xr, tref, extra=9.0))


# ------------------------------------------------------------------------------------------------
# host-pointer calls with a large Y are pipelined in row chunks (copy-in / compute / copy-out on
# three streams): same results as the oracle, from pageable and from pinned host buffers
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,prelu", [(300, False), (1030, True)])
def test_large_host_call_pipelined(tsg, orc, M, prelu):
    import torch
    K, N, s = 512, 14336, 4
    W = orc.generate_sparse_matrix(K, N, s, 21)
    tref = orc.tcsc(W)
    t = tsg.TCSC(W)
    Xi = orc.init_x(M, K, 22)
    rng = np.random.default_rng(23)
    b = rng.integers(-8, 9, N).astype(np.float32)
    al = np.full(N, 0.125, np.float32) if prelu else None
    want = orc.base_tcsc_prelu(Xi, tref, b, al) if prelu else orc.base_tcsc(Xi, tref, b)
    assert M * N * 4 >= (16 << 20)                               # takes the chunked path
    assert np.array_equal(t.spmm(Xi, b, al), want)               # pageable numpy buffers
    Xp, bp = torch.from_numpy(Xi).pin_memory(), torch.from_numpy(b).pin_memory()
    ap = torch.from_numpy(al).pin_memory() if prelu else None
    Yp = torch.empty(M, N).pin_memory()
    for _ in range(2):                                            # twice: buffers and events are reused
        Yp.zero_()
        t.spmm_host_ptr(Xp.data_ptr(), bp.data_ptr(), ap.data_ptr() if prelu else None, Yp.data_ptr(), M)
        assert np.array_equal(Yp.numpy(), want)
    assert np.array_equal(t.spmm(Xi, b, al, algo=tsg.ALGO_GATHER), want)


# ------------------------------------------------------------------------------------------------
# decode-sized host-pointer calls (one staging block, one inline copy, Y through mapped memory):
# changing / repeated bias and alpha between calls, K not a multiple of 4, K beyond one batch
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("K,N,s,seed", [(37, 129, 2, 1), (2048, 300, 4, 2), (2050, 4096, 3, 3), (4096, 4096, 3, 4),
                                        (5000, 1000, 8, 5), (7168, 512, 16, 6), (7170, 512, 16, 7)])
def test_decode_fast_path(tsg, orc, K, N, s, seed):
    W = orc.generate_sparse_matrix(K, N, s, seed)
    tref = orc.tcsc(W)
    t = tsg.TCSC(W)
    rng = np.random.default_rng(seed)
    for it in range(4):
        Xi = orc.init_x(1, K, seed + it)
        b = rng.integers(-8, 9, N).astype(np.float32) if it != 1 else b      # it == 1: same bias again
        al = rng.uniform(0.05, 0.5, N).astype(np.float32) if it >= 2 else None
        want = orc.base_tcsc_prelu(Xi, tref, b, al) if al is not None else orc.base_tcsc(Xi, tref, b)
        for algo in (tsg.ALGO_CODE_GEMV, tsg.ALGO_AUTO):
            assert np.array_equal(t.spmm(Xi, b, al, algo=algo), want), (it, algo)
    Xr = rng.uniform(-1, 1, (1, K)).astype(np.float32)
    b[0] += 1.0                                                               # changed in place: must be noticed
    wr = orc.base_tcsc(Xr, tref, b)
    gr = t.spmm(Xr, b, algo=tsg.ALGO_CODE_GEMV)
    assert rel_err(gr, wr) <= REL_TOL
