"""ctypes/numpy face of the CHECKER (oracle/liboracle.so and, when present, oracle/_ref/libtsgref.so).

TEST INFRASTRUCTURE ONLY.  Import this from tests/, from ``__graft_entry__.smoke()`` and from
bench.py's CPU legs (``cpu_baseline`` / ``--impl reference``) — never from ternary-spgemm_b200/.

``Oracle``    — the C restatement (oracle/tsg_oracle.c), always available after ``make -C oracle``.
``Reference`` — the unmodified reference compiled in place (oracle/ref_shim.cpp); available when
                oracle/_ref/libtsgref.so exists (built in the container that has /root/reference,
                shipped prebuilt to the GPU box).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libtsgref.so")
REF_DRIVER = os.path.join(HERE, "_ref", "sparseGEMM_ref.out")

_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_ll = C.c_longlong
_llp = C.POINTER(C.c_longlong)


def build(force: bool = False) -> None:
    """(Re)build the checker libraries with oracle/Makefile (gcc only; no GPU needed)."""
    if force or not os.path.exists(ORACLE_SO) or (
        os.path.exists("/root/reference/cpp_impl/comp.h") and not os.path.exists(REF_SO)
    ):
        subprocess.run(["make", "-C", HERE, "all"], check=True, capture_output=True)


def have_reference() -> bool:
    return os.path.exists(REF_SO)


class _Tcsc:
    """Plain holder of the four TCSC arrays (names follow cpp_impl/data_structures/TCSC.h:8-11)."""

    def __init__(self, csp, csn, rip, rin, rows, cols):
        self.col_start_pos, self.col_start_neg = csp, csn
        self.row_index_pos, self.row_index_neg = rip, rin
        self.rows, self.cols = rows, cols

    @property
    def arrays(self):
        return self.col_start_pos, self.col_start_neg, self.row_index_pos, self.row_index_neg

    @property
    def nnz(self):
        return int(self.row_index_pos.size + self.row_index_neg.size)


class Oracle:
    def __init__(self):
        build()
        L = self.lib = C.CDLL(ORACLE_SO)
        L.orc_generate_sparse_matrix.argtypes = [_i32p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_init_x.argtypes = [_f32p, _ll, C.c_int, C.c_uint32]
        L.orc_tcsc_count.argtypes = [_i32p, C.c_int, C.c_int, _llp, _llp]
        L.orc_tcsc_build.argtypes = [_i32p, C.c_int, C.c_int, _i32p, _i32p, _i32p, _i32p]
        L.orc_tcsc_to_dense.argtypes = [_i32p, _i32p, _i32p, _i32p, C.c_int, C.c_int, _i32p]
        L.orc_tcsc_size_bytes.argtypes = [C.c_int, _ll, _ll]
        L.orc_tcsc_size_bytes.restype = _ll
        kern = [_f32p, _i32p, _i32p, _i32p, _i32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int]
        L.orc_base_tcsc.argtypes = kern
        L.orc_double_unrolled_tcsc_k4_m4.argtypes = kern
        L.orc_base_tcsr.argtypes = kern
        L.orc_base_blocked.argtypes = kern + [C.c_int]
        L.orc_base_tcsc_prelu.argtypes = [_f32p, _i32p, _i32p, _i32p, _i32p, _f32p, _f32p, _f32p,
                                          C.c_int, C.c_int, C.c_int]
        L.orc_gemm.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int]
        L.orc_gemm_prelu.argtypes = [_f32p, _f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int]
        L.orc_compare_results.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.POINTER(C.c_int),
                                          C.POINTER(C.c_int)]
        L.orc_compare_results.restype = C.c_int
        L.orc_tcsr_count.argtypes = [_i32p, C.c_int, C.c_int, _llp, _llp]
        L.orc_tcsr_build.argtypes = [_i32p, C.c_int, C.c_int, _i32p, _i32p, _i32p, _i32p]
        L.orc_blocked_count.argtypes = [_i32p, C.c_int, C.c_int, C.c_int, _llp, _llp]
        L.orc_blocked_build.argtypes = [_i32p, C.c_int, C.c_int, C.c_int, _i32p, _i32p, _i32p, _i32p]

    # -- inputs ------------------------------------------------------------------------------
    def generate_sparse_matrix(self, K, N, s, seed) -> np.ndarray:
        W = np.empty((K, N), dtype=np.int32)
        self.lib.orc_generate_sparse_matrix(W, K, N, s, seed)
        return W

    def init_x(self, M, K, seed, rng=512) -> np.ndarray:
        X = np.empty((M, K), dtype=np.float32)
        self.lib.orc_init_x(X, M * K, rng, seed)
        return X

    # -- TCSC --------------------------------------------------------------------------------
    def tcsc(self, W: np.ndarray) -> _Tcsc:
        W = np.ascontiguousarray(W, dtype=np.int32)
        K, N = W.shape
        p, q = _ll(), _ll()
        self.lib.orc_tcsc_count(W, K, N, C.byref(p), C.byref(q))
        csp = np.empty(N + 1, np.int32)
        csn = np.empty(N + 1, np.int32)
        rip = np.empty(p.value, np.int32)
        rin = np.empty(q.value, np.int32)
        self.lib.orc_tcsc_build(W, K, N, csp, csn, rip, rin)
        return _Tcsc(csp, csn, rip, rin, K, N)

    def tcsc_to_dense(self, t: _Tcsc) -> np.ndarray:
        W = np.empty((t.rows, t.cols), np.int32)
        self.lib.orc_tcsc_to_dense(*t.arrays, t.rows, t.cols, W)
        return W

    def tcsc_size_bytes(self, t: _Tcsc) -> int:
        return int(self.lib.orc_tcsc_size_bytes(t.cols, t.row_index_pos.size, t.row_index_neg.size))

    # -- kernels -----------------------------------------------------------------------------
    def _run(self, fn, X, t, b, *extra):
        X = np.ascontiguousarray(X, np.float32)
        M, K = X.shape
        Y = np.zeros((M, t.cols), np.float32)
        fn(X, *t.arrays, np.ascontiguousarray(b, np.float32), Y, M, t.cols, K, *extra)
        return Y

    def base_tcsc(self, X, t, b):
        return self._run(self.lib.orc_base_tcsc, X, t, b)

    def double_unrolled_tcsc_k4_m4(self, X, t, b):
        return self._run(self.lib.orc_double_unrolled_tcsc_k4_m4, X, t, b)

    def base_tcsc_prelu(self, X, t, b, alpha):
        X = np.ascontiguousarray(X, np.float32)
        M, K = X.shape
        Y = np.zeros((M, t.cols), np.float32)
        self.lib.orc_base_tcsc_prelu(X, *t.arrays, np.ascontiguousarray(b, np.float32),
                                     np.ascontiguousarray(alpha, np.float32), Y, M, t.cols, K)
        return Y

    def gemm(self, X, W, b):
        X = np.ascontiguousarray(X, np.float32)
        Wf = np.ascontiguousarray(W, np.float32)
        M, K = X.shape
        N = Wf.shape[1]
        Y = np.zeros((M, N), np.float32)
        self.lib.orc_gemm(X, Wf, np.ascontiguousarray(b, np.float32), Y, M, N, K)
        return Y

    def gemm_prelu(self, X, W, b, alpha):
        X = np.ascontiguousarray(X, np.float32)
        Wf = np.ascontiguousarray(W, np.float32)
        M, K = X.shape
        N = Wf.shape[1]
        Y = np.zeros((M, N), np.float32)
        self.lib.orc_gemm_prelu(X, Wf, np.ascontiguousarray(b, np.float32),
                                np.ascontiguousarray(alpha, np.float32), Y, M, N, K)
        return Y

    def compare_results(self, result, truth) -> bool:
        r = np.ascontiguousarray(result, np.float32)
        g = np.ascontiguousarray(truth, np.float32)
        H, W = r.shape
        return bool(self.lib.orc_compare_results(r, g, H, W, None, None))

    # -- TCSR / BlockedTCSC ------------------------------------------------------------------
    def tcsr(self, W: np.ndarray) -> _Tcsc:
        W = np.ascontiguousarray(W, dtype=np.int32)
        K, N = W.shape
        p, q = _ll(), _ll()
        self.lib.orc_tcsr_count(W, K, N, C.byref(p), C.byref(q))
        rsp = np.empty(K + 1, np.int32)
        rsn = np.empty(K + 1, np.int32)
        cip = np.empty(p.value, np.int32)
        cin = np.empty(q.value, np.int32)
        self.lib.orc_tcsr_build(W, K, N, rsp, rsn, cip, cin)
        return _Tcsc(rsp, rsn, cip, cin, K, N)

    def base_tcsr(self, X, t, b):
        return self._run(self.lib.orc_base_tcsr, X, t, b)

    def blocked(self, W: np.ndarray, B: int) -> _Tcsc:
        W = np.ascontiguousarray(W, dtype=np.int32)
        K, N = W.shape
        p, q = _ll(), _ll()
        self.lib.orc_blocked_count(W, K, N, B, C.byref(p), C.byref(q))
        nptr = (K // B) * N + 1
        csp = np.empty(nptr, np.int32)
        csn = np.empty(nptr, np.int32)
        rip = np.empty(p.value, np.int32)
        rin = np.empty(q.value, np.int32)
        self.lib.orc_blocked_build(W, K, N, B, csp, csn, rip, rin)
        return _Tcsc(csp, csn, rip, rin, K, N)

    def base_blocked(self, X, t, b, B):
        return self._run(self.lib.orc_base_blocked, X, t, b, B)


class _RefHandle:
    def __init__(self, lib, h, free, rows, cols):
        self.lib, self.h, self._free, self.rows, self.cols = lib, h, free, rows, cols

    def __del__(self):
        if self.h:
            self._free(self.h)
            self.h = None


class Reference:
    """The unmodified reference behind oracle/ref_shim.cpp."""

    def __init__(self):
        build()
        if not have_reference():
            raise FileNotFoundError(REF_SO)
        L = self.lib = C.CDLL(REF_SO)
        L.ref_generate_sparse_matrix.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _i32p]
        for fam in ("tcsc", "tcsr", "blocked512"):
            getattr(L, f"ref_{fam}_new").argtypes = [_i32p, C.c_int, C.c_int]
            getattr(L, f"ref_{fam}_new").restype = C.c_void_p
            getattr(L, f"ref_{fam}_free").argtypes = [C.c_void_p]
            getattr(L, f"ref_{fam}_counts").argtypes = [C.c_void_p, _llp, _llp, _llp]
            getattr(L, f"ref_{fam}_export").argtypes = [C.c_void_p, _i32p, _i32p, _i32p, _i32p]
        L.ref_tcsc_size_bytes.argtypes = [C.c_void_p]
        kern = [C.c_void_p, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int]
        for name in ("ref_base_tcsc", "ref_unrolled_tcsc_12", "ref_double_unrolled_tcsc_k4_m4",
                     "ref_base_tcsr", "ref_base_blocked512"):
            getattr(L, name).argtypes = kern
        L.ref_base_tcsc_prelu.argtypes = [C.c_void_p, _f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int,
                                          C.c_int]
        L.ref_gemm.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int]
        L.ref_gemm_prelu.argtypes = [_f32p, _f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int]
        L.ref_compare_results.argtypes = [_f32p, _f32p, C.c_int, C.c_int]

    def generate_sparse_matrix(self, K, N, s, seed) -> np.ndarray:
        W = np.empty((K, N), dtype=np.int32)
        self.lib.ref_generate_sparse_matrix(K, N, s, seed, W)
        return W

    def _new(self, fam, W):
        W = np.ascontiguousarray(W, dtype=np.int32)
        K, N = W.shape
        h = getattr(self.lib, f"ref_{fam}_new")(W, K, N)
        return _RefHandle(self.lib, h, getattr(self.lib, f"ref_{fam}_free"), K, N)

    def _export(self, fam, hd) -> _Tcsc:
        a, p, q = _ll(), _ll(), _ll()
        getattr(self.lib, f"ref_{fam}_counts")(hd.h, C.byref(a), C.byref(p), C.byref(q))
        c0 = np.empty(a.value, np.int32)
        c1 = np.empty(a.value, np.int32)
        r0 = np.empty(p.value, np.int32)
        r1 = np.empty(q.value, np.int32)
        getattr(self.lib, f"ref_{fam}_export")(hd.h, c0, c1, r0, r1)
        return _Tcsc(c0, c1, r0, r1, hd.rows, hd.cols)

    def tcsc_handle(self, W):
        return self._new("tcsc", W)

    def tcsc(self, W) -> _Tcsc:
        return self._export("tcsc", self._new("tcsc", W))

    def tcsr(self, W) -> _Tcsc:
        return self._export("tcsr", self._new("tcsr", W))

    def blocked512(self, W) -> _Tcsc:
        return self._export("blocked512", self._new("blocked512", W))

    def tcsc_size_bytes(self, hd) -> int:
        return int(self.lib.ref_tcsc_size_bytes(hd.h))

    def _run(self, name, hd, X, b, alpha=None, zero=True):
        X = np.ascontiguousarray(X, np.float32)
        M, K = X.shape
        Y = np.zeros((M, hd.cols), np.float32)
        b = np.ascontiguousarray(b, np.float32)
        if alpha is None:
            getattr(self.lib, name)(hd.h, X, b, Y, M, hd.cols, K)
        else:
            getattr(self.lib, name)(hd.h, X, b, np.ascontiguousarray(alpha, np.float32), Y, M,
                                    hd.cols, K)
        return Y

    def base_tcsc(self, hd, X, b):
        return self._run("ref_base_tcsc", hd, X, b)

    def base_tcsc_prelu(self, hd, X, b, alpha):
        return self._run("ref_base_tcsc_prelu", hd, X, b, alpha)

    def unrolled_tcsc_12(self, hd, X, b):
        return self._run("ref_unrolled_tcsc_12", hd, X, b)

    def double_unrolled_tcsc_k4_m4(self, hd, X, b):
        return self._run("ref_double_unrolled_tcsc_k4_m4", hd, X, b)

    def base_tcsr(self, W, X, b):
        return self._run("ref_base_tcsr", self._new("tcsr", W), X, b)

    def base_blocked512(self, W, X, b):
        return self._run("ref_base_blocked512", self._new("blocked512", W), X, b)

    def gemm(self, X, W, b):
        X = np.ascontiguousarray(X, np.float32)
        Wf = np.ascontiguousarray(W, np.float32)
        M, K = X.shape
        N = Wf.shape[1]
        Y = np.zeros((M, N), np.float32)
        self.lib.ref_gemm(X, Wf, np.ascontiguousarray(b, np.float32), Y, M, N, K)
        return Y

    def gemm_prelu(self, X, W, b, alpha):
        X = np.ascontiguousarray(X, np.float32)
        Wf = np.ascontiguousarray(W, np.float32)
        M, K = X.shape
        N = Wf.shape[1]
        Y = np.zeros((M, N), np.float32)
        self.lib.ref_gemm_prelu(X, Wf, np.ascontiguousarray(b, np.float32),
                                np.ascontiguousarray(alpha, np.float32), Y, M, N, K)
        return Y

    def compare_results(self, result, truth) -> bool:
        r = np.ascontiguousarray(result, np.float32)
        g = np.ascontiguousarray(truth, np.float32)
        return bool(self.lib.ref_compare_results(r, g, r.shape[0], r.shape[1]))
