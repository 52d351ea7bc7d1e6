"""ternary-spgemm_b200 — Python face of libtsg.so (ctypes over the C ABI in include/tsg.h).

The product is the CUDA library; this module is the thin host mirror of the reference's
interface used by tests/ and bench.py (the C++ mirror the reference's driver links against
lives in host/).  Names follow the reference:

    TCSC(W)                         <- class TCSC            cpp_impl/data_structures/TCSC.h:5-50
      .init / .getVectorRepresentation / .getNumRows / .getNumCols
                                    <- DataStructureInterface.hpp:4-14 + readme.md:62-72
      .col_start_pos/.col_start_neg/.row_index_pos/.row_index_neg, .getDataStructureSize()
    BaseTCSC(X, W, b)               <- BaseTCSC<float>       cpp_impl/comp.h:25-69
    BaseTCSC_PreLU(X, W, b, alpha)  <- BaseTCSC_PreLU<float> cpp_impl/comp_prelu.h:12-70

There is no CPU path here: importing works anywhere (so CPU-only checks can inspect the ABI),
but every operation needs libtsg.so AND an sm_100 GPU and raises TsgError otherwise.
The directory name contains '-', so import it with `load_package()` from __graft_entry__ or
`importlib` (tests/conftest.py registers it as `ternary_spgemm_b200`).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TSG_LIB_PATH") or os.path.join(HERE, "libtsg.so")  # override: A/B runs of two builds
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "tsg.h")

ALGO_AUTO, ALGO_GATHER, ALGO_GATHER_SEQ, ALGO_DENSE_TC, ALGO_CODE_GEMV = 0, 1, 2, 3, 4
ALGO_TCSR_SEQ, ALGO_PCSC_GATHER, ALGO_PCSR_SEQ = 5, 6, 7   # format-native kernels of TCSR / PackedCSC / PackedCSR
ALGO_NAMES = {0: "auto", 1: "gather", 2: "gather_seq", 3: "dense_tc", 4: "code_gemv",
              5: "tcsr_seq", 6: "pcsc_gather", 7: "pcsr_seq"}


class TsgError(RuntimeError):
    def __init__(self, status: int, text: str):
        super().__init__(f"libtsg status {status}: {text}")
        self.status = status


_lib = None


def lib() -> C.CDLL:
    """Load libtsg.so (once).  Missing library is a hard error — never a silent fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TsgError(-100, f"{LIB_PATH} not built; run `make -C ternary-spgemm_b200` "
                             "(or __graft_entry__.build())")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    pp = C.POINTER(C.c_void_p)
    L.tsg_abi_version.restype = i32
    L.tsg_last_error.restype = C.c_char_p
    L.tsg_device_count.argtypes = [C.POINTER(i32)]
    L.tsg_device_info.argtypes = [i32, C.POINTER(i32), C.POINTER(i64), C.POINTER(i64), C.c_char_p, i32]
    L.tsg_tcsc_from_dense.argtypes = [vp, i32, i32, pp]
    L.tsg_tcsc_from_dense_cols.argtypes = [vp, i32, i32, i32, i32, pp]
    L.tsg_tcsc_from_dense_dev.argtypes = [vp, i32, i32, i32, i64, i32, i32, vp, pp]
    L.tsg_tcsc_from_arrays.argtypes = [vp, vp, vp, vp, i32, i32, pp]
    L.tsg_tcsc_slice_cols.argtypes = [vp, i32, i32, pp]
    L.tsg_destroy.argtypes = [vp]
    L.tsg_destroy.restype = None
    L.tsg_rows.argtypes = [vp, C.POINTER(i32)]
    L.tsg_cols.argtypes = [vp, C.POINTER(i32)]
    L.tsg_nnz.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    L.tsg_data_structure_size.argtypes = [vp, C.POINTER(i64)]
    L.tsg_tcsc_export.argtypes = [vp, vp, vp, vp, vp]
    L.tsg_tcsc_to_dense.argtypes = [vp, vp]
    L.tsg_spmm.argtypes = [vp, vp, vp, vp, i32, i32, i32]
    L.tsg_spmm_prelu.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32]
    L.tsg_spmm_algo.argtypes = [vp, i32, vp, vp, vp, vp, i32, i32, i32]
    L.tsg_spmm_dev.argtypes = [vp, i32, vp, i64, vp, vp, vp, i64, i32, vp]
    L.tsg_spmm_pick.argtypes = [vp, i32, C.POINTER(i32)]
    L.tsg_launch_count.restype = i64
    L.tsg_spmm_bytes.argtypes = [vp, i32, i32, C.POINTER(i64)]
    try:
        L.tsg_debug_last_build_device_ms.restype = C.c_double
        L.tsg_set_fast_split.argtypes = [i32]
        L.tsg_set_fast_split.restype = i32
        L.tsg_host_store_release_i64.argtypes = [vp, i64]
        L.tsg_host_store_release_i64.restype = None
        L.tsg_host_load_acquire_i64.argtypes = [vp]
        L.tsg_host_load_acquire_i64.restype = i64
    except AttributeError:
        if not os.environ.get("TSG_LIB_PATH"):   # only an older build loaded for an A/B run may lack them
            raise
    L.tsg_blocked_tcsc_export.argtypes = [vp, i32, C.POINTER(i64), C.POINTER(i64), vp, vp, vp, vp]
    L.tsg_tcsr_from_dense.argtypes = [vp, i32, i32, pp]
    L.tsg_tcsr_destroy.argtypes = [vp]
    L.tsg_tcsr_destroy.restype = None
    L.tsg_tcsr_nnz.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    L.tsg_tcsr_data_structure_size.argtypes = [vp, C.POINTER(i64)]
    L.tsg_tcsr_export.argtypes = [vp, vp, vp, vp, vp]
    L.tsg_tcsr_to_dense.argtypes = [vp, vp]
    L.tsg_tcsr_spmm.argtypes = [vp, i32, vp, vp, vp, vp, i32, i32, i32]
    L.tsg_pcsc_from_dense.argtypes = [vp, i32, i32, pp]
    L.tsg_pcsc_from_dense_dev.argtypes = [vp, i32, i32, i32, vp, pp]
    L.tsg_pcsc_from_arrays.argtypes = [vp, vp, vp, i32, i32, pp]
    L.tsg_pcsc_destroy.argtypes = [vp]
    L.tsg_pcsc_destroy.restype = None
    L.tsg_pcsc_sizes.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    L.tsg_pcsc_data_structure_size.argtypes = [vp, C.POINTER(i64)]
    L.tsg_pcsc_export.argtypes = [vp, vp, vp, vp]
    L.tsg_pcsc_to_dense.argtypes = [vp, vp]
    L.tsg_pcsc_spmm.argtypes = [vp, i32, vp, vp, vp, vp, i32, i32, i32]
    L.tsg_pcsc_spmm_dev.argtypes = [vp, i32, vp, i64, vp, vp, vp, i64, i32, vp]
    L.tsg_pcsc_spmm_pick.argtypes = [vp, i32, C.POINTER(i32)]
    L.tsg_pcsr_from_dense.argtypes = [vp, i32, i32, pp]
    L.tsg_pcsr_from_dense_dev.argtypes = [vp, i32, i32, i32, vp, pp]
    L.tsg_pcsr_from_arrays.argtypes = [vp, vp, vp, i32, i32, pp]
    L.tsg_pcsr_destroy.argtypes = [vp]
    L.tsg_pcsr_destroy.restype = None
    L.tsg_pcsr_sizes.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    L.tsg_pcsr_data_structure_size.argtypes = [vp, C.POINTER(i64)]
    L.tsg_pcsr_export.argtypes = [vp, vp, vp, vp]
    L.tsg_pcsr_to_dense.argtypes = [vp, vp]
    L.tsg_pcsr_spmm.argtypes = [vp, i32, vp, vp, vp, vp, i32, i32, i32]
    _lib = L
    return L


def _check(status: int):
    if status != 0:
        raise TsgError(status, lib().tsg_last_error().decode(errors="replace"))


def device_count() -> int:
    n = C.c_int(0)
    _check(lib().tsg_device_count(C.byref(n)))
    return n.value


def device_info(device: int = 0) -> dict:
    sm, l2, hbm = C.c_int(), C.c_int64(), C.c_int64()
    name = C.create_string_buffer(128)
    _check(lib().tsg_device_info(device, C.byref(sm), C.byref(l2), C.byref(hbm), name, 128))
    return {"sm_count": sm.value, "l2_bytes": l2.value, "hbm_bytes": hbm.value,
            "name": name.value.decode()}


def set_fast_split(on: bool) -> bool:
    """Allow two-fp16-term operand tiles on the tensor-core path (include/tsg.h, numerical contract);
    returns the previous setting."""
    return bool(lib().tsg_set_fast_split(1 if on else 0))


def launch_count() -> int:
    return int(lib().tsg_launch_count())


def _ptr(a):
    """Raw address of a numpy array / torch tensor / int / None (plain pointers cross the ABI)."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError(type(a))


class TCSC:
    """Ternary CSC weight resident on the current GPU; same public surface as the reference class."""

    def __init__(self, matrix=None, rows: int | None = None, cols: int | None = None, *,
                 col_range=None):
        self._h = None
        if matrix is not None:
            self.init(matrix, rows, cols, col_range=col_range)

    # DataStructureInterface::init(const int *matrix, int rows, int cols)
    def init(self, matrix, rows=None, cols=None, *, col_range=None):
        W = np.ascontiguousarray(matrix, dtype=np.int32)
        if rows is None:
            rows, cols = W.shape
        assert W.size == rows * cols
        self.close()
        h = C.c_void_p()
        if col_range is None:
            _check(lib().tsg_tcsc_from_dense(W.ctypes.data, rows, cols, C.byref(h)))
        else:
            _check(lib().tsg_tcsc_from_dense_cols(W.ctypes.data, rows, cols, col_range[0],
                                                  col_range[1], C.byref(h)))
        self._h = h
        return self

    @classmethod
    def from_device_dense(cls, W_dev, K, N, *, elem_bytes=4, ld=None, col_range=None, stream=None):
        """W already in HBM (a torch tensor or a raw pointer), int32 or int8, row-major."""
        self = cls()
        lo, hi = col_range if col_range is not None else (0, N)
        h = C.c_void_p()
        _check(lib().tsg_tcsc_from_dense_dev(_ptr(W_dev), elem_bytes, K, N, ld or N, lo, hi,
                                             _ptr(stream), C.byref(h)))
        self._h = h
        return self

    @classmethod
    def from_arrays(cls, col_start_pos, col_start_neg, row_index_pos, row_index_neg, K, N):
        self = cls()
        a = [np.ascontiguousarray(x, dtype=np.int32)
             for x in (col_start_pos, col_start_neg, row_index_pos, row_index_neg)]
        h = C.c_void_p()
        _check(lib().tsg_tcsc_from_arrays(*[x.ctypes.data for x in a], K, N, C.byref(h)))
        self._h = h
        return self

    def slice_cols(self, lo, hi) -> "TCSC":
        out = TCSC()
        h = C.c_void_p()
        _check(lib().tsg_tcsc_slice_cols(self._h, lo, hi, C.byref(h)))
        out._h = h
        return out

    def close(self):
        if self._h is not None and _lib is not None:
            _lib.tsg_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # README getNumRows / getNumCols
    def getNumRows(self) -> int:
        v = C.c_int()
        _check(lib().tsg_rows(self._h, C.byref(v)))
        return v.value

    def getNumCols(self) -> int:
        v = C.c_int()
        _check(lib().tsg_cols(self._h, C.byref(v)))
        return v.value

    @property
    def nnz(self):
        p, q = C.c_int64(), C.c_int64()
        _check(lib().tsg_nnz(self._h, C.byref(p), C.byref(q)))
        return p.value, q.value

    def getDataStructureSize(self) -> int:
        v = C.c_int64()
        _check(lib().tsg_data_structure_size(self._h, C.byref(v)))
        return v.value

    def spmm_bytes(self, M: int, prelu: bool = False) -> int:
        v = C.c_int64()
        _check(lib().tsg_spmm_bytes(self._h, M, int(prelu), C.byref(v)))
        return v.value

    def export(self):
        N = self.getNumCols()
        p, q = self.nnz
        csp, csn = np.empty(N + 1, np.int32), np.empty(N + 1, np.int32)
        rip, rin = np.empty(p, np.int32), np.empty(q, np.int32)
        _check(lib().tsg_tcsc_export(self._h, csp.ctypes.data, csn.ctypes.data, rip.ctypes.data,
                                     rin.ctypes.data))
        return csp, csn, rip, rin

    col_start_pos = property(lambda s: s.export()[0])
    col_start_neg = property(lambda s: s.export()[1])
    row_index_pos = property(lambda s: s.export()[2])
    row_index_neg = property(lambda s: s.export()[3])

    # DataStructureInterface::getVectorRepresentation(size_t rows, size_t cols)
    def getVectorRepresentation(self, rows=None, cols=None) -> np.ndarray:
        K, N = self.getNumRows(), self.getNumCols()
        if rows is not None and (rows, cols) != (K, N):
            raise TsgError(-1, f"expected shape {(rows, cols)} but matrix is {(K, N)}")
        W = np.empty((K, N), np.int32)
        _check(lib().tsg_tcsc_to_dense(self._h, W.ctypes.data))
        return W

    def blocked(self, B: int = 512):
        """BlockedTCSC<B> arrays of the same W (BlockedTCSC.h:15-43), built on the device."""
        p, q = C.c_int64(), C.c_int64()
        _check(lib().tsg_blocked_tcsc_export(self._h, B, C.byref(p), C.byref(q), None, None, None, None))
        pairs = (self.getNumRows() // B) * self.getNumCols()
        csp, csn = np.empty(pairs + 1, np.int32), np.empty(pairs + 1, np.int32)
        rip, rin = np.empty(p.value, np.int32), np.empty(q.value, np.int32)
        _check(lib().tsg_blocked_tcsc_export(self._h, B, None, None, csp.ctypes.data, csn.ctypes.data,
                                             rip.ctypes.data, rin.ctypes.data))
        return csp, csn, rip, rin

    def pick(self, M: int) -> int:
        v = C.c_int()
        _check(lib().tsg_spmm_pick(self._h, M, C.byref(v)))
        return v.value

    # ---- compute ---------------------------------------------------------------------------
    def spmm(self, X, b, alpha=None, *, algo=ALGO_AUTO, out=None) -> np.ndarray:
        """Host-pointer call (what a comp_func lambda does): numpy in, numpy out, synchronous."""
        X = np.ascontiguousarray(X, np.float32)
        b = np.ascontiguousarray(b, np.float32)
        M, K = X.shape
        N = self.getNumCols()
        Y = out if out is not None else np.empty((M, N), np.float32)
        a = None if alpha is None else np.ascontiguousarray(alpha, np.float32)
        _check(lib().tsg_spmm_algo(self._h, algo, X.ctypes.data, b.ctypes.data,
                                   None if a is None else a.ctypes.data, Y.ctypes.data, M, N, K))
        return Y

    def spmm_host_ptr(self, X_ptr, b_ptr, alpha_ptr, Y_ptr, M, *, algo=ALGO_AUTO):
        """Raw host pointers (e.g. pinned torch tensors) — the end-to-end path bench.py times.
        One C call per step: the shape is looked up once per handle."""
        kn = getattr(self, "_kn", None)
        if kn is None or kn[0] is not self._h:
            kn = self._kn = (self._h, self.getNumRows(), self.getNumCols())
        status = _lib.tsg_spmm_algo(self._h, algo, X_ptr, b_ptr, alpha_ptr, Y_ptr, M, kn[2], kn[1])
        if status != 0:
            _check(status)

    def spmm_dev(self, X, b, Y, M, *, alpha=None, algo=ALGO_AUTO, ldx=None, ldy=None, stream=None):
        """Device pointers (torch CUDA tensors or ints); enqueues on `stream`, does not block."""
        _check(lib().tsg_spmm_dev(self._h, algo, _ptr(X), ldx or self.getNumRows(), _ptr(b),
                                  _ptr(alpha), _ptr(Y), ldy or self.getNumCols(), M, _ptr(stream)))


class _FormatBase:
    """Shared plumbing of the TCSR / PackedCSC handles (same DataStructureInterface surface)."""
    _prefix = ""

    def __init__(self, matrix=None, rows=None, cols=None):
        self._h, self._K, self._N = None, 0, 0
        if matrix is not None:
            self.init(matrix, rows, cols)

    def _fn(self, name):
        return getattr(lib(), f"tsg_{self._prefix}_{name}")

    def init(self, matrix, rows=None, cols=None):
        W = np.ascontiguousarray(matrix, dtype=np.int32)
        if rows is None:
            rows, cols = W.shape
        assert W.size == rows * cols
        self.close()
        h = C.c_void_p()
        _check(self._fn("from_dense")(W.ctypes.data, rows, cols, C.byref(h)))
        self._h, self._K, self._N = h, rows, cols
        return self

    def close(self):
        if self._h is not None and _lib is not None:
            self._fn("destroy")(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def getNumRows(self):
        return self._K

    def getNumCols(self):
        return self._N

    def getDataStructureSize(self) -> int:
        v = C.c_int64()
        _check(self._fn("data_structure_size")(self._h, C.byref(v)))
        return v.value

    def getVectorRepresentation(self, rows=None, cols=None) -> np.ndarray:
        if rows is not None and (rows, cols) != (self._K, self._N):
            raise TsgError(-1, f"expected shape {(rows, cols)} but matrix is {(self._K, self._N)}")
        W = np.empty((self._K, self._N), np.int32)
        _check(self._fn("to_dense")(self._h, W.ctypes.data))
        return W

    def spmm(self, X, b, alpha=None, *, algo=ALGO_AUTO) -> np.ndarray:
        X = np.ascontiguousarray(X, np.float32)
        b = np.ascontiguousarray(b, np.float32)
        M, K = X.shape
        Y = np.empty((M, self._N), np.float32)
        a = None if alpha is None else np.ascontiguousarray(alpha, np.float32)
        _check(self._fn("spmm")(self._h, algo, X.ctypes.data, b.ctypes.data,
                                None if a is None else a.ctypes.data, Y.ctypes.data, M, self._N, K))
        return Y


class TCSR(_FormatBase):
    """Ternary CSR built on the GPU — class TCSR, cpp_impl/data_structures/TCSR.h:5-50."""
    _prefix = "tcsr"

    @property
    def nnz(self):
        p, q = C.c_int64(), C.c_int64()
        _check(lib().tsg_tcsr_nnz(self._h, C.byref(p), C.byref(q)))
        return p.value, q.value

    def export(self):
        p, q = self.nnz
        rsp, rsn = np.empty(self._K + 1, np.int32), np.empty(self._K + 1, np.int32)
        cip, cin = np.empty(p, np.int32), np.empty(q, np.int32)
        _check(lib().tsg_tcsr_export(self._h, rsp.ctypes.data, rsn.ctypes.data, cip.ctypes.data,
                                     cin.ctypes.data))
        return rsp, rsn, cip, cin

    row_start_pos = property(lambda s: s.export()[0])
    row_start_neg = property(lambda s: s.export()[1])
    col_index_pos = property(lambda s: s.export()[2])
    col_index_neg = property(lambda s: s.export()[3])


class PackedCSC(_FormatBase):
    """Packed-value CSC (5 signs per byte; readme.md:108-111) built on the GPU."""
    _prefix = "pcsc"

    @classmethod
    def from_device_dense(cls, W_dev, K, N, *, elem_bytes=4, stream=None):
        self = cls()
        h = C.c_void_p()
        _check(lib().tsg_pcsc_from_dense_dev(_ptr(W_dev), elem_bytes, K, N, _ptr(stream), C.byref(h)))
        self._h, self._K, self._N = h, K, N
        return self

    @classmethod
    def from_arrays(cls, col_ptr, row_idx, vals, K, N):
        self = cls()
        cp = np.ascontiguousarray(col_ptr, np.int32)
        ri = np.ascontiguousarray(row_idx, np.int32)
        vv = np.ascontiguousarray(vals, np.uint8)
        h = C.c_void_p()
        _check(lib().tsg_pcsc_from_arrays(cp.ctypes.data, ri.ctypes.data, vv.ctypes.data, K, N, C.byref(h)))
        self._h, self._K, self._N = h, K, N
        return self

    @property
    def sizes(self):
        n, b = C.c_int64(), C.c_int64()
        _check(lib().tsg_pcsc_sizes(self._h, C.byref(n), C.byref(b)))
        return n.value, b.value

    def export(self):
        nnz, nb = self.sizes
        cp, ri, vv = np.empty(self._N + 1, np.int32), np.empty(nnz, np.int32), np.empty(nb, np.uint8)
        _check(lib().tsg_pcsc_export(self._h, cp.ctypes.data, ri.ctypes.data, vv.ctypes.data))
        return cp, ri, vv

    def pick(self, M: int) -> int:
        v = C.c_int()
        _check(lib().tsg_pcsc_spmm_pick(self._h, M, C.byref(v)))
        return v.value

    def spmm_host_ptr(self, X_ptr, b_ptr, alpha_ptr, Y_ptr, M, *, algo=ALGO_AUTO):
        """Raw host pointers (pinned torch tensors): the end-to-end path bench.py times."""
        status = _lib.tsg_pcsc_spmm(self._h, algo, X_ptr, b_ptr, alpha_ptr, Y_ptr, M, self._N, self._K)
        if status != 0:
            _check(status)

    def spmm_dev(self, X, b, Y, M, *, alpha=None, algo=ALGO_AUTO, ldx=None, ldy=None, stream=None):
        _check(lib().tsg_pcsc_spmm_dev(self._h, algo, _ptr(X), ldx or self._K, _ptr(b), _ptr(alpha),
                                       _ptr(Y), ldy or self._N, M, _ptr(stream)))


class PackedCSR(_FormatBase):
    """Packed-value CSR (the row-major twin of PackedCSC; readme.md:108-111) built on the GPU."""
    _prefix = "pcsr"

    @classmethod
    def from_device_dense(cls, W_dev, K, N, *, elem_bytes=4, stream=None):
        self = cls()
        h = C.c_void_p()
        _check(lib().tsg_pcsr_from_dense_dev(_ptr(W_dev), elem_bytes, K, N, _ptr(stream), C.byref(h)))
        self._h, self._K, self._N = h, K, N
        return self

    @classmethod
    def from_arrays(cls, row_ptr, col_idx, vals, K, N):
        self = cls()
        rp = np.ascontiguousarray(row_ptr, np.int32)
        ci = np.ascontiguousarray(col_idx, np.int32)
        vv = np.ascontiguousarray(vals, np.uint8)
        h = C.c_void_p()
        _check(lib().tsg_pcsr_from_arrays(rp.ctypes.data, ci.ctypes.data, vv.ctypes.data, K, N, C.byref(h)))
        self._h, self._K, self._N = h, K, N
        return self

    @property
    def sizes(self):
        n, b = C.c_int64(), C.c_int64()
        _check(lib().tsg_pcsr_sizes(self._h, C.byref(n), C.byref(b)))
        return n.value, b.value

    def export(self):
        nnz, nb = self.sizes
        rp, ci, vv = np.empty(self._K + 1, np.int32), np.empty(nnz, np.int32), np.empty(nb, np.uint8)
        _check(lib().tsg_pcsr_export(self._h, rp.ctypes.data, ci.ctypes.data, vv.ctypes.data))
        return rp, ci, vv


def BaseTCSR(X, W: TCSR, b, *, algo=ALGO_AUTO) -> np.ndarray:
    """Y = X·W + b  (reference BaseTCSR<float>, comp.h:478-528)."""
    return W.spmm(X, b, algo=algo)


def BaseTCSC(X, W: TCSC, b, *, algo=ALGO_AUTO) -> np.ndarray:
    """Y = X·W + b  (reference BaseTCSC<float>, comp.h:25-69)."""
    return W.spmm(X, b, algo=algo)


def BaseTCSC_PreLU(X, W: TCSC, b, alpha, *, algo=ALGO_AUTO) -> np.ndarray:
    """Fused bias + PReLU (reference BaseTCSC_PreLU<float>, comp_prelu.h:12-70)."""
    return W.spmm(X, b, alpha, algo=algo)


def shard_columns(N: int, world: int, rank: int) -> tuple[int, int]:
    """Column range rank `rank` of `world` owns under N-sharding (contiguous, near-equal)."""
    return (N * rank) // world, (N * (rank + 1)) // world
