/* include/tsg.h — C ABI of libtsg.so, the B200 (sm_100a) ternary sparse-GEMM engine.
 *
 * This is the drop-in boundary for the reference's hot path  Y = X·W + b  (optionally PReLU),
 * W a K×N matrix over {-1,0,+1}.  Plain pointers and sizes only; no C++/torch types.  Every
 * entry point names the reference interface it replaces (file:line under the reference tree
 * alessiomelone/Ternary-spGEMM).  INTEGRATION.md shows the reference-side binding.
 *
 * Conventions
 *   - every function returns tsg_status (0 = ok, <0 = error); tsg_last_error() gives the text of
 *     the calling thread's last failure.  The reference's functions return void and cannot fail
 *     (cpp_impl/common.h:12-13); the C++ host mirror turns a non-zero status into abort().
 *   - a tsg_matrix lives on ONE GPU (the device current when it was created).  Handles are
 *     thread-compatible, not thread-safe — same contract as the reference's single-threaded
 *     driver (cpp_impl/perf.cpp:63-66).
 *   - "host" entry points take host pointers, copy, run and block until Y is complete: that is
 *     what a comp_func lambda must do because the reference reads Y (cpp_impl/main.cpp:214-216)
 *     and the TSC (cpp_impl/perf.cpp:62-69) right after the call returns.
 *   - "_dev" entry points take device pointers and a cudaStream_t (as void*) and do not block.
 *   - there is NO CPU fallback: without a usable sm_100 device every call fails with
 *     TSG_ERR_NO_DEVICE.
 *
 * Numerical contract of the SpMM entry points (what "drop-in for BaseTCSC" means here)
 *   - integer-valued X (the reference's own regime, sparseUtils.h:6-23) with partial sums below
 *     2^24: bit-identical to BaseTCSC for every kernel.
 *   - general fp32 X: TSG_ALGO_GATHER_SEQ is bit-identical to BaseTCSC (same order, one fp32
 *     accumulator); the re-ordered kernels agree within 1e-5 of max|Y| (BASELINE.json) and within
 *     the forward bound (n+2)·eps·(Σ|x| + |b|) per element.  The tensor-core kernel multiplies W by
 *     an EXACT split of x into 16-bit terms, chosen per tile of X (rows of an m-tile x 64 k), and
 *     accumulates in fp32: values exact in fp16 -> one fp16 term; values with at most 16 significant
 *     bits -> one or two bf16 terms; anything else -> three bf16 terms (all products exact for
 *     2^-110 <= |x| < 2^100; smaller magnitudes lose at most 2^-133 per element).  The tensor
 *     core's fp32 accumulator truncates instead of rounding to nearest, so its summation error grows
 *     with the number of 16-deep accumulation steps rather than its square root; the kernel adds the
 *     small split terms first to keep that down.  Measured worst case over random shapes and scales
 *     (tools/fuzz.py): 2.6e-6 of Σ|x||w| + |b|, for K = 8192 with integer tiles of ±512 next to
 *     tiles of magnitude 1e-3; uniform data stays below 4e-7 (the reference-order kernel: 4e-7).
 *     Results are deterministic for a given handle and M (fixed reduction orders everywhere); a
 *     different M may associate the fp32 sums differently (tile height and K-splits follow the grid),
 *     so a host-pointer call that runs M in row chunks agrees with the one-piece device call to
 *     rounding, not bit for bit, on real-valued X — bit for bit on integer-valued X.
 *     Opt-in (tsg_set_fast_split(1), or TSG_TC_FAST=1 in the environment): full-precision values in a
 *     tile whose largest magnitude lies in [2^-4, 65520) travel as TWO fp16 terms instead of three
 *     bf16 terms — x carried with |error| <= max(2^-24 |x|, 2^-25), two thirds of the tensor work;
 *     measured at the full BASELINE sizes the result is as close to BaseTCSC as with the exact
 *     split (1.1e-6 of max|Y| at M=2048 K=8192 N=28672: the noise between two fp32 summation orders).
 *   - non-finite or huge X: the reference's sparse sum (comp.h:44-61) never touches x where W is 0,
 *     a dense product would compute 0·x.  Every kernel that multiplies (dense_tc, code_gemv) tests
 *     its X for inf / NaN / |x| >= 2^100 while staging it and recomputes the affected tile in the
 *     reference's own order, so Y is what BaseTCSC gives — an inf or NaN only where W is non-zero.
 *   - streams: `_dev` calls on ONE stream are ordered by it.  The tensor-core path keeps one
 *     operand scratch per handle; a call arriving on a different stream first waits (on the host)
 *     for the previous stream to drain.  A captured CUDA graph containing calls on a handle must
 *     not be replayed concurrently with other calls on the same handle, and a call that has to
 *     GROW the scratch (larger M than any before) or build the gather kernel's lists (first
 *     TSG_ALGO_GATHER call) cannot be captured: TSG_ERR_UNSUPPORTED — run one such call first.
 */
#ifndef TSG_H
#define TSG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSG_ABI_VERSION 1

typedef enum tsg_status
{
    TSG_OK = 0,
    TSG_ERR_INVALID = -1,     /* bad argument / shape mismatch with the handle            */
    TSG_ERR_CUDA = -2,        /* a CUDA runtime call failed (text in tsg_last_error)       */
    TSG_ERR_NO_DEVICE = -3,   /* no CUDA device, or not compute capability 10.x            */
    TSG_ERR_NOMEM = -4,       /* device or host allocation failed                          */
    TSG_ERR_UNSUPPORTED = -5, /* shape outside what the selected kernel supports           */
    TSG_ERR_OVERFLOW = -6     /* nnz or K*N does not fit the reference's int32 indexing    */
} tsg_status;

/* Kernel selection for the SpMM entry points. */
typedef enum tsg_algo
{
    TSG_ALGO_AUTO = 0,       /* engine picks (recorded crossover)                         */
    TSG_ALGO_GATHER = 1,     /* TCSC index-stream gather-add kernel (small M; HBM-bound; 16-bit row  */
                             /* ids when K <= 65535: half the bytes of the int32 stream)             */
    TSG_ALGO_GATHER_SEQ = 2, /* one thread per Y[m,n], reference summation ORDER: bit-identical to   */
                             /* BaseTCSC for any fp32 input (slow; the on-device parity anchor)      */
    TSG_ALGO_DENSE_TC = 3,   /* 2-bit codes expanded in registers -> TMEM -> tcgen05.mma, fp32       */
                             /* accumulators in TMEM (any M; the default)                            */
    TSG_ALGO_CODE_GEMV = 4,  /* 2-bit code stream on the FMA pipe, for one or two rows of X (decode) */
    TSG_ALGO_TCSR_SEQ = 5,   /* tsg_tcsr only: BaseTCSR's arithmetic and summation order on the GPU  */
    TSG_ALGO_PCSC_GATHER = 6, /* tsg_pcsc only: gather kernel over the packed-value CSC stream itself */
    TSG_ALGO_PCSR_SEQ = 7    /* tsg_pcsr only: BaseTCSR's summation order over the packed-value CSR rows */
} tsg_algo;

typedef struct tsg_matrix tsg_matrix; /* opaque: one ternary weight matrix resident in HBM */

/* ---- library / device ------------------------------------------------------------------- */
int tsg_abi_version(void);
const char *tsg_last_error(void);
/* number of usable sm_100 devices (0 is not an error here) */
int tsg_device_count(int *count);
/* sm count, L2 bytes, total HBM bytes and name of `device`; any out pointer may be NULL */
int tsg_device_info(int device, int *sm_count, int64_t *l2_bytes, int64_t *hbm_bytes,
                    char *name, int name_len);

/* ---- (a1) format construction — replaces TCSC::TCSC(const int*,int,int) ----------------------
 * reference: cpp_impl/data_structures/TCSC.h:13-41.  W is the reference's dense row-major
 * int32 K×N matrix (K = rows, N = cols); values other than +1/-1 count as 0.  The four arrays
 * are built ON THE DEVICE and are bit-identical to the reference constructor's vectors. */
int tsg_tcsc_from_dense(const int32_t *W_host, int K, int N, tsg_matrix **out);

/* Same, but only columns [col_lo, col_hi) of W: the N-column shard one GPU owns in the
 * multi-GPU layout (no reference counterpart; pointers are rebased to start at 0, so the result
 * equals TCSC(W[:, col_lo:col_hi])). */
int tsg_tcsc_from_dense_cols(const int32_t *W_host, int K, int N, int col_lo, int col_hi,
                             tsg_matrix **out);

/* W already in device memory.  elem_bytes is 4 (int32) or 1 (int8); ld = elements per row. */
int tsg_tcsc_from_dense_dev(const void *W_dev, int elem_bytes, int K, int N, int64_t ld,
                            int col_lo, int col_hi, void *stream, tsg_matrix **out);

/* Adopt arrays a caller already holds in the reference's own layout (the public members
 * col_start_pos/col_start_neg/row_index_pos/row_index_neg of class TCSC, TCSC.h:8-11). */
int tsg_tcsc_from_arrays(const int32_t *col_start_pos, const int32_t *col_start_neg,
                         const int32_t *row_index_pos, const int32_t *row_index_neg, int K, int N,
                         tsg_matrix **out);

/* Column slice [col_lo, col_hi) of an existing matrix, sliced on the device (TCSC is column
 * major, so a shard is a contiguous range of all four arrays). */
int tsg_tcsc_slice_cols(const tsg_matrix *m, int col_lo, int col_hi, tsg_matrix **out);

void tsg_destroy(tsg_matrix *m);

/* ---- queries / export (parity and DataStructureInterface) ----------------------------------- */
int tsg_rows(const tsg_matrix *m, int *K); /* README getNumRows, readme.md:62-72 */
int tsg_cols(const tsg_matrix *m, int *N); /* README getNumCols                  */
int tsg_nnz(const tsg_matrix *m, int64_t *npos, int64_t *nneg);
/* TCSC::getDataStructureSize(), TCSC.h:43-49: 4*(2(N+1)+nnz+ + nnz-) bytes */
int tsg_data_structure_size(const tsg_matrix *m, int64_t *bytes);
/* Copy the device arrays to caller-owned host arrays (csp/csn: N+1 ints; rip/rin: nnz+/nnz-).
 * Any pointer may be NULL to skip that array. */
int tsg_tcsc_export(const tsg_matrix *m, int32_t *col_start_pos, int32_t *col_start_neg,
                    int32_t *row_index_pos, int32_t *row_index_neg);
/* DataStructureInterface::getVectorRepresentation (DataStructureInterface.hpp:13): rebuild the
 * dense K×N int32 matrix from the device TCSC arrays into W_host. */
int tsg_tcsc_to_dense(const tsg_matrix *m, int32_t *W_host);

/* ---- (a2) Y = X·W + b — replaces BaseTCSC<float> -------------------------------------------
 * reference: cpp_impl/comp.h:25-69 and the comp_func signature cpp_impl/common.h:12
 * (argument order X, B, Y, M, N, K).  X: M×K row-major, b: N, Y: M×N row-major (overwritten).
 * HOST pointers; synchronous.  N and K must equal the handle's cols/rows. */
int tsg_spmm(tsg_matrix *m, const float *X, const float *b, float *Y, int M, int N, int K);

/* ---- (a3) fused bias + PReLU — replaces BaseTCSC_PreLU<float> -------------------------------
 * reference: cpp_impl/comp_prelu.h:12-70 and comp_func_prelu cpp_impl/common.h:13:
 * y = acc + b[n];  Y = y > 0 ? y : alpha[n]*y. */
int tsg_spmm_prelu(tsg_matrix *m, const float *X, const float *b, const float *alpha, float *Y,
                   int M, int N, int K);

/* Same two operations with an explicit kernel choice (tsg_algo). */
int tsg_spmm_algo(tsg_matrix *m, int algo, const float *X, const float *b, const float *alpha,
                  float *Y, int M, int N, int K);

/* Device-pointer form: X (M×K, row stride ldx), b, alpha (NULL = no PReLU), Y (M×N, row stride
 * ldy) are device pointers on the handle's GPU; enqueued on `stream`, returns immediately. */
int tsg_spmm_dev(tsg_matrix *m, int algo, const float *X_dev, int64_t ldx, const float *b_dev,
                 const float *alpha_dev, float *Y_dev, int64_t ldy, int M, void *stream);

/* Which kernel TSG_ALGO_AUTO resolves to for this handle and M (a tsg_algo value). */
int tsg_spmm_pick(const tsg_matrix *m, int M, int *algo);

/* Kernels this library launched since load (all handles, all threads) — bench.py's
 * "gpu_launches" is read from here, not guessed. */
int64_t tsg_launch_count(void);

/* Allow (on != 0) or forbid (the default) two-fp16-term operand tiles on the tensor-core path — see
 * "Numerical contract" above.  Process-wide; takes effect with the next call.  Returns the previous
 * setting. */
int tsg_set_fast_split(int on);

/* Host-only helpers (no GPU involved): release store / acquire load of a 64-bit word in memory
 * shared between the ranks of one box.  The multi-GPU host path publishes each step's X through
 * POSIX shared memory (ternary-spgemm_b200/shard.py, HostSharedX); these order the payload against
 * the sequence word on any host architecture. */
void tsg_host_store_release_i64(int64_t *addr, int64_t value);
int64_t tsg_host_load_acquire_i64(const int64_t *addr);

/* Algorithmic HBM bytes of one SpMM call in the reference's own accounting
 * ("Total Input Size", cpp_impl/main.cpp:267,289): 4(MK + MN + N [+N alpha]) + data structure. */
int tsg_spmm_bytes(const tsg_matrix *m, int M, int with_prelu, int64_t *bytes);

/* ---- BlockedTCSC<B> — replaces BlockedTCSC<B>::BlockedTCSC(int*,int,int) --------------------------
 * reference: cpp_impl/data_structures/BlockedTCSC.h:15-43 (B = 512 in main.cpp:7,69).  The arrays
 * of the matrix a TCSC handle holds, rebuilt per K-block of B rows on the device, bit-identical to
 * the reference's vectors: col_start_*: (K/B)*N + 1 ints indexed b*N + j, row ids global, rows at
 * or beyond (K/B)*B dropped.  Call once with NULL arrays for the sizes, again for the arrays.
 * B must be a multiple of 32. */
int tsg_blocked_tcsc_export(const tsg_matrix *m, int B, int64_t *npos, int64_t *nneg,
                            int32_t *col_start_pos, int32_t *col_start_neg,
                            int32_t *row_index_pos, int32_t *row_index_neg);

/* ---- TCSR — replaces class TCSR and BaseTCSR ---------------------------------------------------
 * reference: cpp_impl/data_structures/TCSR.h:13-41 (row_start_pos/neg: K+1 ints, col_index_pos/neg
 * ascending n inside each row), BaseTCSR cpp_impl/comp.h:478-528.  Built on the device,
 * bit-identical to the reference constructor's vectors. */
typedef struct tsg_tcsr tsg_tcsr;
int tsg_tcsr_from_dense(const int32_t *W_host, int K, int N, tsg_tcsr **out);
void tsg_tcsr_destroy(tsg_tcsr *h);
int tsg_tcsr_nnz(const tsg_tcsr *h, int64_t *npos, int64_t *nneg);
/* TCSR::getDataStructureSize(), TCSR.h:43-49 */
int tsg_tcsr_data_structure_size(const tsg_tcsr *h, int64_t *bytes);
int tsg_tcsr_export(const tsg_tcsr *h, int32_t *row_start_pos, int32_t *row_start_neg,
                    int32_t *col_index_pos, int32_t *col_index_neg);
/* dense K×N int32 matrix rebuilt from the TCSR arrays (getVectorRepresentation) */
int tsg_tcsr_to_dense(const tsg_tcsr *h, int32_t *W_host);
/* host pointers, synchronous.  algo = TSG_ALGO_TCSR_SEQ: bit-identical to BaseTCSR; any TCSC
 * algo (or AUTO): the engine's kernels on the same W.  alpha may be NULL. */
int tsg_tcsr_spmm(tsg_tcsr *h, int algo, const float *X, const float *b, const float *alpha,
                  float *Y, int M, int N, int K);

/* ---- packed-value CSC — the README's "value compression (5 values into 8 bits)" ----------------
 * reference: readme.md:108-111 names the idea; no code or layout exists there, so the layout is
 * defined by this library:
 *   col_ptr int32[N+1], row_idx int32[nnz] (+1 and -1 merged, rows ascending per column),
 *   vals uint8[ceil(nnz/5)]: byte b = sum_j d(5b+j)·3^j with d = value+1 (0 or 2), pad digit 1. */
typedef struct tsg_pcsc tsg_pcsc;
int tsg_pcsc_from_dense(const int32_t *W_host, int K, int N, tsg_pcsc **out);
int tsg_pcsc_from_dense_dev(const void *W_dev, int elem_bytes, int K, int N, void *stream,
                            tsg_pcsc **out);
int tsg_pcsc_from_arrays(const int32_t *col_ptr, const int32_t *row_idx, const uint8_t *vals,
                         int K, int N, tsg_pcsc **out);
void tsg_pcsc_destroy(tsg_pcsc *h);
int tsg_pcsc_sizes(const tsg_pcsc *h, int64_t *nnz, int64_t *val_bytes);
int tsg_pcsc_data_structure_size(const tsg_pcsc *h, int64_t *bytes);
int tsg_pcsc_export(const tsg_pcsc *h, int32_t *col_ptr, int32_t *row_idx, uint8_t *vals);
int tsg_pcsc_to_dense(const tsg_pcsc *h, int32_t *W_host);
/* algo = TSG_ALGO_PCSC_GATHER: computes from the packed stream; any TCSC algo (or AUTO): the
 * engine's kernels on the same W. */
int tsg_pcsc_spmm(tsg_pcsc *h, int algo, const float *X, const float *b, const float *alpha,
                  float *Y, int M, int N, int K);
int tsg_pcsc_spmm_dev(tsg_pcsc *h, int algo, const float *X_dev, int64_t ldx, const float *b_dev,
                      const float *alpha_dev, float *Y_dev, int64_t ldy, int M, void *stream);
/* which kernel TSG_ALGO_AUTO runs for M rows on this handle (the engine's choice for the same W) */
int tsg_pcsc_spmm_pick(const tsg_pcsc *h, int M, int *algo);

/* ---- packed-value CSR — the same value compression along rows ----------------------------------
 * reference: readme.md:108-111 ("value compression" for the CSC/CSR formats; no code or layout in
 * the reference).  Layout, the row-major twin of the packed CSC:
 *   row_ptr int32[K+1], col_idx int32[nnz] (+1 and -1 merged, columns ascending per row),
 *   vals uint8[ceil(nnz/5)]: byte b = sum_j d(5b+j)·3^j with d = value+1 (0 or 2), pad digit 1.
 * algo = TSG_ALGO_PCSR_SEQ walks the packed rows in BaseTCSR's order (cpp_impl/comp.h:478-528) and
 * is bit-identical to it; any TCSC algo (or AUTO): the engine's kernels on the same W. */
typedef struct tsg_pcsr tsg_pcsr;
int tsg_pcsr_from_dense(const int32_t *W_host, int K, int N, tsg_pcsr **out);
int tsg_pcsr_from_dense_dev(const void *W_dev, int elem_bytes, int K, int N, void *stream,
                            tsg_pcsr **out);
int tsg_pcsr_from_arrays(const int32_t *row_ptr, const int32_t *col_idx, const uint8_t *vals,
                         int K, int N, tsg_pcsr **out);
void tsg_pcsr_destroy(tsg_pcsr *h);
int tsg_pcsr_sizes(const tsg_pcsr *h, int64_t *nnz, int64_t *val_bytes);
int tsg_pcsr_data_structure_size(const tsg_pcsr *h, int64_t *bytes);
int tsg_pcsr_export(const tsg_pcsr *h, int32_t *row_ptr, int32_t *col_idx, uint8_t *vals);
int tsg_pcsr_to_dense(const tsg_pcsr *h, int32_t *W_host);
int tsg_pcsr_spmm(tsg_pcsr *h, int algo, const float *X, const float *b, const float *alpha,
                  float *Y, int M, int N, int K);

#ifdef __cplusplus
}
#endif
#endif /* TSG_H */
