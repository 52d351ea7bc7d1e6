// tsg_internal.cuh — shared between the translation units of libtsg.so (not installed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

#include "../../include/tsg.h"

// ---- error plumbing --------------------------------------------------------------------------
void tsg_set_error(const char *fmt, ...);
extern std::atomic<long long> g_tsg_launches;
extern std::atomic<int> g_tsg_fast_split; // two-fp16-term tiles allowed (tsg_set_fast_split, TSG_TC_FAST=1)

#define TSG_CUDA(call)                                                                          \
    do                                                                                          \
    {                                                                                           \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess)                                                                 \
        {                                                                                       \
            tsg_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__,    \
                          __LINE__);                                                            \
            return (e__ == cudaErrorMemoryAllocation) ? TSG_ERR_NOMEM : TSG_ERR_CUDA;           \
        }                                                                                       \
    } while (0)

#define TSG_CHECK(cond, code, ...)                                                              \
    do                                                                                          \
    {                                                                                           \
        if (!(cond))                                                                            \
        {                                                                                       \
            tsg_set_error(__VA_ARGS__);                                                         \
            return (code);                                                                      \
        }                                                                                       \
    } while (0)

#define TSG_TRY(expr)                                                                           \
    do                                                                                          \
    {                                                                                           \
        int s__ = (expr);                                                                       \
        if (s__ != TSG_OK)                                                                      \
            return s__;                                                                         \
    } while (0)

// Count a kernel launch and surface launch-configuration errors immediately.
#define TSG_LAUNCHED()                                                                          \
    do                                                                                          \
    {                                                                                           \
        g_tsg_launches.fetch_add(1, std::memory_order_relaxed);                                 \
        TSG_CUDA(cudaGetLastError());                                                           \
    } while (0)

// ---- the handle ------------------------------------------------------------------------------
// Data layout in HBM for one ternary K×N matrix:
//   csp, csn : int32[N+1]        column pointers, exactly TCSC::col_start_pos/neg
//   rip, rin : int32[nnz±]       row indices ascending inside each column (+ 64 B zero pad so
//                                 128-bit loads may run past the end)
//   ppos,pneg: uint32[N][Kw]     bit planes, column-major: bit (k&31) of word k>>5 of column n
//                                 is set iff W[k][n] == +1 / -1.  Kw = ceil(K/32) rounded up to 4
//                                 (16-byte rows).  2 bits per matrix element.
struct tsg_matrix
{
    int device = 0;
    int K = 0, N = 0;
    int Kw = 0;                 // words per plane column
    long long npos = 0, nneg = 0;
    int32_t *csp = nullptr, *csn = nullptr, *rip = nullptr, *rin = nullptr;
    uint32_t *ppos = nullptr, *pneg = nullptr;
    // builder-made handles: the six arrays above live in two allocations (blk0: planes, pointers and
    // the scan scratch, sized from K and N; blk1: both index arrays, sized after the scan) and are
    // freed through them; handles assembled elsewhere (from_arrays, slices) own each array separately
    void *blk0 = nullptr, *blk1 = nullptr;
    bool pooled = false; // blk0, blk1 and codes came from cudaMallocAsync (the builder): freed with cudaFreeAsync
    // Kernel-side copy of the index lists for the gather kernel, built by its first launch: every
    // list (one column, one sign) starts on a 16-byte boundary and is padded to whole 16-byte units
    // with the sentinel row index K (the kernel keeps X[K] = 0 in shared memory), so 128-bit loads
    // never straddle two lists and need no masks.  A unit holds 8 uint16 row ids when K <= 65535
    // (idx16: half the HBM bytes of the reference's int32 stream — the README's "denser stream",
    // readme.md:108-111), else 4 int32.  lp/ln: int32[N+1] list pointers in units.
    int32_t *lp = nullptr, *ln = nullptr;
    int32_t *rip4 = nullptr, *rin4 = nullptr;
    long long n4pos = 0, n4neg = 0; // padded lengths in 16-byte units
    bool idx16 = false;
    // Tile-packed 2-bit codes for the tensor-core path: [N/128 tiles][K/64 k-blocks][128 cols][16 B].
    // One uint4 = 64 consecutive k of one column; word (k&63)>>4 holds 16 of them in the
    // shift-and-mask layout of pack_code_word (tsg_build.cu): element e = 2p+h has its non-zero
    // flag at bit 16h+14-2p and its sign one above.  One expander thread turns one uint4 into one
    // 128-byte smem row, fetched with one coalesced 128-bit load.
    uint4 *codes = nullptr;
    int code_tiles = 0, code_kblocks = 0;
    // staging for the host-pointer entry points (grown on demand)
    float *sX = nullptr, *sB = nullptr, *sA = nullptr, *sY = nullptr;
    size_t capX = 0, capB = 0, capA = 0, capY = 0;
    // bias / alpha of the previous small host call: device copy + host shadow (see tsg_spmm_algo)
    float *cB = nullptr, *cA = nullptr, *hB = nullptr, *hA = nullptr;
    bool cB_valid = false, cA_valid = false;
    // scratch for the tensor-core path: per-tile operand flags + 16-bit split copies of X.  One set
    // per handle: a call on a different stream than the previous one first waits until that stream
    // has drained (tsg_dense_tc.cu), so concurrent streams serialise on the scratch instead of
    // overwriting it.
    void *xsplit = nullptr;
    size_t cap_xsplit = 0;
    cudaStream_t scratch_stream = nullptr; // the stream the last call that used the scratch ran on
    bool scratch_used = false;
    cudaStream_t stream = nullptr; // the device's shared internal stream unless owns_stream
    bool owns_stream = false;
    int sm_count = 0;
    size_t smem_optin = 0;
};

// makes `dev` current for the scope (handles live on the device that was current at creation)
struct DeviceGuard
{
    int prev = -1;
    explicit DeviceGuard(int dev)
    {
        cudaGetDevice(&prev);
        if (prev != dev)
            cudaSetDevice(dev);
        else
            prev = -1;
    }
    ~DeviceGuard()
    {
        if (prev >= 0)
            cudaSetDevice(prev);
    }
};

static inline int tsg_kw(int K) { return ((K + 31) / 32 + 3) & ~3; }

// ---- inputs a dense product cannot take --------------------------------------------------------
// The re-ordered kernels that multiply (dense_tc, code_gemv) compute 0·x for the zero entries of W
// and hold 2·W, so x = ±inf / NaN at a position where W is 0 would turn into NaN, and |x| close to
// FLT_MAX would overflow, where the reference's sparse sum (comp.h:44-61) never touches that x.
// Every such kernel therefore tests its X for "non-finite or |x| >= 2^100" while it stages or
// splits it, and a tile that saw one recomputes its outputs with the reference's own sum below.
#define TSG_X_HUGE_BITS 0x71800000u /* fp32 bit pattern of 2^100; inf and NaN compare above it */

#ifdef __CUDACC__
// BaseTCSC's arithmetic for ONE output element (comp.h:41-63): one fp32 accumulator, positives in
// ascending row order, then negatives, bias last.  x[k * xstride] is X[m][k].
__device__ __forceinline__ float tsg_ref_order_sum(const float *x, int64_t xstride, const int32_t *csp,
                                                   const int32_t *csn, const int32_t *rip,
                                                   const int32_t *rin, int n, float bias)
{
    float acc = 0.0f;
    for (int i = csp[n], e = csp[n + 1]; i < e; ++i)
        acc += x[(int64_t)rip[i] * xstride];
    for (int i = csn[n], e = csn[n + 1]; i < e; ++i)
        acc -= x[(int64_t)rin[i] * xstride];
    return acc + bias;
}
#endif

// ---- builders (tsg_build.cu) -----------------------------------------------------------------
// element (k, n) of the matrix being built is W_dev[k*ld + (col_lo+n)*cs]: cs = 1 for the
// reference's row-major W; ld = 1, cs = N builds from the TRANSPOSE (TCSR(W) == TCSC(W^T))
int tsg_build_from_dense_dev(tsg_matrix *m, const void *W_dev, int elem_bytes, int64_t ld,
                             int col_lo, cudaStream_t st, int64_t cs = 1);
// bare handle (device, stream, shape) — tsg_api.cu
int tsg_new_matrix(int K, int N, tsg_matrix **out);
int tsg_build_planes_from_arrays(tsg_matrix *m, cudaStream_t st);
// validation of caller-made arrays (tsg_*_from_arrays): host pointer arrays; device index lists
// (inside [0, bound), strictly ascending per list); no row in both sign lists.  The device checks
// synchronise `st`.
int tsg_validate_pointers(const int32_t *ptr, int n, long long total, const char *what);
int tsg_validate_lists(const int32_t *ptr_dev, const int32_t *idx_dev, int nlists, int bound, cudaStream_t st,
                       const char *what);
int tsg_validate_no_overlap(const tsg_matrix *m, cudaStream_t st);
int tsg_scatter_to_dense(const tsg_matrix *m, int32_t *W_dev, cudaStream_t st);
int tsg_rebase_slice(int32_t *dst, const int32_t *src, int n, cudaStream_t st);
// BlockedTCSC<B> arrays (fresh device allocations, caller cudaFree()s them)
int tsg_build_blocked(const tsg_matrix *m, int B, int32_t **csp, int32_t **csn, int32_t **rip, int32_t **rin,
                      long long *npos, long long *nneg);
// builds lp/ln/rip4/rin4 from csp/csn/rip/rin; synchronises `st`.  Called by the first gather
// launch on the handle (tsg_gather.cu): no other kernel reads these.
int tsg_build_padded_lists(tsg_matrix *m, cudaStream_t st);
// device time of the last tsg_build_from_dense_dev + tsg_build_tile_codes pair when TSG_BUILD_TIMING=1
// (CUDA events around the kernel sequences; allocations and host waits excluded), else 0
extern "C" double tsg_debug_last_build_device_ms(void);
// builds the tile-packed codes from the bit planes; asynchronous on `st`
int tsg_build_tile_codes(tsg_matrix *m, cudaStream_t st);

// ---- kernels (tsg_gather.cu, tsg_dense_tc.cu, tsg_code_gemv.cu) ------------------------------
int tsg_launch_gather(tsg_matrix *m, const float *X, int64_t ldx, const float *b,
                      const float *alpha, float *Y, int64_t ldy, int M, cudaStream_t st);
int tsg_launch_gather_seq(tsg_matrix *m, const float *X, int64_t ldx, const float *b,
                          const float *alpha, float *Y, int64_t ldy, int M, cudaStream_t st);
int tsg_launch_dense_tc(tsg_matrix *m, const float *X, int64_t ldx, const float *b,
                        const float *alpha, float *Y, int64_t ldy, int M, cudaStream_t st);
int tsg_launch_code_gemv(tsg_matrix *m, const float *X, int64_t ldx, const float *b,
                         const float *alpha, float *Y, int64_t ldy, int M, cudaStream_t st);
