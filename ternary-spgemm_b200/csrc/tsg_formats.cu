// tsg_formats.cu — the two "next" storage formats of SURVEY §8f, built on the device and served
// by the same engine as TCSC.
//
//  TCSR  (reference cpp_impl/data_structures/TCSR.h:13-41, kernel BaseTCSR cpp_impl/comp.h:478-528)
//        row_start_pos/neg[K+1], col_index_pos/neg ascending n inside each row.  TCSR(W) is
//        TCSC(Wᵀ): the TCSC builder runs on the transposed strides, so the four arrays are
//        bit-identical to the reference constructor's.
//        tcsr_seq_kernel states BaseTCSR's arithmetic on the device (Y = b, then k ascending:
//        += x to the row's positive columns, -= x to its negative ones): one CTA per row of X,
//        the columns of one k are disjoint so they update in parallel, a barrier between k keeps
//        the order — bit-identical to BaseTCSR for arbitrary fp32 X.
//
//  PCSC  packed-value CSC — the README's "value compression, 5 values in 8 bits"
//        (readme.md:108-111; the reference ships no code or layout for it, so the layout is
//        defined here and the format's parity is UNPINNED by the reference: it is pinned by
//        round trip and by Y parity with BaseTCSC):
//          col_ptr int32[N+1]           merged (+ and -) non-zeros per column
//          row_idx int32[nnz]           rows ascending inside each column
//          vals    uint8[ceil(nnz/5)]   base-3 little-endian digits d = v+1 (0 for -1, 2 for +1)
//                                       of five consecutive entries of row_idx; pad digit 1
//        emit_merged_kernel is a warp-ballot / popc-prefix compaction over the bit planes (one
//        warp per column, lane i owns row 32j+i); pack_vals_kernel folds five signs into a byte.
//        pcsc_gather_kernel computes Y from the packed stream itself (one warp per column,
//        X row tile in shared memory, sign decoded per entry).
//
//  PCSR  packed-value CSR — the same packing along rows (the README lists the value compression
//        for "CSC/CSR"): row_ptr int32[K+1], col_idx int32[nnz] ascending n inside each row,
//        vals uint8[ceil(nnz/5)].  PCSR(W) is PCSC(Wᵀ): the same builder kernels on the planes of
//        the transposed matrix.  pcsr_seq_kernel walks the packed rows in BaseTCSR's order (Y = b,
//        then k ascending, every entry adds ±x to its own column): per output column the
//        operations arrive in the same order as in BaseTCSR, so Y is bit-identical to it.
//
// Both handles also own a regular engine matrix built from the same W, so TSG_ALGO_AUTO and the
// explicit TCSC kernels serve them at full speed; the format-native kernels are the parity anchors.
#include "tsg_internal.cuh"

#include <new>

struct tsg_tcsr
{
    tsg_matrix *fwd = nullptr; // W as the engine holds it (TCSC + planes + codes)
    tsg_matrix *t = nullptr;   // TCSC of W^T == TCSR of W (arrays only)
};

struct tsg_pcsc
{
    tsg_matrix *fwd = nullptr;
    int32_t *col_ptr = nullptr, *row_idx = nullptr;
    uint8_t *vals = nullptr;
    long long nnz = 0, nbytes = 0;
};

struct tsg_pcsr
{
    tsg_matrix *fwd = nullptr;
    int32_t *row_ptr = nullptr, *col_idx = nullptr;
    uint8_t *vals = nullptr;
    long long nnz = 0, nbytes = 0;
};

namespace
{

// ---- TCSR ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tcsr_to_dense_kernel(const int *__restrict__ rsp, const int *__restrict__ rsn, const int *__restrict__ cip,
                     const int *__restrict__ cin, int K, int N, int32_t *__restrict__ W)
{
    const int k = blockIdx.x;
    for (int j = rsp[k] + threadIdx.x; j < rsp[k + 1]; j += blockDim.x)
        W[(int64_t)k * N + cip[j]] = 1;
    for (int j = rsn[k] + threadIdx.x; j < rsn[k + 1]; j += blockDim.x)
        W[(int64_t)k * N + cin[j]] = -1;
}

// BaseTCSR, comp.h:478-528.  One CTA per row m of X; Y row lives in global memory (N can exceed
// shared memory); block-level barriers order the k steps.
__global__ void __launch_bounds__(1024)
tcsr_seq_kernel(const int *__restrict__ rsp, const int *__restrict__ rsn, const int *__restrict__ cip,
                const int *__restrict__ cin, const float *__restrict__ X, int64_t ldx,
                const float *__restrict__ bias, const float *__restrict__ alpha, float *Y, int64_t ldy, int K, int N)
{
    const int m = blockIdx.x;
    float *y = Y + (int64_t)m * ldy;
    const float *x = X + (int64_t)m * ldx;
    for (int n = threadIdx.x; n < N; n += blockDim.x)
        y[n] = bias[n]; // comp.h:491-497
    __syncthreads();
    for (int k = 0; k < K; ++k) // comp.h:506
    {
        const float xv = x[k];
        const int p0 = rsp[k], p1 = rsp[k + 1], q0 = rsn[k], q1 = rsn[k + 1];
        for (int j = p0 + threadIdx.x; j < p1; j += blockDim.x)
            y[cip[j]] += xv; // comp.h:512
        for (int j = q0 + threadIdx.x; j < q1; j += blockDim.x)
            y[cin[j]] -= xv; // comp.h:521
        if (p1 > p0 || q1 > q0)
            __syncthreads(); // uniform: the ranges are the same for every thread
    }
    if (alpha != nullptr)
        for (int n = threadIdx.x; n < N; n += blockDim.x)
        {
            const float v = y[n];
            y[n] = (v > 0.0f) ? v : alpha[n] * v;
        }
}

// ---- PCSC ------------------------------------------------------------------------------------
__global__ void merged_ptr_kernel(const int *__restrict__ csp, const int *__restrict__ csn, int n1, int *__restrict__ cp)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n1)
        cp[i] = csp[i] + csn[i];
}

// one warp per column: rows ascending, signs as digits (0 = -1, 2 = +1)
__global__ void __launch_bounds__(256)
emit_merged_kernel(const uint32_t *__restrict__ ppos, const uint32_t *__restrict__ pneg, const int *__restrict__ cp,
                   int ncols, int Kw, int *__restrict__ row_idx, uint8_t *__restrict__ digit)
{
    const int lane = threadIdx.x & 31;
    const int col = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (col >= ncols)
        return;
    const uint32_t *pp = ppos + (int64_t)col * Kw, *pq = pneg + (int64_t)col * Kw;
    const uint32_t lt = (1u << lane) - 1u;
    int out = cp[col];
    for (int j = 0; j < Kw; ++j)
    {
        const uint32_t p = pp[j], q = pq[j]; // broadcast loads (same address across the warp)
        const uint32_t nz = p | q;
        if ((nz >> lane) & 1u)
        {
            const int slot = out + __popc(nz & lt);
            row_idx[slot] = j * 32 + lane;
            digit[slot] = ((p >> lane) & 1u) ? 2 : 0;
        }
        out += __popc(nz);
    }
}

__global__ void pack_vals_kernel(const uint8_t *__restrict__ digit, long long nnz, long long nbytes, uint8_t *__restrict__ vals)
{
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbytes)
        return;
    unsigned v = 0, mul = 1;
#pragma unroll
    for (int j = 0; j < 5; ++j)
    {
        const long long i = b * 5 + j;
        v += mul * (i < nnz ? (unsigned)digit[i] : 1u); // pad digit 1
        mul *= 3;
    }
    vals[b] = (uint8_t)v;
}

__device__ __forceinline__ int pcsc_sign(const uint8_t *__restrict__ vals, long long i)
{
    const long long b = i / 5;
    const int j = (int)(i - b * 5);
    const unsigned pw[5] = {1, 3, 9, 27, 81};
    return (int)((vals[b] / pw[j]) % 3u) - 1;
}

template <typename T>
__global__ void __launch_bounds__(256)
pcsc_to_dense_kernel(const int *__restrict__ cp, const int *__restrict__ row_idx, const uint8_t *__restrict__ vals,
                     int K, int N, T *__restrict__ W)
{
    const int col = blockIdx.x;
    for (int i = cp[col] + threadIdx.x; i < cp[col + 1]; i += blockDim.x)
        W[(int64_t)row_idx[i] * N + col] = (T)pcsc_sign(vals, i);
}

template <typename T>
__global__ void __launch_bounds__(256)
pcsr_to_dense_kernel(const int *__restrict__ rp, const int *__restrict__ col_idx, const uint8_t *__restrict__ vals,
                     int K, int N, T *__restrict__ W)
{
    const int k = blockIdx.x;
    for (int i = rp[k] + threadIdx.x; i < rp[k + 1]; i += blockDim.x)
        W[(int64_t)k * N + col_idx[i]] = (T)pcsc_sign(vals, i);
}

// BaseTCSR's order (comp.h:478-528) over the packed rows.  One CTA per row m of X; the entries of
// one k hit distinct columns and update in parallel, a barrier between k keeps the order.
__global__ void __launch_bounds__(1024)
pcsr_seq_kernel(const int *__restrict__ rp, const int *__restrict__ col_idx, const uint8_t *__restrict__ vals,
                const float *__restrict__ X, int64_t ldx, const float *__restrict__ bias,
                const float *__restrict__ alpha, float *Y, int64_t ldy, int K, int N)
{
    const int m = blockIdx.x;
    float *y = Y + (int64_t)m * ldy;
    const float *x = X + (int64_t)m * ldx;
    for (int n = threadIdx.x; n < N; n += blockDim.x)
        y[n] = bias[n];
    __syncthreads();
    for (int k = 0; k < K; ++k)
    {
        const float xv = x[k];
        const int j0 = rp[k], j1 = rp[k + 1];
        for (int j = j0 + threadIdx.x; j < j1; j += blockDim.x)
        {
            const int c = col_idx[j];
            y[c] = (pcsc_sign(vals, j) > 0) ? y[c] + xv : y[c] - xv;
        }
        if (j1 > j0)
            __syncthreads(); // uniform: the range is the same for every thread
    }
    if (alpha != nullptr)
        for (int n = threadIdx.x; n < N; n += blockDim.x)
        {
            const float v = y[n];
            y[n] = (v > 0.0f) ? v : alpha[n] * v;
        }
}

// Y from the packed stream: one warp per column, MT = 4 rows of X per pass, X tile k-major in
// shared memory (one LDS.128 per entry), lanes stride over the column's entries (coalesced
// index loads), warp-shuffle reduction, fused bias / PReLU.
constexpr int kPcscWarps = 16;
__global__ void __launch_bounds__(kPcscWarps * 32)
pcsc_gather_kernel(const int *__restrict__ cp, const int *__restrict__ row_idx, const uint8_t *__restrict__ vals,
                   const float *__restrict__ X, int64_t ldx, const float *__restrict__ bias,
                   const float *__restrict__ alpha, float *__restrict__ Y, int64_t ldy, int M, int K, int N)
{
    extern __shared__ __align__(16) float xs[]; // [K][4]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = blockIdx.y * 4;
    for (int i = tid; i < K * 4; i += kPcscWarps * 32)
    {
        const int k = i >> 2, m = i & 3;
        xs[i] = (m0 + m < M) ? X[(int64_t)(m0 + m) * ldx + k] : 0.0f;
    }
    __syncthreads();
    for (int col = blockIdx.x * kPcscWarps + warp; col < N; col += gridDim.x * kPcscWarps)
    {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = cp[col] + lane; i < cp[col + 1]; i += 32)
        {
            const float s = (float)pcsc_sign(vals, i);
            const float4 x = *reinterpret_cast<const float4 *>(xs + 4 * row_idx[i]);
            acc.x = fmaf(s, x.x, acc.x), acc.y = fmaf(s, x.y, acc.y);
            acc.z = fmaf(s, x.z, acc.z), acc.w = fmaf(s, x.w, acc.w);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o), acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
            acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o), acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
        }
        if (lane < 4 && m0 + lane < M)
        {
            float y = (lane == 0 ? acc.x : lane == 1 ? acc.y : lane == 2 ? acc.z : acc.w) + bias[col];
            if (alpha != nullptr)
                y = (y > 0.0f) ? y : alpha[col] * y;
            Y[(int64_t)(m0 + lane) * ldy + col] = y;
        }
    }
}

int upload_dense(const int32_t *W_host, int K, int N, int32_t **dW)
{
    TSG_CHECK(K >= 0 && N >= 0, TSG_ERR_INVALID, "negative shape K=%d N=%d", K, N);
    TSG_CHECK(W_host != nullptr || (long long)K * N == 0, TSG_ERR_INVALID, "W is NULL");
    int usable = 0;
    tsg_device_count(&usable);
    TSG_CHECK(usable > 0, TSG_ERR_NO_DEVICE, "no sm_100 device visible (libtsg has no CPU fallback)");
    TSG_CHECK((long long)K * N <= (long long)INT32_MAX, TSG_ERR_OVERFLOW,
              "K*N = %lld exceeds the reference's int indexing", (long long)K * N);
    const size_t bytes = (size_t)K * N * 4;
    TSG_CUDA(cudaMalloc(dW, bytes ? bytes : 4));
    if (bytes && cudaMemcpy(*dW, W_host, bytes, cudaMemcpyHostToDevice) != cudaSuccess)
    {
        cudaFree(*dW);
        *dW = nullptr;
        tsg_set_error("upload of W failed: %s", cudaGetErrorString(cudaGetLastError()));
        return TSG_ERR_CUDA;
    }
    return TSG_OK;
}

// host-pointer wrapper shared by the format-native kernels: stage in the engine handle's own
// staging buffers (grown on demand, kept for the next call — no allocation per call), run `launch`
// on the handle's stream, copy back
static int grow_staging(float **p, size_t *cap, size_t need_floats)
{
    if (*cap >= need_floats)
        return TSG_OK;
    if (*p)
        cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    TSG_CUDA(cudaMalloc(p, need_floats * sizeof(float) + 64));
    *cap = need_floats;
    return TSG_OK;
}

template <typename F>
int run_host(tsg_matrix *eng, const float *X, const float *b, const float *alpha, float *Y, int M, int N, int K,
             F launch)
{
    cudaStream_t st = eng->stream;
    TSG_TRY(grow_staging(&eng->sX, &eng->capX, (size_t)M * K + 1));
    TSG_TRY(grow_staging(&eng->sB, &eng->capB, (size_t)N));
    TSG_TRY(grow_staging(&eng->sY, &eng->capY, (size_t)M * N + 1));
    if (alpha)
        TSG_TRY(grow_staging(&eng->sA, &eng->capA, (size_t)N));
    TSG_CUDA(cudaMemcpyAsync(eng->sX, X, (size_t)M * K * 4, cudaMemcpyHostToDevice, st));
    TSG_CUDA(cudaMemcpyAsync(eng->sB, b, (size_t)N * 4, cudaMemcpyHostToDevice, st));
    if (alpha)
        TSG_CUDA(cudaMemcpyAsync(eng->sA, alpha, (size_t)N * 4, cudaMemcpyHostToDevice, st));
    TSG_TRY(launch(eng->sX, eng->sB, alpha ? eng->sA : nullptr, eng->sY, st));
    TSG_CUDA(cudaMemcpyAsync(Y, eng->sY, (size_t)M * N * 4, cudaMemcpyDeviceToHost, st));
    const cudaError_t e = cudaStreamSynchronize(st);
    TSG_CHECK(e == cudaSuccess, TSG_ERR_CUDA, "format-native SpMM failed: %s", cudaGetErrorString(e));
    return TSG_OK;
}

} // namespace

extern "C"
{

    // =========================================== TCSR ===========================================
    void tsg_tcsr_destroy(tsg_tcsr *h)
    {
        if (!h)
            return;
        tsg_destroy(h->fwd);
        tsg_destroy(h->t);
        delete h;
    }

    int tsg_tcsr_from_dense(const int32_t *W_host, int K, int N, tsg_tcsr **out)
    {
        TSG_CHECK(out != nullptr, TSG_ERR_INVALID, "out is NULL");
        *out = nullptr;
        int32_t *dW = nullptr;
        TSG_TRY(upload_dense(W_host, K, N, &dW));
        tsg_tcsr *h = new (std::nothrow) tsg_tcsr();
        int s = h ? TSG_OK : TSG_ERR_NOMEM;
        if (s == TSG_OK)
            s = tsg_tcsc_from_dense_dev(dW, 4, K, N, N, 0, N, nullptr, &h->fwd);
        if (s == TSG_OK)
            s = tsg_new_matrix(N, K, &h->t); // W^T: N rows, K columns
        if (s == TSG_OK)
            s = tsg_build_from_dense_dev(h->t, dW, 4, /*ld=*/1, 0, h->t->stream, /*cs=*/N);
        cudaFree(dW);
        if (s != TSG_OK)
        {
            tsg_tcsr_destroy(h);
            return s;
        }
        *out = h;
        return TSG_OK;
    }

    int tsg_tcsr_nnz(const tsg_tcsr *h, int64_t *npos, int64_t *nneg)
    {
        TSG_CHECK(h, TSG_ERR_INVALID, "NULL argument");
        return tsg_nnz(h->t, npos, nneg);
    }

    // TCSR::getDataStructureSize(), TCSR.h:43-49: 4*(2(K+1) + nnz+ + nnz-)
    int tsg_tcsr_data_structure_size(const tsg_tcsr *h, int64_t *bytes)
    {
        TSG_CHECK(h, TSG_ERR_INVALID, "NULL argument");
        return tsg_data_structure_size(h->t, bytes);
    }

    int tsg_tcsr_export(const tsg_tcsr *h, int32_t *row_start_pos, int32_t *row_start_neg, int32_t *col_index_pos,
                        int32_t *col_index_neg)
    {
        TSG_CHECK(h, TSG_ERR_INVALID, "NULL argument");
        DeviceGuard guard_(h->fwd->device); // the handle lives on the device that was current when it was built
        return tsg_tcsc_export(h->t, row_start_pos, row_start_neg, col_index_pos, col_index_neg);
    }

    int tsg_tcsr_to_dense(const tsg_tcsr *h, int32_t *W_host)
    {
        TSG_CHECK(h && (W_host || (long long)h->fwd->K * h->fwd->N == 0), TSG_ERR_INVALID, "NULL argument");
        DeviceGuard guard_(h->fwd->device); // the handle lives on the device that was current when it was built
        const int K = h->fwd->K, N = h->fwd->N;
        const size_t bytes = (size_t)K * N * 4;
        if (!bytes)
            return TSG_OK;
        int32_t *dW = nullptr;
        TSG_CUDA(cudaMalloc(&dW, bytes));
        cudaStream_t st = h->t->stream;
        cudaMemsetAsync(dW, 0, bytes, st);
        tcsr_to_dense_kernel<<<K, 256, 0, st>>>(h->t->csp, h->t->csn, h->t->rip, h->t->rin, K, N, dW);
        g_tsg_launches.fetch_add(1, std::memory_order_relaxed);
        cudaError_t e = cudaMemcpyAsync(W_host, dW, bytes, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess)
            e = cudaStreamSynchronize(st);
        cudaFree(dW);
        TSG_CHECK(e == cudaSuccess, TSG_ERR_CUDA, "TCSR -> dense failed: %s", cudaGetErrorString(e));
        return TSG_OK;
    }

    int tsg_tcsr_spmm(tsg_tcsr *h, int algo, const float *X, const float *b, const float *alpha, float *Y, int M,
                      int N, int K)
    {
        TSG_CHECK(h, TSG_ERR_INVALID, "matrix is NULL");
        DeviceGuard guard_(h->fwd->device); // the handle lives on the device that was current when it was built
        if (algo != TSG_ALGO_TCSR_SEQ)
            return tsg_spmm_algo(h->fwd, algo, X, b, alpha, Y, M, N, K);
        TSG_CHECK(N == h->fwd->N && K == h->fwd->K, TSG_ERR_INVALID, "shape mismatch");
        if (M <= 0 || N == 0)
            return TSG_OK;
        TSG_CHECK(X && b && Y, TSG_ERR_INVALID, "X, b and Y must be non-NULL");
        const tsg_matrix *t = h->t;
        return run_host(h->fwd, X, b, alpha, Y, M, N, K,
                        [&](float *dX, float *dB, float *dA, float *dY, cudaStream_t st) -> int {
                            tcsr_seq_kernel<<<M, 1024, 0, st>>>(t->csp, t->csn, t->rip, t->rin, dX, K, dB, dA, dY, N, K, N);
                            TSG_LAUNCHED();
                            return (int)TSG_OK;
                        });
    }

    // ======================================= BlockedTCSC<B> ======================================
    // reference: cpp_impl/data_structures/BlockedTCSC.h:15-43.  A view of the W a TCSC handle
    // holds, built on the device at export time.  Two calls like the reference's vectors need:
    // sizes first, then the arrays (col_start_*: (K/B)*N + 1 ints).
    int tsg_blocked_tcsc_export(const tsg_matrix *m, int B, int64_t *npos, int64_t *nneg, int32_t *col_start_pos,
                                int32_t *col_start_neg, int32_t *row_index_pos, int32_t *row_index_neg)
    {
        TSG_CHECK(m, TSG_ERR_INVALID, "matrix is NULL");
        DeviceGuard guard_(m->device); // the handle lives on the device that was current when it was built
        int32_t *csp = nullptr, *csn = nullptr, *rip = nullptr, *rin = nullptr;
        long long np = 0, nn = 0;
        TSG_TRY(tsg_build_blocked(m, B, &csp, &csn, &rip, &rin, &np, &nn));
        const size_t pairs = (size_t)(m->K / B) * m->N;
        cudaError_t e = cudaSuccess;
        if (col_start_pos)
            e = cudaMemcpy(col_start_pos, csp, (pairs + 1) * 4, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && col_start_neg)
            e = cudaMemcpy(col_start_neg, csn, (pairs + 1) * 4, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && row_index_pos && np)
            e = cudaMemcpy(row_index_pos, rip, (size_t)np * 4, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && row_index_neg && nn)
            e = cudaMemcpy(row_index_neg, rin, (size_t)nn * 4, cudaMemcpyDeviceToHost);
        cudaFree(csp), cudaFree(csn), cudaFree(rip), cudaFree(rin);
        TSG_CHECK(e == cudaSuccess, TSG_ERR_CUDA, "BlockedTCSC export failed: %s", cudaGetErrorString(e));
        if (npos)
            *npos = np;
        if (nneg)
            *nneg = nn;
        return TSG_OK;
    }

    // =========================================== PCSC ===========================================
    void tsg_pcsc_destroy(tsg_pcsc *h)
    {
        if (!h)
            return;
        tsg_destroy(h->fwd);
        cudaFree(h->col_ptr), cudaFree(h->row_idx), cudaFree(h->vals);
        delete h;
    }

    // merged pointers / indices / packed signs of the matrix `m` holds (columns of m; pass the
    // engine matrix of W^T for the row-major variant)
    static int build_packed(const tsg_matrix *m, int32_t **ptr, int32_t **idx, uint8_t **vals, long long *nnz_out,
                            long long *nbytes_out)
    {
        const int N = m->N;
        const long long nnz = m->npos + m->nneg;
        TSG_CHECK(nnz <= INT32_MAX, TSG_ERR_OVERFLOW, "nnz = %lld exceeds int32 pointers", nnz);
        const long long nbytes = (nnz + 4) / 5;
        *nnz_out = nnz, *nbytes_out = nbytes;
        cudaStream_t st = m->stream;
        uint8_t *digit = nullptr;
        TSG_CUDA(cudaMalloc(ptr, (size_t)(N + 1) * 4));
        TSG_CUDA(cudaMalloc(idx, (size_t)nnz * 4 + 16));
        TSG_CUDA(cudaMalloc(vals, (size_t)nbytes + 16));
        TSG_CUDA(cudaMalloc(&digit, (size_t)nnz + 16));
        merged_ptr_kernel<<<(N + 1 + 255) / 256, 256, 0, st>>>(m->csp, m->csn, N + 1, *ptr);
        g_tsg_launches.fetch_add(1, std::memory_order_relaxed);
        if (N > 0)
        {
            emit_merged_kernel<<<(N + 7) / 8, 256, 0, st>>>(m->ppos, m->pneg, *ptr, N, m->Kw, *idx, digit);
            g_tsg_launches.fetch_add(1, std::memory_order_relaxed);
        }
        if (nbytes > 0)
        {
            pack_vals_kernel<<<(unsigned)((nbytes + 255) / 256), 256, 0, st>>>(digit, nnz, nbytes, *vals);
            g_tsg_launches.fetch_add(1, std::memory_order_relaxed);
        }
        const cudaError_t e = cudaStreamSynchronize(st);
        cudaFree(digit);
        TSG_CHECK(e == cudaSuccess, TSG_ERR_CUDA, "packed-value builder failed: %s", cudaGetErrorString(e));
        return TSG_OK;
    }

    static int pcsc_from_engine(tsg_pcsc *h)
    {
        return build_packed(h->fwd, &h->col_ptr, &h->row_idx, &h->vals, &h->nnz, &h->nbytes);
    }

    int tsg_pcsc_from_dense(const int32_t *W_host, int K, int N, tsg_pcsc **out)
    {
        TSG_CHECK(out != nullptr, TSG_ERR_INVALID, "out is NULL");
        *out = nullptr;
        tsg_pcsc *h = new (std::nothrow) tsg_pcsc();
        TSG_CHECK(h != nullptr, TSG_ERR_NOMEM, "host allocation failed");
        int s = tsg_tcsc_from_dense(W_host, K, N, &h->fwd);
        if (s == TSG_OK)
            s = pcsc_from_engine(h);
        if (s != TSG_OK)
        {
            tsg_pcsc_destroy(h);
            return s;
        }
        *out = h;
        return TSG_OK;
    }

    // W already in HBM (int32 or int8 row-major): the large BASELINE shapes
    int tsg_pcsc_from_dense_dev(const void *W_dev, int elem_bytes, int K, int N, void *stream, tsg_pcsc **out)
    {
        TSG_CHECK(out != nullptr, TSG_ERR_INVALID, "out is NULL");
        *out = nullptr;
        tsg_pcsc *h = new (std::nothrow) tsg_pcsc();
        TSG_CHECK(h != nullptr, TSG_ERR_NOMEM, "host allocation failed");
        int s = tsg_tcsc_from_dense_dev(W_dev, elem_bytes, K, N, N, 0, N, stream, &h->fwd);
        if (s == TSG_OK)
            s = pcsc_from_engine(h);
        if (s != TSG_OK)
        {
            tsg_pcsc_destroy(h);
            return s;
        }
        *out = h;
        return TSG_OK;
    }

    // Adopt arrays in the packed layout (interchange): decoded on the device, then built like any W.
    int tsg_pcsc_from_arrays(const int32_t *col_ptr, const int32_t *row_idx, const uint8_t *vals, int K, int N,
                             tsg_pcsc **out)
    {
        TSG_CHECK(out != nullptr, TSG_ERR_INVALID, "out is NULL");
        *out = nullptr;
        TSG_CHECK(col_ptr && K >= 0 && N >= 0, TSG_ERR_INVALID, "bad arguments");
        int usable = 0;
        tsg_device_count(&usable);
        TSG_CHECK(usable > 0, TSG_ERR_NO_DEVICE, "no sm_100 device visible (libtsg has no CPU fallback)");
        TSG_CHECK((long long)K * N <= (long long)INT32_MAX, TSG_ERR_OVERFLOW, "K*N too large");
        const long long nnz = col_ptr[N];
        TSG_TRY(tsg_validate_pointers(col_ptr, N, nnz, "packed-CSC col_ptr"));
        TSG_CHECK(nnz == 0 || (row_idx && vals), TSG_ERR_INVALID, "packed-CSC row_idx / vals is NULL");
        for (long long i = 0, nb = (nnz + 4) / 5; i < nb; ++i)
            TSG_CHECK(vals[i] < 243, TSG_ERR_INVALID, "packed-CSC vals[%lld] = %d is not five base-3 digits", i, (int)vals[i]);
        int32_t *dcp = nullptr, *dri = nullptr;
        uint8_t *dv = nullptr;
        int8_t *dW = nullptr;
        const long long nbytes = (nnz + 4) / 5;
        const size_t wbytes = (size_t)K * N;
        int s = TSG_OK;
        if (cudaMalloc(&dcp, (size_t)(N + 1) * 4) != cudaSuccess || cudaMalloc(&dri, (size_t)nnz * 4 + 16) != cudaSuccess ||
            cudaMalloc(&dv, (size_t)nbytes + 16) != cudaSuccess || cudaMalloc(&dW, wbytes ? wbytes : 1) != cudaSuccess)
        {
            tsg_set_error("allocation failed");
            s = TSG_ERR_NOMEM;
        }
        if (s == TSG_OK)
        {
            cudaError_t ce = cudaMemcpy(dcp, col_ptr, (size_t)(N + 1) * 4, cudaMemcpyHostToDevice);
            if (ce == cudaSuccess && nnz)
                ce = cudaMemcpy(dri, row_idx, (size_t)nnz * 4, cudaMemcpyHostToDevice);
            if (ce == cudaSuccess && nnz)
                ce = cudaMemcpy(dv, vals, (size_t)nbytes, cudaMemcpyHostToDevice);
            if (ce == cudaSuccess)
                ce = cudaMemset(dW, 0, wbytes);
            if (ce != cudaSuccess)
            {
                tsg_set_error("upload of packed-CSC arrays failed: %s", cudaGetErrorString(ce));
                s = TSG_ERR_CUDA;
            }
            // caller-made arrays: rows inside [0, K) and strictly ascending per column, before the
            // decoder scatters through them
            if (s == TSG_OK)
                s = tsg_validate_lists(dcp, dri, N, K, nullptr, "packed-CSC row_idx");
        }
        if (s == TSG_OK)
        {
            if (N > 0 && K > 0)
                pcsc_to_dense_kernel<int8_t><<<N, 256>>>(dcp, dri, dv, K, N, dW);
            g_tsg_launches.fetch_add(1, std::memory_order_relaxed);
            if (cudaDeviceSynchronize() != cudaSuccess)
            {
                tsg_set_error("decode of packed-CSC arrays failed: %s", cudaGetErrorString(cudaGetLastError()));
                s = TSG_ERR_CUDA;
            }
        }
        if (s == TSG_OK)
            s = tsg_pcsc_from_dense_dev(dW, 1, K, N, nullptr, out);
        cudaFree(dcp), cudaFree(dri), cudaFree(dv), cudaFree(dW);
        return s;
    }

    int tsg_pcsc_sizes(const tsg_pcsc *h, int64_t *nnz, int64_t *val_bytes)
    {
        TSG_CHECK(h, TSG_ERR_INVALID, "NULL argument");
        if (nnz)
            *nnz = h->nnz;
        if (val_bytes)
            *val_bytes = h->nbytes;
        return TSG_OK;
    }

    // bytes of the packed structure: 4(N+1) + 4 nnz + ceil(nnz/5)
    int tsg_pcsc_data_structure_size(const tsg_pcsc *h, int64_t *bytes)
    {
        TSG_CHECK(h && bytes, TSG_ERR_INVALID, "NULL argument");
        *bytes = 4ll * (h->fwd->N + 1) + 4ll * h->nnz + h->nbytes;
        return TSG_OK;
    }

    int tsg_pcsc_export(const tsg_pcsc *h, int32_t *col_ptr, int32_t *row_idx, uint8_t *vals)
    {
        TSG_CHECK(h, TSG_ERR_INVALID, "NULL argument");
        DeviceGuard guard_(h->fwd->device); // the handle lives on the device that was current when it was built
        if (col_ptr)
            TSG_CUDA(cudaMemcpy(col_ptr, h->col_ptr, (size_t)(h->fwd->N + 1) * 4, cudaMemcpyDeviceToHost));
        if (row_idx && h->nnz)
            TSG_CUDA(cudaMemcpy(row_idx, h->row_idx, (size_t)h->nnz * 4, cudaMemcpyDeviceToHost));
        if (vals && h->nbytes)
            TSG_CUDA(cudaMemcpy(vals, h->vals, (size_t)h->nbytes, cudaMemcpyDeviceToHost));
        return TSG_OK;
    }

    int tsg_pcsc_to_dense(const tsg_pcsc *h, int32_t *W_host)
    {
        TSG_CHECK(h && (W_host || (long long)h->fwd->K * h->fwd->N == 0), TSG_ERR_INVALID, "NULL argument");
        DeviceGuard guard_(h->fwd->device); // the handle lives on the device that was current when it was built
        const int K = h->fwd->K, N = h->fwd->N;
        const size_t bytes = (size_t)K * N * 4;
        if (!bytes)
            return TSG_OK;
        int32_t *dW = nullptr;
        TSG_CUDA(cudaMalloc(&dW, bytes));
        cudaStream_t st = h->fwd->stream;
        cudaMemsetAsync(dW, 0, bytes, st);
        pcsc_to_dense_kernel<int32_t><<<N, 256, 0, st>>>(h->col_ptr, h->row_idx, h->vals, K, N, dW);
        g_tsg_launches.fetch_add(1, std::memory_order_relaxed);
        cudaError_t e = cudaMemcpyAsync(W_host, dW, bytes, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess)
            e = cudaStreamSynchronize(st);
        cudaFree(dW);
        TSG_CHECK(e == cudaSuccess, TSG_ERR_CUDA, "packed CSC -> dense failed: %s", cudaGetErrorString(e));
        return TSG_OK;
    }

    int tsg_pcsc_spmm_dev(tsg_pcsc *h, int algo, const float *X_dev, int64_t ldx, const float *b_dev,
                          const float *alpha_dev, float *Y_dev, int64_t ldy, int M, void *stream)
    {
        TSG_CHECK(h, TSG_ERR_INVALID, "matrix is NULL");
        DeviceGuard guard_(h->fwd->device); // the handle lives on the device that was current when it was built
        if (algo != TSG_ALGO_PCSC_GATHER)
            return tsg_spmm_dev(h->fwd, algo, X_dev, ldx, b_dev, alpha_dev, Y_dev, ldy, M, stream);
        const tsg_matrix *m = h->fwd;
        if (M <= 0 || m->N == 0)
            return TSG_OK;
        TSG_CHECK(X_dev && b_dev && Y_dev, TSG_ERR_INVALID, "X, b and Y must be non-NULL");
        const size_t smem = (size_t)m->K * 16 + 16;
        TSG_CHECK(smem <= m->smem_optin, TSG_ERR_UNSUPPORTED, "pcsc_gather: K=%d does not fit shared memory", m->K);
        // largest opt-in granted so far per device (the attribute is per device and function); atomic: two host
        // threads may launch the same kernel — a repeated, equal cudaFuncSetAttribute is harmless, a torn size is not
        static std::atomic<size_t> configured[64];
        std::atomic<size_t> &have = configured[m->device & 63];
        if (have.load(std::memory_order_acquire) < smem)
        {
            TSG_CUDA(cudaFuncSetAttribute(pcsc_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            size_t seen = have.load(std::memory_order_relaxed);
            while (seen < smem && !have.compare_exchange_weak(seen, smem, std::memory_order_release))
                ;
        }
        const int per = kPcscWarps;
        int gx = (m->N + per - 1) / per;
        const int cap = 2 * (m->sm_count > 0 ? m->sm_count : 148);
        if (gx > cap)
            gx = cap;
        dim3 grid(gx, (M + 3) / 4);
        TSG_CHECK(grid.y <= 65535, TSG_ERR_UNSUPPORTED, "pcsc_gather: M too large");
        pcsc_gather_kernel<<<grid, kPcscWarps * 32, smem, (cudaStream_t)stream>>>(
            h->col_ptr, h->row_idx, h->vals, X_dev, ldx, b_dev, alpha_dev, Y_dev, ldy, M, m->K, m->N);
        TSG_LAUNCHED();
        return TSG_OK;
    }

    int tsg_pcsc_spmm_pick(const tsg_pcsc *h, int M, int *algo)
    {
        TSG_CHECK(h, TSG_ERR_INVALID, "matrix is NULL");
        return tsg_spmm_pick(h->fwd, M, algo);
    }

    int tsg_pcsc_spmm(tsg_pcsc *h, int algo, const float *X, const float *b, const float *alpha, float *Y, int M,
                      int N, int K)
    {
        TSG_CHECK(h, TSG_ERR_INVALID, "matrix is NULL");
        DeviceGuard guard_(h->fwd->device); // the handle lives on the device that was current when it was built
        if (algo != TSG_ALGO_PCSC_GATHER)
            return tsg_spmm_algo(h->fwd, algo, X, b, alpha, Y, M, N, K);
        TSG_CHECK(N == h->fwd->N && K == h->fwd->K, TSG_ERR_INVALID, "shape mismatch");
        if (M <= 0 || N == 0)
            return TSG_OK;
        TSG_CHECK(X && b && Y, TSG_ERR_INVALID, "X, b and Y must be non-NULL");
        return run_host(h->fwd, X, b, alpha, Y, M, N, K,
                        [&](float *dX, float *dB, float *dA, float *dY, cudaStream_t st) -> int {
                            return tsg_pcsc_spmm_dev(h, TSG_ALGO_PCSC_GATHER, dX, K, dB, dA, dY, N, M, st);
                        });
    }

    // =========================================== PCSR ===========================================
    void tsg_pcsr_destroy(tsg_pcsr *h)
    {
        if (!h)
            return;
        tsg_destroy(h->fwd);
        cudaFree(h->row_ptr), cudaFree(h->col_idx), cudaFree(h->vals);
        delete h;
    }

    // W in HBM (int32 or int8 row-major): engine matrix of W, then the packed rows from the planes
    // of W^T (a temporary engine matrix built on the transposed strides, as for TCSR)
    int tsg_pcsr_from_dense_dev(const void *W_dev, int elem_bytes, int K, int N, void *stream, tsg_pcsr **out)
    {
        TSG_CHECK(out != nullptr, TSG_ERR_INVALID, "out is NULL");
        *out = nullptr;
        tsg_pcsr *h = new (std::nothrow) tsg_pcsr();
        TSG_CHECK(h != nullptr, TSG_ERR_NOMEM, "host allocation failed");
        tsg_matrix *t = nullptr;
        int s = tsg_tcsc_from_dense_dev(W_dev, elem_bytes, K, N, N, 0, N, stream, &h->fwd);
        if (s == TSG_OK)
            s = tsg_new_matrix(N, K, &t); // W^T: N rows, K columns
        if (s == TSG_OK)
            s = tsg_build_from_dense_dev(t, W_dev, elem_bytes, /*ld=*/1, 0, t->stream, /*cs=*/N);
        if (s == TSG_OK)
            s = build_packed(t, &h->row_ptr, &h->col_idx, &h->vals, &h->nnz, &h->nbytes);
        tsg_destroy(t);
        if (s != TSG_OK)
        {
            tsg_pcsr_destroy(h);
            return s;
        }
        *out = h;
        return TSG_OK;
    }

    int tsg_pcsr_from_dense(const int32_t *W_host, int K, int N, tsg_pcsr **out)
    {
        TSG_CHECK(out != nullptr, TSG_ERR_INVALID, "out is NULL");
        *out = nullptr;
        int32_t *dW = nullptr;
        TSG_TRY(upload_dense(W_host, K, N, &dW));
        const int s = tsg_pcsr_from_dense_dev(dW, 4, K, N, nullptr, out);
        cudaFree(dW);
        return s;
    }

    // Adopt arrays in the packed row-major layout (interchange): decoded on the device, rebuilt.
    int tsg_pcsr_from_arrays(const int32_t *row_ptr, const int32_t *col_idx, const uint8_t *vals, int K, int N,
                             tsg_pcsr **out)
    {
        TSG_CHECK(out != nullptr, TSG_ERR_INVALID, "out is NULL");
        *out = nullptr;
        TSG_CHECK(row_ptr && K >= 0 && N >= 0, TSG_ERR_INVALID, "bad arguments");
        int usable = 0;
        tsg_device_count(&usable);
        TSG_CHECK(usable > 0, TSG_ERR_NO_DEVICE, "no sm_100 device visible (libtsg has no CPU fallback)");
        TSG_CHECK((long long)K * N <= (long long)INT32_MAX, TSG_ERR_OVERFLOW, "K*N too large");
        const long long nnz = row_ptr[K];
        TSG_TRY(tsg_validate_pointers(row_ptr, K, nnz, "packed-CSR row_ptr"));
        TSG_CHECK(nnz == 0 || (col_idx && vals), TSG_ERR_INVALID, "packed-CSR col_idx / vals is NULL");
        for (long long i = 0, nb = (nnz + 4) / 5; i < nb; ++i)
            TSG_CHECK(vals[i] < 243, TSG_ERR_INVALID, "packed-CSR vals[%lld] = %d is not five base-3 digits", i, (int)vals[i]);
        int32_t *drp = nullptr, *dci = nullptr;
        uint8_t *dv = nullptr;
        int8_t *dW = nullptr;
        const long long nbytes = (nnz + 4) / 5;
        const size_t wbytes = (size_t)K * N;
        int s = TSG_OK;
        if (cudaMalloc(&drp, (size_t)(K + 1) * 4) != cudaSuccess || cudaMalloc(&dci, (size_t)nnz * 4 + 16) != cudaSuccess ||
            cudaMalloc(&dv, (size_t)nbytes + 16) != cudaSuccess || cudaMalloc(&dW, wbytes ? wbytes : 1) != cudaSuccess)
        {
            tsg_set_error("allocation failed");
            s = TSG_ERR_NOMEM;
        }
        if (s == TSG_OK)
        {
            cudaError_t ce = cudaMemcpy(drp, row_ptr, (size_t)(K + 1) * 4, cudaMemcpyHostToDevice);
            if (ce == cudaSuccess && nnz)
                ce = cudaMemcpy(dci, col_idx, (size_t)nnz * 4, cudaMemcpyHostToDevice);
            if (ce == cudaSuccess && nnz)
                ce = cudaMemcpy(dv, vals, (size_t)nbytes, cudaMemcpyHostToDevice);
            if (ce == cudaSuccess)
                ce = cudaMemset(dW, 0, wbytes);
            if (ce != cudaSuccess)
            {
                tsg_set_error("upload of packed-CSR arrays failed: %s", cudaGetErrorString(ce));
                s = TSG_ERR_CUDA;
            }
            if (s == TSG_OK)
                s = tsg_validate_lists(drp, dci, K, N, nullptr, "packed-CSR col_idx");
        }
        if (s == TSG_OK)
        {
            if (N > 0 && K > 0)
                pcsr_to_dense_kernel<int8_t><<<K, 256>>>(drp, dci, dv, K, N, dW);
            g_tsg_launches.fetch_add(1, std::memory_order_relaxed);
            if (cudaDeviceSynchronize() != cudaSuccess)
            {
                tsg_set_error("decode of packed-CSR arrays failed: %s", cudaGetErrorString(cudaGetLastError()));
                s = TSG_ERR_CUDA;
            }
        }
        if (s == TSG_OK)
            s = tsg_pcsr_from_dense_dev(dW, 1, K, N, nullptr, out);
        cudaFree(drp), cudaFree(dci), cudaFree(dv), cudaFree(dW);
        return s;
    }

    int tsg_pcsr_sizes(const tsg_pcsr *h, int64_t *nnz, int64_t *val_bytes)
    {
        TSG_CHECK(h, TSG_ERR_INVALID, "NULL argument");
        if (nnz)
            *nnz = h->nnz;
        if (val_bytes)
            *val_bytes = h->nbytes;
        return TSG_OK;
    }

    // bytes of the packed structure: 4(K+1) + 4 nnz + ceil(nnz/5)
    int tsg_pcsr_data_structure_size(const tsg_pcsr *h, int64_t *bytes)
    {
        TSG_CHECK(h && bytes, TSG_ERR_INVALID, "NULL argument");
        *bytes = 4ll * (h->fwd->K + 1) + 4ll * h->nnz + h->nbytes;
        return TSG_OK;
    }

    int tsg_pcsr_export(const tsg_pcsr *h, int32_t *row_ptr, int32_t *col_idx, uint8_t *vals)
    {
        TSG_CHECK(h, TSG_ERR_INVALID, "NULL argument");
        DeviceGuard guard_(h->fwd->device); // the handle lives on the device that was current when it was built
        if (row_ptr)
            TSG_CUDA(cudaMemcpy(row_ptr, h->row_ptr, (size_t)(h->fwd->K + 1) * 4, cudaMemcpyDeviceToHost));
        if (col_idx && h->nnz)
            TSG_CUDA(cudaMemcpy(col_idx, h->col_idx, (size_t)h->nnz * 4, cudaMemcpyDeviceToHost));
        if (vals && h->nbytes)
            TSG_CUDA(cudaMemcpy(vals, h->vals, (size_t)h->nbytes, cudaMemcpyDeviceToHost));
        return TSG_OK;
    }

    int tsg_pcsr_to_dense(const tsg_pcsr *h, int32_t *W_host)
    {
        TSG_CHECK(h && (W_host || (long long)h->fwd->K * h->fwd->N == 0), TSG_ERR_INVALID, "NULL argument");
        DeviceGuard guard_(h->fwd->device); // the handle lives on the device that was current when it was built
        const int K = h->fwd->K, N = h->fwd->N;
        const size_t bytes = (size_t)K * N * 4;
        if (!bytes)
            return TSG_OK;
        int32_t *dW = nullptr;
        TSG_CUDA(cudaMalloc(&dW, bytes));
        cudaStream_t st = h->fwd->stream;
        cudaMemsetAsync(dW, 0, bytes, st);
        pcsr_to_dense_kernel<int32_t><<<K, 256, 0, st>>>(h->row_ptr, h->col_idx, h->vals, K, N, dW);
        g_tsg_launches.fetch_add(1, std::memory_order_relaxed);
        cudaError_t e = cudaMemcpyAsync(W_host, dW, bytes, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess)
            e = cudaStreamSynchronize(st);
        cudaFree(dW);
        TSG_CHECK(e == cudaSuccess, TSG_ERR_CUDA, "packed CSR -> dense failed: %s", cudaGetErrorString(e));
        return TSG_OK;
    }

    int tsg_pcsr_spmm(tsg_pcsr *h, int algo, const float *X, const float *b, const float *alpha, float *Y, int M,
                      int N, int K)
    {
        TSG_CHECK(h, TSG_ERR_INVALID, "matrix is NULL");
        DeviceGuard guard_(h->fwd->device); // the handle lives on the device that was current when it was built
        if (algo != TSG_ALGO_PCSR_SEQ)
            return tsg_spmm_algo(h->fwd, algo, X, b, alpha, Y, M, N, K);
        TSG_CHECK(N == h->fwd->N && K == h->fwd->K, TSG_ERR_INVALID, "shape mismatch");
        if (M <= 0 || N == 0)
            return TSG_OK;
        TSG_CHECK(X && b && Y, TSG_ERR_INVALID, "X, b and Y must be non-NULL");
        return run_host(h->fwd, X, b, alpha, Y, M, N, K,
                        [&](float *dX, float *dB, float *dA, float *dY, cudaStream_t st) -> int {
                            pcsr_seq_kernel<<<M, 1024, 0, st>>>(h->row_ptr, h->col_idx, h->vals, dX, K, dB, dA, dY, N, K, N);
                            TSG_LAUNCHED();
                            return (int)TSG_OK;
                        });
    }

} // extern "C"
