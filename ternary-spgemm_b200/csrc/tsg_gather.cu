// tsg_gather.cu — TCSC gather-add kernels.
//
// Replaces BaseTCSC<float> (reference cpp_impl/comp.h:25-69) and BaseTCSC_PreLU<float>
// (cpp_impl/comp_prelu.h:12-70):   Y[m,n] = Σ_{k∈pos(n)} X[m,k] − Σ_{k∈neg(n)} X[m,k] + b[n]
// and optionally  Y = y > 0 ? y : alpha[n]·y.
//
// Two kernels:
//
//  gather_slices_kernel<MT>  — the production kernel for decode-shaped M.  HBM-bound: the only
//    large operand is the index stream (4 B per non-zero), read exactly once per MT rows of X.
//      * persistent grid, one 1024-thread CTA per SM; each CTA owns a contiguous column range
//        chosen so that all CTAs hold the same number of non-zeros (tsg_matrix::part);
//      * the CTA's X row-tile (MT rows × K) is staged once in shared memory, k-major
//        (Xs[k*MT+m]) so that one LDS.32/64/128 fetches all MT operands of a non-zero;
//      * columns are handled in groups of GC; the group's pos stream and neg stream are each
//        cut into 32 EQUAL slices, one per warp, regardless of column boundaries — every warp
//        streams the same number of bytes with 128-bit ld.global.nc.L1::no_allocate loads,
//        eight of them in flight per lane;
//      * inside a slice a warp walks the (few) columns it intersects; each piece is reduced
//        with warp shuffles and parked in a per-warp table in shared memory;
//      * after one __syncthreads the pieces of a column are summed in warp order (fixed order:
//        results are run-to-run deterministic), bias / PReLU applied, Y written coalesced.
//    Summation order differs from the reference's single accumulator (it is a tree), which is
//    exact for the reference's integer-valued inputs and within 1e-5 relative otherwise.
//
//  gather_seq_kernel — one thread per Y[m,n] walking the lists in the reference's order with one
//    fp32 accumulator: bit-identical to BaseTCSC for arbitrary fp32 X.  Slow (uncoalesced index
//    reads); it exists as the on-device statement of the reference's arithmetic.
#include "tsg_internal.cuh"

namespace
{

constexpr int kWarps = 32;          // warps per CTA (1024 threads)
constexpr int kGC = 32;             // columns per group
constexpr int kUnroll = 8;          // 128-bit index loads in flight per lane

__device__ __forceinline__ int4 ldg_stream(const int *p)
{
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

template <int MT>
struct Acc
{
    float v[MT];
};

template <int MT>
__device__ __forceinline__ void gather_add(const float *__restrict__ Xs, int k, bool valid,
                                           Acc<MT> &a)
{
    if constexpr (MT == 1)
    {
        const float x = Xs[k];
        a.v[0] += valid ? x : 0.0f;
    }
    else if constexpr (MT == 2)
    {
        const float2 x = *reinterpret_cast<const float2 *>(Xs + 2 * k);
        a.v[0] += valid ? x.x : 0.0f;
        a.v[1] += valid ? x.y : 0.0f;
    }
    else
    {
        const float4 x = *reinterpret_cast<const float4 *>(Xs + 4 * k);
        a.v[0] += valid ? x.x : 0.0f;
        a.v[1] += valid ? x.y : 0.0f;
        a.v[2] += valid ? x.z : 0.0f;
        a.v[3] += valid ? x.w : 0.0f;
    }
}

// Sum X over the index range [lo, hi) of `idx` (one column piece), all lanes cooperating.
template <int MT>
__device__ __forceinline__ Acc<MT> piece_sum(const int *__restrict__ idx, int lo, int hi,
                                             const float *__restrict__ Xs, int lane)
{
    Acc<MT> a;
#pragma unroll
    for (int m = 0; m < MT; ++m)
        a.v[m] = 0.0f;
    for (int q0 = (lo & ~3) + lane * 4; q0 < hi; q0 += 128 * kUnroll)
    {
        int4 v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
        {
            const int q = q0 + u * 128;
            v[u] = (q < hi) ? ldg_stream(idx + q) : make_int4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
        {
            const int q = q0 + u * 128;
            gather_add<MT>(Xs, v[u].x, q + 0 >= lo && q + 0 < hi, a);
            gather_add<MT>(Xs, v[u].y, q + 1 >= lo && q + 1 < hi, a);
            gather_add<MT>(Xs, v[u].z, q + 2 >= lo && q + 2 < hi, a);
            gather_add<MT>(Xs, v[u].w, q + 3 >= lo && q + 3 < hi, a);
        }
    }
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            a.v[m] += __shfl_xor_sync(0xffffffffu, a.v[m], o);
    return a;
}

// One warp's slice [a, b) of one sign's stream; cs = smem copy of the group's pointers
// (cs[0..gcols]), table = this warp's row of the piece table for that sign.
template <int MT>
__device__ __forceinline__ void slice_walk(const int *__restrict__ idx, int a, int b,
                                           const int *cs, int gcols,
                                           const float *__restrict__ Xs, float *table, int lane)
{
    if (a >= b)
        return;
    // largest c in [0, gcols) with cs[c] <= a   (cs[0] <= a < cs[gcols] holds)
    int c = 0, hi = gcols;
    while (hi - c > 1)
    {
        const int mid = (c + hi) >> 1;
        if (cs[mid] <= a)
            c = mid;
        else
            hi = mid;
    }
    int pos = a;
    while (pos < b)
    {
        const int cend = min(b, cs[c + 1]);
        if (cend > pos)
        {
            const Acc<MT> s = piece_sum<MT>(idx, pos, cend, Xs, lane);
            if (lane == 0)
            {
#pragma unroll
                for (int m = 0; m < MT; ++m)
                    table[c * MT + m] = s.v[m];
            }
            pos = cend;
        }
        ++c;
    }
}

template <int MT>
__global__ void __launch_bounds__(kWarps * 32, 1)
gather_slices_kernel(const int *__restrict__ csp, const int *__restrict__ csn,
                     const int *__restrict__ rip, const int *__restrict__ rin,
                     const int *__restrict__ part, const float *__restrict__ X, int64_t ldx,
                     const float *__restrict__ bias, const float *__restrict__ alpha,
                     float *__restrict__ Y, int64_t ldy, int M, int K)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *Xs = reinterpret_cast<float *>(smem_raw);                    // K*MT
    float *tab = Xs + (size_t)K * MT;                                   // 2*kWarps*kGC*MT
    int *cs = reinterpret_cast<int *>(tab + 2 * kWarps * kGC * MT);     // 2*(kGC+1)

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int m0 = blockIdx.y * MT;
    const int col_lo = part[blockIdx.x], col_hi = part[blockIdx.x + 1];
    if (col_lo >= col_hi)
        return;

    // stage the X row tile, k-major; rows past M read as zero
    for (int k = tid; k < K; k += kWarps * 32)
    {
#pragma unroll
        for (int m = 0; m < MT; ++m)
            Xs[k * MT + m] = (m0 + m < M) ? X[(int64_t)(m0 + m) * ldx + k] : 0.0f;
    }
    for (int i = tid; i < 2 * kWarps * kGC * MT; i += kWarps * 32)
        tab[i] = 0.0f;

    float *tab_pos = tab + (size_t)wid * kGC * MT;
    float *tab_neg = tab + (size_t)(kWarps + wid) * kGC * MT;

    for (int g0 = col_lo; g0 < col_hi; g0 += kGC)
    {
        const int gcols = min(kGC, col_hi - g0);
        if (tid <= gcols)
            cs[tid] = csp[g0 + tid];
        else if (tid >= 64 && tid - 64 <= gcols)
            cs[kGC + 1 + tid - 64] = csn[g0 + tid - 64];
        __syncthreads(); // also covers Xs / tab initialisation on the first trip
        {
            const int p0 = cs[0], p1 = cs[gcols];
            const long long len = p1 - p0;
            const int a = p0 + (int)((len * wid) / kWarps), b = p0 + (int)((len * (wid + 1)) / kWarps);
            slice_walk<MT>(rip, a, b, cs, gcols, Xs, tab_pos, lane);
        }
        {
            const int *cq = cs + kGC + 1;
            const int p0 = cq[0], p1 = cq[gcols];
            const long long len = p1 - p0;
            const int a = p0 + (int)((len * wid) / kWarps), b = p0 + (int)((len * (wid + 1)) / kWarps);
            slice_walk<MT>(rin, a, b, cq, gcols, Xs, tab_neg, lane);
        }
        __syncthreads();
        // combine: thread -> (m, c), c fastest so that Y stores are coalesced
        if (tid < kGC * MT)
        {
            const int c = tid % kGC, m = tid / kGC;
            float sp = 0.0f, sn = 0.0f;
#pragma unroll 8
            for (int w = 0; w < kWarps; ++w)
            {
                float *tp = tab + ((size_t)w * kGC + c) * MT + m;
                float *tn = tab + ((size_t)(kWarps + w) * kGC + c) * MT + m;
                sp += *tp;
                sn += *tn;
                *tp = 0.0f;
                *tn = 0.0f;
            }
            if (c < gcols && m0 + m < M)
            {
                const int n = g0 + c;
                float y = (sp - sn) + bias[n];
                if (alpha != nullptr)
                    y = (y > 0.0f) ? y : alpha[n] * y;
                Y[(int64_t)(m0 + m) * ldy + n] = y;
            }
        }
        // no barrier needed here: the next trip only writes cs (every warp is past its reads of
        // cs) and the barrier at the top of the loop orders this combine before new table writes.
    }
}

// contiguous, nnz-balanced column ranges: part[i] = first column whose running nnz reaches
// total*i/ctas.
__global__ void partition_kernel(const int *__restrict__ csp, const int *__restrict__ csn, int N,
                                 int ctas, int *__restrict__ part)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > ctas)
        return;
    const long long total = (long long)csp[N] + csn[N];
    const long long target = (total * i) / ctas;
    int lo = 0, hi = N; // smallest c in [0,N] with csp[c]+csn[c] >= target
    while (lo < hi)
    {
        const int mid = (lo + hi) >> 1;
        if ((long long)csp[mid] + csn[mid] >= target)
            hi = mid;
        else
            lo = mid + 1;
    }
    if (i == 0)
        lo = 0;
    if (i == ctas)
        lo = N;
    // never hand a CTA zero nnz but many empty columns at the very end: fine, ranges only need
    // to tile [0,N) monotonically, which the monotone prefix guarantees.
    part[i] = lo;
}

// Reference-order kernel: bit-identical to BaseTCSC / BaseTCSC_PreLU.
__global__ void __launch_bounds__(256)
gather_seq_kernel(const int *__restrict__ csp, const int *__restrict__ csn,
                  const int *__restrict__ rip, const int *__restrict__ rin,
                  const float *__restrict__ X, int64_t ldx, const float *__restrict__ bias,
                  const float *__restrict__ alpha, float *__restrict__ Y, int64_t ldy, int M, int N)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = blockIdx.y;
    if (n >= N || m >= M)
        return;
    const float *x = X + (int64_t)m * ldx;
    float y = 0.0f;
    for (int i = csp[n]; i < csp[n + 1]; ++i) // comp.h:44-51
        y += x[rip[i]];
    for (int i = csn[n]; i < csn[n + 1]; ++i) // comp.h:54-61
        y -= x[rin[i]];
    y = y + bias[n];                          // comp.h:63 / comp_prelu.h:51
    if (alpha != nullptr)
        y = (y > 0.0f) ? y : alpha[n] * y;    // comp_prelu.h:57-64
    Y[(int64_t)m * ldy + n] = y;
}

size_t gather_smem_bytes(int K, int MT)
{
    return (size_t)K * MT * 4 + (size_t)2 * kWarps * kGC * MT * 4 + (size_t)2 * (kGC + 1) * 4;
}

} // namespace

static int ensure_partition(tsg_matrix *m, cudaStream_t st)
{
    const int ctas = m->N < m->sm_count ? (m->N > 0 ? m->N : 1) : m->sm_count;
    if (m->part != nullptr && m->part_ctas == ctas)
        return TSG_OK;
    if (m->part)
        cudaFree(m->part);
    m->part = nullptr;
    TSG_CUDA(cudaMalloc(&m->part, (size_t)(ctas + 1) * 4));
    partition_kernel<<<(ctas + 1 + 127) / 128, 128, 0, st>>>(m->csp, m->csn, m->N, ctas, m->part);
    TSG_LAUNCHED();
    m->part_ctas = ctas;
    return TSG_OK;
}

template <int MT>
static int launch_slices(tsg_matrix *m, const float *X, int64_t ldx, const float *b,
                         const float *alpha, float *Y, int64_t ldy, int M, cudaStream_t st)
{
    const size_t smem = gather_smem_bytes(m->K, MT);
    static size_t configured[64] = {0}; // per device: the attribute is per (device, function)
    size_t &have = configured[m->device & 63];
    if (have < smem)
    {
        TSG_CUDA(cudaFuncSetAttribute(gather_slices_kernel<MT>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        have = smem;
    }
    dim3 grid(m->part_ctas, (M + MT - 1) / MT);
    gather_slices_kernel<MT><<<grid, kWarps * 32, smem, st>>>(m->csp, m->csn, m->rip, m->rin,
                                                              m->part, X, ldx, b, alpha, Y, ldy, M,
                                                              m->K);
    TSG_LAUNCHED();
    return TSG_OK;
}

int tsg_launch_gather(tsg_matrix *m, const float *X, int64_t ldx, const float *b,
                      const float *alpha, float *Y, int64_t ldy, int M, cudaStream_t st)
{
    if (M <= 0 || m->N == 0)
        return TSG_OK;
    TSG_TRY(ensure_partition(m, st));
    // widest row tile whose X staging fits in shared memory
    int MT = (M >= 4) ? 4 : (M >= 2 ? 2 : 1);
    while (MT > 1 && gather_smem_bytes(m->K, MT) > m->smem_optin)
        MT >>= 1;
    TSG_CHECK(gather_smem_bytes(m->K, MT) <= m->smem_optin, TSG_ERR_UNSUPPORTED,
              "gather kernel: K=%d does not fit shared memory (%zu B needed, %zu B available)",
              m->K, gather_smem_bytes(m->K, MT), m->smem_optin);
    switch (MT)
    {
    case 4:
        return launch_slices<4>(m, X, ldx, b, alpha, Y, ldy, M, st);
    case 2:
        return launch_slices<2>(m, X, ldx, b, alpha, Y, ldy, M, st);
    default:
        return launch_slices<1>(m, X, ldx, b, alpha, Y, ldy, M, st);
    }
}

int tsg_launch_gather_seq(tsg_matrix *m, const float *X, int64_t ldx, const float *b,
                          const float *alpha, float *Y, int64_t ldy, int M, cudaStream_t st)
{
    if (M <= 0 || m->N == 0)
        return TSG_OK;
    TSG_CHECK(M <= 65535, TSG_ERR_UNSUPPORTED, "gather_seq: M=%d exceeds grid.y", M);
    dim3 grid((m->N + 255) / 256, M);
    gather_seq_kernel<<<grid, 256, 0, st>>>(m->csp, m->csn, m->rip, m->rin, X, ldx, b, alpha, Y,
                                            ldy, M, m->N);
    TSG_LAUNCHED();
    return TSG_OK;
}
