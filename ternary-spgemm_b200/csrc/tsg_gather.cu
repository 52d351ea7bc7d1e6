// tsg_gather.cu — TCSC gather-add kernels.
//
// Replaces BaseTCSC<float> (reference cpp_impl/comp.h:25-69) and BaseTCSC_PreLU<float>
// (cpp_impl/comp_prelu.h:12-70):   Y[m,n] = Σ_{k∈pos(n)} X[m,k] − Σ_{k∈neg(n)} X[m,k] + b[n]
// and optionally  Y = y > 0 ? y : alpha[n]·y.
//
// Two kernels:
//
//  gather_pieces_kernel<MT>  — the production kernel for decode-shaped M.  HBM-bound: the only
//    large operand is the index stream (4 B per non-zero), read exactly once per MT rows of X.
//      * persistent grid, one 512-thread CTA per SM; CTA i owns the contiguous column range
//        [N·i/G, N·(i+1)/G) (computed from blockIdx alone, so its pointer loads issue at once);
//      * every index list (one column, one sign) is cut into P equal pieces (P a power of two
//        chosen on the host so that a piece is ~256-384 indices).  Pieces are dealt round-robin
//        to the 16 warps: neighbouring warps stream neighbouring memory and all warps carry the
//        same load to within one piece;
//      * a piece is fetched with five 128-bit ld.global.nc.L1::no_allocate loads per lane, all
//        issued together; the NEXT piece's loads are issued before the current one is consumed
//        (two register buffers), so each lane keeps ten 128-bit loads in flight;
//      * the CTA's X row-tile (MT rows × K) is staged once in shared memory, k-major
//        (Xs[k*MT+m]) so that one LDS.32/64/128 fetches all MT operands of a non-zero.  X and
//        the column pointers are requested before any index load: the SM's load path returns
//        in issue order, so the small operands come back first;
//      * a piece never crosses a column, so the inner loop has no segmentation: full 128-index
//        blocks take four unmasked adds per lane, the first/last block of a piece is masked.
//        One warp-shuffle reduction per piece, result parked in a table in shared memory;
//      * after ONE __syncthreads the P pieces of each list are added in piece order (a fixed
//        order: results are run-to-run deterministic), bias / PReLU applied, Y written coalesced.
//    Summation order differs from the reference's single accumulator (it is a tree), which is
//    exact for the reference's integer-valued inputs and within 1e-5 relative otherwise.
//
//  gather_seq_kernel — one thread per Y[m,n] walking the lists in the reference's order with one
//    fp32 accumulator: bit-identical to BaseTCSC for arbitrary fp32 X.  Slow (uncoalesced index
//    reads); it exists as the on-device statement of the reference's arithmetic.
#include "tsg_internal.cuh"

#include <stdlib.h>

namespace
{

constexpr int kWarps = 16;     // warps per CTA (512 threads, one CTA per SM, <=128 regs/thread)
constexpr int kColCap = 1024;  // columns per pass (bounds the pointer staging)
constexpr int kTabFloats = 8192; // piece-sum table (floats) per pass
constexpr int kU = 6;          // 128-bit loads per lane per piece batch (8 lanes x 6 x 4 = 192 indices)

__device__ __forceinline__ int4 ldg_stream(const int *p)
{
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ float ldg_f32_ordered(const float *p)
{
    float r; // volatile asm: keeps its place in program order ahead of the index loads
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

__device__ __forceinline__ int ldg_s32_ordered(const int *p)
{
    int r;
    asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

template <int MT>
struct Acc
{
    float v[MT];
    __device__ __forceinline__ void zero()
    {
#pragma unroll
        for (int m = 0; m < MT; ++m)
            v[m] = 0.0f;
    }
    __device__ __forceinline__ void warp_sum()
    {
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
                v[m] += __shfl_xor_sync(0xffffffffu, v[m], o);
    }
};

// X operand fetch: explicit shared-window address (32-bit), one LDS.32/64/128 per non-zero.
template <int MT>
__device__ __forceinline__ Acc<MT> gather(uint32_t xs_base, int k)
{
    Acc<MT> a;
    const uint32_t addr = xs_base + (uint32_t)k * (4u * MT);
    if constexpr (MT == 1)
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a.v[0]) : "r"(addr));
    else if constexpr (MT == 2)
        asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(a.v[0]), "=f"(a.v[1]) : "r"(addr));
    else
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(a.v[0]), "=f"(a.v[1]), "=f"(a.v[2]), "=f"(a.v[3])
                     : "r"(addr));
    return a;
}

// One piece = int4 range [a4, b4) of one padded list (tsg_matrix::rip4/rin4).  A piece is worked
// on by a TEAM of 8 lanes, so a warp carries four pieces at a time and the fixed per-piece cost
// (decode, reduction, store) is paid once per four pieces in issue slots.  Lists are padded to
// whole int4s with the sentinel row K (Xs[K] = 0), so there are no masks anywhere.
struct Piece
{
    const int4 *idx;
    int a4, b4;
};

// item -> piece.  Items enumerate (column, sign, part) with part fastest:
//   item = ((c*2 + sign) << logP) + part
__device__ __forceinline__ Piece decode_item(int item, int nitems, int logP, const int *ls_pos,
                                             const int *ls_neg, const int4 *rip4, const int4 *rin4)
{
    Piece p;
    p.idx = rip4;
    p.a4 = p.b4 = 0;
    if (item < nitems)
    {
        const int part = item & ((1 << logP) - 1);
        const int list = item >> logP;
        const int c = list >> 1;
        const int *ls = (list & 1) ? ls_neg : ls_pos;
        const int lo = ls[c];
        const int len = ls[c + 1] - lo;       // int4s in the list
        const int q = len >> logP, r = len & ((1 << logP) - 1);
        p.idx = (list & 1) ? rin4 : rip4;
        p.a4 = lo + q * part + ((r * part) >> logP);
        p.b4 = lo + q * (part + 1) + ((r * (part + 1)) >> logP);
    }
    return p;
}

// lane `sub` (0..7) of the team fetches unit number u*8+sub of the piece (128 B per team-load);
// `fill` is a unit of sentinels (K in every int32 slot, or K|K<<16 for 16-bit row ids)
__device__ __forceinline__ void issue_piece(const Piece &p, int sub, int off, int fill, int4 (&v)[kU])
{
#pragma unroll
    for (int u = 0; u < kU; ++u)
    {
        const int q = p.a4 + off + u * 8 + sub;
        if (q < p.b4)
            v[u] = ldg_stream(reinterpret_cast<const int *>(p.idx + q));
        else
            v[u] = make_int4(fill, fill, fill, fill);
    }
}

// Sum X over the piece; v holds its first kU*8 int4s.  Result valid in the team's lane 0.
template <int MT, bool I16>
__device__ __forceinline__ Acc<MT> consume_piece(const Piece &p, int sub, int K, int4 (&v)[kU],
                                                 uint32_t xs_base)
{
    Acc<MT> acc;
    acc.zero();
    for (int off = 0;;)
    {
#pragma unroll
        for (int u = 0; u < kU; ++u)
        {
            if constexpr (I16)
            {
                // eight 16-bit row ids per unit
                const uint32_t w[4] = {(uint32_t)v[u].x, (uint32_t)v[u].y, (uint32_t)v[u].z, (uint32_t)v[u].w};
                Acc<MT> x[8];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                {
                    x[2 * j] = gather<MT>(xs_base, (int)(w[j] & 0xFFFFu));
                    x[2 * j + 1] = gather<MT>(xs_base, (int)(w[j] >> 16));
                }
#pragma unroll
                for (int m = 0; m < MT; ++m)
                    acc.v[m] += ((x[0].v[m] + x[1].v[m]) + (x[2].v[m] + x[3].v[m])) +
                                ((x[4].v[m] + x[5].v[m]) + (x[6].v[m] + x[7].v[m]));
            }
            else
            {
                const Acc<MT> x0 = gather<MT>(xs_base, v[u].x), x1 = gather<MT>(xs_base, v[u].y),
                              x2 = gather<MT>(xs_base, v[u].z), x3 = gather<MT>(xs_base, v[u].w);
#pragma unroll
                for (int m = 0; m < MT; ++m)
                    acc.v[m] += (x0.v[m] + x1.v[m]) + (x2.v[m] + x3.v[m]);
            }
        }
        // pieces longer than one batch (very uneven column lengths): keep going, warp-uniformly
        off += 8 * kU;
        if (!__any_sync(0xffffffffu, p.a4 + off < p.b4))
            break;
        issue_piece(p, sub, off, I16 ? (K | (K << 16)) : K, v);
    }
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int o = 4; o > 0; o >>= 1)
            acc.v[m] += __shfl_xor_sync(0xffffffffu, acc.v[m], o);
    return acc;
}

template <int MT, bool I16>
__global__ void __launch_bounds__(kWarps * 32, 1)
gather_pieces_kernel(const int *__restrict__ lp, const int *__restrict__ ln,
                     const int4 *__restrict__ rip4, const int4 *__restrict__ rin4,
                     const float *__restrict__ X, int64_t ldx, const float *__restrict__ bias,
                     const float *__restrict__ alpha, float *__restrict__ Y, int64_t ldy, int M,
                     int K, int N, int logP, int cols_per_pass,
                     unsigned long long *__restrict__ trace, int tma_x)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *Xs = reinterpret_cast<float *>(smem_raw);          // (K+1)*MT, k-major; row K is zero
    float *tab = Xs + (size_t)(K + 4) * MT;                   // kTabFloats
    int *ls_pos = reinterpret_cast<int *>(tab + kTabFloats);  // kColCap+1
    int *ls_neg = ls_pos + kColCap + 1;                       // kColCap+1
    const uint32_t xs_base = (uint32_t)__cvta_generic_to_shared(Xs);
    // MT == 1 with a 16-byte aligned row of X: the row tile is one TMA bulk copy into shared memory
    // (cp.async.bulk, completion on an mbarrier) instead of eight loads and stores per thread
    const bool use_tma = (MT == 1) && tma_x;
    const uint32_t xbar = (uint32_t)__cvta_generic_to_shared(
        reinterpret_cast<unsigned long long *>((reinterpret_cast<uintptr_t>(ls_neg + kColCap + 1) + 7) & ~(uintptr_t)7));

    const int tid = threadIdx.x, lane = tid & 31;
    const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0); // provably warp-uniform
    const int m0 = blockIdx.y * MT;
    const int col_lo = (int)(((long long)N * blockIdx.x) / gridDim.x);
    const int col_hi = (int)(((long long)N * (blockIdx.x + 1)) / gridDim.x);
    if (col_lo >= col_hi)
        return;
    // developer trace (TSG_GATHER_TRACE=1): %globaltimer at the phase boundaries of each warp
#define TSG_TRACE(slot)                                                                         \
    do                                                                                          \
    {                                                                                           \
        if (trace != nullptr && lane == 0 && blockIdx.y == 0)                                   \
        {                                                                                       \
            unsigned long long t__;                                                             \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));                             \
            trace[((size_t)blockIdx.x * kWarps + wid) * 8 + (slot)] = t__;                      \
        }                                                                                       \
    } while (0)
    TSG_TRACE(0);

    bool first_pass = true;
    for (int g0 = col_lo; g0 < col_hi; g0 += cols_per_pass)
    {
        const int ncols = min(cols_per_pass, col_hi - g0);
        const int nitems = (ncols * 2) << logP;
        // ---- stage X (first pass) and this pass's list pointers -----------------------------
        // X: thread handles k = tid + i*512 for every row m of the tile (coalesced per row)
        int cr0[3] = {0, 0, 0}, cr1[3] = {0, 0, 0}; // 3 x 512 >= kColCap + 1
#pragma unroll
        for (int i = 0; i < 3; ++i)
        {
            const int j = tid + i * kWarps * 32;
            if (j <= ncols)
            {
                cr0[i] = ldg_s32_ordered(lp + g0 + j);
                cr1[i] = ldg_s32_ordered(ln + g0 + j);
            }
        }
        // Programmatic dependent launch: the list pointers above belong to the weight, which no kernel
        // in front of us writes; X, bias and Y may depend on it, so wait for it here (returns at once
        // for an ordinary serialised launch).
        if (first_pass)
        {
            if (use_tma && tid == 0)
            {
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(xbar));
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            asm volatile("griddepcontrol.wait;" ::: "memory");
            if (use_tma && tid == 0)
            {
                const uint32_t bytes = (uint32_t)K * 4u;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(xbar), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                                 "r"(xs_base), "l"(X + (int64_t)m0 * ldx), "r"(bytes), "r"(xbar)
                             : "memory");
            }
        }
        constexpr int XR = 8 / MT; // k positions per thread held in registers (K <= XR*512 fast)
        float xr[XR][MT];
#pragma unroll
        for (int i = 0; i < XR; ++i)
        {
            const int k = tid + i * kWarps * 32;
#pragma unroll
            for (int m = 0; m < MT; ++m)
                xr[i][m] = (first_pass && !use_tma && k < K && m0 + m < M)
                               ? ldg_f32_ordered(X + (int64_t)(m0 + m) * ldx + k)
                               : 0.0f;
        }
        // epilogue operands of this thread's first output, requested now so that their latency
        // is hidden behind the streaming phase
        float bias0 = 0.0f, alpha0 = 0.0f;
        const int c_first = tid % ncols, m_first = tid / ncols; // off the critical path
        if (tid < ncols * MT)
        {
            bias0 = bias[g0 + c_first];
            if (alpha != nullptr)
                alpha0 = alpha[g0 + c_first];
        }
        if (!first_pass)
            __syncthreads(); // previous pass's combine is done with tab / ls
#pragma unroll
        for (int i = 0; i < 3; ++i)
        {
            const int j = tid + i * kWarps * 32;
            if (j <= ncols)
            {
                ls_pos[j] = cr0[i];
                ls_neg[j] = cr1[i];
            }
        }
        if (first_pass && !use_tma)
        {
#pragma unroll
            for (int i = 0; i < XR; ++i)
            {
                const int k = tid + i * kWarps * 32;
                if (k < K)
                {
#pragma unroll
                    for (int m = 0; m < MT; ++m)
                        Xs[k * MT + m] = xr[i][m];
                }
            }
            for (int k = tid + XR * kWarps * 32; k < K; k += kWarps * 32) // large-K remainder
            {
#pragma unroll
                for (int m = 0; m < MT; ++m)
                    Xs[k * MT + m] = (m0 + m < M) ? X[(int64_t)(m0 + m) * ldx + k] : 0.0f;
            }
            if (tid < MT)
                Xs[K * MT + tid] = 0.0f; // the sentinel row
        }
        if (first_pass && use_tma && tid == 32)
            Xs[K] = 0.0f; // the sentinel row (outside the bulk copy's K*4 bytes)
        TSG_TRACE(1);
        __syncthreads();
        if (first_pass && use_tma)
        {
            // every thread observes the completion itself: that is what makes the bulk copy's
            // writes visible to it
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\t"
                             "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                             "selp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done)
                             : "r"(xbar)
                             : "memory");
        }
        TSG_TRACE(2);

        // ---- stream the pieces: two register buffers, next piece always in flight ------------
        // team t (8 lanes) of warp w takes items (r*kWarps + w)*4 + t, r = 0, 1, ...
        const int sub = lane & 7;
        const int stride = kWarps * 4;
        int4 va[kU], vb[kU];
        int item = wid * 4 + (lane >> 3);
        const int rounds = (nitems + stride - 1) / stride; // warp-uniform trip count
        const int fill = I16 ? (K | (K << 16)) : K;
        Piece pa = decode_item(item, nitems, logP, ls_pos, ls_neg, rip4, rin4);
        issue_piece(pa, sub, 0, fill, va);
        for (int r = 0; r < rounds; r += 2)
        {
            const Piece pb = decode_item(item + stride, nitems, logP, ls_pos, ls_neg, rip4, rin4);
            issue_piece(pb, sub, 0, fill, vb);
            {
                const Acc<MT> s = consume_piece<MT, I16>(pa, sub, K, va, xs_base);
                if (sub == 0 && item < nitems)
                {
#pragma unroll
                    for (int m = 0; m < MT; ++m)
                        tab[item * MT + m] = s.v[m];
                }
            }
            item += stride;
            if (r + 1 >= rounds)
                break;
            pa = decode_item(item + stride, nitems, logP, ls_pos, ls_neg, rip4, rin4);
            issue_piece(pa, sub, 0, fill, va);
            {
                const Acc<MT> s = consume_piece<MT, I16>(pb, sub, K, vb, xs_base);
                if (sub == 0 && item < nitems)
                {
#pragma unroll
                    for (int m = 0; m < MT; ++m)
                        tab[item * MT + m] = s.v[m];
                }
            }
            item += stride;
        }
        TSG_TRACE(3);
        if (g0 + cols_per_pass >= col_hi) // last pass streamed: the next kernel may begin launching
            asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        __syncthreads();
        TSG_TRACE(4);

        // ---- combine: thread -> (m, c), c fastest (coalesced Y stores); pieces in piece order --
        const int P = 1 << logP;
        for (int i = tid; i < ncols * MT; i += kWarps * 32)
        {
            const bool first = (i == tid);
            const int c = first ? c_first : i % ncols, m = first ? m_first : i / ncols;
            const float *tp = tab + (size_t)(((c * 2) << logP) * MT) + m;
            const float *tn = tab + (size_t)(((c * 2 + 1) << logP) * MT) + m;
            float sp = 0.0f, sn = 0.0f;
            for (int j = 0; j < P; ++j)
            {
                sp += tp[j * MT];
                sn += tn[j * MT];
            }
            if (m0 + m < M)
            {
                const int n = g0 + c;
                float y = (sp - sn) + (first ? bias0 : bias[n]);
                if (alpha != nullptr)
                    y = (y > 0.0f) ? y : (first ? alpha0 : alpha[n]) * y;
                Y[(int64_t)(m0 + m) * ldy + n] = y;
            }
        }
        first_pass = false;
    }
    TSG_TRACE(5);
#undef TSG_TRACE
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

template <int MT>
size_t gather_smem_bytes(int K)
{
    return (size_t)(K + 4) * MT * 4 + (size_t)kTabFloats * 4 + (size_t)2 * (kColCap + 1) * 4 + 16;
}

} // namespace

// TSG_GATHER_TRACE=1 in the environment turns on per-warp phase timestamps (developer tool).
static unsigned long long *g_trace_buf = nullptr;
static unsigned long long *gather_trace_buffer()
{
    static int enabled = -1;
    if (enabled < 0)
    {
        const char *e = getenv("TSG_GATHER_TRACE");
        enabled = (e && e[0] == '1') ? 1 : 0;
        if (enabled && cudaMalloc(&g_trace_buf, 148 * 16 * 8 * sizeof(unsigned long long)) != cudaSuccess)
            enabled = 0;
        if (enabled)
            cudaMemset(g_trace_buf, 0, 148 * 16 * 8 * sizeof(unsigned long long));
    }
    return enabled ? g_trace_buf : nullptr;
}

extern "C" int tsg_debug_gather_trace(unsigned long long *out, int max_entries)
{
    if (!g_trace_buf || !out)
        return 0;
    const int n = max_entries < 148 * 16 * 8 ? max_entries : 148 * 16 * 8;
    cudaDeviceSynchronize();
    cudaMemcpy(out, g_trace_buf, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    return n;
}

template <int MT, bool I16>
static int launch_pieces(tsg_matrix *m, const float *X, int64_t ldx, const float *b,
                         const float *alpha, float *Y, int64_t ldy, int M, cudaStream_t st)
{
    const size_t smem = gather_smem_bytes<MT>(m->K);
    // largest opt-in granted so far per device (the attribute is per device and function); atomic: two host
    // threads may launch the same kernel — a repeated, equal cudaFuncSetAttribute is harmless, a torn size is not
    static std::atomic<size_t> configured[64];
    std::atomic<size_t> &have = configured[m->device & 63];
    if (have.load(std::memory_order_acquire) < smem)
    {
        TSG_CUDA(cudaFuncSetAttribute(gather_pieces_kernel<MT, I16>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        size_t seen = have.load(std::memory_order_relaxed);
        while (seen < smem && !have.compare_exchange_weak(seen, smem, std::memory_order_release))
            ;
    }
    // pieces per list: power of two such that a piece is <= ~176 indices on average (one batch
    // of an 8-lane team covers 192)
    const long long lists = 2ll * m->N;
    const long long avg4 = lists ? (m->n4pos + m->n4neg + lists - 1) / lists : 0; // int4s per list
    int logP = 0;
    while (logP < 6 && (avg4 >> logP) > 44) // one batch of an 8-lane team covers 48 int4s
        ++logP;
    if (const char *e = getenv("TSG_GATHER_LOGP")) // developer override for tuning
        logP = atoi(e) < 0 ? 0 : (atoi(e) > 6 ? 6 : atoi(e));
    int cols_per_pass = kTabFloats / ((2 << logP) * MT);
    if (cols_per_pass > kColCap)
        cols_per_pass = kColCap;
    const int ctas = m->N < m->sm_count ? m->N : m->sm_count;
    dim3 grid(ctas, (M + MT - 1) / MT);
    // one-row tiles whose rows of X are 16-byte aligned and a whole number of 16-byte units are
    // staged by one TMA bulk copy per CTA (TSG_GATHER_NO_TMA=1: developer override)
    static const bool no_tma = getenv("TSG_GATHER_NO_TMA") != nullptr;
    const int tma_x = (MT == 1 && !no_tma && (m->K & 3) == 0 && (ldx & 3) == 0 &&
                       (reinterpret_cast<uintptr_t>(X) & 15) == 0)
                          ? 1
                          : 0;
    // programmatic dependent launch (see griddepcontrol in the kernel): back-to-back calls overlap the
    // next launch with this kernel's combine phase
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kWarps * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    TSG_CUDA(cudaLaunchKernelEx(&cfg, gather_pieces_kernel<MT, I16>, (const int *)m->lp, (const int *)m->ln,
                                (const int4 *)m->rip4, (const int4 *)m->rin4, X, ldx, b, alpha, Y, ldy, M, m->K, m->N,
                                logP, cols_per_pass, gather_trace_buffer(), tma_x));
    TSG_LAUNCHED();
    return TSG_OK;
}

int tsg_launch_gather(tsg_matrix *m, const float *X, int64_t ldx, const float *b,
                      const float *alpha, float *Y, int64_t ldy, int M, cudaStream_t st)
{
    if (M <= 0 || m->N == 0)
        return TSG_OK;
    TSG_CHECK(gather_smem_bytes<1>(m->K) <= m->smem_optin, TSG_ERR_UNSUPPORTED,
              "gather kernel: K=%d does not fit shared memory (%zu B needed, %zu B available)",
              m->K, gather_smem_bytes<1>(m->K), m->smem_optin);
    if (!m->rip4)
    {
        // first gather call on this handle: build the 16-byte aligned, sentinel-padded copy of the
        // index lists (allocates and synchronises the stream once; not possible inside a capture)
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        TSG_CUDA(cudaStreamIsCapturing(st, &cap));
        TSG_CHECK(cap == cudaStreamCaptureStatusNone, TSG_ERR_UNSUPPORTED,
                  "gather: the first call on a handle builds its padded index lists and cannot be captured; run one "
                  "gather call outside the capture first");
        TSG_TRY(tsg_build_padded_lists(m, st));
    }
    // widest row tile whose X staging fits in shared memory; 16-bit row ids when the handle has them
    if (M >= 4 && gather_smem_bytes<4>(m->K) <= m->smem_optin)
        return m->idx16 ? launch_pieces<4, true>(m, X, ldx, b, alpha, Y, ldy, M, st)
                        : launch_pieces<4, false>(m, X, ldx, b, alpha, Y, ldy, M, st);
    if (M >= 2 && gather_smem_bytes<2>(m->K) <= m->smem_optin)
        return m->idx16 ? launch_pieces<2, true>(m, X, ldx, b, alpha, Y, ldy, M, st)
                        : launch_pieces<2, false>(m, X, ldx, b, alpha, Y, ldy, M, st);
    return m->idx16 ? launch_pieces<1, true>(m, X, ldx, b, alpha, Y, ldy, M, st)
                    : launch_pieces<1, false>(m, X, ldx, b, alpha, Y, ldy, M, st);
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
