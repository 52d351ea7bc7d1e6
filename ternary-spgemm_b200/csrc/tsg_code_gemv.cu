// tsg_code_gemv.cu — decode-shaped M (1 or 2 rows of X): Y = X·W + b straight from the 2-bit
// code stream on the CUDA cores.
//
// Same contract as BaseTCSC / BaseTCSC_PreLU (reference cpp_impl/comp.h:25-69,
// cpp_impl/comp_prelu.h:12-70).  For one or two rows the tensor-core path is bounded by fixed
// costs (TMEM allocation, cluster barrier for split-K, ≈64 clk per 128×16 slice of W fed to the
// tensor core whatever the row count) and the TCSC gather by its 4-byte index per non-zero;
// the code stream is K·N/4 bytes — 5.3x less HBM traffic than the index stream at s = 3 — and
// on the FMA pipe one matrix element costs three instructions however sparse W is:
//
//      v   = (word << sh) & 0xC0000000      // sign at bit 31, "non-zero" at bit 30:
//                                           // the fp32 patterns of 0, +2.0, -2.0
//      acc = fmaf(v, x[k], acc)             // acc ± 2·x, exactly one rounding like the add
//
// (the factor 2 leaves exactly in the epilogue).  The word layout is the dense kernel's
// (tsg_build.cu pack_code_word), so both kernels share one stream.
//
//   grid   one CTA per 32 columns of W; 16 warps split K into 16 contiguous ranges
//   lane   one column: its 16-byte code (64 k) per k-block is part of a 512-byte contiguous
//          warp load; every load of the warp's range is issued before anything is consumed
//   X      staged once per CTA in shared memory; all lanes of a warp read the same k, so the
//          128-bit LDS is a broadcast (one wavefront for four operands)
//   PDL    launched with programmatic stream serialisation; griddepcontrol.launch_dependents after
//          the compute loop, griddepcontrol.wait before X / bias / Y are touched
//   sum    four independent accumulators per row, combined in a fixed order; the 16 partial sums
//          of a column meet in shared memory and are added warp 0 .. 15 (deterministic)
#include "tsg_internal.cuh"

#include <stdlib.h>

namespace
{

constexpr int kWarps = 16;
constexpr int kBatch = 4; // k-blocks per batch: two batches of 4 x uint4 per lane in registers

__device__ __forceinline__ uint4 ldg_v4_ordered(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// two fp32 FMAs in one instruction (sm_100 FFMA2): d.lo = a.lo*b.lo + d.lo, d.hi likewise
__device__ __forceinline__ void fma2(float2 &d, uint32_t a_lo, uint32_t a_hi, float b_lo, float b_hi)
{
    unsigned long long dd, aa, bb;
    asm("mov.b64 %0, {%1, %2};" : "=l"(dd) : "f"(d.x), "f"(d.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "r"(a_lo), "r"(a_hi));
    asm("mov.b64 %0, {%1, %2};" : "=l"(bb) : "f"(b_lo), "f"(b_hi));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(dd));
}

// 16 elements of one code word against x[0..15].
//   MR = 1: X is [Kp]; one broadcast LDS.128 brings four k; the packed FMA pairs two k.
//   MR = 2: X is interleaved [Kp][2]; one LDS.128 brings two k of both rows; the packed FMA pairs
//           the two ROWS (same multiplier in both halves), so a second row costs one more LDS per
//           two k instead of doubling the FMA work.
template <int MR>
__device__ __forceinline__ void word_fma(uint32_t w, const float *xs, float2 (&acc)[MR][2])
{
    constexpr uint32_t kMask = 0xC0000000u;
#pragma unroll
    for (int g = 0; g < 4; ++g) // elements 4g .. 4g+3  = pairs 2g, 2g+1
    {
        // element e = 2p + h: flag bits at 16h + 14 - 2p (tsg_build.cu)
        const uint32_t v0 = (w << (16 + 4 * g)) & kMask; // p = 2g,   h = 0
        const uint32_t v1 = (w << (4 * g)) & kMask;      // p = 2g,   h = 1
        const uint32_t v2 = (w << (18 + 4 * g)) & kMask; // p = 2g+1, h = 0
        const uint32_t v3 = (w << (2 + 4 * g)) & kMask;  // p = 2g+1, h = 1
        if constexpr (MR == 1)
        {
            const float4 x = *reinterpret_cast<const float4 *>(xs + 4 * g);
            fma2(acc[0][0], v0, v1, x.x, x.y);
            fma2(acc[0][1], v2, v3, x.z, x.w);
        }
        else
        {
            const float4 xa = *reinterpret_cast<const float4 *>(xs + 8 * g);     // k = 4g, 4g+1 (rows 0,1 each)
            const float4 xb = *reinterpret_cast<const float4 *>(xs + 8 * g + 4); // k = 4g+2, 4g+3
            fma2(acc[0][0], v0, v0, xa.x, xa.y);
            fma2(acc[0][1], v1, v1, xa.z, xa.w);
            fma2(acc[1][0], v2, v2, xb.x, xb.y);
            fma2(acc[1][1], v3, v3, xb.z, xb.w);
        }
    }
}

// One batch of this warp's K range: kBatch k-blocks of one column block, 16 bytes per lane each.
struct Unit
{
    int j; // which of this CTA's column blocks (blockIdx.x + j * gridDim.x)
    int b; // which batch of the warp's K range
};

template <int MR>
__global__ void __launch_bounds__(kWarps * 32, 2)
code_gemv_kernel(const uint4 *__restrict__ codes, int nkb, const float *__restrict__ X, int64_t ldx,
                 const float *__restrict__ bias, const float *__restrict__ alpha,
                 float *__restrict__ Y, int64_t ldy, int M, int K, int N, int pdl,
                 const int32_t *__restrict__ csp, const int32_t *__restrict__ csn,
                 const int32_t *__restrict__ rip, const int32_t *__restrict__ rin)
{
    extern __shared__ __align__(16) float smem[];
    const int Kp = nkb * 64;
    float *xs = smem;                 // MR = 1: [Kp];  MR = 2: [Kp][2] (rows interleaved)
    float *part = smem + MR * Kp;     // [2][kWarps][MR][32], double-buffered over column blocks
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int m0 = blockIdx.y * MR;
    // this warp's k-blocks, in batches of kBatch
    const int kb_lo = (int)(((long long)nkb * warp) / kWarps), kb_hi = (int)(((long long)nkb * (warp + 1)) / kWarps);
    const int nb = max(1, (((nkb + kWarps - 1) / kWarps) + kBatch - 1) / kBatch); // same for every warp
    // this CTA's column blocks: blockIdx.x, blockIdx.x + gridDim.x, ...
    const int ncb = (N + 31) >> 5;
    const int mine = (ncb - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    // Code stream first: everything else hides behind its HBM latency.  Columns beyond N read the
    // zero codes of the tile padding (tiles are always 128 columns wide).
    auto load_unit = [&](uint4(&c)[kBatch], const Unit &u) {
        const int n = (((int)blockIdx.x + u.j * (int)gridDim.x) << 5) + lane;
        const int k0 = kb_lo + u.b * kBatch;
        const uint4 *src = codes + ((size_t)(n >> 7) * nkb + k0) * 128 + (n & 127);
#pragma unroll
        for (int i = 0; i < kBatch; ++i)
            c[i] = (u.j < mine && k0 + i < kb_hi) ? ldg_v4_ordered(src + (size_t)i * 128) : make_uint4(0, 0, 0, 0);
    };
    auto advance = [&](Unit &u) {
        if (++u.b == nb)
            u.b = 0, ++u.j;
    };
    uint4 ca[kBatch], cb[kBatch];
    Unit cur = {0, 0}, nxt = {0, 0};
    load_unit(ca, cur);
    advance(nxt);
    // pdl != 0: launched with programmatic stream serialisation.  The code loads above touch only the
    // weight stream, which no kernel in front of us writes; wait for that kernel to complete (and
    // flush) before reading X / bias or writing Y.  (After a kernel that never triggers, or a copy,
    // the launch is an ordinary serialised one and the wait returns at once.)
    if (pdl == 1)
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (pdl)
        asm volatile("griddepcontrol.wait;" ::: "memory");
    // stage X once per CTA (zero beyond K and beyond M): 128-bit loads when the rows allow it.
    // `huge`: a value the multiply formulation cannot take (non-finite or >= 2^100, tsg_internal.cuh)
    uint32_t huge = 0;
    auto look = [&](float x) { huge = max(huge, __float_as_uint(x) & 0x7FFFFFFFu); }; // inf / NaN sort above finite
    {
        const float *x0 = X + (int64_t)m0 * ldx;
        const bool two = MR == 2 && m0 + 1 < M;
        const float *x1 = x0 + (two ? ldx : 0);
        const bool vec = ((reinterpret_cast<uintptr_t>(X) | (uintptr_t)(ldx * 4)) & 15) == 0;
        const int K4 = vec ? (K & ~3) : 0;
        for (int k = tid * 4; k < K4; k += kWarps * 32 * 4)
        {
            const float4 a = *reinterpret_cast<const float4 *>(x0 + k);
            look(a.x), look(a.y), look(a.z), look(a.w);
            if constexpr (MR == 1)
                *reinterpret_cast<float4 *>(xs + k) = a;
            else
            {
                const float4 b = two ? *reinterpret_cast<const float4 *>(x1 + k) : make_float4(0.f, 0.f, 0.f, 0.f);
                look(b.x), look(b.y), look(b.z), look(b.w);
                *reinterpret_cast<float4 *>(xs + 2 * k) = make_float4(a.x, b.x, a.y, b.y);
                *reinterpret_cast<float4 *>(xs + 2 * k + 4) = make_float4(a.z, b.z, a.w, b.w);
            }
        }
        for (int k = K4 + tid; k < Kp; k += kWarps * 32) // unaligned rows, the K tail and the zero padding
        {
            xs[k * MR] = (k < K) ? x0[k] : 0.0f;
            look(xs[k * MR]);
            if constexpr (MR == 2)
            {
                xs[k * MR + 1] = (k < K && two) ? x1[k] : 0.0f;
                look(xs[k * MR + 1]);
            }
        }
    }
    if (__syncthreads_or((int)(huge >= TSG_X_HUGE_BITS)))
    {
        // X holds inf / NaN / |x| >= 2^100: 0·x and 2·x are not what the reference's sparse sum
        // computes (comp.h:44-61 never touches x where W is 0).  This CTA's columns in the
        // reference's own order, from the staged rows — slow, exact, only ever for such input.
        for (int j = 0; j < mine; ++j)
            for (int t = tid; t < 32 * MR; t += kWarps * 32)
            {
                const int n = (((int)blockIdx.x + j * (int)gridDim.x) << 5) + (t & 31), mr = t >> 5;
                if (n < N && m0 + mr < M)
                {
                    float y = tsg_ref_order_sum(xs + mr, MR, csp, csn, rip, rin, n, bias[n]);
                    if (alpha != nullptr)
                        y = (y > 0.0f) ? y : alpha[n] * y;
                    Y[(int64_t)(m0 + mr) * ldy + n] = y;
                }
            }
        return;
    }

    float2 acc[MR][2];
#pragma unroll
    for (int m = 0; m < MR; ++m)
        acc[m][0] = acc[m][1] = make_float2(0.0f, 0.0f);
    float bn = 0.0f, an = 0.0f; // epilogue operands, fetched by the warp that will reduce the block

    // one batch against X; at the end of a column block: partial sums to shared memory, one
    // barrier, warp (j mod 16) adds them in warp order (deterministic) and stores Y
    auto consume = [&](const uint4(&c)[kBatch], const Unit &u) {
        const int n = (((int)blockIdx.x + u.j * (int)gridDim.x) << 5) + lane;
        const bool reducer = warp == (u.j & (kWarps - 1));
        if (u.b == 0 && reducer && n < N)
        {
            bn = bias[n];
            if (alpha != nullptr)
                an = alpha[n];
        }
        const int k0 = kb_lo + u.b * kBatch;
#pragma unroll
        for (int i = 0; i < kBatch; ++i)
        {
            if (k0 + i < kb_hi) // warp-uniform
            {
                const float *xk = xs + (k0 + i) * 64 * MR;
                word_fma<MR>(c[i].x, xk, acc);
                word_fma<MR>(c[i].y, xk + 16 * MR, acc);
                word_fma<MR>(c[i].z, xk + 32 * MR, acc);
                word_fma<MR>(c[i].w, xk + 48 * MR, acc);
            }
        }
        if (u.b != nb - 1)
            return;
        if (pdl == 2 && u.j == mine - 1) // last block computed: the next kernel may begin launching
            asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        float *pb = part + (u.j & 1) * (kWarps * MR * 32);
        if constexpr (MR == 1)
            pb[warp * 32 + lane] = (acc[0][0].x + acc[0][0].y) + (acc[0][1].x + acc[0][1].y);
        else
        {
            pb[(warp * 2 + 0) * 32 + lane] = (acc[0][0].x + acc[0][1].x) + (acc[1][0].x + acc[1][1].x);
            pb[(warp * 2 + 1) * 32 + lane] = (acc[0][0].y + acc[0][1].y) + (acc[1][0].y + acc[1][1].y);
        }
#pragma unroll
        for (int m = 0; m < MR; ++m)
            acc[m][0] = acc[m][1] = make_float2(0.0f, 0.0f);
        // the buffer written two blocks ago is free: its reducer passed the previous barrier only
        // after it had finished reading
        __syncthreads();
        if (reducer && n < N)
        {
#pragma unroll
            for (int m = 0; m < MR; ++m)
            {
                float s = 0.0f;
#pragma unroll
                for (int w = 0; w < kWarps; ++w) // fixed order: deterministic
                    s += pb[(w * MR + m) * 32 + lane];
                float y = 0.5f * s + bn; // the codes expand to 2·W
                if (alpha != nullptr)
                    y = (y > 0.0f) ? y : an * y;
                if (m0 + m < M)
                    Y[(int64_t)(m0 + m) * ldy + n] = y;
            }
        }
    };

    // two register buffers: the next batch is always in flight while the current one is consumed
    for (;;)
    {
        load_unit(cb, nxt);
        consume(ca, cur);
        cur = nxt;
        advance(nxt);
        if (cur.j >= mine)
            break;
        load_unit(ca, nxt);
        consume(cb, cur);
        cur = nxt;
        advance(nxt);
        if (cur.j >= mine)
            break;
    }
}

template <int MR>
int launch(tsg_matrix *m, const float *X, int64_t ldx, const float *b, const float *alpha, float *Y,
           int64_t ldy, int M, cudaStream_t st)
{
    const int nkb = m->code_kblocks;
    const size_t smem = ((size_t)MR * nkb * 64 + (size_t)2 * kWarps * MR * 32) * sizeof(float);
    TSG_CHECK(smem <= m->smem_optin, TSG_ERR_UNSUPPORTED,
              "code_gemv: K=%d does not fit shared memory (%zu B needed)", m->K, smem);
    // largest opt-in granted so far per device (the attribute is per device and function); atomic: two host
    // threads may launch the same kernel — a repeated, equal cudaFuncSetAttribute is harmless, a torn size is not
    static std::atomic<size_t> configured[64];
    std::atomic<size_t> &have = configured[m->device & 63];
    if (have.load(std::memory_order_acquire) < smem)
    {
        TSG_CUDA(cudaFuncSetAttribute(code_gemv_kernel<MR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        size_t seen = have.load(std::memory_order_relaxed);
        while (seen < smem && !have.compare_exchange_weak(seen, smem, std::memory_order_release))
            ;
    }
    // one CTA per 32 columns while that is at most two CTAs per SM (what the register file holds);
    // beyond that a CTA walks several column blocks with its code loads software-pipelined
    const int ncb = (m->N + 31) / 32, resident = 2 * (m->sm_count > 0 ? m->sm_count : 148);
    dim3 grid(ncb < resident ? ncb : resident, (M + MR - 1) / MR);
    TSG_CHECK(grid.y <= 65535, TSG_ERR_UNSUPPORTED, "code_gemv: M too large");
    // Programmatic dependent launch, trigger AFTER the compute loop: the next kernel's launch latency
    // overlaps this kernel's reduction and stores (measured at c2, back-to-back calls in a graph:
    // 4.98 µs without, 4.28 µs with the late trigger, 6.03 µs with a trigger at kernel entry, which
    // lets the next grid compete for issue slots during the FMA-bound loop).  TSG_GEMV_PDL=0 turns it off.
    static const int pdl = getenv("TSG_GEMV_PDL") ? atoi(getenv("TSG_GEMV_PDL")) : 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kWarps * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    TSG_CUDA(cudaLaunchKernelEx(&cfg, code_gemv_kernel<MR>, (const uint4 *)m->codes, nkb, X, ldx, b, alpha, Y, ldy, M,
                                m->K, m->N, pdl, (const int32_t *)m->csp, (const int32_t *)m->csn,
                                (const int32_t *)m->rip, (const int32_t *)m->rin));
    TSG_LAUNCHED();
    return TSG_OK;
}

} // namespace

int tsg_launch_code_gemv(tsg_matrix *m, const float *X, int64_t ldx, const float *b,
                         const float *alpha, float *Y, int64_t ldy, int M, cudaStream_t st)
{
    if (M <= 0 || m->N == 0)
        return TSG_OK;
    TSG_CHECK(m->codes != nullptr && m->code_kblocks > 0, TSG_ERR_UNSUPPORTED, "code_gemv: tile codes missing");
    if (M >= 2)
        return launch<2>(m, X, ldx, b, alpha, Y, ldy, M, st);
    return launch<1>(m, X, ldx, b, alpha, Y, ldy, M, st);
}
