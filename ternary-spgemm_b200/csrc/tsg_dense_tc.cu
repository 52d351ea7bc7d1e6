// tsg_dense_tc.cu — dense-expand tensor-core path (north-star subsystem 3) for larger M.
//
// Same contract as BaseTCSC / BaseTCSC_PreLU (reference cpp_impl/comp.h:25-69,
// cpp_impl/comp_prelu.h:12-70) — Y = X·W + b, optional PReLU — computed as a dense GEMM on the
// 5th-generation tensor cores:
//
//      D[n, m] = Σ_k  Wt[n, k] · X[m, k]            (UMMA: D = A·Bᵀ, both operands K-major)
//
//   A = Wᵀ tile, 128 columns of W × 64 k, bf16, EXPANDED ON THE FLY in shared memory from the
//       tile-packed 2-bit codes (tsg_matrix::codes; +1 -> 0x3F80, -1 -> 0xBF80, 0 -> 0) straight
//       into the 128-byte-swizzled K-major layout tcgen05.mma reads.  W is never materialised
//       as bf16 in HBM: HBM sees K·N/4 bytes, fetched with coalesced 128-bit loads.
//   B = X tile, NT rows × 64 k, bf16, loaded by TMA (cp.async.bulk.tensor, SWIZZLE_128B).
//       fp32 X is split EXACTLY into three bf16 terms x = x1 + x2 + x3 (8+8+8 mantissa bits) by
//       split_x_kernel, and the three products accumulate into the same fp32 accumulator, so
//       every product W·x_i is exact and only the fp32 accumulation rounds — like the
//       reference's fp32 adds.  Terms that are identically zero for the whole X (the
//       reference's integer-valued inputs need only two) are skipped.
//   D = 128 × NT fp32 accumulator in TMEM (tcgen05.alloc), read back with tcgen05.ld.
//
// One CTA (640 threads) per (128-column tile of W, m-tile, K-split), warp-specialised:
//   warp 16     TMA producer for the X tiles (one elected lane)
//   warp 19     tcgen05.mma issuer (one elected lane); tcgen05.commit releases smem stages
//   warp 18     TMEM allocator / deallocator
//   warps 0-15  expanders: four groups of four warps take every fourth k-block; thread -> one W
//               column (one 128-byte smem row); afterwards the same warps run the epilogue
//               (TMEM -> registers -> bias/PReLU -> coalesced stores, lane = W column)
// mbarrier pipeline: full[s] (TMA bytes + 4 expander warps), empty[s] (tcgen05.commit),
// tmem_full (last commit).
// Split-K: the K-splits of one tile form a thread-block CLUSTER (<= 8 CTAs).  Non-leader CTAs
// park their accumulators in their own shared memory; after a cluster barrier the leader adds
// them through distributed shared memory in rank order (deterministic), applies bias / PReLU
// and writes Y — no partial sums in HBM, no second kernel.
#include "tsg_internal.cuh"

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace
{

constexpr int kTileN = 128;   // W columns per CTA  (UMMA M)
constexpr int kBlockK = 64;   // k per pipeline stage (128 bytes of bf16 per row)
constexpr int kThreads = 640;   // 4 role warps + 16 expander/epilogue warps
constexpr int kExpGroups = 4;
// Warp roles.  The SM's issue arbiter favours the highest warp id among eligible warps, so the
// two single-thread critical-path roles get the top ids and are never starved by the sixteen
// expander warps sharing their schedulers.
constexpr int kExpWarps = 16, kTmaWarp = 16, kAllocWarp = 18, kMmaWarp = 19;
constexpr int kMaxSplits = 3;

// ---- PTX wrappers ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a pipeline bug must surface as a trapped kernel (an error code through the C
// ABI), never as a hung GPU.  try_wait itself sleeps in hardware, so the bound is seconds.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 26))
            __trap();
}
__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
                     "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] · B[smem]ᵀ, 16-bit float in (format in the instruction descriptor), fp32 out
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128-byte swizzle: rows of 128 B, 8-row atoms of 1024 B (SBO), version 1 (sm_100).
// Field layout: cute/arch/mma_sm100_desc.hpp SmemDescriptor.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16: D = F32 (bits 4-5 = 1), A and B K-major, N>>3 at bit 17, M>>4 at bit 24; the 16-bit
// input format (0 = F16, 1 = BF16, bits 7-9 for A and 10-12 for B) is OR-ed in by the issuer
// (InstrDescriptor in the same header).
__host__ __device__ constexpr uint32_t make_idesc(int n)
{
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileN >> 4) << 24);
}

// One code word (16 k of one column, layout of pack_code_word in tsg_build.cu) -> eight packed
// pairs of 16-bit floats in {0, +2, -2}: (word << 2p) & 0xC000C000.  Two 16-byte chunks of the
// smem row.  0x4000 is 2.0 in bf16 and in fp16, so the A tile is the same for both X formats;
// the factor 2 is taken out again, exactly, in the epilogue.
__device__ __forceinline__ void expand_word(uint32_t w, uint32_t addr_lo, uint32_t addr_hi)
{
    constexpr uint32_t kMask = 0xC000C000u;
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr_lo), "r"(w & kMask), "r"((w << 2) & kMask),
                 "r"((w << 4) & kMask), "r"((w << 6) & kMask)
                 : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr_hi), "r"((w << 8) & kMask), "r"((w << 10) & kMask),
                 "r"((w << 12) & kMask), "r"((w << 14) & kMask)
                 : "memory");
}

// exactly one lane of a converged warp (lets ptxas issue the single-thread tcgen05/TMA
// instructions without a per-instruction divergence loop)
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .b32 rx;\n\t"
        ".reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, px;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem_f32(uint32_t local_addr, uint32_t rank)
{
    uint32_t remote;
    float v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
    return v;
}

struct DenseParams
{
    const uint4 *codes; // [tiles][nkb][128] tile-packed 2-bit codes
    int N, M, K;
    int nkb;         // k-blocks in total (Kp / 64)
    int ksplit;      // K-splits (= cluster size along z)
    int Mp;          // padded rows per split term in the pre-split X buffer (TMA path)
    const int *flags; // TMA path: bit0 term 2 non-zero, bit1 term 3 non-zero, bit2 X not exact in fp16
    const float *X;  // in-kernel conversion path: fp32 X
    int64_t ldx;
    const float *bias, *alpha;
    float *Y;        // M×N
    int64_t ldy;
    int stage_budget; // shared-memory bytes available for pipeline stages
};

constexpr int kABytes = kTileN * 128;  // one expanded A stage: 128 columns x 64 k x 2 B
constexpr int kBarBytes = 1024;        // barriers + TMEM slot live in front of the stages

// fp32 -> three bf16 terms, two elements at a time (exact: x == t1 + t2 + t3).
__device__ __forceinline__ void split3_pair(float x0, float x1, uint32_t &t1, uint32_t &t2, uint32_t &t3)
{
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(t1) : "f"(x1), "f"(x0));
    const float r0 = x0 - __uint_as_float(t1 << 16), r1 = x1 - __uint_as_float(t1 & 0xFFFF0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(t2) : "f"(r1), "f"(r0));
    const float q0 = r0 - __uint_as_float(t2 << 16), q1 = r1 - __uint_as_float(t2 & 0xFFFF0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(t3) : "f"(q1), "f"(q0));
}

// NT  : accumulator columns per split term (rows of X per m-tile), multiple of 16
// XK  : true  -> X is converted to its bf16 terms inside the kernel (small M: one launch, no
//                scratch; always three terms),
//       false -> X tiles come by TMA from the buffer split_x_kernel wrote (1-3 bf16 terms, or one
//                fp16 term when every x is exactly representable in fp16 — the reference's
//                integer-valued inputs are).
template <int NT, bool XK>
__global__ void __launch_bounds__(kThreads, 1)
dense_tc_kernel(const __grid_constant__ CUtensorMap xmap, const DenseParams p)
{
    constexpr int kTmemCols = NT * kMaxSplits <= 32 ? 32 : (NT * kMaxSplits <= 64 ? 64 : (NT * kMaxSplits <= 128 ? 128 : (NT * kMaxSplits <= 256 ? 256 : 512)));
    constexpr int kBBytes = NT * 128; // one split term of one stage
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char *smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_al);
    const uint32_t stage0 = smem_base + kBarBytes;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int n0 = blockIdx.x * kTileN;
    const int mtile = blockIdx.y;
    const int split = blockIdx.z;
    const int kb_lo = (int)(((long long)p.nkb * split) / p.ksplit);
    const int kb_hi = (int)(((long long)p.nkb * (split + 1)) / p.ksplit);
    const int iters = kb_hi - kb_lo;

    // how X arrives: number of split terms, their 16-bit format, first row in the split buffer
    int nterms = kMaxSplits, fmt = 1, row0 = 0;
    if constexpr (!XK)
    {
        const int fl = *p.flags;
        if (!(fl & 4))
            nterms = 1, fmt = 0, row0 = kMaxSplits * p.Mp; // one fp16 term
        else
            nterms = (fl & 2) ? 3 : ((fl & 1) ? 2 : 1);
    }
    const int stage_bytes = kABytes + nterms * kBBytes;
    int S = p.stage_budget / stage_bytes;
    S = S > 8 ? 8 : S;
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * 8, tmem_full = empty0 + 8 * 8;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 17);

    if (warp == kMmaWarp && lane == 0)
    {
        for (int s = 0; s < S; ++s)
        {
            mbar_init(full0 + 8 * s, XK ? 4 : 5); // 4 expander warps (+ the TMA producer's expect_tx arrive)
            mbar_init(empty0 + 8 * s, 1);         // one tcgen05.commit
        }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    else if (warp == kAllocWarp)
    {
        tmem_alloc(smem_u32(tmem_slot), kTmemCols);
    }
    else if (!XK && warp == kTmaWarp && lane == 0)
    {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
    }
    if constexpr (XK)
    {
        // rows of the X tiles at or beyond M are never written again: zero all B regions once
        const int per_stage = nterms * kBBytes / 16;
        for (int i = tid; i < S * per_stage; i += kThreads)
        {
            const int s = i / per_stage, o = i - s * per_stage;
            asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(stage0 + s * stage_bytes + kABytes + o * 16), "r"(0)
                         : "memory");
        }
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;

    if (!XK && warp == kTmaWarp)
    {
        // ===== TMA producer: X tiles of the split terms =====
        if (elect_one())
        {
            uint32_t eb = empty0, fb = full0, bdst = stage0 + kABytes, ph = 0;
            int kcoord = kb_lo * kBlockK, st = 0;
            const int row = row0 + mtile * NT;
            for (int it = 0; it < iters; ++it)
            {
                mbar_wait(eb, ph ^ 1);
                mbar_arrive_expect_tx(fb, (uint32_t)(nterms * kBBytes));
                tma_load_2d(bdst, &xmap, fb, kcoord, row);
                if (nterms > 1)
                    tma_load_2d(bdst + kBBytes, &xmap, fb, kcoord, p.Mp + row);
                if (nterms > 2)
                    tma_load_2d(bdst + 2 * kBBytes, &xmap, fb, kcoord, 2 * p.Mp + row);
                kcoord += kBlockK;
                eb += 8, fb += 8, bdst += stage_bytes;
                if (++st == S)
                    st = 0, eb = empty0, fb = full0, bdst = stage0 + kABytes, ph ^= 1;
            }
        }
    }
    else if (warp == kMmaWarp)
    {
        // ===== MMA issuer: the whole warp runs the loop (uniform registers), one lane issues =====
        // One MMA per 16-k step covers all split terms at once: their X tiles are adjacent in smem
        // (rows [t*NT, (t+1)*NT)), so B is simply nterms*NT rows tall and term t lands in
        // accumulator columns [t*NT, (t+1)*NT).  (N <= 256 per instruction: NT=128 with three
        // terms issues 256 + 128.)  The terms are added in the epilogue.
        const int nrows = nterms * NT;
        const uint32_t fbits = ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10);
        const uint32_t idesc_a = make_idesc(nrows > 256 ? 256 : nrows) | fbits;
        const uint32_t idesc_b = make_idesc(NT) | fbits; // only used when nrows == 384
        const uint64_t adesc0 = make_smem_desc(stage0);
        constexpr uint64_t kBOff = kABytes >> 4, kBStep = kBBytes >> 4;
        const uint64_t stage_step = (uint64_t)(stage_bytes >> 4);
        uint64_t adesc = adesc0;
        uint32_t fb = full0, eb = empty0, ph = 0;
        int st = 0;
        for (int it = 0; it < iters; ++it)
        {
            mbar_wait(fb, ph);
            tc_fence_after();
            if (elect_one())
            {
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k) // UMMA_K = 16 x 2 B = 32 B: +2 in the address field
                    umma_f16(tmem_d, adesc + 2 * k, adesc + kBOff + 2 * k, idesc_a, (it | k) != 0);
                if (nrows > 256)
                {
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k)
                        umma_f16(tmem_d + 256, adesc + 2 * k, adesc + kBOff + 2 * kBStep + 2 * k, idesc_b,
                                 (it | k) != 0);
                }
                umma_commit(eb); // frees the stage when these MMAs retire
            }
            __syncwarp();
            adesc += stage_step, fb += 8, eb += 8;
            if (++st == S)
                st = 0, adesc = adesc0, fb = full0, eb = empty0, ph ^= 1;
        }
        if (elect_one())
            umma_commit(tmem_full);
        __syncwarp();
    }
    // accumulators of this warp's 16-column chunks (chunks slice, slice+4, ... of the m-tile)
    constexpr int kChunks = NT / 16;
    constexpr int kMyChunks = (kChunks + kExpGroups - 1) / kExpGroups;
    uint32_t acc[kMyChunks][16];
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int slice = warp < kExpWarps ? warp >> 2 : 0;
    const int erow = q * 32 + lane;               // accumulator lane = W column inside the tile
    if (warp < kExpWarps)
    {
        // ===== expanders: tile-packed codes -> swizzled 16-bit A tile =====
        // A group may only run one barrier phase ahead of the MMA issuer (mbarrier parity is one
        // bit), which holds iff #groups <= #stages.
        const int groups = S < kExpGroups ? S : kExpGroups;
        const int grp = slice;                       // k-blocks with it % groups == grp
        const uint4 *src = p.codes + ((size_t)blockIdx.x * p.nkb + kb_lo) * 128 + erow;
        // smem byte offsets of this thread's eight 16-byte chunks inside a stage (128-byte swizzle)
        uint32_t off[8];
#pragma unroll
        for (int c = 0; c < 8; ++c)
            off[c] = (uint32_t)(erow * 128 + ((c ^ (erow & 7)) << 4));
        // The code stream is the kernel's HBM stream (2 KB per k-block and tile): each thread
        // pulls its 16 bytes into L2 kPrefetch k-blocks ahead (prefetch.global.L2) and into
        // registers two of its own iterations ahead, so the expansion never waits on DRAM.
        constexpr int kPrefetch = 32;
        uint4 nxt = make_uint4(0, 0, 0, 0), nxt2 = make_uint4(0, 0, 0, 0);
        // in-kernel X conversion: this thread's pairs of the next k-block of its group.
        // pair (m_local = q + 4j, k = 2*lane, 2*lane+1), j < NT/4... only rows < M are touched.
        constexpr int kPairs = XK ? NT / 4 : 1;
        float2 xv[kPairs];
        auto load_x = [&](int it) {
            if constexpr (XK)
            {
                const int k = (kb_lo + it) * kBlockK + 2 * lane;
#pragma unroll
                for (int j = 0; j < kPairs; ++j)
                {
                    const int m = mtile * NT + q + 4 * j;
                    xv[j] = make_float2(0.0f, 0.0f);
                    if (m < p.M)
                    {
                        const float *xp = p.X + (int64_t)m * p.ldx + k;
                        if (k < p.K)
                            xv[j].x = __ldg(xp);
                        if (k + 1 < p.K)
                            xv[j].y = __ldg(xp + 1);
                    }
                }
            }
        };
        if (grp < groups)
        {
            for (int it = grp; it < iters && it < kPrefetch; it += groups)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(src + (size_t)it * 128));
            if (grp < iters)
            {
                nxt = __ldg(src + (size_t)grp * 128);
                load_x(grp);
            }
            if (grp + groups < iters)
                nxt2 = __ldg(src + (size_t)(grp + groups) * 128);
        }
        int st = grp;            // grp < groups <= S
        uint32_t ph = 0;
        for (int it = (grp < groups ? grp : iters); it < iters; it += groups)
        {
            const uint4 cur = nxt;
            nxt = nxt2;
            if (it + kPrefetch < iters)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(src + (size_t)(it + kPrefetch) * 128));
            if (it + 2 * groups < iters) // codes two iterations ahead, in flight during this expansion
                nxt2 = __ldg(src + (size_t)(it + 2 * groups) * 128);
            uint32_t xt[kPairs][3];
            if constexpr (XK)
            {
#pragma unroll
                for (int j = 0; j < kPairs; ++j)
                    split3_pair(xv[j].x, xv[j].y, xt[j][0], xt[j][1], xt[j][2]);
                if (it + groups < iters)
                    load_x(it + groups);
            }
            mbar_wait(empty0 + 8 * st, ph ^ 1);
            const uint32_t sbase = stage0 + st * stage_bytes;
            expand_word(cur.x, sbase + off[0], sbase + off[1]);
            expand_word(cur.y, sbase + off[2], sbase + off[3]);
            expand_word(cur.z, sbase + off[4], sbase + off[5]);
            expand_word(cur.w, sbase + off[6], sbase + off[7]);
            if constexpr (XK)
            {
#pragma unroll
                for (int j = 0; j < kPairs; ++j)
                {
                    const int ml = q + 4 * j;
                    if (mtile * NT + ml < p.M)
                    {
                        const uint32_t a = sbase + kABytes + ml * 128 + (((lane >> 2) ^ (ml & 7)) << 4) + (lane & 3) * 4;
#pragma unroll
                        for (int t = 0; t < 3; ++t)
                            asm volatile("st.shared.b32 [%0], %1;" ::"r"(a + t * kBBytes), "r"(xt[j][t]) : "memory");
                    }
                }
            }
            fence_proxy_async(); // generic-proxy smem writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0)
                mbar_arrive(full0 + 8 * st);
            st += groups;
            if (st >= S)
                st -= S, ph ^= 1;
        }

        // ===== epilogue part 1: TMEM -> registers =====
        mbar_wait(tmem_full, 0);
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < kMyChunks; ++j)
        {
            const int ch = slice + j * kExpGroups;
            if (ch < kChunks)
            {
                tmem_ld16(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 16), acc[j]);
                for (int t = 1; t < nterms; ++t) // x1 + x2 + x3 terms, fixed order
                {
                    uint32_t more[16];
                    tmem_ld16(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * NT + ch * 16), more);
#pragma unroll
                    for (int c = 0; c < 16; ++c)
                        acc[j][c] = __float_as_uint(__uint_as_float(acc[j][c]) + __uint_as_float(more[c]));
                }
            }
        }
    }

    // ===== split-K reduction across the cluster (ranks = K-splits), then the output =====
    const uint32_t crank = (p.ksplit > 1) ? blockIdx.z : 0;
    float *park = reinterpret_cast<float *>(smem_al + kBarBytes); // [NT][128] column-major, reuses the stages
    if (p.ksplit > 1)
    {
        if (warp < kExpWarps && crank != 0)
        {
#pragma unroll
            for (int j = 0; j < kMyChunks; ++j)
            {
                const int ch = slice + j * kExpGroups;
                if (ch < kChunks)
                {
#pragma unroll
                    for (int c = 0; c < 16; ++c)
                        park[(ch * 16 + c) * 128 + erow] = __uint_as_float(acc[j][c]);
                }
            }
        }
        cluster_sync_all();
    }
    if (warp < kExpWarps && crank == 0)
    {
        const int en = n0 + erow;
        float bn = 0.0f, an = 0.0f;
        if (en < p.N)
        {
            bn = p.bias[en];
            if (p.alpha)
                an = p.alpha[en];
        }
#pragma unroll
        for (int j = 0; j < kMyChunks; ++j)
        {
            const int ch = slice + j * kExpGroups;
            if (ch < kChunks)
            {
#pragma unroll
                for (int c = 0; c < 16; ++c)
                {
                    const int m = mtile * NT + ch * 16 + c;
                    if (m >= p.M)
                        break; // rows are ascending in c: nothing further in this chunk
                    float y = __uint_as_float(acc[j][c]);
                    for (int r = 1; r < p.ksplit; ++r) // rank order: deterministic
                        y += ld_dsmem_f32(smem_u32(park + (ch * 16 + c) * 128 + erow), (uint32_t)r);
                    if (en < p.N)
                    {
                        y = 0.5f * y + bn; // the A tile holds 2·W (exact power-of-two scaling)
                        if (p.alpha)
                            y = (y > 0.0f) ? y : an * y;
                        p.Y[(int64_t)m * p.ldy + en] = y;
                    }
                }
            }
        }
    }
    if (p.ksplit > 1)
        cluster_sync_all(); // peers keep their smem alive until the leader has read it
    tc_fence_before();
    __syncthreads();
    if (warp == kAllocWarp)
        tmem_dealloc(tmem_d, kTmemCols);
}

// fp32 X -> three bf16 terms (exact: x == x1 + x2 + x3) and one fp16 copy, zero padded to
// [Mp][Kp] each: rows [0,Mp) [Mp,2Mp) [2Mp,3Mp) bf16 terms, [3Mp,4Mp) fp16.
__global__ void __launch_bounds__(256)
split_x_kernel(const float *__restrict__ X, int64_t ldx, int M, int K, int Mp, int Kp,
               uint16_t *__restrict__ out, int *__restrict__ flags)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)Mp * Kp;
    int used = 0;
    if (i < total)
    {
        const int m = (int)(i / Kp), k = (int)(i - (long long)m * Kp);
        const float x = (m < M && k < K) ? X[(int64_t)m * ldx + k] : 0.0f;
        const __nv_bfloat16 x1 = __float2bfloat16_rn(x);
        const float r1 = x - __bfloat162float(x1);
        const __nv_bfloat16 x2 = __float2bfloat16_rn(r1);
        const float r2 = r1 - __bfloat162float(x2);
        const __nv_bfloat16 x3 = __float2bfloat16_rn(r2);
        const __half h = __float2half_rn(x);
        out[i] = __bfloat16_as_ushort(x1);
        out[total + i] = __bfloat16_as_ushort(x2);
        out[2 * total + i] = __bfloat16_as_ushort(x3);
        out[3 * total + i] = __half_as_ushort(h);
        used = (r1 != 0.0f ? 1 : 0) | (r2 != 0.0f ? 2 : 0) | (__half2float(h) == x ? 0 : 4);
    }
    // one atomic per warp at most
    for (int o = 16; o > 0; o >>= 1)
        used |= __shfl_xor_sync(0xffffffffu, used, o);
    if ((threadIdx.x & 31) == 0 && used)
        atomicOr(flags, used);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn)
    {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

template <int NT, bool XK>
int launch_nt(const CUtensorMap &map, const DenseParams &p, dim3 grid, size_t smem, int device, cudaStream_t st)
{
    static size_t configured[64] = {0};
    size_t &have = configured[device & 63];
    if (have < smem)
    {
        TSG_CUDA(cudaFuncSetAttribute(dense_tc_kernel<NT, XK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        have = smem;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = grid.z; // the K-splits of a tile form one cluster
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    TSG_CUDA(cudaLaunchKernelEx(&cfg, dense_tc_kernel<NT, XK>, map, p));
    TSG_LAUNCHED();
    return TSG_OK;
}

// K-split: smallest factor that fills the machine to >= 85 % in whole waves
int choose_ksplit(long long tiles, int nkb, int sms)
{
    int ksplit = 1;
    const int max_split = nkb / 4 > 0 ? (nkb / 4 > 8 ? 8 : nkb / 4) : 1; // portable cluster size
    double best = -1.0;
    for (int ks = 1; ks <= max_split; ++ks)
    {
        const long long ctas = tiles * ks;
        const long long waves = (ctas + sms - 1) / sms;
        const double eff = (double)ctas / (double)(waves * sms);
        if (eff > best + 0.03) // prefer the smaller split unless clearly better
        {
            best = eff;
            ksplit = ks;
        }
        if (eff >= 0.85)
            break;
    }
    return ksplit;
}

} // namespace

int tsg_launch_dense_tc(tsg_matrix *m, const float *X, int64_t ldx, const float *b,
                        const float *alpha, float *Y, int64_t ldy, int M, cudaStream_t st)
{
    if (M <= 0 || m->N == 0)
        return TSG_OK;
    const int K = m->K, N = m->N;
    const int Kp = (K + kBlockK - 1) / kBlockK * kBlockK;
    TSG_CHECK(Kp > 0, TSG_ERR_UNSUPPORTED, "dense_tc: K == 0");
    TSG_CHECK(m->codes != nullptr && m->code_kblocks == Kp / kBlockK, TSG_ERR_UNSUPPORTED,
              "dense_tc: tile codes missing");
    const int nkb = Kp / kBlockK;
    const int ntiles = (N + kTileN - 1) / kTileN;
    const int sms = m->sm_count > 0 ? m->sm_count : 148;

    // Small M: X is converted inside the kernel, 16 rows per m-tile, ONE launch.  Each extra
    // m-tile repeats the expansion of W (~3.7e-8 µs per matrix element); worth it while that
    // stays below the ~3 µs the two extra launches of the TMA path cost.
    const int mt16 = (M + 15) / 16;
    const double expand_us = 3.7e-8 * (double)K * (double)N;
    const bool xk = M <= 16 || (M <= 64 && (mt16 - 1) * expand_us < 3.0);

    DenseParams p = {};
    p.codes = m->codes;
    p.N = N;
    p.M = M;
    p.K = K;
    p.nkb = nkb;
    p.bias = b;
    p.alpha = alpha;
    p.Y = Y;
    p.ldy = ldy;
    p.X = X;
    p.ldx = ldx;
    p.stage_budget = (int)(m->smem_optin - 1024 - kBarBytes);
    const size_t smem = m->smem_optin; // the kernel sizes its stage ring from the budget at run time
    CUtensorMap map = {};

    if (xk)
    {
        p.ksplit = choose_ksplit((long long)ntiles * mt16, nkb, sms);
        dim3 grid(ntiles, mt16, p.ksplit);
        TSG_CHECK(mt16 <= 65535, TSG_ERR_UNSUPPORTED, "dense_tc: grid too large");
        return launch_nt<16, true>(map, p, grid, smem, m->device, st);
    }

    EncodeTiledFn encode = get_encode();
    TSG_CHECK(encode != nullptr, TSG_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    const int NT = M <= 32 ? 32 : (M <= 64 ? 64 : 128);
    const int mtiles = (M + NT - 1) / NT;
    const int Mp = mtiles * NT;
    p.Mp = Mp;
    p.ksplit = choose_ksplit((long long)ntiles * mtiles, nkb, sms);

    // scratch: flags + split terms of X (16-bit [4][Mp][Kp]: three bf16 terms and one fp16 copy)
    const size_t xs_bytes = (size_t)(kMaxSplits + 1) * Mp * Kp * sizeof(uint16_t);
    const size_t need = 256 + xs_bytes + 256;
    if (m->cap_xsplit < need)
    {
        if (m->xsplit)
            cudaFree(m->xsplit);
        m->xsplit = nullptr;
        m->cap_xsplit = 0;
        TSG_CUDA(cudaMalloc(&m->xsplit, need));
        m->cap_xsplit = need;
    }
    int *flags = reinterpret_cast<int *>(m->xsplit);
    uint16_t *xs = reinterpret_cast<uint16_t *>((char *)m->xsplit + 256);
    p.flags = flags;

    TSG_CUDA(cudaMemsetAsync(flags, 0, 4, st));
    {
        const long long total = (long long)Mp * Kp;
        split_x_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(X, ldx, M, K, Mp, Kp, xs, flags);
        TSG_LAUNCHED();
    }

    // tensor map over the split buffer: 2-D [4*Mp rows][Kp], box 64 x NT, 128-byte swizzle
    {
        const cuuint64_t gdim[2] = {(cuuint64_t)Kp, (cuuint64_t)(kMaxSplits + 1) * Mp};
        const cuuint64_t gstride[1] = {(cuuint64_t)Kp * sizeof(uint16_t)};
        const cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)NT};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, xs, gdim, gstride, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        TSG_CHECK(r == CUDA_SUCCESS, TSG_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    }

    dim3 grid(ntiles, mtiles, p.ksplit);
    TSG_CHECK(mtiles <= 65535, TSG_ERR_UNSUPPORTED, "dense_tc: grid too large");
    switch (NT)
    {
    case 32:
        return launch_nt<32, false>(map, p, grid, smem, m->device, st);
    case 64:
        return launch_nt<64, false>(map, p, grid, smem, m->device, st);
    default:
        return launch_nt<128, false>(map, p, grid, smem, m->device, st);
    }
}
