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
// D[tmem] (+)= A[smem] · B[smem]ᵀ, bf16 in, fp32 out
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
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
// kind::f16: D = F32 (bits 4-5 = 1), A = B = BF16 (bits 7-9, 10-12 = 1), both K-major,
// N>>3 at bit 17, M>>4 at bit 24 (InstrDescriptor in the same header).
__host__ __device__ constexpr uint32_t make_idesc(int n)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileN >> 4) << 24);
}

// One code nibble [nz0, neg0, nz1, neg1] -> two packed bf16 in {0, +1, -1}.
//   y * (2^7+2^14+2^21+2^28) drops the four flags at bits 7, 15, 23, 31 (the partial products do
//   not overlap); a flag at bit 7/23 times 0x7F is the |1.0| pattern 0x3F80; bits 15/31 are the
//   signs.
__device__ __forceinline__ uint32_t expand_nibble(uint32_t code, int sh)
{
    const uint32_t t = ((code >> sh) & 0xFu) * 0x10204080u;
    return ((t & 0x00800080u) * 0x7Fu) | (t & 0x80008000u);
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
    int NT;          // accumulator columns (m-tile), multiple of 16
    int nkb;         // k-blocks in total (Kp / 64)
    int ksplit;      // K-splits
    int Mp;          // padded rows per split term in the X buffer
    const int *flags; // bit0: term 2 non-zero, bit1: term 3 non-zero
    const float *bias, *alpha;
    float *Y;        // M×N
    int64_t ldy;
    int stages;
};

template <int NT>
__global__ void __launch_bounds__(kThreads, 1)
dense_tc_kernel(const __grid_constant__ CUtensorMap xmap, const DenseParams p)
{
    // the (up to) three split terms accumulate side by side: columns [t*NT, (t+1)*NT)
    constexpr int kTmemCols = NT * kMaxSplits <= 32 ? 32 : (NT * kMaxSplits <= 64 ? 64 : (NT * kMaxSplits <= 128 ? 128 : (NT * kMaxSplits <= 256 ? 256 : 512)));
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // [stages][A 16 KB | B kMaxSplits*NT*128], then barriers
    constexpr int kABytes = kTileN * 128;
    constexpr int kBBytes = NT * 128;
    constexpr int kStageBytes = kABytes + kMaxSplits * kBBytes;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char *smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    const int S = p.stages;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_al + (size_t)S * kStageBytes);
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * S, tmem_full = empty0 + 8 * S;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * S + 1);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int n0 = blockIdx.x * kTileN;
    const int mtile = blockIdx.y;
    const int split = blockIdx.z;
    const int kb_lo = (int)(((long long)p.nkb * split) / p.ksplit);
    const int kb_hi = (int)(((long long)p.nkb * (split + 1)) / p.ksplit);
    const int iters = kb_hi - kb_lo;
    const int fl = *p.flags;
    const int nsplit = (fl & 2) ? 3 : ((fl & 1) ? 2 : 1);

    if (warp == kMmaWarp && lane == 0)
    {
        for (int s = 0; s < S; ++s)
        {
            mbar_init(full0 + 8 * s, 5);  // 4 expander warps + the producer's expect_tx arrive
            mbar_init(empty0 + 8 * s, 1); // one tcgen05.commit
        }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    else if (warp == kAllocWarp)
    {
        tmem_alloc(smem_u32(tmem_slot), kTmemCols);
    }
    else if (warp == kTmaWarp && lane == 0)
    {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;

    if (warp == kTmaWarp)
    {
        // ===== TMA producer: X tiles of the (up to) three split terms =====
        if (elect_one())
        {
            // stage index / phase advance by increments: this is a single thread on the critical
            // path, integer division by the runtime stage count would dominate its loop
            uint32_t eb = empty0, fb = full0, bdst = smem_base + kABytes, ph = 0;
            int kcoord = kb_lo * kBlockK, st = 0;
            for (int it = 0; it < iters; ++it)
            {
                mbar_wait(eb, ph ^ 1);
                mbar_arrive_expect_tx(fb, (uint32_t)(nsplit * kBBytes));
                tma_load_2d(bdst, &xmap, fb, kcoord, mtile * NT);
                if (nsplit > 1)
                    tma_load_2d(bdst + kBBytes, &xmap, fb, kcoord, p.Mp + mtile * NT);
                if (nsplit > 2)
                    tma_load_2d(bdst + 2 * kBBytes, &xmap, fb, kcoord, 2 * p.Mp + mtile * NT);
                kcoord += kBlockK;
                eb += 8, fb += 8, bdst += kStageBytes;
                if (++st == S)
                    st = 0, eb = empty0, fb = full0, bdst = smem_base + kABytes, ph ^= 1;
            }
        }
    }
    else if (warp == kMmaWarp)
    {
        // ===== MMA issuer: the whole warp runs the loop (uniform registers), one lane issues =====
        {
            // One MMA per 16-k step covers all split terms at once: their X tiles are adjacent in
            // smem (rows [t*NT, (t+1)*NT)), so B is simply nsplit*NT rows tall and term t lands in
            // accumulator columns [t*NT, (t+1)*NT).  (N <= 256 per instruction: NT=128 with three
            // terms issues 256 + 128.)  The terms are added in the epilogue.
            const int nrows = nsplit * NT;
            const uint32_t idesc_a = make_idesc(nrows > 256 ? 256 : nrows);
            const uint32_t idesc_b = make_idesc(NT); // only used when nrows == 384
            const uint64_t adesc0 = make_smem_desc(smem_base);
            constexpr uint64_t kStageStep = kStageBytes >> 4, kBOff = kABytes >> 4, kBStep = kBBytes >> 4;
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
                    for (int k = 0; k < kBlockK / 16; ++k) // UMMA_K = 16 bf16 = 32 B: +2 in the address field
                        umma_bf16(tmem_d, adesc + 2 * k, adesc + kBOff + 2 * k, idesc_a, (it | k) != 0);
                    if (nrows > 256)
                    {
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k)
                            umma_bf16(tmem_d + 256, adesc + 2 * k, adesc + kBOff + 2 * kBStep + 2 * k, idesc_b,
                                      (it | k) != 0);
                    }
                    umma_commit(eb); // frees the stage when these MMAs retire
                }
                __syncwarp();
                adesc += kStageStep, fb += 8, eb += 8;
                if (++st == S)
                    st = 0, adesc = adesc0, fb = full0, eb = empty0, ph ^= 1;
            }
            if (elect_one())
                umma_commit(tmem_full);
            __syncwarp();
        }
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
        // ===== expanders: tile-packed codes -> swizzled bf16 A tile =====
        // A group may only run one barrier phase ahead of the MMA issuer (mbarrier parity is one
        // bit), which holds iff #groups <= #stages; with the 3-stage NT=128 pipeline the fourth
        // group sits the main loop out (that shape is MMA-bound anyway).
        const int groups = S < kExpGroups ? S : kExpGroups;
        const int grp = slice;                       // k-blocks with it % groups == grp
        const int row = erow;                        // smem row of the A tile
        const uint4 *src = p.codes + ((size_t)blockIdx.x * p.nkb + kb_lo) * 128 + row;
        // The code stream is the kernel's HBM stream (2 KB per k-block and tile): each thread
        // pulls its 16 bytes into L2 kPrefetch k-blocks ahead (prefetch.global.L2) and into
        // registers two of its own iterations ahead, so the expansion never waits on DRAM.
        constexpr int kPrefetch = 32;
        uint4 nxt = make_uint4(0, 0, 0, 0), nxt2 = make_uint4(0, 0, 0, 0);
        if (grp < groups)
        {
            for (int it = grp; it < iters && it < kPrefetch; it += groups)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(src + (size_t)it * 128));
            if (grp < iters)
                nxt = __ldg(src + (size_t)grp * 128);
            if (grp + groups < iters)
                nxt2 = __ldg(src + (size_t)(grp + groups) * 128);
        }
        int st = grp;            // grp < groups <= S
        uint32_t ph = 0;
        for (int it = (grp < groups ? grp : iters); it < iters; it += groups)
        {
            const int s = st;
            const uint4 cur = nxt;
            nxt = nxt2;
            if (it + kPrefetch < iters)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(src + (size_t)(it + kPrefetch) * 128));
            if (it + 2 * groups < iters) // codes two iterations ahead, in flight during this expansion
                nxt2 = __ldg(src + (size_t)(it + 2 * groups) * 128);
            mbar_wait(empty0 + 8 * s, ph ^ 1);
            const uint32_t rowaddr = smem_base + s * kStageBytes + row * 128;
            const uint32_t sw = (uint32_t)(row & 7);
            const uint32_t cw[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
            for (int w = 0; w < 4; ++w)
            {
#pragma unroll
                for (int h = 0; h < 2; ++h) // 16-byte chunk = 8 elements = 4 nibbles
                {
                    const int sh = h * 16;
                    const uint32_t w0 = expand_nibble(cw[w], sh), w1 = expand_nibble(cw[w], sh + 4),
                                   w2 = expand_nibble(cw[w], sh + 8), w3 = expand_nibble(cw[w], sh + 12);
                    const uint32_t chunk = (uint32_t)(w * 2 + h);
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(rowaddr + ((chunk ^ sw) << 4)),
                                 "r"(w0), "r"(w1), "r"(w2), "r"(w3)
                                 : "memory");
                }
            }
            fence_proxy_async(); // generic-proxy smem writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0)
                mbar_arrive(full0 + 8 * s);
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
                for (int t = 1; t < nsplit; ++t) // x1 + x2 + x3 terms, fixed order
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
    float *park = reinterpret_cast<float *>(smem_al); // [NT][128] column-major, reuses the stages
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
                    float y = __uint_as_float(acc[j][c]);
                    for (int r = 1; r < p.ksplit; ++r) // rank order: deterministic
                        y += ld_dsmem_f32(smem_u32(park + (ch * 16 + c) * 128 + erow), (uint32_t)r);
                    const int m = mtile * NT + ch * 16 + c;
                    if (m < p.M && en < p.N)
                    {
                        y = y + bn;
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

// fp32 X -> three bf16 terms (exact: x == x1 + x2 + x3), zero padded to [Mp][Kp] each.
__global__ void __launch_bounds__(256)
split_x_kernel(const float *__restrict__ X, int64_t ldx, int M, int K, int Mp, int Kp,
               __nv_bfloat16 *__restrict__ out, int *__restrict__ flags)
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
        out[i] = x1;
        out[total + i] = x2;
        out[2 * total + i] = x3;
        used = (r1 != 0.0f ? 1 : 0) | (r2 != 0.0f ? 2 : 0);
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

template <int NT>
int launch_nt(const CUtensorMap &map, const DenseParams &p, dim3 grid, size_t smem, int device, cudaStream_t st)
{
    static size_t configured[64] = {0};
    size_t &have = configured[device & 63];
    if (have < smem)
    {
        TSG_CUDA(cudaFuncSetAttribute(dense_tc_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
    TSG_CUDA(cudaLaunchKernelEx(&cfg, dense_tc_kernel<NT>, map, p));
    TSG_LAUNCHED();
    return TSG_OK;
}

} // namespace

int tsg_launch_dense_tc(tsg_matrix *m, const float *X, int64_t ldx, const float *b,
                        const float *alpha, float *Y, int64_t ldy, int M, cudaStream_t st)
{
    if (M <= 0 || m->N == 0)
        return TSG_OK;
    EncodeTiledFn encode = get_encode();
    TSG_CHECK(encode != nullptr, TSG_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    const int K = m->K, N = m->N;
    const int Kp = (K + kBlockK - 1) / kBlockK * kBlockK;
    TSG_CHECK(Kp > 0, TSG_ERR_UNSUPPORTED, "dense_tc: K == 0");
    TSG_CHECK(m->codes != nullptr && m->code_kblocks == Kp / kBlockK, TSG_ERR_UNSUPPORTED,
              "dense_tc: tile codes missing");
    const int NT = M <= 16 ? 16 : (M <= 32 ? 32 : (M <= 64 ? 64 : 128));
    const int mtiles = (M + NT - 1) / NT;
    const int Mp = mtiles * NT;
    const int nkb = Kp / kBlockK;
    const int ntiles = (N + kTileN - 1) / kTileN;

    // K-split: smallest factor that fills the machine to >= 85 % in whole waves
    const long long tiles = (long long)ntiles * mtiles;
    int ksplit = 1;
    {
        const int sms = m->sm_count > 0 ? m->sm_count : 148;
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
    }

    // scratch: flags + split terms of X (bf16 [3][Mp][Kp])
    const size_t xs_bytes = (size_t)kMaxSplits * Mp * Kp * sizeof(__nv_bfloat16);
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
    __nv_bfloat16 *xs = reinterpret_cast<__nv_bfloat16 *>((char *)m->xsplit + 256);

    TSG_CUDA(cudaMemsetAsync(flags, 0, 4, st));
    {
        const long long total = (long long)Mp * Kp;
        split_x_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(X, ldx, M, K, Mp, Kp, xs, flags);
        TSG_LAUNCHED();
    }

    // tensor map over the split buffer: 2-D [3*Mp rows][Kp], box 64 x NT, 128-byte swizzle
    CUtensorMap map;
    {
        const cuuint64_t gdim[2] = {(cuuint64_t)Kp, (cuuint64_t)kMaxSplits * Mp};
        const cuuint64_t gstride[1] = {(cuuint64_t)Kp * sizeof(__nv_bfloat16)};
        const cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)NT};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, xs, gdim, gstride, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        TSG_CHECK(r == CUDA_SUCCESS, TSG_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    }

    const size_t stage_bytes = (size_t)kTileN * 128 + (size_t)kMaxSplits * NT * 128;
    int stages = (int)((m->smem_optin - 2048) / stage_bytes);
    if (stages > 8)
        stages = 8;
    TSG_CHECK(stages >= 2, TSG_ERR_UNSUPPORTED, "dense_tc: shared memory too small for two stages");
    const size_t smem = 1024 + (size_t)stages * stage_bytes + (2 * stages + 2) * 8 + 16;

    DenseParams p;
    p.codes = m->codes;
    p.N = N;
    p.M = M;
    p.K = K;
    p.NT = NT;
    p.nkb = nkb;
    p.ksplit = ksplit;
    p.Mp = Mp;
    p.flags = flags;
    p.bias = b;
    p.alpha = alpha;
    p.Y = Y;
    p.ldy = ldy;
    p.stages = stages;
    dim3 grid(ntiles, mtiles, ksplit);
    TSG_CHECK(mtiles <= 65535 && ksplit <= 65535, TSG_ERR_UNSUPPORTED, "dense_tc: grid too large");
    int s = TSG_OK;
    switch (NT)
    {
    case 16:
        s = launch_nt<16>(map, p, grid, smem, m->device, st);
        break;
    case 32:
        s = launch_nt<32>(map, p, grid, smem, m->device, st);
        break;
    case 64:
        s = launch_nt<64>(map, p, grid, smem, m->device, st);
        break;
    default:
        s = launch_nt<128>(map, p, grid, smem, m->device, st);
        break;
    }
    TSG_TRY(s);
    return TSG_OK;
}
