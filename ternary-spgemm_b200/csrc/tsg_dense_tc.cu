// tsg_dense_tc.cu — dense-expand tensor-core path (north-star subsystem 3).
//
// Same contract as BaseTCSC / BaseTCSC_PreLU (reference cpp_impl/comp.h:25-69,
// cpp_impl/comp_prelu.h:12-70) — Y = X·W + b, optional PReLU — computed as a dense GEMM on the
// 5th-generation tensor cores:
//
//      D[n, m] = Σ_k  Wt[n, k] · X[m, k]            (UMMA: D = A·Bᵀ, K-major)
//
//   A = Wᵀ tile, 128 columns of W × 64 k per sub-block, 16-bit floats in {0, +2, -2}, EXPANDED
//       ON THE FLY from the tile-packed 2-bit codes (tsg_matrix::codes) in registers — one shift
//       and one AND per two matrix elements — and written with tcgen05.st straight into TENSOR
//       MEMORY, from where tcgen05.mma reads it as its A operand.  W never exists as 16-bit data
//       in HBM (HBM sees K·N/4 bytes, coalesced 128-bit loads) nor in shared memory: an A tile
//       in shared memory would cost 16 KB of writes plus 16 KB of tensor-core reads per
//       sub-block through the SM's 128 B/clk shared-memory port, which is what bounded the
//       previous revision for small M.  0x4000 is 2.0 in bf16 and in fp16, so the tile is the
//       same for both X formats; the factor 2 is taken out, exactly, in the epilogue.
//   B = X tile, NT rows × 64 k, in shared memory (128-byte swizzle).  fp32 X is split EXACTLY into
//       three bf16 terms x = x1 + x2 + x3 (8+8+8 mantissa bits); the three products accumulate
//       side by side in fp32, so every product W·x_i is exact and only the fp32 accumulation
//       rounds — like the reference's fp32 adds.
//         small M (<= 16 rows per m-tile): converted inside the kernel by the expander warps —
//           ONE launch, no scratch, always three terms;
//         larger M: split_tiles_kernel converts X once, TMA (cp.async.bulk.tensor) loads the
//           tiles.  The operand format is chosen PER TILE (rows of one m-tile x 64 k): a tile
//           whose values are all exact in fp16 (the reference's integer-valued inputs are) is
//           written and multiplied as ONE fp16 term, any other tile as one to three bf16 terms
//           (terms that are identically zero in the tile are neither written nor multiplied).
//           tcgen05.mma takes the format of B per instruction and 0x4000 is 2.0 in both formats,
//           so tiles of both kinds accumulate into the same fp32 columns.  One byte per tile
//           tells the TMA producer and the MMA issuer what the split kernel wrote.
//   D = 128 × (terms·NT) fp32 accumulators in TMEM, read back with tcgen05.ld.
//
// One CTA per (128-column tile of W, m-tile, K-split), warp-specialised (EW = 16 or 8 expander
// warps; 16: one CTA per SM with all 512 TMEM columns, 8: two CTAs per SM with 256 each):
//   warps 0..EW-1  expanders: 4-warp groups expand the sub-blocks of every stage (stage = 4
//                  sub-blocks = 256 k); thread -> one W column = one TMEM lane.  Afterwards the
//                  same warps run the epilogue (TMEM -> registers -> bias/PReLU -> coalesced stores).
//   warp EW        TMA producer for the X tiles (one elected lane), own ring of sub-block tiles
//   warp EW+2      TMEM allocator / deallocator
//   warp EW+3      tcgen05.mma issuer (one elected lane): 16 MMAs per stage, tcgen05.commit
//                  releases the A stage in TMEM and the X tiles in shared memory
// Tile heights: 16 rows of X (in-kernel conversion), 32 / 64 / 128 (TMA), 256 (TMA, one split term).
// mbarriers: afull[s] (EW expander warps), aempty[s] (commit), bfull/bempty per X tile (TMA
// path), tmem_full (last commit).
// Split-K: the K-splits of one tile form a thread-block CLUSTER (<= 8 CTAs).  Non-leader CTAs
// push their accumulators into the leader's shared memory (st.shared::cluster); after a cluster
// barrier the leader adds them in rank order (deterministic), applies bias / PReLU and writes
// Y — no partial sums in HBM, no second kernel.
// Tail launch: the tiles of a tall-tile grid's partial last wave are launched separately, K-split
// over clusters that meet in an L2-resident workspace instead (256-row peers do not fit the
// leader's shared memory); every rank then finishes 1/ksplit of the tile's rows (plan_tail).
#include "tsg_internal.cuh"

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include <type_traits>

namespace
{

constexpr int kTileN = 128;   // W columns per CTA  (UMMA M)
constexpr int kBlockK = 64;   // k per sub-block (128 bytes of 16-bit floats per row, 32 TMEM columns)
constexpr int kSub = 4;       // sub-blocks per stage: one per expander group
// Warp roles (template parameter EW = expander warps): warps [0, EW) expand and run the epilogue,
// warp EW is the TMA producer, EW+2 allocates TMEM, EW+3 issues the MMAs.  The SM's issue arbiter
// favours the highest warp id among eligible warps, so the single-thread critical-path roles get
// the top ids and are never starved by the expander warps sharing their schedulers.
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
// pairs of 16-bit floats in {0, +2, -2}: (word << 2p) & 0xC000C000.
__device__ __forceinline__ void expand_word(uint32_t w, uint32_t (&r)[8])
{
    constexpr uint32_t kMask = 0xC000C000u;
#pragma unroll
    for (int p = 0; p < 8; ++p)
        r[p] = (w << (2 * p)) & kMask;
}
// 16 consecutive 32-bit TMEM columns of this thread's lane <- 16 registers
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&a)[8], const uint32_t (&b)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::
                     "r"(taddr), "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
                 "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] · B[smem]ᵀ
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// volatile loads keep their place in program order (the compiler must not sink them to their use)
__device__ __forceinline__ uint4 ldg_v4_ordered(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_f32_ordered(const float *p)
{
    float r;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
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
__device__ __forceinline__ uint32_t mapa_rank(uint32_t local_addr, uint32_t rank)
{
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
    return remote;
}
__device__ __forceinline__ void st_dsmem_u32(uint32_t remote_addr, uint32_t v)
{
    asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(remote_addr), "r"(v) : "memory");
}

struct DenseParams
{
    const uint4 *codes; // [tiles][nkb][128] tile-packed 2-bit codes
    int N, M, K;
    int nkb;         // k-blocks in total (Kp / 64)
    int ksplit;      // K-splits (= cluster size along z)
    int ntiles;      // 128-column tiles of W; CTA blockIdx.x works on tile (tile0 + blockIdx.x): n-tile = % ntiles, m-tile = / ntiles
    int tile0;
    float *gws;      // split-K through global memory (tail launch, see tsg_launch_dense_tc): [tile][rank][nt][128] partial sums, else NULL
    int Mp;          // padded rows per split term in the pre-split X buffer (TMA path)
    const uint8_t *tflags; // TMA path, one byte per (m-tile, k-block), [mtiles][nkb] (kTile* bits below)
    const int32_t *csp, *csn, *rip, *rin; // TCSC arrays: the reference-order fallback for non-finite X
    int nt;           // rows of X per m-tile for the run-time-height instantiation (NT = 256)
    const float *X;  // in-kernel conversion path: fp32 X
    int64_t ldx;
    const float *bias, *alpha;
    float *Y;        // M×N
    int64_t ldy;
    int smem_budget; // shared-memory bytes available for X tiles + the split-K landing zone
    unsigned long long *trace; // developer trace (TSG_TC_TRACE=1): 16 clock stamps per CTA, else NULL
};

// stamp slot `slot` of this CTA with the SM cycle counter relative to the CTA's first stamp
#define TC_TRACE(slot)                                                                          \
    do                                                                                          \
    {                                                                                           \
        if (p.trace != nullptr)                                                                 \
            p.trace[(((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16 + (slot)] = \
                (unsigned long long)clock64();                                                  \
    } while (0)

constexpr int kABytes = kTileN * 128;  // one expanded A stage: 128 columns x 64 k x 2 B
constexpr int kBarBytes = 1024;        // barriers + TMEM slot live in front of the stages
constexpr int kFlagBytes = 4096;       // this CTA's tile flags (one byte per k-block of its K range)

// what split_tiles_kernel found in one X tile (rows of an m-tile x 64 k) and therefore wrote
constexpr uint32_t kTileTerm2 = 1;   // some x is not exact in bf16: second bf16 term written
constexpr uint32_t kTileTerm3 = 2;   // some x has more than 16 significant bits: third term written
constexpr uint32_t kTileBf16 = 4;    // some x is not exact in fp16: bf16 terms (else ONE fp16 term)
constexpr uint32_t kTileHuge = 8;    // some x is non-finite or >= 2^100: reference-order fallback
constexpr uint32_t kTileF16x2 = 16;  // full-precision values inside fp16's range: TWO fp16 terms (x1 in the fp16
                                     // plane, the remainder in bf16 plane 1's storage) — 22 significant bits

// fp32 -> three bf16 terms, two elements at a time (exact: x == t1 + t2 + t3).
__device__ __forceinline__ void split3_pair(float x0, float x1, uint32_t &t1, uint32_t &t2, uint32_t &t3)
{
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(t1) : "f"(x1), "f"(x0));
    const float r0 = x0 - __uint_as_float(t1 << 16), r1 = x1 - __uint_as_float(t1 & 0xFFFF0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(t2) : "f"(r1), "f"(r0));
    const float q0 = r0 - __uint_as_float(t2 << 16), q1 = r1 - __uint_as_float(t2 & 0xFFFF0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(t3) : "f"(q1), "f"(q0));
}

// NT : accumulator columns per split term (rows of X per m-tile), multiple of 16
// XK : true  -> X is converted to its bf16 terms inside the kernel (always three terms),
//      false -> X tiles come by TMA from the buffer split_tiles_kernel wrote.
// EW : expander/epilogue warps.  16: one CTA per SM with all 512 TMEM columns.  8: TWO CTAs per SM
//      (256 TMEM columns and half the shared memory each): while one CTA sits in its prologue
//      (TMEM alloc, first HBM latency) or epilogue, the other keeps the tensor core fed; needs
//      terms*NT <= 128 accumulator columns, which the host can only promise for NT <= 32.
// NT = 256 halves the A-operand feeds, expansions and code reads per flop: at N = 128 a 128x16
// slice of W takes about as long to enter the tensor core from TMEM (64 B/clk) as its MMA takes.
template <int NT, bool XK, int EW>
__global__ void __launch_bounds__((EW + 4) * 32, EW == 8 ? 2 : 1)
dense_tc_kernel(const __grid_constant__ CUtensorMap xmap, const DenseParams p)
{
    constexpr int kThreadsT = (EW + 4) * 32;
    constexpr int G = EW / 4;                  // expander groups (4 warps = 128 TMEM lanes each)
    constexpr int kMine = kSub / G;            // sub-blocks of a stage one group expands
    constexpr int kTmem = EW == 8 ? 256 : 512; // TMEM columns of this CTA
    constexpr int kTmaW = EW, kAllocW = EW + 2, kMmaW = EW + 3;
    // rows of X per m-tile: a template constant, except for the 256 instantiation where it is the
    // run-time p.nt (any multiple of 16 up to 256, chosen by the host to fill whole waves)
    const int nt = (NT == 256) ? p.nt : NT;
    const int kBBytes = nt * 128;              // X tile of one split term and one sub-block
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char *smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_al);
    const uint8_t *sflags = smem_al + kBarBytes;  // this CTA's tile flags (TMA path)
    const uint32_t xs0 = smem_base + kBarBytes + kFlagBytes;   // X tiles start here

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

    const int lin = p.tile0 + (int)blockIdx.x;   // n-tiles fastest: CTAs running together share an m-tile of X
    const int ntile = lin % p.ntiles, mtile = lin / p.ntiles;
    const int n0 = ntile * kTileN;
    const int split = blockIdx.z;
    // this CTA's K range in stages of kSub sub-blocks (p.nkb is a multiple of kSub: the builder
    // pads the code stream with zero codes)
    const int nst = p.nkb / kSub;
    const int st_lo = (int)(((long long)nst * split) / p.ksplit);
    const int st_hi = (int)(((long long)nst * (split + 1)) / p.ksplit);
    const int iters = st_hi - st_lo;
    if (tid == 0)
        TC_TRACE(0);

    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int grp = warp < EW ? warp >> 2 : 0;    // expander group
    const int erow = q * 32 + lane;               // W column inside the tile = TMEM lane

    // ---- requests that do not depend on the prologue go out first ---------------------------
    // code stream: one uint4 (64 k of this thread's column) per sub-block, 2 KB per sub-block and
    // tile, coalesced; registers hold this group's sub-blocks of the current stage and the next
    // kRing-1, an L2 prefetch runs kPrefetch stages ahead.  Group g owns sub-blocks g, g+G, ...
    const uint4 *src = p.codes + ((size_t)ntile * p.nkb + (size_t)st_lo * kSub + grp) * 128 + erow;
    constexpr int kPrefetch = 12, kRing = EW == 8 ? 3 : 4;
    uint4 ring[kRing][kMine];
#pragma unroll
    for (int i = 0; i < kRing; ++i)
#pragma unroll
        for (int u = 0; u < kMine; ++u)
            ring[i][u] = make_uint4(0, 0, 0, 0);
    // in-kernel X conversion: pairs (row q + 4j, k = 2*lane, 2*lane+1) of this group's sub-blocks
    constexpr int kPairs = XK ? NT / 4 : 1;
    float2 xv[kMine][kPairs];
    uint32_t xhuge = 0; // in-kernel conversion: largest |x| bit pattern this thread saw (inf / NaN sort above finite)
    auto load_x = [&](int it) {
        if constexpr (XK)
        {
#pragma unroll
            for (int u = 0; u < kMine; ++u)
            {
                const int k = ((st_lo + it) * kSub + grp + u * G) * kBlockK + 2 * lane;
#pragma unroll
                for (int j = 0; j < kPairs; ++j)
                {
                    const int m = mtile * nt + q + 4 * j;
                    xv[u][j] = make_float2(0.0f, 0.0f);
                    if (m < p.M)
                    {
                        const float *xp = p.X + (int64_t)m * p.ldx + k;
                        if (k < p.K)
                            xv[u][j].x = __ldg(xp);
                        if (k + 1 < p.K)
                            xv[u][j].y = __ldg(xp + 1);
                    }
                }
            }
        }
    };
    float bn = 0.0f, an = 0.0f; // epilogue operands of this thread's column
    if (warp < EW)
    {
#pragma unroll
        for (int i = 0; i < kRing; ++i)
            if (i < iters)
            {
#pragma unroll
                for (int u = 0; u < kMine; ++u)
                    ring[i][u] = ldg_v4_ordered(src + ((size_t)i * kSub + u * G) * 128);
            }
        for (int it = kRing; it < iters && it < kPrefetch; ++it)
#pragma unroll
            for (int u = 0; u < kMine; ++u)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(src + ((size_t)it * kSub + u * G) * 128));
    }
    // Programmatic dependent launch (behind split_tiles_kernel on the TMA path, behind the previous call's
    // kernel otherwise): everything above reads only the weight stream, which no kernel in front of
    // us writes.  From here on we touch what the previous kernel may have produced (X, split X,
    // flags, bias) or still be reading (Y), so wait for it.  A launch without the attribute, or one
    // behind a kernel that never triggers, is an ordinary serialised launch and this returns at once.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (warp < EW)
    {
        load_x(0);
        if (n0 + erow < p.N)
        {
            bn = ldg_f32_ordered(p.bias + n0 + erow);
            if (p.alpha)
                an = ldg_f32_ordered(p.alpha + n0 + erow);
        }
    }

    // TMA path: what split_tiles_kernel wrote for every (m-tile, k-block) tile of this CTA's K range —
    // one byte per tile, copied to shared memory once; the TMA producer, the MMA issuer and the
    // epilogue all read the same bytes, so they agree on terms and formats without talking.  The OR
    // of all of them (one word per warp, combined after the prologue barrier) sizes the pipeline.
    const uint32_t afull0 = smem_u32(bars), aempty0 = afull0 + 8 * 4, bfull0 = aempty0 + 8 * 4,
                   bempty0 = bfull0 + 8 * 16, tmem_full = bempty0 + 8 * 16;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 48);
    uint32_t *huge_local = reinterpret_cast<uint32_t *>(bars + 49);     // in-kernel conversion: CTA-wide OR of xhuge
    uint32_t *huge_ranks = reinterpret_cast<uint32_t *>(bars + 50);     // [8]: the cluster ranks' verdicts, pushed to the leader
    uint32_t *warp_or = reinterpret_cast<uint32_t *>(bars + 54);        // [EW + 4]: OR of the flag words each warp copied
    uint32_t *warp_and = reinterpret_cast<uint32_t *>(bars + 64);       // [EW + 4]: AND of them
    if constexpr (!XK)
    {
        const uint32_t *fsrc = reinterpret_cast<const uint32_t *>(p.tflags + (size_t)mtile * p.nkb) + st_lo; // kSub == 4 flags per word
        uint32_t mine_or = 0, mine_and = 0xFFFFFFFFu;
        for (int i = tid; i < iters; i += kThreadsT)
        {
            uint32_t w;
            asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(w) : "l"(fsrc + i));
            reinterpret_cast<uint32_t *>(smem_al + kBarBytes)[i] = w;
            mine_or |= w, mine_and &= w;
        }
        mine_or = __reduce_or_sync(0xffffffffu, mine_or);
        mine_and = __reduce_and_sync(0xffffffffu, mine_and);
        if (lane == 0)
            warp_or[warp] = mine_or, warp_and[warp] = mine_and;
    }
    // a tile's flag byte -> number of 16-bit terms, their format (0 = fp16, 1 = bf16) and the plane
    // of the split buffer the first term lives in (planes 0..2 bf16 terms, plane 3 the fp16 copy)
    // (term t of a tile lives in plane plane0 + t*pstep: bf16 terms in planes 0, 1, 2; the fp16 copy in
    // plane 3; the second fp16 term of a kTileF16x2 tile in plane 1)
    auto tile_terms = [](uint32_t f, int &nterms, uint32_t &fmt, int &plane0, int &pstep) {
        fmt = (f >> 2) & 1u;                                             // kTileBf16
        nterms = fmt ? ((f & kTileTerm3) ? 3 : 1 + (int)(f & kTileTerm2)) : 1 + (int)((f >> 4) & 1u); // kTileTerm2 == 1
        plane0 = fmt ? 0 : kMaxSplits;
        pstep = fmt ? 1 : -2;
    };
    const uint32_t *sflags32 = reinterpret_cast<const uint32_t *>(sflags); // one word = the kSub tiles of a stage
    constexpr bool kSeq = NT >= 128;

    if (warp == kMmaW && lane == 0)
    {
        for (int s = 0; s < 4; ++s)
        {
            mbar_init(afull0 + 8 * s, EW);        // every expander warp arrives once per stage
            mbar_init(aempty0 + 8 * s, 1);        // one tcgen05.commit
        }
        if constexpr (!XK)
            for (int s = 0; s < 16; ++s)
            {
                mbar_init(bfull0 + 8 * s, 1);     // the producer's expect_tx arrive
                mbar_init(bempty0 + 8 * s, 1);    // one tcgen05.commit
            }
        mbar_init(tmem_full, 1);
        *huge_local = 0;
        fence_barrier_init();
    }
    else if (warp == kAllocW)
    {
        tmem_alloc(smem_u32(tmem_slot), kTmem);
    }
    else if (!XK && warp == kTmaW && lane == 0)
    {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
    }
    if constexpr (XK)
    {
        // rows of the X tiles at or beyond M are never written again: zero all tiles once (three
        // terms of NT rows x 128 B per sub-block, kSub sub-blocks per A stage, at most 3 stages)
        constexpr int kSx = (kTmem - kMaxSplits * NT) / (kSub * 32) > 3 ? 3 : (kTmem - kMaxSplits * NT) / (kSub * 32);
        for (int i = tid; i < kSx * kSub * kMaxSplits * NT * 128 / 16; i += kThreadsT)
            asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(xs0 + i * 16), "r"(0) : "memory");
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;
    if (tid == 0)
        TC_TRACE(1);

    // ---- geometry, from what this CTA's tiles hold ---------------------------------------------
    uint32_t fl_all = 0, fl_and = 0xFFFFFFFFu;
    if constexpr (!XK)
        for (int w = 0; w < EW + 4; ++w)
            fl_all |= warp_or[w], fl_and &= warp_and[w];
    fl_all |= fl_all >> 16, fl_all |= fl_all >> 8;
    fl_and &= fl_and >> 16, fl_and &= fl_and >> 8;
    // every tile of this CTA carries the same flag byte (all-integer X, all full-precision X, ...): the
    // producer and the MMA issuer then run loops with the decode hoisted out — the per-tile decode
    // on the issue path costs a feed-bound small-tile shape 15 % (c5a: 64 -> 75 us)
    const bool uniform = XK || ((fl_all ^ fl_and) & 0xFFu) == 0;
    // most terms any tile of this CTA needs (in-kernel conversion: always three)
    const int tmax = XK ? kMaxSplits
                        : ((fl_all & kTileBf16) && (fl_all & kTileTerm3)
                               ? 3
                               : ((((fl_all & kTileBf16) && (fl_all & kTileTerm2)) || (fl_all & kTileF16x2)) ? 2 : 1));
    // Accumulators.  NT <= 64: the split terms sit side by side (columns [t*NT, (t+1)*NT)) and one
    // wide MMA per 16-k step covers all terms of a tile — the A operand is fed once for all terms,
    // which is what bounds small tiles; the columns are added in the epilogue.  NT >= 128 (kSeq): one
    // set of NT columns, and the terms run TERM-MAJOR: a whole pass over K with the third terms,
    // then one with the second, the first terms last.  The tensor core aligns every product to the
    // accumulator and drops what falls below its last bit; interleaved, each small term would lose
    // its low bits against the already large sum at every one of its K/16 MMAs (a drift of ~1 ulp
    // per MMA: 1.2e-5 of max|Y| at c4).  Term-major, each pass adds 8-bit values to a sum of its own
    // magnitude — exact until the sum outgrows them — so the result carries a few ulp in total.
    // The MMA count is unchanged; W is expanded once per pass by warps that are otherwise idle.
    const int acc_cols = kSeq ? nt : tmax * nt;
    const int npass = kSeq ? tmax : 1;

    // TMEM: accumulators in columns [0, acc_cols), A stages of 128 columns at the top
    int S = (kTmem - acc_cols) / (kSub * 32);    // A stages in TMEM (>= 1)
    S = S > 3 ? 3 : S;
    const int a_col0 = kTmem - S * kSub * 32;
    // shared memory: X tiles.  In-kernel conversion: one set of kSub tiles per A stage (filled by
    // the expanders, published by the same barrier).  TMA: an independent ring — one TERM tile per
    // slot in term-major mode, all terms of one sub-block adjacent otherwise (one wide B operand)
    const int xtile = kSeq ? kBBytes : tmax * kBBytes;
    // K-splits of a tall tile meet in global memory (the leader's shared memory cannot hold peers of
    // nt = 256 rows next to the X ring): no landing zone then
    const bool gsplit = kSeq && !XK && p.gws != nullptr && p.ksplit > 1;
    const int park_bytes = gsplit ? 0 : (p.ksplit - 1) * nt * 512;
    int SB = XK ? S * kSub : (p.smem_budget - park_bytes) / xtile;
    SB = SB > 16 ? 16 : SB;

    if (!XK && warp == kTmaW)
    {
        // ===== TMA producer: X tiles of the split terms, one ring slot per sub-block =====
        if (elect_one())
        {
            uint32_t eb = bempty0, fb = bfull0, dst = xs0, ph = 0;
            int slot = 0;
            const int row = mtile * nt;
            auto advance = [&]() {
                eb += 8, fb += 8, dst += xtile;
                if (++slot == SB)
                    slot = 0, eb = bempty0, fb = bfull0, dst = xs0, ph ^= 1;
            };
            int nterms, plane0, pstep;
            uint32_t fmt;
            tile_terms(fl_all & 0xFFu, nterms, fmt, plane0, pstep); // stays as it is when the tiles are uniform
            if constexpr (kSeq)
            {
                for (int t = npass - 1; t >= 0; --t) // term-major: third terms first, first terms last
                {
                    int kcoord = st_lo * kSub * kBlockK;
                    uint32_t fw = 0;
                    for (int sb = 0; sb < iters * kSub; ++sb, kcoord += kBlockK, fw >>= 8)
                    {
                        if (!uniform)
                        {
                            if ((sb & (kSub - 1)) == 0)
                                fw = sflags32[sb / kSub];
                            tile_terms(fw & 0xFFu, nterms, fmt, plane0, pstep);
                        }
                        if (nterms <= t)
                            continue; // this tile has no such term
                        mbar_wait(eb, ph ^ 1);
                        mbar_arrive_expect_tx(fb, (uint32_t)kBBytes);
                        tma_load_2d(dst, &xmap, fb, kcoord, (plane0 + t * pstep) * p.Mp + row);
                        advance();
                    }
                }
            }
            else
            {
                // (straight-line, predicated loads: a run-time loop around the bulk-tensor copy costs
                // the feed-bound small-tile shapes 15 % — c5a 64 -> 76 us)
                int kcoord = st_lo * kSub * kBlockK;
                uint32_t fw = 0;
                int row_t = plane0 * p.Mp + row, row_s = pstep * p.Mp;
                for (int sb = 0; sb < iters * kSub; ++sb, kcoord += kBlockK, fw >>= 8)
                {
                    if (!uniform)
                    {
                        if ((sb & (kSub - 1)) == 0)
                            fw = sflags32[sb / kSub];
                        tile_terms(fw & 0xFFu, nterms, fmt, plane0, pstep);
                        row_t = plane0 * p.Mp + row, row_s = pstep * p.Mp;
                    }
                    mbar_wait(eb, ph ^ 1);
                    mbar_arrive_expect_tx(fb, (uint32_t)(nterms * kBBytes));
                    tma_load_2d(dst, &xmap, fb, kcoord, row_t);
                    if (nterms > 1)
                        tma_load_2d(dst + kBBytes, &xmap, fb, kcoord, row_t + row_s);
                    if (nterms > 2)
                        tma_load_2d(dst + 2 * kBBytes, &xmap, fb, kcoord, row_t + 2 * row_s);
                    eb += 8, fb += 8, dst += xtile;
                    if (++slot == SB)
                        slot = 0, eb = bempty0, fb = bfull0, dst = xs0, ph ^= 1;
                }
            }
        }
    }
    else if (warp == kMmaW)
    {
        // ===== MMA issuer: the whole warp runs the loop (uniform registers), one lane issues =====
        // NT <= 64: one MMA per 16-k step covers all split terms at once — their X tiles are adjacent
        // in smem (rows [t*NT, (t+1)*NT)), so B is simply nterms*NT rows tall and term t lands in
        // accumulator columns [t*NT, (t+1)*NT); the terms are added in the epilogue.
        // NT >= 128: one MMA per term and 16-k step, all into the same accumulator columns.
        // (the TMA path takes terms and format from the tile's flag byte; tiles in fp16 and in bf16
        // accumulate into the same columns — 0x4000 is 2.0 in both, so A is the same for both)
        constexpr uint32_t kBf16Bits = (1u << 7) | (1u << 10);
        const uint64_t bdesc0 = make_smem_desc(xs0);
        const uint64_t xstep = (uint64_t)(xtile >> 4);
        uint32_t started = 0; // kSeq: the very first MMA overwrites the accumulator
        uint64_t bdesc = bdesc0;
        uint32_t afb = afull0, aeb = aempty0, aph = 0, bfb = bfull0, beb = bempty0, bph = 0;
        uint32_t acol = tmem_d + (uint32_t)a_col0;
        int st = 0, slot = 0;
        auto advance_b = [&]() {
            bdesc += xstep, bfb += 8, beb += 8;
            if (++slot == SB)
                slot = 0, bdesc = bdesc0, bfb = bfull0, beb = bempty0, bph ^= 1;
        };
        // instruction descriptors for 1, 2, 3 terms' worth of accumulator columns (side by side) —
        // the tile's flag byte only selects among them and ORs the B format in
        const uint32_t idesc1 = make_idesc(nt), idesc2 = make_idesc(kSeq ? nt : 2 * nt), idesc3 = make_idesc(kSeq ? nt : 3 * nt);
        auto issue_all = [&](auto uni_tag) {
        constexpr bool kUni = decltype(uni_tag)::value; // decode hoisted: one flag byte for all tiles
        int nterms = kMaxSplits, plane0 = 0, pstep = 1;
        uint32_t fmt = 1;
        if constexpr (!XK)
            tile_terms(fl_all & 0xFFu, nterms, fmt, plane0, pstep);
        uint32_t fbits = fmt ? kBf16Bits : 0u;
        uint32_t idesc_w = (nterms == 1 ? idesc1 : (nterms == 2 ? idesc2 : idesc3)) | fbits;
        for (int pass = npass - 1; pass >= 0; --pass) // kSeq: term-major (npass == 1 otherwise)
            for (int it = 0; it < iters; ++it)
            {
                uint32_t fw = 0;
                if constexpr (!kUni)
                    fw = sflags32[it]; // the stage's four flag bytes, fetched before the wait below
                mbar_wait(afb, aph);
                tc_fence_after();
                if (it == 0 && lane == 0)
                    TC_TRACE(5);
#pragma unroll
                for (int u = 0; u < kSub; ++u)
                {
                    if constexpr (!kUni)
                    {
                        tile_terms((fw >> (8 * u)) & 0xFFu, nterms, fmt, plane0, pstep);
                        fbits = fmt ? kBf16Bits : 0u;
                        idesc_w = (nterms == 1 ? idesc1 : (nterms == 2 ? idesc2 : idesc3)) | fbits;
                    }
                    if constexpr (kSeq)
                    {
                        // term `pass` of this tile, if it has one: its own ring slot, the same nt columns
                        if (nterms <= pass)
                            continue;
                        const uint32_t idesc = idesc1 | fbits;
                        mbar_wait(bfb, bph);
                        tc_fence_after();
                        if (elect_one())
                        {
#pragma unroll
                            for (int k = 0; k < kBlockK / 16; ++k) // 16 k = 8 TMEM columns of A = 32 B of each X row
                                umma_f16_ts(tmem_d, acol + u * 32 + k * 8, bdesc + 2 * k, idesc, started | (uint32_t)k);
                            umma_commit(beb); // frees the term tile when these MMAs retire
                        }
                        __syncwarp();
                        started = 1;
                        advance_b();
                    }
                    else
                    {
                        // one wide B operand: the nterms term tiles of the sub-block are adjacent, term t
                        // lands in accumulator columns [t*nt, (t+1)*nt).  TMA path: the expanders zeroed
                        // the column groups in use, so every MMA accumulates.
                        if constexpr (!XK)
                        {
                            mbar_wait(bfb, bph);
                            tc_fence_after();
                        }
                        const uint32_t idesc = idesc_w;
                        if (elect_one())
                        {
#pragma unroll
                            for (int k = 0; k < kBlockK / 16; ++k)
                                umma_f16_ts(tmem_d, acol + u * 32 + k * 8, bdesc + 2 * k, idesc,
                                            XK ? (uint32_t)((it | u | k) != 0) : 1u);
                            if constexpr (!XK)
                                umma_commit(beb); // frees the X tile when these MMAs retire
                        }
                        __syncwarp();
                        advance_b();
                    }
                }
                if (elect_one())
                    umma_commit(aeb); // frees the A stage (and, in-kernel conversion, its X tiles)
                __syncwarp();
                afb += 8, aeb += 8, acol += kSub * 32;
                if (++st == S)
                    st = 0, afb = afull0, aeb = aempty0, acol = tmem_d + (uint32_t)a_col0, aph ^= 1;
            }
        };
        if (uniform)
            issue_all(std::true_type{});
        else
            issue_all(std::false_type{});
        if (elect_one())
            umma_commit(tmem_full);
        __syncwarp();
        if (lane == 0)
            TC_TRACE(6);
    }
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    if (warp < EW)
    {
        // ===== expanders: tile-packed codes -> 16-bit A sub-blocks in TMEM =====
        int st = 0;
        uint32_t ph = 0;
        if constexpr (!XK && !kSeq)
        {
            // side-by-side accumulators: tiles differ in how many of the three column groups their
            // MMA touches, so all groups start from zero here and every MMA accumulates.  Published
            // to the MMA issuer by the first stage's tcgen05.wait::st + barrier arrive below.
            const uint32_t zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int ch = grp; ch < acc_cols / 16; ch += G)
                tmem_st16(tmem_d + lane_base + (uint32_t)(ch * 16), zero, zero);
        }
        // term-major mode walks K once per term (npass passes); the code ring simply wraps around
        const int total = npass * iters;
        auto stage_of = [&](int j) { // stage j of the whole walk -> stage inside this CTA's K range
            if (j >= iters) j -= iters;
            if (j >= iters) j -= iters;
            return j;
        };
        if (npass > 1)
        {
            // the prologue filled the ring for the first pass only
#pragma unroll
            for (int i = 0; i < kRing; ++i)
                if (i >= iters && i < total)
                {
#pragma unroll
                    for (int u = 0; u < kMine; ++u)
                        ring[i][u] = ldg_v4_ordered(src + ((size_t)stage_of(i) * kSub + u * G) * 128);
                }
        }
        for (int j = 0; j < total; ++j)
        {
            const int it = stage_of(j);
            uint4 cur[kMine];
#pragma unroll
            for (int u = 0; u < kMine; ++u)
                cur[u] = ring[0][u];
#pragma unroll
            for (int i = 0; i + 1 < kRing; ++i)
#pragma unroll
                for (int u = 0; u < kMine; ++u)
                    ring[i][u] = ring[i + 1][u];
            if (j + kPrefetch < iters) // first pass only: later passes find the codes in L2
#pragma unroll
                for (int u = 0; u < kMine; ++u)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(src + ((size_t)(j + kPrefetch) * kSub + u * G) * 128));
            if (j + kRing < total) // codes kRing stages ahead, in flight during this expansion
#pragma unroll
                for (int u = 0; u < kMine; ++u)
                    ring[kRing - 1][u] = ldg_v4_ordered(src + ((size_t)stage_of(j + kRing) * kSub + u * G) * 128);
            uint32_t xt[kMine][kPairs][3];
            if constexpr (XK)
            {
#pragma unroll
                for (int u = 0; u < kMine; ++u)
#pragma unroll
                    for (int j = 0; j < kPairs; ++j)
                    {
                        // (tested where the values are consumed: a test next to the loads would wait for them)
                        xhuge = max(xhuge, max(__float_as_uint(xv[u][j].x) & 0x7FFFFFFFu, __float_as_uint(xv[u][j].y) & 0x7FFFFFFFu));
                        split3_pair(xv[u][j].x, xv[u][j].y, xt[u][j][0], xt[u][j][1], xt[u][j][2]);
                    }
                if (it + 1 < iters)
                    load_x(it + 1);
            }
            mbar_wait(aempty0 + 8 * st, ph ^ 1);
            tc_fence_after();
#pragma unroll
            for (int u = 0; u < kMine; ++u)
            {
                const int sub = grp + u * G;
                uint32_t ra[8], rb[8];
                const uint32_t ta = tmem_d + lane_base + (uint32_t)(a_col0 + (st * kSub + sub) * 32);
                // 32 columns: word w -> columns 8w .. 8w+7 (pairs of consecutive k)
                expand_word(cur[u].x, ra), expand_word(cur[u].y, rb);
                tmem_st16(ta, ra, rb);
                expand_word(cur[u].z, ra), expand_word(cur[u].w, rb);
                tmem_st16(ta + 16, ra, rb);
                if constexpr (XK)
                {
                    const uint32_t xb = xs0 + (uint32_t)((st * kSub + sub) * xtile);
#pragma unroll
                    for (int j = 0; j < kPairs; ++j)
                    {
                        const int ml = q + 4 * j;
                        if (mtile * nt + ml < p.M)
                        {
                            const uint32_t a = xb + ml * 128 + (((lane >> 2) ^ (ml & 7)) << 4) + (lane & 3) * 4;
#pragma unroll
                            for (int t = 0; t < 3; ++t)
                                asm volatile("st.shared.b32 [%0], %1;" ::"r"(a + t * kBBytes), "r"(xt[u][j][t]) : "memory");
                        }
                    }
                }
            }
            if (j == 0 && tid == 0)
                TC_TRACE(2);
            if constexpr (XK)
                fence_proxy_async(); // generic-proxy smem writes -> visible to the tensor core
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0)
                mbar_arrive(afull0 + 8 * st);
            if (j == 0 && tid == 0)
                TC_TRACE(3);
            if (++st == S)
                st = 0, ph ^= 1;
        }
        if (tid == 0)
            TC_TRACE(4);
        if constexpr (XK)
        {
            // non-finite / huge X seen by any expander thread -> CTA-wide verdict
            if (__any_sync(0xffffffffu, xhuge >= TSG_X_HUGE_BITS) && lane == 0)
                atomicOr(huge_local, 1u);
            asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
        }
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        if (tid == 0)
            TC_TRACE(7);
    }

    // the main loop is over (expanders) or was never ours (role warps): let the next kernel of the
    // stream begin launching while the epilogue runs
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    // ===== epilogue: TMEM -> registers 16 columns (rows of X) at a time =====
    // Split-K across the cluster (ranks = K-splits): peers PUSH their accumulators into the
    // leader's shared memory (st.shared::cluster is fire and forget: no DSMEM round trips), one
    // release/acquire cluster barrier publishes them, the leader adds them in rank order
    // (deterministic), applies bias / PReLU and writes Y — no partial sums in HBM, no second
    // kernel.  The landing zone lies behind the X tiles.  Nothing is held in registers across the
    // barrier: the leader reads its own accumulators from TMEM afterwards.
    const int kChunks = nt / 16;
    const uint32_t crank = (p.ksplit > 1) ? blockIdx.z : 0;
    // (at the END of the budget: the ring in front of it is sized from this rank's own tiles, and
    // the ranks of a cluster must agree on where the leader's landing zone lies)
    float *park = reinterpret_cast<float *>(smem_al + kBarBytes + kFlagBytes + ((p.smem_budget - park_bytes) & ~127)); // [rank-1][NT][128]
    // what this CTA's K range held: column groups in use (side-by-side accumulators) and whether X
    // had a value the dense product cannot take (non-finite or >= 2^100)
    const int acc_terms = tmax;
    uint32_t huge = 0;
    if (warp < EW)
    {
        if constexpr (XK)
            huge = *reinterpret_cast<volatile uint32_t *>(huge_local);
        else
            huge = (fl_all & kTileHuge) ? 1u : 0u;
    }
    auto load_chunk = [&](int ch, uint32_t (&acc)[16]) {
        tmem_ld16(tmem_d + lane_base + (uint32_t)(ch * 16), acc);
        for (int t = 1; t < (kSeq ? 1 : acc_terms); ++t) // x1 + x2 + x3 terms, fixed order
        {
            uint32_t more[16];
            tmem_ld16(tmem_d + lane_base + (uint32_t)(t * nt + ch * 16), more);
#pragma unroll
            for (int c = 0; c < 16; ++c)
                acc[c] = __float_as_uint(__uint_as_float(acc[c]) + __uint_as_float(more[c]));
        }
    };
    if (p.ksplit > 1)
    {
        if (gsplit)
        {
            // Split-K through global memory (the tail launch of a tall-tile grid, tsg_launch_dense_tc): every
            // rank writes its partial tile to the workspace (it stays in L2), one release/acquire cluster
            // barrier publishes them, and every rank then finishes 1/ksplit of the tile's rows — the
            // reduction and the epilogue are spread over the cluster's SMs.
            if (warp < EW)
            {
                float *ws = p.gws + ((size_t)blockIdx.x * p.ksplit + crank) * (size_t)nt * 128 + erow;
#pragma unroll 1
                for (int ch = grp; ch < kChunks; ch += G)
                {
                    if (mtile * nt + ch * 16 >= p.M)
                        break; // rows beyond M are never read
                    uint32_t acc[16];
                    load_chunk(ch, acc);
#pragma unroll
                    for (int c = 0; c < 16; ++c)
                        __stcg(ws + (size_t)(ch * 16 + c) * 128, __uint_as_float(acc[c]));
                }
            }
            if (tid == 0) // every rank needs every rank's verdict
                for (int r = 0; r < p.ksplit; ++r)
                    st_dsmem_u32(mapa_rank(smem_u32(huge_ranks + crank), (uint32_t)r), huge);
        }
        else if (warp < EW && crank != 0)
        {
            const uint32_t remote = mapa_rank(smem_u32(park + (size_t)(crank - 1) * nt * 128 + erow), 0);
#pragma unroll 1
            for (int ch = grp; ch < kChunks; ch += G)
            {
                if (mtile * nt + ch * 16 >= p.M)
                    break; // rows beyond M are never read
                uint32_t acc[16];
                load_chunk(ch, acc);
#pragma unroll
                for (int c = 0; c < 16; ++c)
                    st_dsmem_u32(remote + (uint32_t)((ch * 16 + c) * 512), acc[c]);
            }
        }
        if (!gsplit && tid == 0) // every rank reports (no initialisation of the leader's words needed)
            st_dsmem_u32(mapa_rank(smem_u32(huge_ranks + crank), 0), huge);
        cluster_sync_all();
        if (tid == 0)
            TC_TRACE(10);
        if (warp < EW && (crank == 0 || gsplit))
        {
#pragma unroll
            for (int r = 0; r < 8; ++r) // the cluster barrier above acquired the peers' words
                if (r < p.ksplit && r != (int)crank)
                    huge |= huge_ranks[r];
        }
    }
    if (warp < EW && crank == 0 && huge)
    {
        // X of this tile holds inf / NaN / |x| >= 2^100: 0·x and 2·x are not what the reference's
        // sparse sum computes (comp.h:44-61 never touches x where W is 0).  Recompute the tile in
        // the reference's own order from the TCSC arrays — slow, exact, and only ever for such input.
        const int en = n0 + erow;
        if (en < p.N)
            for (int ch = grp; ch < kChunks; ch += G)
                for (int c = 0; c < 16; ++c)
                {
                    const int mrow = mtile * nt + ch * 16 + c;
                    if (mrow >= p.M)
                        break;
                    float y = tsg_ref_order_sum(p.X + (int64_t)mrow * p.ldx, 1, p.csp, p.csn, p.rip, p.rin, en, bn);
                    if (p.alpha)
                        y = (y > 0.0f) ? y : an * y;
                    p.Y[(int64_t)mrow * p.ldy + en] = y;
                }
    }
    else if (gsplit)
    {
        // this rank's share of the tile: rows in units of 4, unit u belongs to rank u % ksplit; the
        // partial sums are added in rank order (deterministic, the same order as the shared-memory path)
        if (warp < EW && !huge)
        {
            const int en = n0 + erow;
            const float *wsr = p.gws + (size_t)blockIdx.x * p.ksplit * (size_t)nt * 128 + erow;
#pragma unroll 1
            for (int u = (int)crank + grp * p.ksplit; u < nt / 4; u += G * p.ksplit)
            {
                const int rows = p.M - (mtile * nt + u * 4);
                if (rows <= 0)
                    break;
                float v[8][4]; // all loads in flight before the first add (ksplit <= 8: the portable cluster size)
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        v[r][c] = r < p.ksplit ? __ldcg(wsr + ((size_t)r * nt + u * 4 + c) * 128) : 0.0f;
                float s[4];
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    s[c] = v[0][c];
#pragma unroll
                for (int r = 1; r < 8; ++r)
                    if (r < p.ksplit)
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            s[c] += v[r][c];
                if (en < p.N)
                {
                    float *yp = p.Y + (int64_t)(mtile * nt + u * 4) * p.ldy + en;
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                    {
                        float y = 0.5f * s[c] + bn;
                        if (p.alpha)
                            y = (y > 0.0f) ? y : an * y;
                        if (c < rows)
                            yp[(int64_t)c * p.ldy] = y;
                    }
                }
            }
        }
    }
    else if (warp < EW && crank == 0)
    {
        const int en = n0 + erow;
#pragma unroll 1
        for (int ch = grp; ch < kChunks; ch += G)
        {
            const int rows = p.M - (mtile * nt + ch * 16);
            if (rows <= 0)
                break;
            uint32_t acc[16];
            load_chunk(ch, acc);
#pragma unroll 1
            for (int r = 1; r < p.ksplit; ++r) // rank order: deterministic
            {
                const float *pr = park + ((size_t)(r - 1) * nt + ch * 16) * 128 + erow;
#pragma unroll
                for (int c = 0; c < 16; ++c)
                    acc[c] = __float_as_uint(__uint_as_float(acc[c]) + pr[c * 128]);
            }
            if (en < p.N)
            {
                float *yp = p.Y + (int64_t)(mtile * nt + ch * 16) * p.ldy + en;
#pragma unroll
                for (int c = 0; c < 16; ++c)
                {
                    float y = 0.5f * __uint_as_float(acc[c]) + bn; // the A tile holds 2·W (exact power-of-two scaling)
                    if (p.alpha)
                        y = (y > 0.0f) ? y : an * y;
                    if (c < rows)
                        yp[(int64_t)c * p.ldy] = y;
                }
            }
        }
    }
    if (tid == 0)
        TC_TRACE(11);
    tc_fence_before();
    __syncthreads();
    if (warp == kAllocW)
        tmem_dealloc(tmem_d, kTmem);
    if (tid == 0)
        TC_TRACE(12);
}

// fp32 X -> 16-bit operand tiles for the TMA path, one CTA per tile (nt rows of one m-tile x 64 k).
// The buffer holds four planes of [Mp][Kp] 16-bit values — bf16 terms x1, x2, x3 (x == x1+x2+x3
// exactly) and one fp16 copy — but a tile writes only what its own values need:
//   every x exact in fp16 (integers up to 2048, fp16-born activations) -> the fp16 plane only;
//   otherwise bf16 term 1, term 2 if some x is not exact in bf16, term 3 if some x has more than
//   16 significant bits.
// The decision is taken on the fp32 bit patterns (no conversions): low 13 mantissa bits and the
// exponent range for fp16, low 16 / low 8 mantissa bits for the bf16 terms (low 8 bits clear means
// the remainder after term 1 is a multiple of 2^8 ulp below 2^16 ulp: exact in bf16) — conservative
// where it is not sharp, which costs a zero term, never accuracy.  One flag byte per tile
// (kTile* bits) tells the dense kernel what was written.  Traffic: 4 B read + 2 B written per
// element for the reference's integer-valued X (was 4 + 8), 4 + 6 for full-precision fp32.
// A tile holding a non-finite or >= 2^100 value is flagged (kTileHuge): the dense kernel then
// recomputes that m-tile's outputs in the reference's order.
__global__ void __launch_bounds__(512, 2)
split_tiles_kernel(const float *__restrict__ X, int64_t ldx, int M, int K, int nt, int Mp, int Kp, int nkb,
                   uint16_t *__restrict__ out, uint8_t *__restrict__ tflags, int exact)
{
    __shared__ uint32_t s_or, s_bad, s_max;
    // programmatic dependent launch on both sides: the kernel in front (the previous call's dense
    // kernel, still reading the buffer we are about to overwrite) must have completed; the dense
    // kernel behind us may start its prologue (weight-stream prefetch, nothing of ours)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int kb = blockIdx.x, mtile = blockIdx.y, tid = threadIdx.x;
    if (kb * kBlockK >= Kp) // zero padding of the code stream: the TMA box is out of bounds = zeros
    {
        if (tid == 0)
            tflags[(size_t)mtile * nkb + kb] = 0;
        return;
    }
    if (tid == 0)
        s_or = 0, s_bad = 0, s_max = 0;
    __syncthreads();
    // 512 threads: 16 threads x 4 k per row, 32 rows per pass, up to 8 passes (nt <= 256) — eight
    // 128-bit loads per thread in flight, 32 warps per SM at two CTAs (the 256-thread version with
    // 16 loads per thread ran at 16 warps per SM and 45 % issue utilisation)
    const int c4 = tid & 15, r = tid >> 4;
    const int k = kb * kBlockK + c4 * 4;
    // all loads of the tile first, the tests after
    const bool vec = ((reinterpret_cast<uintptr_t>(X) | (uintptr_t)(ldx * 4)) & 15) == 0 &&
                     kb * kBlockK + kBlockK <= K; // block-uniform
    float4 v[8];
    if (vec)
    {
#pragma unroll
        for (int i = 0; i < 8; ++i)
        {
            const int m = mtile * nt + r + 32 * i;
            v[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (r + 32 * i < nt && m < M)
                asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                             : "=f"(v[i].x), "=f"(v[i].y), "=f"(v[i].z), "=f"(v[i].w)
                             : "l"(X + (int64_t)m * ldx + k));
        }
    }
    else
    {
#pragma unroll
        for (int i = 0; i < 8; ++i)
        {
            const int m = mtile * nt + r + 32 * i;
            v[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (r + 32 * i < nt && m < M)
            {
                const float *xp = X + (int64_t)m * ldx + k;
                if (k < K) v[i].x = __ldg(xp);
                if (k + 1 < K) v[i].y = __ldg(xp + 1);
                if (k + 2 < K) v[i].z = __ldg(xp + 2);
                if (k + 3 < K) v[i].w = __ldg(xp + 3);
            }
        }
    }
    uint32_t orbits = 0, bad = 0, amax = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
    {
        const uint32_t u[4] = {__float_as_uint(v[i].x), __float_as_uint(v[i].y), __float_as_uint(v[i].z),
                               __float_as_uint(v[i].w)};
#pragma unroll
        for (int j = 0; j < 4; ++j) // (rows beyond nt hold zeros: neutral)
        {
            const uint32_t a = u[j] & 0x7FFFFFFFu;
            orbits |= u[j];
            amax = max(amax, a);
            // fp16 holds 2^-14 <= |x| < 65536 with the low 13 mantissa bits clear (and zero)
            bad |= (a != 0u && (a - 0x38800000u) >= 0x0F000000u) ? 1u : 0u;
        }
    }
    orbits = __reduce_or_sync(0xffffffffu, orbits & 0xFFFFu);
    bad = __reduce_or_sync(0xffffffffu, bad);
    amax = __reduce_max_sync(0xffffffffu, amax);
    if ((tid & 31) == 0)
    {
        if (orbits)
            atomicOr(&s_or, orbits);
        if (bad)
            atomicOr(&s_bad, bad);
        atomicMax(&s_max, amax);
    }
    __syncthreads();
    const uint32_t o = s_or, b = s_bad, mx = s_max;
    // Which operand the tile becomes:
    //   every x exact in fp16                          -> ONE fp16 term (flag 0);
    //   16 significant bits suffice (low 8 bits clear) -> one or two bf16 terms, exact;
    //   full-precision values -> three bf16 terms, exact — unless the caller opted into the fast
    //   split (`exact` == 0: tsg_set_fast_split(1) / TSG_TC_FAST=1) and the tile's largest magnitude
    //   lies inside [2^-4, 65520) (fp16's range with room for the remainder): then TWO fp16 terms,
    //   x1 = fp16(x), x2 = fp16(x - x1): x is carried with |error| <= max(2^-24 |x|, 2^-25), two
    //   thirds of the tensor work (include/tsg.h states the contract).
    const bool f16_exact = !((o & 0x1FFFu) || (b & 1u));
    const bool need3 = (o & 0xFFu) != 0u;
    const bool huge = mx >= TSG_X_HUGE_BITS;
    const bool f16x2 = !f16_exact && need3 && !exact && mx >= 0x3D800000u /* 2^-4 */ && mx < 0x477FF000u /* 65520 */;
    const uint32_t flag = f16_exact ? 0u
                          : (f16x2 ? kTileF16x2
                                   : (kTileBf16 | ((o & 0xFFFFu) ? kTileTerm2 : 0u) | (need3 ? kTileTerm3 : 0u))) |
                                (huge ? kTileHuge : 0u);
    if (tid == 0)
        tflags[(size_t)mtile * nkb + kb] = (uint8_t)flag;
    const size_t plane = (size_t)Mp * Kp;
#pragma unroll
    for (int i = 0; i < 8; ++i)
    {
        if (r + 32 * i < nt)
        {
            uint16_t *dst = out + (size_t)(mtile * nt + r + 32 * i) * Kp + k;
            if (!(flag & kTileBf16))
            {
                uint32_t h0, h1;
                asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h0) : "f"(v[i].y), "f"(v[i].x));
                asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h1) : "f"(v[i].w), "f"(v[i].z));
                *reinterpret_cast<uint2 *>(dst + kMaxSplits * plane) = make_uint2(h0, h1);
                if (flag & kTileF16x2)
                {
                    // the remainder after the first fp16 term, again in fp16, into plane 1's storage
                    const __half2 a0 = *reinterpret_cast<const __half2 *>(&h0), a1 = *reinterpret_cast<const __half2 *>(&h1);
                    const float2 f0 = __half22float2(a0), f1 = __half22float2(a1);
                    uint32_t r0, r1;
                    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r0) : "f"(v[i].y - f0.y), "f"(v[i].x - f0.x));
                    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r1) : "f"(v[i].w - f1.y), "f"(v[i].z - f1.x));
                    *reinterpret_cast<uint2 *>(dst + plane) = make_uint2(r0, r1);
                }
            }
            else
            {
                uint32_t a1, a2, a3, b1, b2, b3;
                split3_pair(v[i].x, v[i].y, a1, a2, a3);
                split3_pair(v[i].z, v[i].w, b1, b2, b3);
                *reinterpret_cast<uint2 *>(dst) = make_uint2(a1, b1);
                if (flag & (kTileTerm2 | kTileTerm3))
                    *reinterpret_cast<uint2 *>(dst + plane) = make_uint2(a2, b2);
                if (flag & kTileTerm3)
                    *reinterpret_cast<uint2 *>(dst + 2 * plane) = make_uint2(a3, b3);
            }
        }
    }
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

// programmatic dependent launch of the dense kernel (behind split_tiles_kernel: 1-5 % on the small and
// mid-sized shapes; behind the previous call on the in-kernel-conversion path); TSG_TC_PDL=0 turns it off
static const bool g_pdl = !(getenv("TSG_TC_PDL") && getenv("TSG_TC_PDL")[0] == '0');

template <int NT, bool XK, int EW>
int launch_nt(const CUtensorMap &map, const DenseParams &p, dim3 grid, size_t smem, int device, cudaStream_t st)
{
    // largest opt-in granted so far per device (the attribute is per device and function); atomic: two host
    // threads may launch the same kernel — a repeated, equal cudaFuncSetAttribute is harmless, a torn size is not
    static std::atomic<size_t> configured[64];
    std::atomic<size_t> &have = configured[device & 63];
    if (have.load(std::memory_order_acquire) < smem)
    {
        TSG_CUDA(cudaFuncSetAttribute(dense_tc_kernel<NT, XK, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        size_t seen = have.load(std::memory_order_relaxed);
        while (seen < smem && !have.compare_exchange_weak(seen, smem, std::memory_order_release))
            ;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3((EW + 4) * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = grid.z; // the K-splits of a tile form one cluster
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; // see griddepcontrol.wait in the kernel
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_pdl ? 2 : 1;
    TSG_CUDA(cudaLaunchKernelEx(&cfg, dense_tc_kernel<NT, XK, EW>, map, p));
    TSG_LAUNCHED();
    return TSG_OK;
}

// K-split: the factor with the smallest estimated makespan — whole waves of CTAs, each costing its
// share of the stages plus a fixed part (prologue, first HBM latency, epilogue: ~5k clk; a cluster
// reduction adds ~1.5k).  Filling the last wave is not worth it when the fixed part dominates.
int choose_ksplit(long long tiles, int nst, int slots, int cap, double stage_clk = 1230.0)
{
    int max_split = nst > 8 ? 8 : (nst > 0 ? nst : 1); // >= 1 stage (256 k) per CTA; portable cluster size
    if (max_split > cap)
        max_split = cap; // the leader's landing zone for the peers' accumulators must fit in smem
    int ksplit = 1;
    double best = 1e300;
    for (int ks = 1; ks <= max_split; ++ks)
    {
        const long long ctas = tiles * ks, waves = (ctas + slots - 1) / slots;
        const double per_cta = (double)((nst + ks - 1) / ks) * stage_clk + 5000.0 + (ks > 1 ? 1500.0 : 0.0);
        const double t = (double)waves * per_cta;
        if (t < best * 0.97) // prefer the smaller split unless clearly better
            best = t, ksplit = ks;
    }
    return ksplit;
}

// A grid of tall tiles (one CTA per SM) that does not fill its last wave leaves most SMs idle for a
// whole tile's time (c4: 1792 tiles = 12 waves + 16 tiles).  The tiles of that partial wave are
// launched separately, each K-split over a cluster of ks = min(8, SMs / tiles) CTAs that meet in
// global memory (DenseParams::gws): the last wave then takes 1/ks of a tile's time.  A grid smaller
// than one wave is all tail.  TSG_TC_TAIL=0 turns it off, TSG_TC_TAIL=k forces ks <= k.
struct TailPlan
{
    long long tiles;
    int ks;
};
TailPlan plan_tail(long long tiles, int nst, int sms, double stage_clk)
{
    static const int knob = getenv("TSG_TC_TAIL") ? atoi(getenv("TSG_TC_TAIL")) : -1;
    const long long r = tiles % sms;
    if (r == 0 || knob == 0 || knob == 1)
        return {0, 1};
    int ks = (int)(sms / r);
    ks = ks > 8 ? 8 : ks;         // portable cluster size
    ks = ks > nst ? nst : ks;     // >= 1 stage per CTA
    if (knob > 1 && ks > knob)
        ks = knob;
    if (ks < 2)
        return {0, 1};
    // worth it when the stages saved outweigh the trip through L2 and the second launch (~8k clk)
    if (knob < 0 && (double)(nst - (nst + ks - 1) / ks) * stage_clk < 8000.0)
        return {0, 1};
    return {r, ks};
}

// TSG_TC_TRACE=1: 16 SM-clock stamps per CTA of the last dense_tc launch (developer tool)
unsigned long long *g_tc_trace = nullptr;
size_t g_tc_trace_cap = 0, g_tc_trace_ctas = 0;
unsigned long long *tc_trace_buffer(size_t ctas)
{
    static int enabled = -1;
    if (enabled < 0)
    {
        const char *e = getenv("TSG_TC_TRACE");
        enabled = (e && e[0] == '1') ? 1 : 0;
    }
    if (!enabled)
        return nullptr;
    if (g_tc_trace_cap < ctas)
    {
        if (g_tc_trace)
            cudaFree(g_tc_trace);
        g_tc_trace = nullptr;
        g_tc_trace_cap = 0;
        if (cudaMalloc(&g_tc_trace, ctas * 16 * sizeof(unsigned long long)) != cudaSuccess)
            return nullptr;
        g_tc_trace_cap = ctas;
    }
    g_tc_trace_ctas = ctas;
    return g_tc_trace;
}

} // namespace

// the tail plan for a grid of `tiles` tall tiles with `nst` stages each on `sms` SMs (host logic only:
// tests/test_abi.py checks it without a GPU); returns the tiles of the tail launch, *ks its K-split
extern "C" int tsg_debug_plan_tail(long long tiles, int nst, int sms, int nt, int *ks)
{
    const TailPlan t = plan_tail(tiles, nst, sms, 16.0 * (nt / 2 > 77 ? nt / 2 : 77));
    if (ks)
        *ks = t.ks;
    return (int)t.tiles;
}

extern "C" int tsg_debug_tc_trace(unsigned long long *out, int max_ctas)
{
    if (!g_tc_trace || !out)
        return 0;
    const int n = (size_t)max_ctas < g_tc_trace_ctas ? max_ctas : (int)g_tc_trace_ctas;
    cudaDeviceSynchronize();
    cudaMemcpy(out, g_tc_trace, (size_t)n * 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    return n;
}

int tsg_launch_dense_tc(tsg_matrix *m, const float *X, int64_t ldx, const float *b,
                        const float *alpha, float *Y, int64_t ldy, int M, cudaStream_t st)
{
    if (M <= 0 || m->N == 0)
        return TSG_OK;
    const int K = m->K, N = m->N;
    const int Kp = (K + kBlockK - 1) / kBlockK * kBlockK;
    TSG_CHECK(Kp > 0, TSG_ERR_UNSUPPORTED, "dense_tc: K == 0");
    TSG_CHECK(m->codes != nullptr && m->code_kblocks >= Kp / kBlockK && m->code_kblocks % kSub == 0,
              TSG_ERR_UNSUPPORTED, "dense_tc: tile codes missing");
    const int nkb = m->code_kblocks; // padded to whole stages by the builder (zero codes)
    const int ntiles = (N + kTileN - 1) / kTileN;
    const int sms = m->sm_count > 0 ? m->sm_count : 148;

    // Small M: X is converted inside the kernel, 16 rows per m-tile, ONE launch.  Each extra
    // m-tile repeats the pass over W (~0.29 ps per matrix element); worth it while that
    // stays below the ~3 µs the two extra launches of the TMA path cost.
    const int mt16 = (M + 15) / 16;
    const double pass_us = 0.29e-6 * (double)K * (double)N; // one pass over the code stream (measured)
    const bool xk = M <= 16 || (M <= 64 && (mt16 - 1) * pass_us < 3.0);

    // shared memory per CTA: everything for the one-CTA-per-SM variant, half of the SM for the
    // two-CTA variant (the kernel sizes its X-tile ring from the budget at run time)
    const size_t smem_full = m->smem_optin;
    const size_t smem_half = ((m->smem_optin + 1024) / 2 - 1024) & ~(size_t)1023;
    int force_ew = 0;
    if (const char *e = getenv("TSG_TC_EW")) // developer override for tuning: 8 or 16
        force_ew = atoi(e);

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
    p.csp = m->csp, p.csn = m->csn, p.rip = m->rip, p.rin = m->rin;
    p.ntiles = ntiles;
    p.tile0 = 0;
    p.gws = nullptr;
    TSG_CHECK((long long)ntiles * ((M + 15) / 16) < (1ll << 31), TSG_ERR_UNSUPPORTED, "dense_tc: grid too large");
    CUtensorMap map = {};
    auto budget = [](size_t smem) { return (int)(smem - 1024 - kBarBytes - kFlagBytes); };

    if (xk)
    {
        // two CTAs per SM pay off once the grid runs to several waves (measured: c5a-sized 119 ->
        // 84 µs); a single wave is faster with all 16 expander warps in one CTA (c1 6.3 vs 7.7 µs)
        const bool half = force_ew ? force_ew == 8 : (long long)ntiles * mt16 >= 2ll * sms;
        p.smem_budget = budget(half ? smem_half : smem_full);
        p.ksplit = choose_ksplit((long long)ntiles * mt16, nkb / kSub, half ? 2 * sms : sms, 8);
        dim3 grid((unsigned)(ntiles * mt16), 1, p.ksplit);
        p.trace = tc_trace_buffer((size_t)ntiles * mt16 * p.ksplit);
        return half ? launch_nt<16, true, 8>(map, p, grid, smem_half, m->device, st)
                    : launch_nt<16, true, 16>(map, p, grid, smem_full, m->device, st);
    }

    EncodeTiledFn encode = get_encode();
    TSG_CHECK(encode != nullptr, TSG_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    // Tile height (rows of X per CTA).  Up to 64 rows: one tile.  Above: any multiple of 16 up to
    // 256 (the 256 instantiation takes its height at run time) — the candidate with the
    // smallest estimated makespan — whole waves x (stages x 16 MMAs x max(A-operand feed ~64 clk,
    // math NT/2 clk) + ~7k clk of fixed per-CTA cost), estimated for one split term.
    int NT = M <= 32 ? 32 : 64;
    if (M > 64)
    {
        double best = 1e300;
        for (int nt = 64; nt <= 256; nt += 16)
        {
            const int mt = (M + nt - 1) / nt;
            const double stage = 16.0 * (nt / 2 > 77 ? nt / 2 : 77); // measured: 77 clk per MMA when feed-bound
            const int nst = nkb / kSub;
            const int ks = choose_ksplit((long long)ntiles * mt, nst, sms, 1 + budget(smem_full) / 2 / (nt * 512), stage);
            const long long ctas = (long long)ntiles * mt * ks, waves = (ctas + sms - 1) / sms;
            double t = (double)waves * ((double)nst / ks * stage + 7000.0 + (ks > 1 ? 3000.0 : 0.0));
            if (ks == 1 && nt >= 80) // a partial last wave of tall tiles is K-split (plan_tail)
            {
                const TailPlan tl = plan_tail((long long)ntiles * mt, nst, sms, stage);
                if (tl.tiles)
                    t = (double)(ctas / sms) * ((double)nst * stage + 7000.0) + (double)((nst + tl.ks - 1) / tl.ks) * stage + 11000.0;
            }
            if (t < best * 0.999) // ties go to the smaller tile
                best = t, NT = nt;
        }
    }
    if (const char *e = getenv("TSG_TC_NT")) // developer override for tuning: any multiple of 16 in [64, 256]
        if (M > 64 && atoi(e) >= 64 && atoi(e) <= 256 && atoi(e) % 16 == 0)
            NT = atoi(e);
    const int mtiles = (M + NT - 1) / NT;
    const int Mp = mtiles * NT;
    p.Mp = Mp;
    TSG_CHECK(mtiles <= 65535, TSG_ERR_UNSUPPORTED, "dense_tc: grid too large");

    // Variants.  Two CTAs per SM (8 expander warps, 256 TMEM columns) need terms*NT <= 128
    // accumulator columns: NT = 32 always fits (c5a: 119 -> 84 µs); for NT >= 64 the gain was ~4 %,
    // so those take the one-CTA variant.
    const bool half = NT == 32 && force_ew != 16;
    const size_t smem = half ? smem_half : smem_full;
    const long long tiles = (long long)ntiles * mtiles;
    const double stage_clk = 16.0 * (NT / 2 > 77 ? NT / 2 : 77);
    // the landing zone of the peers' accumulators may take at most half of the shared memory
    const int ksplit = choose_ksplit(tiles, nkb / kSub, half ? 2 * sms : sms, 1 + budget(smem) / 2 / (NT * 512), stage_clk);
    TailPlan tail = {0, 1};
    if (NT >= 80 && ksplit == 1) // the run-time-height instantiation (term-major accumulators)
        tail = plan_tail(tiles, nkb / kSub, sms, stage_clk);

    // scratch: tile flags [mtiles][nkb] + 16-bit planes of X ([4][Mp][Kp]: three bf16 terms and one
    // fp16 copy; a tile writes only the planes its values need) + the tail launch's partial tiles
    const size_t fl_bytes = ((size_t)mtiles * nkb + 255) & ~(size_t)255;
    const size_t xs_bytes = (size_t)(kMaxSplits + 1) * Mp * Kp * sizeof(uint16_t);
    const size_t ws_bytes = (size_t)tail.tiles * tail.ks * NT * 512;
    const size_t need = fl_bytes + xs_bytes + ws_bytes + 256;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    TSG_CUDA(cudaStreamIsCapturing(st, &cap));
    const bool capturing = cap != cudaStreamCaptureStatusNone;
    if (m->cap_xsplit < need)
    {
        // growing frees the old block: not while a capture is recording launches that use it, and
        // only once every kernel that may still read it has finished
        TSG_CHECK(!capturing, TSG_ERR_UNSUPPORTED,
                  "dense_tc: the handle's scratch must grow (%zu B) but the stream is being captured; run one "
                  "call of this size outside the capture first", need);
        TSG_CUDA(cudaDeviceSynchronize());
        if (m->xsplit)
            cudaFree(m->xsplit);
        m->xsplit = nullptr;
        m->cap_xsplit = 0;
        TSG_CUDA(cudaMalloc(&m->xsplit, need));
        m->cap_xsplit = need;
        m->scratch_used = false;
    }
    // One scratch per handle: calls on one stream are ordered by the stream; a call on ANOTHER stream
    // first waits (on the host: rare, and an event between two kernels would undo their programmatic
    // overlap) until the previous stream's kernels have finished with the scratch.  Launches being
    // captured are ordered by their graph; a graph must not be replayed concurrently with other
    // calls on the same handle (include/tsg.h).
    if (!capturing && m->scratch_used && m->scratch_stream != st)
        TSG_CUDA(cudaStreamSynchronize(m->scratch_stream));
    if (!capturing)
        m->scratch_stream = st, m->scratch_used = true;
    TSG_CHECK(nkb / kSub <= kFlagBytes / 4, TSG_ERR_UNSUPPORTED, "dense_tc: K=%d too large for the TMA path", K);
    uint8_t *tflags = reinterpret_cast<uint8_t *>(m->xsplit);
    uint16_t *xs = reinterpret_cast<uint16_t *>((char *)m->xsplit + fl_bytes);
    p.tflags = tflags;
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)nkb, (unsigned)mtiles);
        cfg.blockDim = dim3(512);
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = g_pdl ? 1 : 0;
        // exact split unless the caller opted into two-fp16-term tiles (tsg_set_fast_split / TSG_TC_FAST=1)
        const int exact_split = g_tsg_fast_split.load(std::memory_order_relaxed) ? 0 : 1;
        TSG_CUDA(cudaLaunchKernelEx(&cfg, split_tiles_kernel, X, ldx, M, K, NT, Mp, Kp, nkb, xs, tflags, exact_split));
        TSG_LAUNCHED();
    }

    // tensor maps over the split buffer: 2-D [4*Mp rows][Kp], box 64 x tile rows, 128-byte swizzle
    auto make_map = [&](CUtensorMap &tm, int rows) -> int {
        const cuuint64_t gdim[2] = {(cuuint64_t)Kp, (cuuint64_t)(kMaxSplits + 1) * Mp};
        const cuuint64_t gstride[1] = {(cuuint64_t)Kp * sizeof(uint16_t)};
        const cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)rows};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, xs, gdim, gstride, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        TSG_CHECK(r == CUDA_SUCCESS, TSG_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
        return TSG_OK;
    };
    TSG_TRY(make_map(map, NT));

    {
        DenseParams q = p;
        q.smem_budget = budget(smem);
        q.ksplit = ksplit;
        q.nt = NT;
        const long long main_tiles = tiles - tail.tiles;
        dim3 grid((unsigned)main_tiles, 1, q.ksplit);
        if (NT == 32)
        {
            q.trace = tc_trace_buffer((size_t)tiles * q.ksplit);
            return half ? launch_nt<32, false, 8>(map, q, grid, smem, m->device, st)
                        : launch_nt<32, false, 16>(map, q, grid, smem, m->device, st);
        }
        if (NT == 64)
        {
            q.trace = tc_trace_buffer((size_t)tiles * q.ksplit);
            return launch_nt<64, false, 16>(map, q, grid, smem, m->device, st);
        }
        // run-time height 80..256: whole waves, then the partial wave K-split through global memory
        if (main_tiles > 0)
        {
            q.trace = tc_trace_buffer((size_t)main_tiles * q.ksplit);
            TSG_TRY((launch_nt<256, false, 16>(map, q, grid, smem, m->device, st)));
        }
        if (tail.tiles > 0)
        {
            q.tile0 = (int)main_tiles;
            q.ksplit = tail.ks;
            q.gws = reinterpret_cast<float *>((char *)m->xsplit + fl_bytes + xs_bytes);
            q.trace = main_tiles > 0 ? nullptr : tc_trace_buffer((size_t)tail.tiles * tail.ks);
            TSG_TRY((launch_nt<256, false, 16>(map, q, dim3((unsigned)tail.tiles, 1, tail.ks), smem, m->device, st)));
        }
        return TSG_OK;
    }
}
