// tsg_build.cu — device-side TCSC builder.
//
// Replaces TCSC::TCSC(const int *matrix, int rows, int cols)
// (reference cpp_impl/data_structures/TCSC.h:13-41): for each column n, rows k ascending,
// +1 -> row_index_pos, -1 -> row_index_neg, col_start_{pos,neg}[n] = running counts.  The
// reference walks W column-wise on one core (stride 4N bytes) with push_back; here W is read
// exactly once, row-wise and coalesced, and everything else works on a 2-bit/element
// intermediate that is kept as the engine's second storage format (the bit planes).
//
//   pass 1  encode_planes_kernel   W (int32|int8, row-major)  ->  ppos/pneg bit planes
//                                  (column-major, 32 rows per word) + per-column counts.
//                                  Thread (col, word): 32 coalesced row reads build one word;
//                                  a 32×32 smem transpose makes the plane writes coalesced too;
//                                  a warp __popc reduction gives the column's count.
//   pass 2  scan_counts_kernel     exclusive prefix sum of the counts -> col_start_pos/neg.
//   pass 3  emit_indices_kernel    one warp per column: each plane word is broadcast with
//                                  __shfl_sync, lane i owns row 32j+i, its slot is
//                                  __popc(word & lanemask_lt) — a ballot-style compaction whose
//                                  stores are contiguous, ascending in k by construction.
//
// HBM traffic: 4·K·N (W read once) + 2·K·N/8·2 (planes written, read) + 4·nnz (indices written).
#include "tsg_internal.cuh"

#include <stdlib.h>

namespace
{

template <typename T>
__global__ void __launch_bounds__(1024)
encode_planes_kernel(const T *__restrict__ W, int K, int64_t ld, int64_t cs, int col_lo, int ncols, int Kw,
                     uint32_t *__restrict__ ppos, uint32_t *__restrict__ pneg,
                     int *__restrict__ cnt_pos, int *__restrict__ cnt_neg)
{
    __shared__ uint32_t tp[32][33];
    __shared__ uint32_t tq[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    {
        const int n = blockIdx.x * 32 + tx;  // column (lanes -> consecutive columns: coalesced)
        const int kw = blockIdx.y * 32 + ty; // plane word = rows [32kw, 32kw+32)
        uint32_t p = 0, q = 0;
        if (n < ncols && kw * 32 < K)
        {
            const T *src = W + (int64_t)kw * 32 * ld + (int64_t)(col_lo + n) * cs; // cs = 1: row-major W
            const int rows = min(32, K - kw * 32);
#pragma unroll 8
            for (int i = 0; i < rows; ++i)
            {
                const int v = (int)src[(int64_t)i * ld];
                p |= (uint32_t)(v == 1) << i;
                q |= (uint32_t)(v == -1) << i;
            }
        }
        tp[ty][tx] = p;
        tq[ty][tx] = q;
    }
    __syncthreads();
    // transposed: this warp (ty) now owns column blockIdx.x*32+ty, lanes -> consecutive words
    const int col = blockIdx.x * 32 + ty;
    const int word = blockIdx.y * 32 + tx;
    const uint32_t p = tp[tx][ty], q = tq[tx][ty];
    if (col < ncols && word < Kw)
    {
        ppos[(int64_t)col * Kw + word] = p;
        pneg[(int64_t)col * Kw + word] = q;
    }
    int cp = __popc(p), cq = __popc(q);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        cp += __shfl_xor_sync(0xffffffffu, cp, o);
        cq += __shfl_xor_sync(0xffffffffu, cq, o);
    }
    if (tx == 0 && col < ncols)
    {
        if (gridDim.y == 1)
        {
            cnt_pos[col] = cp;
            cnt_neg[col] = cq;
        }
        else
        {
            atomicAdd(&cnt_pos[col], cp); // integer: order-independent, exact
            atomicAdd(&cnt_neg[col], cq);
        }
    }
}

// int32 W (the reference's own element type, `const int *matrix`) whose rows allow 16-byte loads:
// warp w of the block owns plane word kw (rows 32kw .. 32kw+31) of a 32-column group, LANE i LOADS
// ROW 32kw+i — its 32 columns are 128 contiguous bytes, whole sectors — and one __ballot_sync per
// column and sign assembles that column's plane word from the 32 lanes' predicates (bit i = lane i
// = row 32kw+i).  The transposed, coalesced plane store and the counts are those of the scalar
// kernel above.  (For int8 W the same idea ran no faster than the scalar kernel — 0.4 warp
// instructions per element either way; the byte-SIMD encoder below serves that case.)
__global__ void __launch_bounds__(1024)
encode_planes_ballot_kernel(const int32_t *__restrict__ W, int K, int64_t ld, int col_lo, int ncols, int Kw,
                            uint32_t *__restrict__ ppos, uint32_t *__restrict__ pneg,
                            int *__restrict__ cnt_pos, int *__restrict__ cnt_neg)
{
    __shared__ uint32_t tp[32][33];
    __shared__ uint32_t tq[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y; // tx = lane, ty = warp
    {
        const int n0 = blockIdx.x * 32;       // first column of the group (whole group inside ncols: host guarantees)
        const int kw = blockIdx.y * 32 + ty;  // plane word of this warp
        const int r = kw * 32 + tx;           // this lane's row
        uint4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
            v[i] = make_uint4(0, 0, 0, 0);
        if (r < K)
        {
            const uint4 *src = reinterpret_cast<const uint4 *>(W + (int64_t)r * ld + col_lo + n0);
#pragma unroll
            for (int i = 0; i < 8; ++i)
                v[i] = __ldg(src + i);
        }
        uint32_t p = 0, q = 0;
#pragma unroll
        for (int c = 0; c < 32; ++c)
        {
            const uint4 &g = v[c >> 2];
            const int e = (int)((c & 3) == 0 ? g.x : ((c & 3) == 1 ? g.y : ((c & 3) == 2 ? g.z : g.w)));
            const uint32_t bp = __ballot_sync(0xffffffffu, e == 1);
            const uint32_t bq = __ballot_sync(0xffffffffu, e == -1);
            if (tx == c)
                p = bp, q = bq;
        }
        tp[ty][tx] = p; // [plane word][column]
        tq[ty][tx] = q;
    }
    __syncthreads();
    // transposed: this warp (ty) now owns column blockIdx.x*32+ty, lanes -> consecutive words
    const int col = blockIdx.x * 32 + ty;
    const int word = blockIdx.y * 32 + tx;
    const uint32_t p = tp[tx][ty], q = tq[tx][ty];
    if (col < ncols && word < Kw)
    {
        ppos[(int64_t)col * Kw + word] = p;
        pneg[(int64_t)col * Kw + word] = q;
    }
    int cp = __popc(p), cq = __popc(q);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        cp += __shfl_xor_sync(0xffffffffu, cp, o);
        cq += __shfl_xor_sync(0xffffffffu, cq, o);
    }
    if (tx == 0 && col < ncols)
    {
        if (gridDim.y == 1)
            cnt_pos[col] = cp, cnt_neg[col] = cq;
        else
        {
            atomicAdd(&cnt_pos[col], cp); // integer: order-independent, exact
            atomicAdd(&cnt_neg[col], cq);
        }
    }
}

// int8 W, the layout model loaders hold: SIMD inside a register.  A thread owns FOUR adjacent columns
// (one 32-bit load per row: a warp reads 128 contiguous bytes of a row) and walks the 32 rows of its
// plane word; per row __vcmpeq4 marks the +1 (and the -1) bytes, and (mask & 0x01010101) << (row & 7)
// drops one bit per column into byte lane c of an accumulator, so after 8 rows each byte lane holds 8
// bits of its column's plane word.  1.5 instructions per matrix element instead of 7; four byte
// permutes per column reassemble the words.  Transposed, coalesced plane stores and counts as above.
__global__ void __launch_bounds__(1024)
encode_planes_bytes_kernel(const int8_t *__restrict__ W, int K, int64_t ld, int col_lo, int ncols, int Kw,
                           uint32_t *__restrict__ ppos, uint32_t *__restrict__ pneg,
                           int *__restrict__ cnt_pos, int *__restrict__ cnt_neg)
{
    __shared__ uint32_t tp[32][129];
    __shared__ uint32_t tq[32][129];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int c0 = blockIdx.x * 128 + tx * 4; // first of this thread's four columns (whole groups of 128: host)
    const int kw = blockIdx.y * 32 + ty;      // plane word = rows [32kw, 32kw+32)
    uint32_t v[32];
    {
        const int8_t *src = W + (int64_t)kw * 32 * ld + col_lo + c0;
        const int rows = min(32, K - kw * 32);
#pragma unroll
        for (int i = 0; i < 32; ++i)
            v[i] = (i < rows) ? __ldg(reinterpret_cast<const uint32_t *>(src + (int64_t)i * ld)) : 0u;
    }
    uint32_t ap[4] = {0, 0, 0, 0}, aq[4] = {0, 0, 0, 0}; // [rows 8j..8j+7]: byte lane c = those bits of column c
#pragma unroll
    for (int i = 0; i < 32; ++i)
    {
        ap[i >> 3] |= (__vcmpeq4(v[i], 0x01010101u) & 0x01010101u) << (i & 7);
        aq[i >> 3] |= (__vcmpeq4(v[i], 0xFFFFFFFFu) & 0x01010101u) << (i & 7);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
    {
        // byte c of ap[0..3] -> one word, rows ascending from bit 0
        const uint32_t sel = 0x0040u + 0x0011u * c; // bytes: {a.c, b.c} = selectors c and 4+c
        const uint32_t lo = __byte_perm(ap[0], ap[1], sel), hi = __byte_perm(ap[2], ap[3], sel);
        tp[ty][tx * 4 + c] = (lo & 0xFFFFu) | (hi << 16);
        const uint32_t lq = __byte_perm(aq[0], aq[1], sel), hq = __byte_perm(aq[2], aq[3], sel);
        tq[ty][tx * 4 + c] = (lq & 0xFFFFu) | (hq << 16);
    }
    __syncthreads();
    // transposed: warp ty owns columns blockIdx.x*128 + 4ty .. +3, lanes -> consecutive plane words
    const int word = blockIdx.y * 32 + tx;
#pragma unroll
    for (int c = 0; c < 4; ++c)
    {
        const int col = blockIdx.x * 128 + ty * 4 + c;
        const uint32_t p = tp[tx][ty * 4 + c], q = tq[tx][ty * 4 + c];
        if (col < ncols && word < Kw)
        {
            ppos[(int64_t)col * Kw + word] = p;
            pneg[(int64_t)col * Kw + word] = q;
        }
        int cp = __popc(p), cq = __popc(q);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
            cp += __shfl_xor_sync(0xffffffffu, cp, o);
            cq += __shfl_xor_sync(0xffffffffu, cq, o);
        }
        if (tx == 0 && col < ncols)
        {
            if (gridDim.y == 1)
                cnt_pos[col] = cp, cnt_neg[col] = cq;
            else
            {
                atomicAdd(&cnt_pos[col], cp); // integer: order-independent, exact
                atomicAdd(&cnt_neg[col], cq);
            }
        }
    }
}

// Exclusive scan of two count arrays (n entries) into two pointer arrays (n+1 entries).  One
// 1024-thread block; warp w owns the contiguous segment [w*L, (w+1)*L), L = ceil(n/32) rounded up
// to whole warps.  Pass 1: every warp sums its segment (coalesced 128-byte loads, no barriers);
// one scan of the 32 segment totals; pass 2: every warp walks its segment again (L2 hits) with a
// warp-level scan per 32 entries and a running carry.  Two block barriers in all.  Totals beyond
// int32 are reported through `totals` (the reference's pointers are int, TCSC.h:8-9).
__global__ void __launch_bounds__(1024)
scan_counts_kernel(const int *__restrict__ cnt_pos, const int *__restrict__ cnt_neg, int n,
                   int *__restrict__ csp, int *__restrict__ csn, long long *__restrict__ totals)
{
    __shared__ long long seg_p[32], seg_q[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int L = ((n + 31) / 32 + 31) & ~31;
    const int lo = min(n, wid * L), hi = min(n, lo + L);
    long long p = 0, q = 0;
    for (int i = lo + lane; i < hi; i += 32)
        p += cnt_pos[i], q += cnt_neg[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        p += __shfl_xor_sync(0xffffffffu, p, o), q += __shfl_xor_sync(0xffffffffu, q, o);
    if (lane == 0)
        seg_p[wid] = p, seg_q[wid] = q;
    __syncthreads();
    if (wid == 0)
    {
        long long wp = seg_p[lane], wq = seg_q[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const long long tp = __shfl_up_sync(0xffffffffu, wp, o), tq = __shfl_up_sync(0xffffffffu, wq, o);
            if (lane >= o)
                wp += tp, wq += tq;
        }
        seg_p[lane] = wp, seg_q[lane] = wq; // inclusive over segments
    }
    __syncthreads();
    long long carry_p = wid ? seg_p[wid - 1] : 0, carry_q = wid ? seg_q[wid - 1] : 0;
    for (int i0 = lo; i0 < hi; i0 += 32)
    {
        const int i = i0 + lane;
        const int cp = (i < hi) ? cnt_pos[i] : 0, cq = (i < hi) ? cnt_neg[i] : 0;
        int ip = cp, iq = cq; // inclusive within these 32 entries (a column holds < 2^31 / 32 entries)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const int tp = __shfl_up_sync(0xffffffffu, ip, o), tq = __shfl_up_sync(0xffffffffu, iq, o);
            if (lane >= o)
                ip += tp, iq += tq;
        }
        if (i < hi)
            csp[i] = (int)(carry_p + ip - cp), csn[i] = (int)(carry_q + iq - cq);
        carry_p += __shfl_sync(0xffffffffu, ip, 31), carry_q += __shfl_sync(0xffffffffu, iq, 31);
    }
    if (threadIdx.x == 0)
    {
        csp[n] = (int)seg_p[31], csn[n] = (int)seg_q[31];
        totals[0] = seg_p[31], totals[1] = seg_q[31];
    }
}

// One warp per (column, sign).  blockDim = 256 (8 warps).  Lane i owns plane word j0 + i (rows
// 32(j0+i) .. +31): a warp-wide exclusive scan of the words' population counts gives every lane the
// slot of its first row, and it then writes its own set bits in ascending order — rows ascend across
// lanes and inside a word, so the list comes out ascending (TCSC.h:24-36) with no serial walk over
// the words (the first version walked them one at a time: 405 us at c4, this one is bounded by the
// longest word).
__global__ void __launch_bounds__(256)
emit_indices_kernel(const uint32_t *__restrict__ ppos, const uint32_t *__restrict__ pneg,
                    const int *__restrict__ csp, const int *__restrict__ csn, int ncols, int Kw,
                    int *__restrict__ rip, int *__restrict__ rin)
{
    const int lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int col = (int)(gw >> 1);
    if (col >= ncols)
        return;
    const bool neg = gw & 1;
    const uint32_t *plane = (neg ? pneg : ppos) + (int64_t)col * Kw;
    int *out = (neg ? rin + csn[col] : rip + csp[col]);
    int written = 0;
    uint32_t nxt = (lane < Kw) ? plane[lane] : 0u;
    for (int j0 = 0; j0 < Kw; j0 += 32)
    {
        uint32_t mine = nxt;                                        // coalesced 128 B
        nxt = (j0 + 32 + lane < Kw) ? plane[j0 + 32 + lane] : 0u;   // next chunk in flight
        const int cnt = __popc(mine);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o)
                incl += t;
        }
        int slot = written + incl - cnt;
        const int base = (j0 + lane) * 32;
        while (mine)
        {
            const int bit = __ffs(mine) - 1;
            mine &= mine - 1;
            out[slot++] = base + bit;
        }
        written += __shfl_sync(0xffffffffu, incl, 31);
    }
}

// Inverse direction for tsg_tcsc_from_arrays: planes from index arrays.  One warp per column.
__global__ void __launch_bounds__(256)
planes_from_arrays_kernel(const int *__restrict__ csp, const int *__restrict__ csn,
                          const int *__restrict__ rip, const int *__restrict__ rin, int ncols,
                          int Kw, uint32_t *__restrict__ ppos, uint32_t *__restrict__ pneg)
{
    const int lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int col = (int)(gw >> 1);
    if (col >= ncols)
        return;
    const bool neg = gw & 1;
    uint32_t *plane = (neg ? pneg : ppos) + (int64_t)col * Kw;
    const int *idx = neg ? rin : rip;
    const int lo = neg ? csn[col] : csp[col], hi = neg ? csn[col + 1] : csp[col + 1];
    for (int i = lo + lane; i < hi; i += 32)
    {
        const int k = idx[i];
        atomicOr(&plane[k >> 5], 1u << (k & 31)); // planes were zero-filled by the caller
    }
}

// Dense reconstruction (getVectorRepresentation): W was zero-filled; one warp per column.
__global__ void __launch_bounds__(256)
scatter_dense_kernel(const int *__restrict__ csp, const int *__restrict__ csn,
                     const int *__restrict__ rip, const int *__restrict__ rin, int ncols,
                     int32_t *__restrict__ W)
{
    const int lane = threadIdx.x & 31;
    const long long col = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (col >= ncols)
        return;
    for (int i = csp[col] + lane; i < csp[col + 1]; i += 32)
        W[(int64_t)rip[i] * ncols + col] = 1;
    for (int i = csn[col] + lane; i < csn[col + 1]; i += 32)
        W[(int64_t)rin[i] * ncols + col] = -1;
}

// ---- kernel-side padded copy of the index lists (see tsg_matrix::lp) --------------------------
// list lengths in 16-byte units of `per` indices (4 x int32, or 8 x uint16 when K fits 16 bits)
__global__ void padded_counts_kernel(const int *__restrict__ csp, const int *__restrict__ csn,
                                     int n, int per, int *__restrict__ c4p, int *__restrict__ c4n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
    {
        c4p[i] = (csp[i + 1] - csp[i] + per - 1) / per;
        c4n[i] = (csn[i + 1] - csn[i] + per - 1) / per;
    }
}

// one warp per (column, sign): copy the list, fill the last int4 with the sentinel K
__global__ void __launch_bounds__(256)
pad_lists_kernel(const int *__restrict__ csp, const int *__restrict__ csn,
                 const int *__restrict__ rip, const int *__restrict__ rin,
                 const int *__restrict__ lp, const int *__restrict__ ln, int ncols, int K,
                 int *__restrict__ rip4, int *__restrict__ rin4)
{
    const int lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int col = (int)(gw >> 1);
    if (col >= ncols)
        return;
    const bool neg = gw & 1;
    const int *src = neg ? rin : rip;
    int *dst = neg ? rin4 : rip4;
    const int lo = neg ? csn[col] : csp[col], len = (neg ? csn[col + 1] : csp[col + 1]) - lo;
    const long long o = 4ll * (neg ? ln[col] : lp[col]);
    const int len4 = (len + 3) & ~3;
    for (int i = lane; i < len4; i += 32)
        dst[o + i] = (i < len) ? src[lo + i] : K;
}

// the same with 16-bit row ids (K <= 65535, true for every BASELINE shape; readme.md:108-111 asks for
// a denser index stream): 8 indices per 16-byte unit, half the HBM bytes of the gather kernel
__global__ void __launch_bounds__(256)
pad_lists16_kernel(const int *__restrict__ csp, const int *__restrict__ csn,
                   const int *__restrict__ rip, const int *__restrict__ rin,
                   const int *__restrict__ lp, const int *__restrict__ ln, int ncols, int K,
                   uint16_t *__restrict__ rip8, uint16_t *__restrict__ rin8)
{
    const int lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int col = (int)(gw >> 1);
    if (col >= ncols)
        return;
    const bool neg = gw & 1;
    const int *src = neg ? rin : rip;
    uint16_t *dst = neg ? rin8 : rip8;
    const int lo = neg ? csn[col] : csp[col], len = (neg ? csn[col + 1] : csp[col + 1]) - lo;
    const long long o = 8ll * (neg ? ln[col] : lp[col]);
    const int len8 = (len + 7) & ~7;
    for (int i = lane; i < len8; i += 32)
        dst[o + i] = (uint16_t)((i < len) ? src[lo + i] : K);
}

// ---- tile-packed codes for the tensor-core path (see tsg_matrix::codes) -------------------------
// 16 consecutive k of one column -> one 32-bit code word laid out for a shift-and-mask expansion:
// element e = 2p + h (pair p = 0..7, h = 0 low / 1 high half of an output register) keeps its
// "non-zero" flag at bit 16h + 14 - 2p and its "negative" flag one above, so that
//     (word << 2p) & 0xC000C000
// is the packed pair (W[2p], W[2p+1]) as two 16-bit floats of value 0 / +2 / -2 (0x4000 is 2.0 in
// bf16 AND in fp16; the sign is bit 15): one shift and one AND per two matrix elements.
__device__ __forceinline__ uint32_t pack_code_word(uint32_t pos16, uint32_t neg16)
{
    // element e = 2p + h keeps its non-zero flag at bit 16h + 14 - 2p and its sign one above: the
    // even elements descend from bit 14, the odd ones from bit 30.  A bit reversal puts element e at
    // bit 31 - e, which IS 30 - 2p for the odd elements and 14 - 2p after a shift by 17 for the even
    // ones — two masks per plane instead of a 16-step loop (the packing kernel was bound by that
    // loop: 66 us at c4).
    const uint32_t bz = __brev((pos16 | neg16) & 0xFFFFu), bq = __brev(neg16 & 0xFFFFu);
    return ((bz >> 17) & 0x00005555u) | (bz & 0x55550000u) | ((bq >> 16) & 0x0000AAAAu) | ((bq << 1) & 0xAAAA0000u);
}

// thread -> (column row of a 128-column tile, 4 consecutive k-blocks): reads one 32-byte sector of
// each plane, writes four coalesced uint4.
__global__ void __launch_bounds__(128)
tile_codes_kernel(const uint32_t *__restrict__ ppos, const uint32_t *__restrict__ pneg, int N, int Kw,
                  int nkb, uint4 *__restrict__ codes)
{
    const int row = threadIdx.x, tile = blockIdx.y, kb0 = blockIdx.x * 4;
    const int n = tile * 128 + row;
    uint32_t P[8] = {0, 0, 0, 0, 0, 0, 0, 0}, Q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (n < N)
    {
        // Kw is a multiple of 4 and 2*kb0 of 8: two aligned 128-bit loads per plane
#pragma unroll
        for (int j = 0; j < 8; j += 4)
        {
            const int w = 2 * kb0 + j;
            if (w < Kw)
            {
                const uint4 a = __ldg(reinterpret_cast<const uint4 *>(ppos + (int64_t)n * Kw + w));
                const uint4 c = __ldg(reinterpret_cast<const uint4 *>(pneg + (int64_t)n * Kw + w));
                P[j] = a.x, P[j + 1] = a.y, P[j + 2] = a.z, P[j + 3] = a.w;
                Q[j] = c.x, Q[j + 1] = c.y, Q[j + 2] = c.z, Q[j + 3] = c.w;
            }
        }
    }
#pragma unroll
    for (int b = 0; b < 4; ++b)
    {
        if (kb0 + b >= nkb)
            break;
        uint32_t c[4];
#pragma unroll
        for (int w = 0; w < 4; ++w)
        {
            const uint32_t p = P[2 * b + (w >> 1)] >> (16 * (w & 1));
            const uint32_t q = Q[2 * b + (w >> 1)] >> (16 * (w & 1));
            c[w] = pack_code_word(p & 0xFFFFu, q & 0xFFFFu);
        }
        codes[((int64_t)tile * nkb + kb0 + b) * 128 + row] = make_uint4(c[0], c[1], c[2], c[3]);
    }
}

__global__ void rebase_kernel(int *__restrict__ dst, const int *__restrict__ src, int n)
{
    const int base = src[0];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        dst[i] = src[i] - base;
}


// ---- BlockedTCSC<B> (reference cpp_impl/data_structures/BlockedTCSC.h:15-43) -----------------------
// The same TCSC construction restricted to K-blocks of B rows: pointer entry b*N + j covers rows
// [b*B, (b+1)*B) of column j, row ids stay global, rows at or beyond (K/B)*B are dropped
// (BlockedTCSC.h:5,17).  B is a multiple of 32, so a (block, column) pair is a run of whole plane
// words.  Counts by popc, the same scan, the same ballot-style emit.
__global__ void blocked_counts_kernel(const uint32_t *__restrict__ ppos, const uint32_t *__restrict__ pneg, int N,
                                      int Kw, int nb, int wpb, int *__restrict__ cnt_pos, int *__restrict__ cnt_neg)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)nb * N)
        return;
    const int b = (int)(i / N), j = (int)(i - (long long)b * N);
    const uint32_t *pp = ppos + (int64_t)j * Kw + (int64_t)b * wpb, *pq = pneg + (int64_t)j * Kw + (int64_t)b * wpb;
    int cp = 0, cq = 0;
    for (int w = 0; w < wpb; ++w)
        cp += __popc(pp[w]), cq += __popc(pq[w]);
    cnt_pos[i] = cp;
    cnt_neg[i] = cq;
}

// one warp per (block, column, sign)
__global__ void __launch_bounds__(256)
blocked_emit_kernel(const uint32_t *__restrict__ ppos, const uint32_t *__restrict__ pneg, const int *__restrict__ csp,
                    const int *__restrict__ csn, int N, int Kw, int nb, int wpb, int *__restrict__ rip,
                    int *__restrict__ rin)
{
    const int lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const long long pair = gw >> 1;
    if (pair >= (long long)nb * N)
        return;
    const bool neg = gw & 1;
    const int b = (int)(pair / N), j = (int)(pair - (long long)b * N);
    const uint32_t *plane = (neg ? pneg : ppos) + (int64_t)j * Kw + (int64_t)b * wpb;
    int *out = neg ? rin + csn[pair] : rip + csp[pair];
    const uint32_t lt = (1u << lane) - 1u;
    int written = 0;
    for (int w = 0; w < wpb; ++w)
    {
        const uint32_t word = plane[w]; // broadcast
        if ((word >> lane) & 1u)
            out[written + __popc(word & lt)] = (b * wpb + w) * 32 + lane;
        written += __popc(word);
    }
}

// Interchange entry points (tsg_*_from_arrays) adopt caller-made index lists: every list must lie
// in [0, bound) and ascend strictly (what the reference constructors produce, TCSC.h:24-36), or the
// kernels that consume them would write outside their buffers.  One warp per list; the first
// violation found is reported (list, position).
__global__ void __launch_bounds__(256)
validate_lists_kernel(const int *__restrict__ ptr, const int *__restrict__ idx, int nlists, int bound,
                      int *__restrict__ bad)
{
    const int lane = threadIdx.x & 31;
    const long long list = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (list >= nlists)
        return;
    const int lo = ptr[list], hi = ptr[list + 1];
    for (int i = lo + lane; i < hi; i += 32)
    {
        const int k = idx[i];
        if (k < 0 || k >= bound || (i > lo && idx[i - 1] >= k))
        {
            if (atomicCAS(&bad[0], 0, 1) == 0)
                bad[1] = (int)list, bad[2] = i - lo;
            return;
        }
    }
}

// the same row listed as +1 and as -1 in one column
__global__ void overlap_kernel(const uint32_t *__restrict__ ppos, const uint32_t *__restrict__ pneg, size_t words,
                               int *__restrict__ bad)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (size_t)gridDim.x * blockDim.x)
        if (ppos[i] & pneg[i])
            bad[0] = 2;
}

} // namespace

static const size_t kIndexPad = 64; // bytes of zero padding after rip/rin (vector loads overrun)

// TSG_BUILD_TIMING=1: CUDA events around the builder's kernel sequences (developer / bench use)
namespace
{
struct BuildTimer
{
    bool on = false;
    cudaEvent_t ev[6] = {};
    int n = 0;
    BuildTimer()
    {
        static const bool enabled = getenv("TSG_BUILD_TIMING") != nullptr;
        on = enabled;
        if (on)
            for (cudaEvent_t &e : ev)
                if (cudaEventCreate(&e) != cudaSuccess)
                    on = false;
    }
    void mark(cudaStream_t st)
    {
        if (on && n < 6)
            cudaEventRecord(ev[n++], st);
    }
    double total_ms()
    {
        double t = 0.0;
        for (int i = 0; on && i + 1 < n; i += 2)
        {
            float ms = 0.0f;
            if (cudaEventElapsedTime(&ms, ev[i], ev[i + 1]) == cudaSuccess)
                t += ms;
        }
        return t;
    }
    ~BuildTimer()
    {
        for (cudaEvent_t &e : ev)
            if (e)
                cudaEventDestroy(e);
    }
};
double g_last_build_ms = 0.0, g_build_ms_open = 0.0;
} // namespace

extern "C" double tsg_debug_last_build_device_ms(void) { return g_last_build_ms; }

// Device-side TCSC construction (reference TCSC::TCSC, TCSC.h:13-41).  Two device allocations:
// blk0 — bit planes, both pointer arrays and the scan scratch, sized from K and N before anything
// runs; blk1 — both row-index arrays, sized from the scan's totals (the one host wait the format
// needs: nnz is data).  encode_planes + scan_counts, wait, emit_indices.
int tsg_build_from_dense_dev(tsg_matrix *m, const void *W_dev, int elem_bytes, int64_t ld,
                             int col_lo, cudaStream_t st, int64_t cs)
{
    const int K = m->K, N = m->N, Kw = m->Kw;
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t plane_bytes = up((size_t)N * Kw * sizeof(uint32_t) + 4), ptr_bytes = up((size_t)(N + 1) * 4),
                 cnt_bytes = up((size_t)(2 * N + 2) * 4);
    // Stream-ordered allocations from the device's default pool (up to 2 GB of freed blocks stay
    // cached there): building matrix after matrix — a model's layers, a benchmark loop — then pays
    // for mapping device memory once, not per matrix (c4: 1.7 -> 0.5 ms of wall time per build).
    {
        static std::atomic<unsigned long long> configured{0};
        const unsigned long long bit = 1ull << (m->device & 63);
        if (!(configured.fetch_or(bit) & bit))
        {
            cudaMemPool_t pool = nullptr;
            unsigned long long keep = 2ull << 30;
            if (cudaDeviceGetDefaultMemPool(&pool, m->device) == cudaSuccess)
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            cudaGetLastError();
        }
    }
    char *blk = nullptr;
    TSG_CUDA(cudaMallocAsync(&blk, 2 * plane_bytes + 2 * ptr_bytes + cnt_bytes + 256, st));
    m->blk0 = blk;
    m->pooled = true;
    m->ppos = reinterpret_cast<uint32_t *>(blk);
    m->pneg = reinterpret_cast<uint32_t *>(blk + plane_bytes);
    m->csp = reinterpret_cast<int32_t *>(blk + 2 * plane_bytes);
    m->csn = reinterpret_cast<int32_t *>(blk + 2 * plane_bytes + ptr_bytes);
    int *cnt = reinterpret_cast<int *>(blk + 2 * plane_bytes + 2 * ptr_bytes);
    long long *totals = reinterpret_cast<long long *>(blk + 2 * plane_bytes + 2 * ptr_bytes + cnt_bytes);
    BuildTimer tm;
    g_last_build_ms = g_build_ms_open = 0.0;
    TSG_CUDA(cudaMemsetAsync(cnt, 0, (size_t)(2 * N + 2) * 4, st));
    tm.mark(st);
    if (N > 0 && K > 0)
    {
        dim3 blkdim(32, 32), grd((N + 31) / 32, (Kw + 31) / 32);
        // int8 W whose rows allow 32-bit loads: the byte-SIMD encoder on whole groups of 128 columns;
        // int32 W whose rows allow 16-byte loads: the ballot encoder on whole groups of 32 columns;
        // the remaining columns, unaligned or transposed (cs != 1) input: the scalar encoder
        const bool scalar_only = getenv("TSG_BUILD_SCALAR") != nullptr || cs != 1;
        const bool bytes_ok = !scalar_only && elem_bytes == 1 && ((uintptr_t)W_dev & 3) == 0 && ld % 4 == 0 && col_lo % 4 == 0;
        const bool vec_ok = !scalar_only && elem_bytes == 4 && ((uintptr_t)W_dev & 15) == 0 && ld % 4 == 0 && col_lo % 4 == 0;
        const int done_cols = bytes_ok ? (N / 128) * 128 : (vec_ok ? (N / 32) * 32 : 0);
        const int groups = done_cols / 32, rest = N - done_cols;
        if (bytes_ok && done_cols > 0)
        {
            encode_planes_bytes_kernel<<<dim3(done_cols / 128, grd.y), blkdim, 0, st>>>(
                (const int8_t *)W_dev, K, ld, col_lo, N, Kw, m->ppos, m->pneg, cnt, cnt + N);
            g_tsg_launches.fetch_add(1, std::memory_order_relaxed);
        }
        else if (vec_ok && done_cols > 0)
        {
            encode_planes_ballot_kernel<<<dim3(groups, grd.y), blkdim, 0, st>>>(
                (const int32_t *)W_dev, K, ld, col_lo, N, Kw, m->ppos, m->pneg, cnt, cnt + N);
            g_tsg_launches.fetch_add(1, std::memory_order_relaxed);
        }
        if (rest > 0)
        {
            // columns [groups*32, N): the scalar kernel on the tail (planes and counts offset accordingly)
            dim3 g3((rest + 31) / 32, grd.y);
            const int c0 = groups * 32;
            if (elem_bytes == 4)
                encode_planes_kernel<int32_t><<<g3, blkdim, 0, st>>>(
                    (const int32_t *)W_dev, K, ld, cs, col_lo + c0, rest, Kw, m->ppos + (size_t)c0 * Kw,
                    m->pneg + (size_t)c0 * Kw, cnt + c0, cnt + N + c0);
            else
                encode_planes_kernel<int8_t><<<g3, blkdim, 0, st>>>(
                    (const int8_t *)W_dev, K, ld, cs, col_lo + c0, rest, Kw, m->ppos + (size_t)c0 * Kw,
                    m->pneg + (size_t)c0 * Kw, cnt + c0, cnt + N + c0);
            g_tsg_launches.fetch_add(1, std::memory_order_relaxed);
        }
    }
    scan_counts_kernel<<<1, 1024, 0, st>>>(cnt, cnt + N, N, m->csp, m->csn, totals);
    g_tsg_launches.fetch_add(1, std::memory_order_relaxed);
    tm.mark(st);
    long long h_tot[2] = {0, 0};
    cudaError_t e = cudaMemcpyAsync(h_tot, totals, 16, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(st);
    TSG_CHECK(e == cudaSuccess, TSG_ERR_CUDA, "TCSC builder (encode/scan) failed: %s", cudaGetErrorString(e));
    TSG_CHECK(h_tot[0] <= INT32_MAX && h_tot[1] <= INT32_MAX, TSG_ERR_OVERFLOW,
              "nnz+ = %lld / nnz- = %lld exceeds the reference's int32 pointers", h_tot[0], h_tot[1]);
    m->npos = h_tot[0];
    m->nneg = h_tot[1];
    const size_t bp = up((size_t)m->npos * 4 + kIndexPad), bq = up((size_t)m->nneg * 4 + kIndexPad);
    char *idx = nullptr;
    if (cudaMallocAsync(&idx, bp + bq, st) != cudaSuccess)
    {
        cudaGetLastError();
        tsg_set_error("cudaMalloc of %zu index bytes failed", bp + bq);
        return TSG_ERR_NOMEM;
    }
    m->blk1 = idx;
    m->rip = reinterpret_cast<int32_t *>(idx);
    m->rin = reinterpret_cast<int32_t *>(idx + bp);
    cudaMemsetAsync((char *)m->rip + (size_t)m->npos * 4, 0, kIndexPad, st);
    cudaMemsetAsync((char *)m->rin + (size_t)m->nneg * 4, 0, kIndexPad, st);
    tm.mark(st);
    if (N > 0)
    {
        const long long warps = 2ll * N;
        emit_indices_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(
            m->ppos, m->pneg, m->csp, m->csn, N, Kw, m->rip, m->rin);
        g_tsg_launches.fetch_add(1, std::memory_order_relaxed);
    }
    tm.mark(st);
    e = cudaStreamSynchronize(st);
    if (e == cudaSuccess)
        e = cudaGetLastError();
    TSG_CHECK(e == cudaSuccess, TSG_ERR_CUDA, "TCSC builder (emit) failed: %s", cudaGetErrorString(e));
    g_build_ms_open = tm.total_ms();
    g_last_build_ms = g_build_ms_open;
    return TSG_OK;
}

int tsg_build_planes_from_arrays(tsg_matrix *m, cudaStream_t st)
{
    const size_t plane_bytes = (size_t)m->N * m->Kw * sizeof(uint32_t);
    TSG_CUDA(cudaMalloc(&m->ppos, plane_bytes ? plane_bytes : 4));
    TSG_CUDA(cudaMalloc(&m->pneg, plane_bytes ? plane_bytes : 4));
    TSG_CUDA(cudaMemsetAsync(m->ppos, 0, plane_bytes, st));
    TSG_CUDA(cudaMemsetAsync(m->pneg, 0, plane_bytes, st));
    if (m->N > 0)
    {
        const long long warps = 2ll * m->N;
        planes_from_arrays_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(
            m->csp, m->csn, m->rip, m->rin, m->N, m->Kw, m->ppos, m->pneg);
        TSG_LAUNCHED();
    }
    return TSG_OK;
}

int tsg_scatter_to_dense(const tsg_matrix *m, int32_t *W_dev, cudaStream_t st)
{
    TSG_CUDA(cudaMemsetAsync(W_dev, 0, (size_t)m->K * m->N * 4, st));
    if (m->N > 0)
    {
        scatter_dense_kernel<<<(m->N + 7) / 8, 256, 0, st>>>(m->csp, m->csn, m->rip, m->rin, m->N,
                                                             W_dev);
        TSG_LAUNCHED();
    }
    return TSG_OK;
}

int tsg_build_padded_lists(tsg_matrix *m, cudaStream_t st)
{
    const int N = m->N;
    for (int32_t **p : {&m->lp, &m->ln, &m->rip4, &m->rin4})
        if (*p)
        {
            cudaFree(*p);
            *p = nullptr;
        }
    TSG_CUDA(cudaMalloc(&m->lp, (size_t)(N + 1) * 4));
    TSG_CUDA(cudaMalloc(&m->ln, (size_t)(N + 1) * 4));
    // 16-bit row ids whenever the sentinel K fits (TSG_GATHER_IDX32=1: developer override)
    static const bool force32 = getenv("TSG_GATHER_IDX32") != nullptr;
    m->idx16 = !force32 && m->K <= 65535;
    int *cnt = nullptr;
    long long *totals = nullptr;
    TSG_CUDA(cudaMalloc(&cnt, (size_t)(2 * N + 2) * 4));
    TSG_CUDA(cudaMalloc(&totals, 16));
    int status = TSG_OK;
    do
    {
        if (N > 0)
        {
            padded_counts_kernel<<<(N + 255) / 256, 256, 0, st>>>(m->csp, m->csn, N, m->idx16 ? 8 : 4, cnt, cnt + N);
            g_tsg_launches.fetch_add(1, std::memory_order_relaxed);
        }
        scan_counts_kernel<<<1, 1024, 0, st>>>(cnt, cnt + N, N, m->lp, m->ln, totals);
        g_tsg_launches.fetch_add(1, std::memory_order_relaxed);
        long long h_tot[2] = {0, 0};
        cudaError_t e = cudaMemcpyAsync(h_tot, totals, 16, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess)
            e = cudaStreamSynchronize(st);
        if (e != cudaSuccess)
        {
            tsg_set_error("padded list build failed: %s", cudaGetErrorString(e));
            status = TSG_ERR_CUDA;
            break;
        }
        if (h_tot[0] * 4 > INT32_MAX || h_tot[1] * 4 > INT32_MAX)
        {
            tsg_set_error("padded index stream exceeds int32 addressing");
            status = TSG_ERR_OVERFLOW;
            break;
        }
        m->n4pos = h_tot[0];
        m->n4neg = h_tot[1];
        // + one batch of slack: the kernel's predicated loads never read past a list, but keep
        // the allocation generous and defined
        const size_t bp = (size_t)m->n4pos * 16 + 256, bq = (size_t)m->n4neg * 16 + 256;
        if (cudaMalloc(&m->rip4, bp) != cudaSuccess || cudaMalloc(&m->rin4, bq) != cudaSuccess)
        {
            tsg_set_error("cudaMalloc of %zu padded index bytes failed", bp + bq);
            status = TSG_ERR_NOMEM;
            break;
        }
        cudaMemsetAsync((char *)m->rip4 + (size_t)m->n4pos * 16, 0, 256, st);
        cudaMemsetAsync((char *)m->rin4 + (size_t)m->n4neg * 16, 0, 256, st);
        if (N > 0)
        {
            const long long warps = 2ll * N;
            if (m->idx16)
                pad_lists16_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(
                    m->csp, m->csn, m->rip, m->rin, m->lp, m->ln, N, m->K, (uint16_t *)m->rip4, (uint16_t *)m->rin4);
            else
                pad_lists_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(
                    m->csp, m->csn, m->rip, m->rin, m->lp, m->ln, N, m->K, m->rip4, m->rin4);
            g_tsg_launches.fetch_add(1, std::memory_order_relaxed);
        }
        e = cudaStreamSynchronize(st);
        if (e == cudaSuccess)
            e = cudaGetLastError();
        if (e != cudaSuccess)
        {
            tsg_set_error("padded list build failed: %s", cudaGetErrorString(e));
            status = TSG_ERR_CUDA;
        }
    } while (0);
    cudaFree(cnt);
    cudaFree(totals);
    return status;
}

int tsg_build_tile_codes(tsg_matrix *m, cudaStream_t st)
{
    if (m->codes)
    {
        if (m->pooled)
            cudaFreeAsync(m->codes, st);
        else
            cudaFree(m->codes);
        m->codes = nullptr;
    }
    const int tiles = (m->N + 127) / 128, nkb = (((m->K + 63) / 64) + 3) & ~3; // whole stages of 4 sub-blocks: zero-code padding
    m->code_tiles = tiles;
    m->code_kblocks = nkb;
    const size_t bytes = (size_t)tiles * nkb * 128 * sizeof(uint4);
    if (m->pooled)
        TSG_CUDA(cudaMallocAsync(&m->codes, bytes ? bytes : 16, st));
    else
        TSG_CUDA(cudaMalloc(&m->codes, bytes ? bytes : 16));
    if (tiles > 0 && nkb > 0)
    {
        TSG_CHECK(tiles <= 65535, TSG_ERR_UNSUPPORTED, "N too large for the tile-code builder");
        dim3 grid((nkb + 3) / 4, tiles);
        BuildTimer tm;
        tm.mark(st);
        tile_codes_kernel<<<grid, 128, 0, st>>>(m->ppos, m->pneg, m->N, m->Kw, nkb, m->codes);
        TSG_LAUNCHED();
        tm.mark(st);
        if (tm.on && cudaStreamSynchronize(st) == cudaSuccess)
            g_last_build_ms = g_build_ms_open + tm.total_ms();
    }
    return TSG_OK;
}

int tsg_rebase_slice(int32_t *dst, const int32_t *src, int n, cudaStream_t st)
{
    rebase_kernel<<<(n + 255) / 256 > 1024 ? 1024 : (n + 255) / 256, 256, 0, st>>>(dst, src, n);
    TSG_LAUNCHED();
    return TSG_OK;
}

// BlockedTCSC<B> arrays of the matrix held by `m`, built on the device into fresh allocations
// (caller frees with cudaFree).  Pointer arrays have (K/B)*N + 1 entries.
int tsg_build_blocked(const tsg_matrix *m, int B, int32_t **csp, int32_t **csn, int32_t **rip, int32_t **rin,
                      long long *npos, long long *nneg)
{
    TSG_CHECK(B > 0 && B % 32 == 0, TSG_ERR_UNSUPPORTED, "BlockedTCSC: block size %d is not a multiple of 32", B);
    const int N = m->N, nb = m->K / B, wpb = B / 32;
    const long long pairs = (long long)nb * N;
    TSG_CHECK(pairs + 1 <= INT32_MAX, TSG_ERR_OVERFLOW, "BlockedTCSC: too many (block, column) pairs");
    *csp = *csn = *rip = *rin = nullptr;
    cudaStream_t st = m->stream;
    int *cnt = nullptr;
    long long *totals = nullptr;
    TSG_CUDA(cudaMalloc(&cnt, (size_t)(2 * pairs + 2) * 4));
    TSG_CUDA(cudaMalloc(&totals, 16));
    TSG_CUDA(cudaMalloc(csp, (size_t)(pairs + 1) * 4));
    TSG_CUDA(cudaMalloc(csn, (size_t)(pairs + 1) * 4));
    int status = TSG_OK;
    do
    {
        if (pairs > 0)
        {
            blocked_counts_kernel<<<(unsigned)((pairs + 255) / 256), 256, 0, st>>>(m->ppos, m->pneg, N, m->Kw, nb, wpb, cnt,
                                                                                   cnt + pairs);
            g_tsg_launches.fetch_add(1, std::memory_order_relaxed);
        }
        scan_counts_kernel<<<1, 1024, 0, st>>>(cnt, cnt + pairs, (int)pairs, *csp, *csn, totals);
        g_tsg_launches.fetch_add(1, std::memory_order_relaxed);
        long long h_tot[2] = {0, 0};
        cudaError_t e = cudaMemcpyAsync(h_tot, totals, 16, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess)
            e = cudaStreamSynchronize(st);
        if (e != cudaSuccess)
        {
            tsg_set_error("BlockedTCSC builder (count/scan) failed: %s", cudaGetErrorString(e));
            status = TSG_ERR_CUDA;
            break;
        }
        *npos = h_tot[0], *nneg = h_tot[1];
        if (cudaMalloc(rip, (size_t)*npos * 4 + 16) != cudaSuccess || cudaMalloc(rin, (size_t)*nneg * 4 + 16) != cudaSuccess)
        {
            tsg_set_error("BlockedTCSC: index allocation failed");
            status = TSG_ERR_NOMEM;
            break;
        }
        if (pairs > 0)
        {
            blocked_emit_kernel<<<(unsigned)((2 * pairs + 7) / 8), 256, 0, st>>>(m->ppos, m->pneg, *csp, *csn, N, m->Kw, nb,
                                                                                wpb, *rip, *rin);
            g_tsg_launches.fetch_add(1, std::memory_order_relaxed);
        }
        e = cudaStreamSynchronize(st);
        if (e != cudaSuccess)
        {
            tsg_set_error("BlockedTCSC builder (emit) failed: %s", cudaGetErrorString(e));
            status = TSG_ERR_CUDA;
        }
    } while (0);
    cudaFree(cnt);
    cudaFree(totals);
    if (status != TSG_OK)
    {
        cudaFree(*csp), cudaFree(*csn), cudaFree(*rip), cudaFree(*rin);
        *csp = *csn = *rip = *rin = nullptr;
    }
    return status;
}

// Host pointers (n + 1 entries): start at 0, never decrease, end at `total`.
int tsg_validate_pointers(const int32_t *ptr, int n, long long total, const char *what)
{
    TSG_CHECK(ptr[0] == 0, TSG_ERR_INVALID, "%s: pointer array starts at %d, not 0", what, ptr[0]);
    for (int i = 0; i < n; ++i)
        TSG_CHECK(ptr[i + 1] >= ptr[i], TSG_ERR_INVALID, "%s: pointer array decreases at entry %d (%d -> %d)", what,
                  i + 1, ptr[i], ptr[i + 1]);
    TSG_CHECK((long long)ptr[n] == total, TSG_ERR_INVALID, "%s: last pointer %d != number of entries %lld", what, ptr[n],
              total);
    return TSG_OK;
}

// Device arrays: every list inside [0, bound) and strictly ascending.  Synchronises `st`.
int tsg_validate_lists(const int32_t *ptr_dev, const int32_t *idx_dev, int nlists, int bound, cudaStream_t st,
                       const char *what)
{
    if (nlists <= 0)
        return TSG_OK;
    int *bad = nullptr;
    TSG_CUDA(cudaMalloc(&bad, 16));
    cudaMemsetAsync(bad, 0, 16, st);
    validate_lists_kernel<<<(unsigned)((nlists + 7) / 8), 256, 0, st>>>(ptr_dev, idx_dev, nlists, bound, bad);
    g_tsg_launches.fetch_add(1, std::memory_order_relaxed);
    int h[4] = {0, 0, 0, 0};
    cudaError_t e = cudaMemcpyAsync(h, bad, 16, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(st);
    cudaFree(bad);
    TSG_CHECK(e == cudaSuccess, TSG_ERR_CUDA, "%s: validation failed to run: %s", what, cudaGetErrorString(e));
    TSG_CHECK(h[0] == 0, TSG_ERR_INVALID,
              "%s: list %d, entry %d is outside [0,%d) or not strictly ascending", what, h[1], h[2], bound);
    return TSG_OK;
}

// a row index present in both sign lists of a column.  Synchronises `st`.
int tsg_validate_no_overlap(const tsg_matrix *m, cudaStream_t st)
{
    const size_t words = (size_t)m->N * m->Kw;
    if (!words)
        return TSG_OK;
    int *bad = nullptr;
    TSG_CUDA(cudaMalloc(&bad, 16));
    cudaMemsetAsync(bad, 0, 16, st);
    overlap_kernel<<<296, 256, 0, st>>>(m->ppos, m->pneg, words, bad);
    g_tsg_launches.fetch_add(1, std::memory_order_relaxed);
    int h = 0;
    cudaError_t e = cudaMemcpyAsync(&h, bad, 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(st);
    cudaFree(bad);
    TSG_CHECK(e == cudaSuccess, TSG_ERR_CUDA, "overlap check failed to run: %s", cudaGetErrorString(e));
    TSG_CHECK(h == 0, TSG_ERR_INVALID, "a row index appears in both the +1 and the -1 list of a column");
    return TSG_OK;
}
