/* oracle/tsg_oracle.c — see tsg_oracle.h.  TEST INFRASTRUCTURE ONLY; parity PINNED against
 * oracle/_ref (the unmodified reference) by tests/test_oracle_pinning.py and tests/golden/.
 *
 * Every function names the reference lines it restates.  Scalar C on purpose. */
#include "tsg_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ============================================================================================
 * std::mt19937 (ISO C++ [rand.eng.mers], parameters of mersenne_twister_engine<uint32,32,624,
 * 397,31,0x9908b0df,11,0xffffffff,7,0x9d2c5680,15,0xefc60000,18,1812433253>)
 * ============================================================================================ */
void orc_mt_seed(orc_mt19937 *g, uint32_t seed)
{
    g->s[0] = seed;
    for (int i = 1; i < 624; ++i)
        g->s[i] = 1812433253u * (g->s[i - 1] ^ (g->s[i - 1] >> 30)) + (uint32_t)i;
    g->idx = 624;
}

static void mt_refill(orc_mt19937 *g)
{
    for (int i = 0; i < 624; ++i)
    {
        uint32_t y = (g->s[i] & 0x80000000u) | (g->s[(i + 1) % 624] & 0x7fffffffu);
        uint32_t v = g->s[(i + 397) % 624] ^ (y >> 1);
        if (y & 1u)
            v ^= 0x9908b0dfu;
        g->s[i] = v;
    }
    g->idx = 0;
}

uint32_t orc_mt_next(orc_mt19937 *g)
{
    if (g->idx >= 624)
        mt_refill(g);
    uint32_t y = g->s[g->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

/* libstdc++ (GCC 11..13) std::uniform_int_distribution<int>(lo,hi)(mt19937&):
 * bits/uniform_int_dist.h, operator() "downscaling" branch for a 32-bit engine, which calls
 * _S_nd<uint64_t> (Lemire 2019).  hi-lo == 0xffffffff cannot occur for int bounds used here. */
int orc_uniform_int(orc_mt19937 *g, int lo, int hi)
{
    uint32_t range = (uint32_t)hi - (uint32_t)lo + 1u;
    uint64_t product = (uint64_t)orc_mt_next(g) * (uint64_t)range;
    uint32_t low = (uint32_t)product;
    if (low < range)
    {
        uint32_t threshold = (0u - range) % range;
        while (low < threshold)
        {
            product = (uint64_t)orc_mt_next(g) * (uint64_t)range;
            low = (uint32_t)product;
        }
    }
    return (int)((uint32_t)(product >> 32) + (uint32_t)lo);
}

/* cpp_impl/sparseUtils.h:52-87 (non-uniform branch, the one main.cpp:60 uses). */
void orc_generate_sparse_matrix(int *out, int H, int W, int nonZero, int seed)
{
    orc_mt19937 g;
    orc_mt_seed(&g, (uint32_t)seed); /* :54 */
    memset(out, 0, (size_t)H * (size_t)W * sizeof(int));
    const int vari_hi = W / nonZero / 20 + 1; /* :56 */
    for (int h = 0; h < H; ++h)
    {
        int *row = out + (size_t)h * W;
        int pos_vari = orc_uniform_int(&g, 0, vari_hi); /* :59 */
        int limit_pos = (W / nonZero) / 2 + pos_vari;   /* :60 */
        int limit_neg = (W / nonZero) / 2 - pos_vari;   /* :61 */
        for (int placed = 0; placed < limit_pos;)       /* :64-73 rejection sampling */
        {
            int c = orc_uniform_int(&g, 0, W - 1);
            if (row[c] == 0)
            {
                row[c] = 1;
                ++placed;
            }
        }
        for (int placed = 0; placed < limit_neg;) /* :76-85 */
        {
            int c = orc_uniform_int(&g, 0, W - 1);
            if (row[c] == 0)
            {
                row[c] = -1;
                ++placed;
            }
        }
    }
}

/* cpp_impl/sparseUtils.h:6-23 (non-uniform branch): integers in [-Range, Range] stored as T. */
void orc_init_x(float *X, long long len, int range, uint32_t seed)
{
    orc_mt19937 g;
    orc_mt_seed(&g, seed);
    for (long long i = 0; i < len; ++i)
        X[i] = (float)orc_uniform_int(&g, -range, range);
}

/* ============================================================================================
 * TCSC — cpp_impl/data_structures/TCSC.h:13-41
 * column by column; within a column rows ascending; values other than +1/-1 ignored.
 * ============================================================================================ */
void orc_tcsc_count(const int *W, int rows, int cols, long long *npos, long long *nneg)
{
    long long p = 0, q = 0;
    for (long long i = 0; i < (long long)rows * cols; ++i)
    {
        p += (W[i] == 1);
        q += (W[i] == -1);
    }
    *npos = p;
    *nneg = q;
}

void orc_tcsc_build(const int *W, int rows, int cols, int *csp, int *csn, int *rip, int *rin)
{
    int p = 0, q = 0;
    for (int n = 0; n < cols; ++n)
    {
        csp[n] = p; /* TCSC.h:20-21 */
        csn[n] = q;
        for (int k = 0; k < rows; ++k)
        {
            int v = W[(size_t)k * cols + n]; /* :25 */
            if (v == 1)
                rip[p++] = k;
            else if (v == -1)
                rin[q++] = k;
        }
    }
    csp[cols] = p; /* :39-40 */
    csn[cols] = q;
}

void orc_tcsc_to_dense(const int *csp, const int *csn, const int *rip, const int *rin, int rows,
                       int cols, int *W)
{
    memset(W, 0, (size_t)rows * (size_t)cols * sizeof(int));
    for (int n = 0; n < cols; ++n)
    {
        for (int i = csp[n]; i < csp[n + 1]; ++i)
            W[(size_t)rip[i] * cols + n] = 1;
        for (int i = csn[n]; i < csn[n + 1]; ++i)
            W[(size_t)rin[i] * cols + n] = -1;
    }
}

long long orc_tcsc_size_bytes(int cols, long long npos, long long nneg)
{
    return 4ll * (2ll * (cols + 1) + npos + nneg); /* TCSC.h:43-49 */
}

/* ============================================================================================
 * kernels
 * ============================================================================================ */
/* cpp_impl/comp.h:37-68 — one fp32 accumulator: +pos (ascending k), -neg (ascending k), +b[n]. */
void orc_base_tcsc(const float *X, const int *csp, const int *csn, const int *rip, const int *rin,
                   const float *b, float *Y, int M, int N, int K)
{
    for (int m = 0; m < M; ++m)
    {
        const float *x = X + (size_t)m * K;
        for (int n = 0; n < N; ++n)
        {
            float y = 0.0f;
            for (int i = csp[n]; i < csp[n + 1]; ++i)
                y += x[rip[i]];
            for (int i = csn[n]; i < csn[n + 1]; ++i)
                y -= x[rin[i]];
            Y[(size_t)m * N + n] = y + b[n];
        }
    }
}

/* cpp_impl/comp_prelu.h:24-69 — same, then strict y>0 ? y : alpha[n]*y. */
void orc_base_tcsc_prelu(const float *X, const int *csp, const int *csn, const int *rip,
                         const int *rin, const float *b, const float *alpha, float *Y, int M,
                         int N, int K)
{
    for (int m = 0; m < M; ++m)
    {
        const float *x = X + (size_t)m * K;
        for (int n = 0; n < N; ++n)
        {
            float y = 0.0f;
            for (int i = csp[n]; i < csp[n + 1]; ++i)
                y += x[rip[i]];
            for (int i = csn[n]; i < csn[n + 1]; ++i)
                y -= x[rin[i]];
            y = y + b[n];
            Y[(size_t)m * N + n] = (y > 0.0f) ? y : alpha[n] * y;
        }
    }
}

/* One sign of one column for one row, in the DoubleUnrolled<4,4> order (comp.h:1260-1303 and
 * :1367-1397): four striped partial sums over the largest multiple-of-4 prefix, reduced
 * 0,1,2,3 into a fresh zero, then the remainder added one by one. */
static float striped4_sum(const float *x, const int *idx, int lo, int hi)
{
    float part[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    int i = lo;
    for (; i + 4 <= hi; i += 4)
        for (int u = 0; u < 4; ++u)
            part[u] += x[idx[i + u]];
    float total = 0.0f;
    for (int u = 0; u < 4; ++u)
        total += part[u];
    for (; i < hi; ++i)
        total += x[idx[i]];
    return total;
}

/* cpp_impl/comp.h:1227-1438.  The 4-row main loop and the 1-row cleanup loop use the same
 * per-row arithmetic, so one helper covers both; Y = (pos - neg) + b (:1354, :1432). */
void orc_double_unrolled_tcsc_k4_m4(const float *X, const int *csp, const int *csn, const int *rip,
                                    const int *rin, const float *b, float *Y, int M, int N, int K)
{
    for (int m = 0; m < M; ++m)
    {
        const float *x = X + (size_t)m * K;
        for (int n = 0; n < N; ++n)
        {
            float p = striped4_sum(x, rip, csp[n], csp[n + 1]);
            float q = striped4_sum(x, rin, csn[n], csn[n + 1]);
            Y[(size_t)m * N + n] = (p - q) + b[n];
        }
    }
}

/* cpp_impl/sparseUtils.h:92-108 */
void orc_gemm(const float *X, const float *W, const float *b, float *Y, int M, int N, int K)
{
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n)
        {
            float y = 0.0f;
            for (int k = 0; k < K; ++k)
                y += X[(size_t)m * K + k] * W[(size_t)k * N + n];
            Y[(size_t)m * N + n] = y + b[n];
        }
}

/* cpp_impl/sparseUtils.h:110-137 (note >= here, > in the sparse kernel; same result) */
void orc_gemm_prelu(const float *X, const float *W, const float *b, const float *alpha, float *Y,
                    int M, int N, int K)
{
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n)
        {
            float y = 0.0f;
            for (int k = 0; k < K; ++k)
                y += X[(size_t)m * K + k] * W[(size_t)k * N + n];
            float pre = y + b[n];
            Y[(size_t)m * N + n] = (pre >= 0.0f) ? pre : alpha[n] * pre;
        }
}

/* cpp_impl/sparseUtils.h:139-156.  The reference calls the integer abs() overload-resolved
 * std::abs(float) via <cmath>; tolerance literal 10e-6 (double). */
int orc_compare_results(const float *result, const float *truth, int H, int W, int *bad_h,
                        int *bad_w)
{
    for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w)
        {
            size_t i = (size_t)h * W + w;
            float d = result[i] - truth[i];
            if (d < 0)
                d = -d;
            if ((double)d > 10e-6 || d != d)
            {
                if (bad_h)
                    *bad_h = h;
                if (bad_w)
                    *bad_w = w;
                return 0;
            }
        }
    return 1;
}

/* ============================================================================================
 * TCSR — cpp_impl/data_structures/TCSR.h:13-41 ; BaseTCSR — cpp_impl/comp.h:478-528
 * ============================================================================================ */
void orc_tcsr_count(const int *W, int rows, int cols, long long *npos, long long *nneg)
{
    orc_tcsc_count(W, rows, cols, npos, nneg);
}

void orc_tcsr_build(const int *W, int rows, int cols, int *rsp, int *rsn, int *cip, int *cin)
{
    int p = 0, q = 0;
    for (int k = 0; k < rows; ++k)
    {
        rsp[k] = p;
        rsn[k] = q;
        for (int n = 0; n < cols; ++n)
        {
            int v = W[(size_t)k * cols + n];
            if (v == 1)
                cip[p++] = n;
            else if (v == -1)
                cin[q++] = n;
        }
    }
    rsp[rows] = p;
    rsn[rows] = q;
}

/* Y = b, then for each m, each k ascending: += x to pos columns, -= x to neg columns. */
void orc_base_tcsr(const float *X, const int *rsp, const int *rsn, const int *cip, const int *cin,
                   const float *b, float *Y, int M, int N, int K)
{
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n)
            Y[(size_t)m * N + n] = b[n];
    for (int m = 0; m < M; ++m)
    {
        float *y = Y + (size_t)m * N;
        for (int k = 0; k < K; ++k)
        {
            float xv = X[(size_t)m * K + k];
            for (int j = rsp[k]; j < rsp[k + 1]; ++j)
                y[cip[j]] += xv;
            for (int j = rsn[k]; j < rsn[k + 1]; ++j)
                y[cin[j]] -= xv;
        }
    }
}

/* ============================================================================================
 * BlockedTCSC<B> — cpp_impl/data_structures/BlockedTCSC.h:15-43 ; kernel cpp_impl/comp.h:607-658
 * ============================================================================================ */
void orc_blocked_count(const int *W, int K, int N, int B, long long *npos, long long *nneg)
{
    orc_tcsc_count(W, (K / B) * B, N, npos, nneg); /* tail rows dropped (BlockedTCSC.h:5,17) */
}

void orc_blocked_build(const int *W, int K, int N, int B, int *csp, int *csn, int *rip, int *rin)
{
    int p = 0, q = 0;
    const int nblk = K / B;
    for (int blk = 0; blk < nblk; ++blk)
        for (int j = 0; j < N; ++j)
        {
            csp[(size_t)blk * N + j] = p;
            csn[(size_t)blk * N + j] = q;
            for (int i = 0; i < B; ++i)
            {
                int row = blk * B + i;
                int v = W[(size_t)row * N + j];
                if (v == 1)
                    rip[p++] = row;
                else if (v == -1)
                    rin[q++] = row;
            }
        }
    csp[(size_t)nblk * N] = p;
    csn[(size_t)nblk * N] = q;
}

/* comp.h:620-657: per block a fresh accumulator is added INTO Y (Y must come in zeroed),
 * bias added after all blocks. */
void orc_base_blocked(const float *X, const int *csp, const int *csn, const int *rip,
                      const int *rin, const float *b, float *Y, int M, int N, int K, int B)
{
    const int nblk = K / B;
    for (int m = 0; m < M; ++m)
    {
        const float *x = X + (size_t)m * K;
        for (int blk = 0; blk < nblk; ++blk)
            for (int j = 0; j < N; ++j)
            {
                size_t c = (size_t)blk * N + j;
                float y = 0.0f;
                for (int i = csp[c]; i < csp[c + 1]; ++i)
                    y += x[rip[i]];
                for (int i = csn[c]; i < csn[c + 1]; ++i)
                    y -= x[rin[i]];
                Y[(size_t)m * N + j] += y;
            }
        for (int n = 0; n < N; ++n)
            Y[(size_t)m * N + n] += b[n];
    }
}
