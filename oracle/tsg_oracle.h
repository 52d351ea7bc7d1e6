/* oracle/tsg_oracle.h — CPU restatement of the reference's ternary sparse-GEMM hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the checker, never the product: only tests/,
 * __graft_entry__.smoke() and bench.py's CPU legs (cpu_baseline, --impl reference fallback)
 * may load liboracle.so.  Nothing under ternary-spgemm_b200/ links or calls it.
 *
 * Parity status: PINNED.  Every function below is checked bit-for-bit against the unmodified
 * reference compiled from /root/reference (oracle/_ref/libtsgref.so, built by oracle/Makefile)
 * in tests/test_oracle_pinning.py, and against the committed fixtures in tests/golden/ that
 * were generated from that library by tests/golden/make_golden.py.
 *
 * Plain C, scalar, single-threaded on purpose: it states WHAT the reference computes
 * (including its summation order) in the most literal form.
 */
#ifndef TSG_ORACLE_H
#define TSG_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- random inputs -------------------------------------------------------------------------
 * std::mt19937 + libstdc++'s std::uniform_int_distribution<int> (GCC >= 11: Lemire's
 * nearly-divisionless reduction for a 32-bit engine).  The distribution is implementation
 * defined in ISO C++, so this restates the libstdc++ one the goldens were made with. */
typedef struct orc_mt19937
{
    uint32_t s[624];
    int idx;
} orc_mt19937;

void orc_mt_seed(orc_mt19937 *g, uint32_t seed);
uint32_t orc_mt_next(orc_mt19937 *g);
int orc_uniform_int(orc_mt19937 *g, int lo, int hi);

/* generateSparseMatrix<int>(H, W, nonZero, false, seed), seed != -1
 *   reference: cpp_impl/sparseUtils.h:25-33,52-90.  out is H*W ints, overwritten. */
void orc_generate_sparse_matrix(int *out, int H, int W, int nonZero, int seed);

/* initX<float>(LEN, Range, false) with an explicit engine seed (the reference seeds with time(0))
 *   reference: cpp_impl/sparseUtils.h:6-23. */
void orc_init_x(float *X, long long len, int range, uint32_t seed);

/* ---- TCSC (a1) -----------------------------------------------------------------------------
 * reference: cpp_impl/data_structures/TCSC.h:13-41.  Two calls: count, then fill caller-owned
 * arrays (csp/csn have cols+1 entries). */
void orc_tcsc_count(const int *W, int rows, int cols, long long *npos, long long *nneg);
void orc_tcsc_build(const int *W, int rows, int cols, int *csp, int *csn, int *rip, int *rin);
/* inverse (what DataStructureInterface::getVectorRepresentation must return) */
void orc_tcsc_to_dense(const int *csp, const int *csn, const int *rip, const int *rin,
                       int rows, int cols, int *W);
/* TCSC::getDataStructureSize, TCSC.h:43-49 */
long long orc_tcsc_size_bytes(int cols, long long npos, long long nneg);

/* ---- kernels (a2, a3, a8) ------------------------------------------------------------------ */
/* BaseTCSC<float>: cpp_impl/comp.h:25-69 */
void orc_base_tcsc(const float *X, const int *csp, const int *csn, const int *rip, const int *rin,
                   const float *b, float *Y, int M, int N, int K);
/* BaseTCSC_PreLU<float>: cpp_impl/comp_prelu.h:12-70 */
void orc_base_tcsc_prelu(const float *X, const int *csp, const int *csn, const int *rip,
                         const int *rin, const float *b, const float *alpha, float *Y, int M,
                         int N, int K);
/* DoubleUnrolledTCSC<float,4,4> summation order: cpp_impl/comp.h:1227-1438 */
void orc_double_unrolled_tcsc_k4_m4(const float *X, const int *csp, const int *csn, const int *rip,
                                    const int *rin, const float *b, float *Y, int M, int N, int K);

/* ---- dense checker (a7) -------------------------------------------------------------------- */
/* GEMM / GEMM_PreLU: cpp_impl/sparseUtils.h:92-137 (W as fp32 K x N) */
void orc_gemm(const float *X, const float *W, const float *b, float *Y, int M, int N, int K);
void orc_gemm_prelu(const float *X, const float *W, const float *b, const float *alpha, float *Y,
                    int M, int N, int K);
/* compare_results: cpp_impl/sparseUtils.h:139-156; returns 1 on pass, fills first mismatch */
int orc_compare_results(const float *result, const float *truth, int H, int W, int *bad_h,
                        int *bad_w);

/* ---- "next" rows of SURVEY §8f --------------------------------------------------------------- */
/* TCSR: cpp_impl/data_structures/TCSR.h:13-41 ; BaseTCSR: cpp_impl/comp.h:478-528 */
void orc_tcsr_count(const int *W, int rows, int cols, long long *npos, long long *nneg);
void orc_tcsr_build(const int *W, int rows, int cols, int *rsp, int *rsn, int *cip, int *cin);
void orc_base_tcsr(const float *X, const int *rsp, const int *rsn, const int *cip, const int *cin,
                   const float *b, float *Y, int M, int N, int K);
/* BlockedTCSC<B>: cpp_impl/data_structures/BlockedTCSC.h:15-43 (rows >= (K/B)*B are dropped);
 * BaseBlockedTCSC<B>: cpp_impl/comp.h:607-658 (accumulates into Y: caller pre-zeroes Y) */
void orc_blocked_count(const int *W, int K, int N, int B, long long *npos, long long *nneg);
void orc_blocked_build(const int *W, int K, int N, int B, int *csp, int *csn, int *rip, int *rin);
void orc_base_blocked(const float *X, const int *csp, const int *csn, const int *rip,
                      const int *rin, const float *b, float *Y, int M, int N, int K, int B);

#ifdef __cplusplus
}
#endif
#endif
