// oracle/ref_shim.cpp — C-ABI window onto the UNMODIFIED reference, compiled where it lies.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under ternary-spgemm_b200/ links, loads or calls this.
// Allowed users: tests/, __graft_entry__.smoke() (as the checker) and bench.py's CPU legs
// (`cpu_baseline`, `--impl reference`).
//
// This translation unit contains no reference code: it #includes the reference headers
// straight from $(REF)/cpp_impl (default /root/reference/cpp_impl, see oracle/Makefile) and
// forwards to them, so that
//   * the C restatement in oracle/tsg_oracle.c can be pinned against the real thing, and
//   * the reference's own kernels can be timed as the CPU baseline (kind "reference").
// Output goes to oracle/_ref/libtsgref.so (git-ignored, but it travels to the GPU box).
//
// Reference entry points forwarded (file:line under /root/reference):
//   generateSparseMatrix<int>      cpp_impl/sparseUtils.h:25-90
//   GEMM / GEMM_PreLU              cpp_impl/sparseUtils.h:92-137
//   compare_results                cpp_impl/sparseUtils.h:139-156
//   TCSC::TCSC                     cpp_impl/data_structures/TCSC.h:13-41
//   TCSR::TCSR                     cpp_impl/data_structures/TCSR.h:13-41
//   BlockedTCSC<512>               cpp_impl/data_structures/BlockedTCSC.h:15-43
//   BaseTCSC<float>                cpp_impl/comp.h:25-69
//   UnrolledTCSC<float,12>         cpp_impl/comp.h:179-265
//   DoubleUnrolledTCSC<float,4,4>  cpp_impl/comp.h:1227-1438   (fastest registered, main.cpp:125-130)
//   BaseTCSR<float>                cpp_impl/comp.h:478-528
//   BaseBlockedTCSC<float,512>     cpp_impl/comp.h:607-658
//   BaseTCSC_PreLU<float>          cpp_impl/comp_prelu.h:12-70
#include <memory>
#include <cstring>
#include <sstream>
#include <iostream>

#include "sparseUtils.h"
#include "common.h"
#include "comp.h"
#include "comp_prelu.h"

// common.h declares the registry; nothing here odr-uses it, but keep the linker happy if a
// future reference revision does.
void add_function(comp_func, std::string) {}
void add_prelu_function(comp_func_prelu, std::string) {}

namespace
{
template <typename V>
void copy_out(const V &v, int *dst)
{
    if (dst && !v.empty())
        std::memcpy(dst, v.data(), v.size() * sizeof(int));
}
} // namespace

extern "C"
{
    int ref_abi_version(void) { return 1; }

    // ---- inputs -----------------------------------------------------------------------------
    void ref_generate_sparse_matrix(int K, int N, int nonZero, int seed, int *out)
    {
        std::vector<int> w = generateSparseMatrix<int>(K, N, nonZero, false, seed);
        std::memcpy(out, w.data(), (size_t)K * N * sizeof(int));
    }

    // ---- TCSC ---------------------------------------------------------------------------------
    void *ref_tcsc_new(const int *W, int K, int N) { return new TCSC(W, K, N); }
    void ref_tcsc_free(void *h) { delete static_cast<TCSC *>(h); }
    void ref_tcsc_counts(void *h, long long *ncol_ptr, long long *npos, long long *nneg)
    {
        TCSC *t = static_cast<TCSC *>(h);
        *ncol_ptr = (long long)t->col_start_pos.size();
        *npos = (long long)t->row_index_pos.size();
        *nneg = (long long)t->row_index_neg.size();
    }
    void ref_tcsc_export(void *h, int *csp, int *csn, int *rip, int *rin)
    {
        TCSC *t = static_cast<TCSC *>(h);
        copy_out(t->col_start_pos, csp);
        copy_out(t->col_start_neg, csn);
        copy_out(t->row_index_pos, rip);
        copy_out(t->row_index_neg, rin);
    }
    int ref_tcsc_size_bytes(void *h) { return static_cast<TCSC *>(h)->getDataStructureSize(); }

    void ref_base_tcsc(void *h, float *X, float *b, float *Y, int M, int N, int K)
    {
        BaseTCSC<float>(X, *static_cast<TCSC *>(h), b, Y, M, N, K);
    }
    void ref_base_tcsc_prelu(void *h, float *X, float *b, float *alpha, float *Y, int M, int N, int K)
    {
        BaseTCSC_PreLU<float>(X, *static_cast<TCSC *>(h), b, alpha, Y, M, N, K);
    }
    void ref_unrolled_tcsc_12(void *h, float *X, float *b, float *Y, int M, int N, int K)
    {
        UnrolledTCSC<float, 12>(X, *static_cast<TCSC *>(h), b, Y, M, N, K);
    }
    void ref_double_unrolled_tcsc_k4_m4(void *h, float *X, float *b, float *Y, int M, int N, int K)
    {
        DoubleUnrolledTCSC<float, 4, 4>(X, *static_cast<TCSC *>(h), b, Y, M, N, K);
    }

    // ---- TCSR ---------------------------------------------------------------------------------
    void *ref_tcsr_new(const int *W, int K, int N) { return new TCSR(W, K, N); }
    void ref_tcsr_free(void *h) { delete static_cast<TCSR *>(h); }
    void ref_tcsr_counts(void *h, long long *nrow_ptr, long long *npos, long long *nneg)
    {
        TCSR *t = static_cast<TCSR *>(h);
        *nrow_ptr = (long long)t->row_start_pos.size();
        *npos = (long long)t->col_index_pos.size();
        *nneg = (long long)t->col_index_neg.size();
    }
    void ref_tcsr_export(void *h, int *rsp, int *rsn, int *cip, int *cin)
    {
        TCSR *t = static_cast<TCSR *>(h);
        copy_out(t->row_start_pos, rsp);
        copy_out(t->row_start_neg, rsn);
        copy_out(t->col_index_pos, cip);
        copy_out(t->col_index_neg, cin);
    }
    void ref_base_tcsr(void *h, float *X, float *b, float *Y, int M, int N, int K)
    {
        BaseTCSR<float>(X, *static_cast<TCSR *>(h), b, Y, M, N, K);
    }

    // ---- BlockedTCSC<512> (main.cpp:7,69) -----------------------------------------------------
    void *ref_blocked512_new(const int *W, int K, int N)
    {
        return new BlockedTCSC<512>(const_cast<int *>(W), K, N);
    }
    void ref_blocked512_free(void *h) { delete static_cast<BlockedTCSC<512> *>(h); }
    void ref_blocked512_counts(void *h, long long *nptr, long long *npos, long long *nneg)
    {
        auto *t = static_cast<BlockedTCSC<512> *>(h);
        *nptr = (long long)t->col_start_pos.size();
        *npos = (long long)t->row_index_pos.size();
        *nneg = (long long)t->row_index_neg.size();
    }
    void ref_blocked512_export(void *h, int *csp, int *csn, int *rip, int *rin)
    {
        auto *t = static_cast<BlockedTCSC<512> *>(h);
        copy_out(t->col_start_pos, csp);
        copy_out(t->col_start_neg, csn);
        copy_out(t->row_index_pos, rip);
        copy_out(t->row_index_neg, rin);
    }
    void ref_base_blocked512(void *h, float *X, float *b, float *Y, int M, int N, int K)
    {
        BaseBlockedTCSC<float, 512>(X, *static_cast<BlockedTCSC<512> *>(h), b, Y, M, N, K);
    }

    // ---- dense oracle + checker -----------------------------------------------------------------
    void ref_gemm(float *X, float *W, float *b, float *Y, int M, int N, int K)
    {
        GEMM<float>(X, W, b, Y, M, N, K);
    }
    void ref_gemm_prelu(float *X, float *W, float *b, float *alpha, float *Y, int M, int N, int K)
    {
        GEMM_PreLU<float>(X, W, b, alpha, Y, M, N, K);
    }
    // returns 1 on pass; the reference prints the first mismatch to cout — silence it here.
    int ref_compare_results(float *result, float *truth, int H, int W)
    {
        std::ostringstream sink;
        std::streambuf *old = std::cout.rdbuf(sink.rdbuf());
        bool ok = compare_results<float>(result, truth, H, W);
        std::cout.rdbuf(old);
        return ok ? 1 : 0;
    }
}
