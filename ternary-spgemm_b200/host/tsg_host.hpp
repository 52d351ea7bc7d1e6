// tsg_host.hpp — C++ host mirror of the reference's plugin interface, backed by libtsg.so.
//
// What a maintainer of alessiomelone/Ternary-spGEMM includes to register CUDA-backed functions
// in the existing driver.  Everything here is a thin veneer over the C ABI
// in include/tsg.h; no arithmetic happens on the host.
//
//   comp_func / comp_func_prelu / add_function / add_prelu_function
//                         <- cpp_impl/common.h:12-16 (same signatures, argument order X,B,Y,M,N,K)
//   DataStructureInterface <- cpp_impl/data_structures/DataStructureInterface.hpp:4-14, plus the
//                            README's getNumRows/getNumCols (readme.md:62-72)
//   CudaTCSC              <- class TCSC, cpp_impl/data_structures/TCSC.h:5-50: same constructor,
//                            same public vectors, same getDataStructureSize(); and, unlike any
//                            format shipped by the reference, it DOES implement the interface, so
//                            cpp_impl/test_data_structure.cpp's test<T>() compiles against it.
//   CudaBaseTCSC / CudaBaseTCSC_PreLU
//                         <- BaseTCSC<T> cpp_impl/comp.h:25-26, BaseTCSC_PreLU<T> comp_prelu.h:12-13
//
// Error behaviour: the reference's functions return void and cannot fail; a failing libtsg call
// (no GPU, shape mismatch, CUDA error) prints tsg_last_error() and aborts — never a silent
// fallback to CPU code.
#pragma once

#include <cstdio>
#include <cstdlib>
#include <functional>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/tsg.h"

#ifdef TSG_IN_REFERENCE_TREE
// built inside the reference: its own headers provide these
#include "common.h"
#include "data_structures/DataStructureInterface.hpp"
#else
using comp_func = std::function<void(float *X, float *B, float *Y, int M, int N, int K)>;
using comp_func_prelu =
    std::function<void(float *X, float *B, float *alpha, float *Y, int M, int N, int K)>;

void add_function(comp_func f, std::string name);
void add_prelu_function(comp_func_prelu f, std::string name);

class DataStructureInterface
{
public:
    virtual ~DataStructureInterface() = default;
    virtual void init(const int *matrix, int rows, int cols) = 0;
    virtual std::vector<int> getVectorRepresentation(size_t rows, size_t cols) = 0;
};
#endif

namespace tsg
{
[[noreturn]] inline void die(const char *what, int status)
{
    std::fprintf(stderr, "libtsg: %s failed (status %d): %s\n", what, status, tsg_last_error());
    std::abort();
}
inline void check(int status, const char *what)
{
    if (status != TSG_OK)
        die(what, status);
}
} // namespace tsg

// Ternary CSC weight living in HBM.  Build it once (like the reference builds its formats in
// main.cpp:63-74), capture a shared_ptr to it in the registered lambda.
class CudaTCSC : public DataStructureInterface
{
public:
    // mirrors of the device arrays, filled by init() unless mirror_on_host is false
    std::vector<int> col_start_pos, col_start_neg, row_index_pos, row_index_neg;

    CudaTCSC() = default;
    CudaTCSC(const int *matrix, int rows, int cols, bool mirror_on_host = true)
        : mirror_(mirror_on_host)
    {
        init(matrix, rows, cols);
    }
    // the N-column shard [col_lo, col_hi) of the matrix (multi-GPU layout)
    CudaTCSC(const int *matrix, int rows, int cols, int col_lo, int col_hi, bool mirror_on_host = true)
        : mirror_(mirror_on_host)
    {
        reset();
        tsg::check(tsg_tcsc_from_dense_cols(matrix, rows, cols, col_lo, col_hi, &h_), "tsg_tcsc_from_dense_cols");
        pull();
    }
    CudaTCSC(const CudaTCSC &) = delete;
    CudaTCSC &operator=(const CudaTCSC &) = delete;
    ~CudaTCSC() override { reset(); }

    void init(const int *matrix, int rows, int cols) override
    {
        reset();
        tsg::check(tsg_tcsc_from_dense(matrix, rows, cols, &h_), "tsg_tcsc_from_dense");
        pull();
    }

    std::vector<int> getVectorRepresentation(size_t rows, size_t cols) override
    {
        if ((int)rows != getNumRows() || (int)cols != getNumCols())
        {
            std::fprintf(stderr, "CudaTCSC::getVectorRepresentation: asked for %zux%zu, matrix is %dx%d\n",
                         rows, cols, getNumRows(), getNumCols());
            std::abort();
        }
        std::vector<int> dense(rows * cols);
        tsg::check(tsg_tcsc_to_dense(h_, dense.data()), "tsg_tcsc_to_dense");
        return dense;
    }

    int getNumRows() const
    {
        int k = 0;
        tsg::check(tsg_rows(h_, &k), "tsg_rows");
        return k;
    }
    int getNumCols() const
    {
        int n = 0;
        tsg::check(tsg_cols(h_, &n), "tsg_cols");
        return n;
    }
    // TCSC::getDataStructureSize(), TCSC.h:43-49
    int getDataStructureSize() const
    {
        int64_t b = 0;
        tsg::check(tsg_data_structure_size(h_, &b), "tsg_data_structure_size");
        return (int)b;
    }
    long long nnz() const
    {
        int64_t p = 0, q = 0;
        tsg::check(tsg_nnz(h_, &p, &q), "tsg_nnz");
        return p + q;
    }
    tsg_matrix *handle() const { return h_; }

private:
    void reset()
    {
        if (h_)
            tsg_destroy(h_);
        h_ = nullptr;
    }
    void pull()
    {
        if (!mirror_)
            return;
        int64_t p = 0, q = 0;
        tsg::check(tsg_nnz(h_, &p, &q), "tsg_nnz");
        const int n = getNumCols();
        col_start_pos.resize(n + 1);
        col_start_neg.resize(n + 1);
        row_index_pos.resize((size_t)p);
        row_index_neg.resize((size_t)q);
        tsg::check(tsg_tcsc_export(h_, col_start_pos.data(), col_start_neg.data(), row_index_pos.data(),
                                   row_index_neg.data()),
                   "tsg_tcsc_export");
    }
    tsg_matrix *h_ = nullptr;
    bool mirror_ = true;
};

// Ternary CSR built on the GPU — class TCSR (cpp_impl/data_structures/TCSR.h:5-50): same public
// vectors and converting constructor, and it implements DataStructureInterface.
class CudaTCSR : public DataStructureInterface
{
public:
    std::vector<int> row_start_pos, row_start_neg, col_index_pos, col_index_neg;

    CudaTCSR() = default;
    CudaTCSR(const int *matrix, int rows, int cols) { init(matrix, rows, cols); }
    CudaTCSR(const CudaTCSR &) = delete;
    CudaTCSR &operator=(const CudaTCSR &) = delete;
    ~CudaTCSR() override { tsg_tcsr_destroy(h_); }

    void init(const int *matrix, int rows, int cols) override
    {
        tsg_tcsr_destroy(h_);
        h_ = nullptr;
        rows_ = rows, cols_ = cols;
        tsg::check(tsg_tcsr_from_dense(matrix, rows, cols, &h_), "tsg_tcsr_from_dense");
        int64_t p = 0, q = 0;
        tsg::check(tsg_tcsr_nnz(h_, &p, &q), "tsg_tcsr_nnz");
        row_start_pos.resize(rows + 1), row_start_neg.resize(rows + 1);
        col_index_pos.resize((size_t)p), col_index_neg.resize((size_t)q);
        tsg::check(tsg_tcsr_export(h_, row_start_pos.data(), row_start_neg.data(), col_index_pos.data(),
                                   col_index_neg.data()),
                   "tsg_tcsr_export");
    }
    std::vector<int> getVectorRepresentation(size_t rows, size_t cols) override
    {
        if ((int)rows != rows_ || (int)cols != cols_)
            tsg::die("CudaTCSR::getVectorRepresentation (shape mismatch)", TSG_ERR_INVALID);
        std::vector<int> dense(rows * cols);
        tsg::check(tsg_tcsr_to_dense(h_, dense.data()), "tsg_tcsr_to_dense");
        return dense;
    }
    int getNumRows() const { return rows_; }
    int getNumCols() const { return cols_; }
    int getDataStructureSize() const // TCSR.h:43-49
    {
        int64_t b = 0;
        tsg::check(tsg_tcsr_data_structure_size(h_, &b), "tsg_tcsr_data_structure_size");
        return (int)b;
    }
    tsg_tcsr *handle() const { return h_; }

private:
    tsg_tcsr *h_ = nullptr;
    int rows_ = 0, cols_ = 0;
};

// Packed-value formats (README "5 values into 8 bits", readme.md:108-111) built on the GPU: one
// class over the C-ABI family it wraps.  `ptr` is col_ptr (CSC, cols+1 entries) or row_ptr (CSR,
// rows+1), `idx` the merged row / column ids, `vals` five base-3 digits per byte (include/tsg.h).
struct PackedCscApi
{
    using handle_t = tsg_pcsc;
    static constexpr const char *name = "CudaPackedCSC";
    static int from_dense(const int *m, int r, int c, handle_t **h) { return tsg_pcsc_from_dense(m, r, c, h); }
    static void destroy(handle_t *h) { tsg_pcsc_destroy(h); }
    static int sizes(const handle_t *h, int64_t *n, int64_t *b) { return tsg_pcsc_sizes(h, n, b); }
    static int export_(const handle_t *h, int *p, int *i, unsigned char *v) { return tsg_pcsc_export(h, p, i, v); }
    static int to_dense(const handle_t *h, int *w) { return tsg_pcsc_to_dense(h, w); }
    static int ds_size(const handle_t *h, int64_t *b) { return tsg_pcsc_data_structure_size(h, b); }
    static int lists(int /*rows*/, int cols) { return cols; }
};
struct PackedCsrApi
{
    using handle_t = tsg_pcsr;
    static constexpr const char *name = "CudaPackedCSR";
    static int from_dense(const int *m, int r, int c, handle_t **h) { return tsg_pcsr_from_dense(m, r, c, h); }
    static void destroy(handle_t *h) { tsg_pcsr_destroy(h); }
    static int sizes(const handle_t *h, int64_t *n, int64_t *b) { return tsg_pcsr_sizes(h, n, b); }
    static int export_(const handle_t *h, int *p, int *i, unsigned char *v) { return tsg_pcsr_export(h, p, i, v); }
    static int to_dense(const handle_t *h, int *w) { return tsg_pcsr_to_dense(h, w); }
    static int ds_size(const handle_t *h, int64_t *b) { return tsg_pcsr_data_structure_size(h, b); }
    static int lists(int rows, int /*cols*/) { return rows; }
};

template <typename Api>
class CudaPacked : public DataStructureInterface
{
public:
    std::vector<int> ptr, idx;
    std::vector<unsigned char> vals;

    CudaPacked() = default;
    CudaPacked(const int *matrix, int rows, int cols) { init(matrix, rows, cols); }
    CudaPacked(const CudaPacked &) = delete;
    CudaPacked &operator=(const CudaPacked &) = delete;
    ~CudaPacked() override { Api::destroy(h_); }

    void init(const int *matrix, int rows, int cols) override
    {
        Api::destroy(h_);
        h_ = nullptr;
        rows_ = rows, cols_ = cols;
        tsg::check(Api::from_dense(matrix, rows, cols, &h_), Api::name);
        int64_t nnz = 0, nb = 0;
        tsg::check(Api::sizes(h_, &nnz, &nb), Api::name);
        ptr.resize(Api::lists(rows, cols) + 1), idx.resize((size_t)nnz), vals.resize((size_t)nb);
        tsg::check(Api::export_(h_, ptr.data(), idx.data(), vals.data()), Api::name);
    }
    std::vector<int> getVectorRepresentation(size_t rows, size_t cols) override
    {
        if ((int)rows != rows_ || (int)cols != cols_)
            tsg::die("CudaPacked::getVectorRepresentation (shape mismatch)", TSG_ERR_INVALID);
        std::vector<int> dense(rows * cols);
        tsg::check(Api::to_dense(h_, dense.data()), Api::name);
        return dense;
    }
    int getNumRows() const { return rows_; }
    int getNumCols() const { return cols_; }
    int getDataStructureSize() const
    {
        int64_t b = 0;
        tsg::check(Api::ds_size(h_, &b), Api::name);
        return (int)b;
    }
    typename Api::handle_t *handle() const { return h_; }

private:
    typename Api::handle_t *h_ = nullptr;
    int rows_ = 0, cols_ = 0;
};
using CudaPackedCSC = CudaPacked<PackedCscApi>;
using CudaPackedCSR = CudaPacked<PackedCsrApi>;

// Y = X·W + b on the GPU.  Same call shape as BaseTCSC<T>(X, W_csc, b, Y, M, N, K).
template <typename T, int ALGO = TSG_ALGO_AUTO>
void CudaBaseTCSC(T *X, const CudaTCSC &W, T *b, T *Y, int M, int N, int K)
{
    static_assert(std::is_same<T, float>::value, "libtsg computes in fp32 like the reference's registered kernels");
    tsg::check(tsg_spmm_algo(W.handle(), ALGO, X, b, nullptr, Y, M, N, K), "tsg_spmm");
}

// Same call shape as BaseTCSR<T>(X, W_csr, b, Y, M, N, K) (comp.h:478-479).
template <typename T, int ALGO = TSG_ALGO_AUTO>
void CudaBaseTCSR(T *X, const CudaTCSR &W, T *b, T *Y, int M, int N, int K)
{
    static_assert(std::is_same<T, float>::value, "libtsg computes in fp32");
    tsg::check(tsg_tcsr_spmm(W.handle(), ALGO, X, b, nullptr, Y, M, N, K), "tsg_tcsr_spmm");
}

// Y = X·W + b from the packed-value CSC handle.
template <typename T, int ALGO = TSG_ALGO_AUTO>
void CudaPackedCSC_spmm(T *X, const CudaPackedCSC &W, T *b, T *Y, int M, int N, int K)
{
    static_assert(std::is_same<T, float>::value, "libtsg computes in fp32");
    tsg::check(tsg_pcsc_spmm(W.handle(), ALGO, X, b, nullptr, Y, M, N, K), "tsg_pcsc_spmm");
}

// Y = X·W + b from the packed-value CSR handle (TSG_ALGO_PCSR_SEQ: BaseTCSR's order).
template <typename T, int ALGO = TSG_ALGO_AUTO>
void CudaPackedCSR_spmm(T *X, const CudaPackedCSR &W, T *b, T *Y, int M, int N, int K)
{
    static_assert(std::is_same<T, float>::value, "libtsg computes in fp32");
    tsg::check(tsg_pcsr_spmm(W.handle(), ALGO, X, b, nullptr, Y, M, N, K), "tsg_pcsr_spmm");
}

// Fused bias + PReLU.  Same call shape as BaseTCSC_PreLU<T>(X, W_csc, b, alpha, Y, M, N, K).
template <typename T, int ALGO = TSG_ALGO_AUTO>
void CudaBaseTCSC_PreLU(T *X, const CudaTCSC &W, T *b, T *alpha, T *Y, int M, int N, int K)
{
    static_assert(std::is_same<T, float>::value, "libtsg computes in fp32 like the reference's registered kernels");
    if (alpha == nullptr)
        tsg::die("tsg_spmm_prelu (alpha is NULL)", TSG_ERR_INVALID);
    tsg::check(tsg_spmm_algo(W.handle(), ALGO, X, b, alpha, Y, M, N, K), "tsg_spmm_prelu");
}
