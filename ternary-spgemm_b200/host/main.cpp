// main.cpp — the benchmark driver of the B200 engine: same CLI, same registry, same stdout
// grammar as the reference's cpp_impl/main.cpp, so its harnesses (plots/run_benchmark.py,
// run_benchmark.py) parse our output unchanged.
//
//   ./sparseGEMM.out -M <int> -K <int> -N <int> -s <int> [-correctness]
//
// Positional like the reference (argv[2],[4],[6],[8]; argv[9] == "-correctness";
// main.cpp:43-57).  Differences, all additive:
//   * the registered functions are CUDA-backed lambdas (CudaBaseTCSC<...> over a CudaTCSC built
//     on the device).  The function named "BaseTCSC" — the Speedup denominator, main.cpp:10,259 —
//     is the reference-ORDER kernel (TSG_ALGO_GATHER_SEQ, bit-identical to the CPU BaseTCSC);
//   * built with REF=<reference tree> (TSG_WITH_REFERENCE) the reference's own CPU BaseTCSC and
//     DoubleUnrolledTCSC_K4_M4 are registered first, from its headers compiled in place, so one
//     binary prints CPU and GPU numbers side by side;
//   * TSG_SEED=<int> in the environment makes W and X reproducible (the reference seeds with
//     time(0), sparseUtils.h:10,54);
//   * the dense O(MKN) checker runs only under -correctness (the reference runs it always,
//     main.cpp:201-204) and none of the unused CPU formats is built (VectorTCSC's constructor
//     is O(N²K), main.cpp:67);
//   * -DINSTRUMENTATION_RUN (make INSTRUMENT=1) prints Flops / Performance / Total Input Size /
//     Operational Intensity / Data Structure Size like main.cpp:264-271, with the flop count
//     M·(nnz + N) that the instrumented BaseTCSC counts (comp.h:48-66).
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "perf_timer.hpp"
#include "tsg_host.hpp"

#ifdef TSG_WITH_REFERENCE
#include "comp.h"
#include "comp_prelu.h"
#include "sparseUtils.h"
#else
#include "sparse_utils.hpp"
#endif

#define BENCHMARK_FUNCTION_NAME "BaseTCSC"

std::vector<comp_func> userFuncs;
std::vector<std::string> funcNames;
int numFuncs = 0;

std::vector<comp_func_prelu> userFuncs_prelu;
std::vector<std::string> funcNames_prelu;
int numFuncs_prelu = 0;

void add_function(comp_func f, std::string name)
{
    userFuncs.push_back(std::move(f));
    funcNames.emplace_back(std::move(name));
    numFuncs++;
}

void add_prelu_function(comp_func_prelu f, std::string name)
{
    userFuncs_prelu.push_back(std::move(f));
    funcNames_prelu.emplace_back(std::move(name));
    numFuncs_prelu++;
}

namespace
{
template <int ALGO>
void add_cuda(const std::shared_ptr<CudaTCSC> &w, const std::string &name)
{
    int picked = ALGO;
    // register only kernels that support this shape (probe with a tiny call is not needed:
    // unsupported shapes abort loudly, so ask the library first)
    if (ALGO == TSG_ALGO_AUTO)
        tsg::check(tsg_spmm_pick(w->handle(), 1, &picked), "tsg_spmm_pick");
    add_function([w](float *X, float *B, float *Y, int M, int N, int K)
                 { CudaBaseTCSC<float, ALGO>(X, *w, B, Y, M, N, K); },
                 name);
    add_prelu_function([w](float *X, float *B, float *alpha, float *Y, int M, int N, int K)
                       { CudaBaseTCSC_PreLU<float, ALGO>(X, *w, B, alpha, Y, M, N, K); },
                       name + "_PreLU");
}

void report_instrumented(long long flops, float cycles, int M, int K, int N, int ds_bytes, bool prelu)
{
#ifdef INSTRUMENTATION_RUN
    std::cout << "Flops: " << flops << std::endl;
    std::cout << "Performance: " << (float)flops / cycles << " flops/cycle" << std::endl;
    float total_bytes = sizeof(float) * ((float)(M * K + M * N + N + (prelu ? N : 0))) + (float)ds_bytes;
    std::cout << "Total Input Size: " << (int)total_bytes << " Bytes" << std::endl;
    std::cout << "Operational Intensity: " << (float)flops / total_bytes << " Flops/Byte" << std::endl;
    std::cout << "Data Structure Size: " << ds_bytes << " Bytes" << std::endl;
#else
    (void)flops, (void)cycles, (void)M, (void)K, (void)N, (void)ds_bytes, (void)prelu;
#endif
}
} // namespace

int main(int argc, char **argv)
{
    std::cout << "Starting program. ";
    float perf_val;
    int i_loop;

    if (argc < 9)
    {
        fprintf(stderr, "Usage: %s -M <int> -K <int> -N <int> -s <int>\n", argv[0]);
        return 1;
    }
    const int M = atoi(argv[2]), K = atoi(argv[4]), N = atoi(argv[6]), nonZero = atoi(argv[8]);
    const bool check_correctness = argc > 9 && std::string(argv[9]) == "-correctness";
    const char *seed_env = getenv("TSG_SEED");
    const int seed = seed_env ? atoi(seed_env) : -1;

    std::vector<int> W_raw = generateSparseMatrix<int>(K, N, nonZero, false, seed);

    // the one format the hot path needs, built on the device
    auto sf_cuda = std::make_shared<CudaTCSC>(W_raw.data(), K, N, /*mirror_on_host=*/false);
    const int ds_bytes = sf_cuda->getDataStructureSize();
    const long long nnz = sf_cuda->nnz();

#ifdef TSG_WITH_REFERENCE
    // the reference's own CPU functions, from its headers, registered first (main.cpp:76-81,125-130)
    auto sf_csc = std::make_shared<TCSC>(W_raw.data(), K, N);
    add_function([sf_csc](float *X, float *B, float *Y, int Ma, int Na, int Ka)
                 { BaseTCSC<float>(X, *sf_csc, B, Y, Ma, Na, Ka); },
                 "BaseTCSC");
    add_function([sf_csc](float *X, float *B, float *Y, int Ma, int Na, int Ka)
                 { DoubleUnrolledTCSC<float, 4, 4>(X, *sf_csc, B, Y, Ma, Na, Ka); },
                 "DoubleUnrolledTCSC_K4_M4");
    add_prelu_function([sf_csc](float *X, float *B, float *al, float *Y, int Ma, int Na, int Ka)
                       { BaseTCSC_PreLU<float>(X, *sf_csc, B, al, Y, Ma, Na, Ka); },
                       "BaseTCSC_PreLU");
    add_cuda<TSG_ALGO_GATHER_SEQ>(sf_cuda, "CudaTCSC_seq");
#else
    add_cuda<TSG_ALGO_GATHER_SEQ>(sf_cuda, "BaseTCSC"); // reference order on the GPU; the Speedup base
#endif
    add_cuda<TSG_ALGO_GATHER>(sf_cuda, "CudaTCSC_gather");
    add_cuda<TSG_ALGO_DENSE_TC>(sf_cuda, "CudaTCSC_denseTC");
    if (M <= 4)
        add_cuda<TSG_ALGO_CODE_GEMV>(sf_cuda, "CudaTCSC_codeGEMV");
    add_cuda<TSG_ALGO_AUTO>(sf_cuda, "CudaTCSC_auto");
    // the two other formats (TSG_FORMATS=1 in the environment: they are built from the same W)
    if (getenv("TSG_FORMATS"))
    {
        auto sf_csr = std::make_shared<CudaTCSR>(W_raw.data(), K, N);
        add_function([sf_csr](float *X, float *B, float *Y, int Ma, int Na, int Ka)
                     { CudaBaseTCSR<float, TSG_ALGO_TCSR_SEQ>(X, *sf_csr, B, Y, Ma, Na, Ka); },
                     "CudaTCSR_seq");
        auto sf_pcsc = std::make_shared<CudaPackedCSC>(W_raw.data(), K, N);
        add_function([sf_pcsc](float *X, float *B, float *Y, int Ma, int Na, int Ka)
                     { CudaPackedCSC_spmm<float, TSG_ALGO_PCSC_GATHER>(X, *sf_pcsc, B, Y, Ma, Na, Ka); },
                     "CudaPackedCSC_gather");
        auto sf_pcsr = std::make_shared<CudaPackedCSR>(W_raw.data(), K, N);
        add_function([sf_pcsr](float *X, float *B, float *Y, int Ma, int Na, int Ka)
                     { CudaPackedCSR_spmm<float, TSG_ALGO_PCSR_SEQ>(X, *sf_pcsr, B, Y, Ma, Na, Ka); },
                     "CudaPackedCSR_seq");
    }

    if (numFuncs == 0 && numFuncs_prelu == 0)
    {
        std::cout << std::endl;
        std::cout << "No functions registered - nothing for driver to do" << std::endl;
        return 0;
    }

    std::cout << numFuncs << " regular functions and " << numFuncs_prelu << " PrelU functions registered." << std::endl;

    if (check_correctness)
    {
        std::vector<float> X_main = initX<float>(M * K, 512);
        std::vector<float> W_FP32_main(W_raw.begin(), W_raw.end());
        std::vector<float> B_main(N, 2);
        std::vector<float> alpha_main(N, 0.1);
        std::vector<float> Y_main((size_t)M * N, 0);
        std::vector<float> refY_main((size_t)M * N, 0);
        std::vector<float> refY_prelu_main((size_t)M * N, 0);
        GEMM(X_main.data(), W_FP32_main.data(), B_main.data(), refY_main.data(), M, N, K);
        GEMM_PreLU(X_main.data(), W_FP32_main.data(), B_main.data(), alpha_main.data(), refY_prelu_main.data(), M, N, K);

        for (i_loop = 0; i_loop < numFuncs; i_loop++)
        {
            std::fill(Y_main.begin(), Y_main.end(), 0);
            userFuncs[i_loop](X_main.data(), B_main.data(), Y_main.data(), M, N, K);
            if (compare_results(Y_main.data(), refY_main.data(), M, N))
                std::cout << "Test case " << funcNames[i_loop] << " passed!" << std::endl;
            else
            {
                std::cout << "Test case " << "\x1b[31m" << funcNames[i_loop] << " failed!" << "\x1b[0m" << std::endl;
                std::cout << "\n  The result differs from the dense GEMM check (compare_results, abs 1e-5); stopping.\n"
                          << std::endl;
                exit(1);
            }
        }
        for (i_loop = 0; i_loop < numFuncs_prelu; i_loop++)
        {
            std::fill(Y_main.begin(), Y_main.end(), 0);
            userFuncs_prelu[i_loop](X_main.data(), B_main.data(), alpha_main.data(), Y_main.data(), M, N, K);
            if (compare_results(Y_main.data(), refY_prelu_main.data(), M, N))
                std::cout << "Test case " << funcNames_prelu[i_loop] << " passed!" << std::endl;
            else
            {
                std::cout << "Test case " << "\x1b[31m" << funcNames_prelu[i_loop] << " failed!" << "\x1b[0m" << std::endl;
                std::cout << "\n  The result differs from the dense GEMM check (compare_results, abs 1e-5); stopping.\n"
                          << std::endl;
                exit(1);
            }
        }
    }

    const long long flops = (long long)M * (nnz + N); // what instrumented BaseTCSC counts
    float base_cycles = 0;
    for (i_loop = 0; i_loop < numFuncs; i_loop++)
    {
        perf_val = perf_test(userFuncs[i_loop], M, K, N, nonZero);
        std::cout << "\nRunning: " << "\x1b[31m" << funcNames[i_loop] << "\x1b[0m" << std::endl;
        std::cout << perf_val << " cycles" << std::endl;
        if (funcNames[i_loop] == BENCHMARK_FUNCTION_NAME)
            base_cycles = perf_val;
        std::cout << "Speedup is: " << "\x1b[32m" << base_cycles / perf_val << "\x1b[0m" << std::endl;
        report_instrumented(flops, perf_val, M, K, N, ds_bytes, false);
    }

    float base_cycles_prelu = 0;
    for (i_loop = 0; i_loop < numFuncs_prelu; i_loop++)
    {
        perf_val = perf_test_prelu(userFuncs_prelu[i_loop], M, K, N, nonZero);
        std::cout << "\nRunning: " << "\x1b[31m" << funcNames_prelu[i_loop] << "\x1b[0m" << std::endl;
        std::cout << perf_val << " cycles" << std::endl;
        if (funcNames_prelu[i_loop] == "BaseTCSC_PreLU")
            base_cycles_prelu = perf_val;
        std::cout << "Speedup is: " << "\x1b[32m" << base_cycles_prelu / perf_val << "\x1b[0m" << std::endl;
        report_instrumented(flops, perf_val, M, K, N, ds_bytes, true);
    }
    return 0;
}
