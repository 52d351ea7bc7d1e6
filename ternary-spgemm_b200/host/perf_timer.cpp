// perf_timer.cpp — see perf_timer.hpp.  Calibration follows the reference's scheme
// (cpp_impl/perf.cpp:28-30,45-60): double the repeat count until one batch takes at least
// CYCLES_REQUIRED cycles, capped at 2^14 repeats; compile with -DNO_CALIBRATE for a single run.
#include "perf_timer.hpp"

#include <chrono>
#include <cstdint>
#include <vector>

#ifdef TSG_WITH_REFERENCE
#include "sparseUtils.h"
#else
#include "sparse_utils.hpp"
#endif

#if defined(__x86_64__)
#include <x86intrin.h>
static inline uint64_t cycles_now()
{
    unsigned aux;
    _mm_lfence();
    uint64_t t = __rdtscp(&aux);
    _mm_lfence();
    return t;
}
#else
static inline uint64_t cycles_now()
{
    // no cycle counter: report nanoseconds scaled to a nominal 3.2 GHz like perf.cpp:30,337
    using namespace std::chrono;
    return (uint64_t)(duration_cast<nanoseconds>(steady_clock::now().time_since_epoch()).count() * 3.2);
}
#endif

namespace
{
constexpr double kCyclesRequired = 1e8;
constexpr int kMaxRuns = 1 << 14;
double g_last_seconds = 0.0;

template <typename Call>
float timed(Call &&call)
{
    int runs = 1;
#ifndef NO_CALIBRATE
    while (runs < kMaxRuns)
    {
        const uint64_t t0 = cycles_now();
        for (int i = 0; i < runs; ++i)
            call();
        if ((double)(cycles_now() - t0) >= kCyclesRequired)
            break;
        runs *= 2;
    }
#endif
    const auto w0 = std::chrono::steady_clock::now();
    const uint64_t t0 = cycles_now();
    for (int i = 0; i < runs; ++i)
        call();
    const uint64_t cyc = (cycles_now() - t0) / (uint64_t)runs;
    g_last_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count() / runs;
    return (float)cyc;
}
} // namespace

double perf_last_seconds_per_call() { return g_last_seconds; }

float perf_test(comp_func f, int M, int K, int N, int /*nonZero*/)
{
    std::vector<float> X = initX<float>(M * K, 512);
    std::vector<float> Y((size_t)M * N + 10, 0.0f);
    std::vector<float> B((size_t)N, 2.0f);
    X.insert(X.end(), 10, 0.0f);
    return timed([&] { f(X.data(), B.data(), Y.data(), M, N, K); });
}

float perf_test_prelu(comp_func_prelu f, int M, int K, int N, int /*nonZero*/)
{
    std::vector<float> X = initX<float>(M * K, 512);
    std::vector<float> Y((size_t)M * N + 10, 0.0f);
    std::vector<float> B((size_t)N, 2.0f);
    std::vector<float> A((size_t)N, 0.1f);
    X.insert(X.end(), 10, 0.0f);
    return timed([&] { f(X.data(), B.data(), A.data(), Y.data(), M, N, K); });
}
