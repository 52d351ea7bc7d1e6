// sparse_utils.hpp — input generators and the dense checker used by OUR driver (host/main.cpp).
//
// Same contracts as the reference's cpp_impl/sparseUtils.h so that seeds reproduce its matrices
// bit-for-bit under libstdc++ (std::mt19937 + std::uniform_int_distribution call sequence):
//   initX                 sparseUtils.h:6-23    integers in [-Range, Range] stored as T
//   generateSparseMatrix  sparseUtils.h:25-90   per row: N/s/2 ± v entries of +1 / -1
//   GEMM / GEMM_PreLU     sparseUtils.h:92-137  dense checker  y = Σ_k X[m,k]·W[k,n] (+b, PReLU)
//   compare_results       sparseUtils.h:139-156 |r - g| > 10e-6 -> report first mismatch, fail
// When the driver is built inside the reference tree (TSG_WITH_REFERENCE) the reference's own
// header is used instead and this file is not included.
#pragma once
#include <cmath>
#include <ctime>
#include <iostream>
#include <random>
#include <vector>

template <typename T>
std::vector<T> initX(int LEN, int Range, bool uniformDistribution = false, long seed = -1)
{
    std::vector<T> x((size_t)LEN, T(0));
    if (uniformDistribution)
    {
        for (auto &v : x)
            v = T(rand() % Range);
        return x;
    }
    std::mt19937 eng(static_cast<unsigned int>(seed < 0 ? time(0) : seed));
    std::uniform_int_distribution<int> pick(-Range, Range);
    for (auto &v : x)
        v = T(pick(eng));
    return x;
}

template <typename T>
std::vector<T> generateSparseMatrix(int H, int W, int nonZero, bool uniformDistribution, int seed = -1)
{
    if (seed != -1)
        srand(seed);
    std::vector<T> w((size_t)H * W, T(0));
    if (uniformDistribution)
    {
        // one +1 and one -1 in every window of 2*nonZero columns
        for (int h = 0; h < H; ++h)
            for (int c0 = 0; c0 < W; c0 += nonZero * 2)
            {
                int a = rand() % nonZero * 2, b = rand() % nonZero * 2;
                w[(size_t)h * W + c0 + a] = T(1);
                while (b == a)
                    b = rand() % nonZero * 2;
                w[(size_t)h * W + c0 + b] = T(-1);
            }
        return w;
    }
    std::mt19937 eng(static_cast<unsigned int>(seed == -1 ? time(0) : seed));
    std::uniform_int_distribution<int> column(0, W - 1);
    std::uniform_int_distribution<int> skew(0, int(W / nonZero / 20 + 1));
    for (int h = 0; h < H; ++h)
    {
        T *row = w.data() + (size_t)h * W;
        const int v = skew(eng);
        const int want[2] = {(W / nonZero) / 2 + v, (W / nonZero) / 2 - v};
        const T value[2] = {T(1), T(-1)};
        for (int sgn = 0; sgn < 2; ++sgn)
            for (int placed = 0; placed < want[sgn];)
            {
                const int c = column(eng);
                if (row[c] == T(0))
                {
                    row[c] = value[sgn];
                    ++placed;
                }
            }
    }
    return w;
}

template <typename T>
void GEMM(T *X, T *W, T *b, T *Y, int M, int N, int K)
{
#pragma omp parallel for
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n)
        {
            T acc = 0;
            for (int k = 0; k < K; ++k)
                acc += X[(size_t)m * K + k] * W[(size_t)k * N + n];
            Y[(size_t)m * N + n] = acc + b[n];
        }
}

template <typename T>
void GEMM_PreLU(T *X, T *W, T *b, T *alpha, T *Y, int M, int N, int K)
{
#pragma omp parallel for
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n)
        {
            T acc = 0;
            for (int k = 0; k < K; ++k)
                acc += X[(size_t)m * K + k] * W[(size_t)k * N + n];
            const T pre = acc + b[n];
            Y[(size_t)m * N + n] = (pre >= 0) ? pre : alpha[n] * pre;
        }
}

template <typename T>
bool compare_results(T *result, T *groundTruth, int H, int W)
{
    for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w)
        {
            const size_t i = (size_t)h * W + w;
            if (std::abs(result[i] - groundTruth[i]) > 10e-6)
            {
                std::cout << "Error at: H=" << h << ", W=" << w << ", result=" << result[i]
                          << ", groundTruth=" << groundTruth[i] << std::endl;
                return false;
            }
        }
    return true;
}
