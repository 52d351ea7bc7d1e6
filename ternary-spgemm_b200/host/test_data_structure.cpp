// test_data_structure.cpp — round-trip test of the device-built format through
// DataStructureInterface, modelled on the reference's cpp_impl/test_data_structure.cpp:
//   W -> T.init(W,K,N) -> T.getVectorRepresentation(K,N) == W          (its test<T>(), :47-74)
// over the reference's own fixture set: the tiny verbose case (K=3,N=4,s=2,seed=0, :149), the
// exhaustive small sweep (testMany, :76-108) and the benchmark shapes × s∈{2,4,8,16}
// (testRequired, :110-145).  Exit code 0 iff everything matches.
#include <cstdio>
#include <type_traits>

#include "sparse_utils.hpp"
#include "tsg_host.hpp"

template <typename T>
bool test(int K, int N, int nonZero, int seed, bool verbose = false)
{
    static_assert(std::is_base_of<DataStructureInterface, T>::value, "T must inherit from DataStructureInterface");
    auto W_truth = generateSparseMatrix<int>(K, N, nonZero, false, seed);
    T m;
    m.init(&W_truth[0], K, N);
    auto back = m.getVectorRepresentation(K, N);
    if (verbose)
        std::printf(back == W_truth ? "pass\n" : "fail\n");
    return back == W_truth;
}

int main(int argc, char **argv)
{
    const bool full = argc > 1; // any argument: also the large benchmark shapes
    int bad = 0, ran = 0;
    bad += !test<CudaTCSC>(3, 4, 2, 0, true), ++ran;
    bad += !test<CudaTCSR>(3, 4, 2, 0, true), ++ran;
    bad += !test<CudaPackedCSC>(3, 4, 2, 0, true), ++ran;
    bad += !test<CudaPackedCSR>(3, 4, 2, 0, true), ++ran;
    for (int k = 1; k < 10; ++k)
        for (int n = 2; n < 10; ++n)
            for (int seed = 0; seed < 3; ++seed)
                if (n / 2 >= 1) // generator needs room for its +1/-1 quota
                    bad += !test<CudaTCSC>(k, n, 2, seed), ++ran;
    const int Ks[] = {512, 1024, 2048, 4096, 2048, 4096, 8192, 16384};
    const int Ns[] = {2048, 4096, 8192, 16384, 512, 1024, 2048, 4096};
    const int ss[] = {2, 4, 8, 16};
    for (int i = 0; i < (full ? 8 : 2); ++i)
        for (int s : ss)
        {
            const bool ok = test<CudaTCSC>(Ks[i], Ns[i], s, i) && test<CudaTCSR>(Ks[i], Ns[i], s, i) &&
                            test<CudaPackedCSC>(Ks[i], Ns[i], s, i) && test<CudaPackedCSR>(Ks[i], Ns[i], s, i);
            if (!ok)
                std::printf("Mismatch at k=%d, n=%d, nonZero=%d\n", Ks[i], Ns[i], s);
            bad += !ok, ++ran;
        }
    std::printf(bad ? "%d of %d round trips FAILED\n" : "All vectors match! (%d of %d failed)\n", bad, ran);
    return bad ? 1 : 0;
}
