// tsg_dense_tc.cu — placeholder until the tcgen05 dense-expand kernel lands (next milestone).
#include "tsg_internal.cuh"
int tsg_launch_dense_tc(tsg_matrix *, const float *, int64_t, const float *, const float *,
                        float *, int64_t, int, cudaStream_t)
{
    tsg_set_error("TSG_ALGO_DENSE_TC is not built in this revision");
    return TSG_ERR_UNSUPPORTED;
}
