// tsg_bitplane.cu — placeholder until the bit-plane kernel lands (next milestone).
#include "tsg_internal.cuh"
int tsg_launch_bitplane(tsg_matrix *, const float *, int64_t, const float *, const float *,
                        float *, int64_t, int, cudaStream_t)
{
    tsg_set_error("TSG_ALGO_BITPLANE is not built in this revision");
    return TSG_ERR_UNSUPPORTED;
}
