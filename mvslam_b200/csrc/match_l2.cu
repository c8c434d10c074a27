// match_l2.cu — placeholder until the tcgen05 L2 matcher lands (see l2.h).
#include "l2.h"

namespace mvs {

void L2Workspace::release()
{
    for (int i = 0; i < 8; ++i) { if (buf[i]) cudaFree(buf[i]); buf[i] = nullptr; cap[i] = 0; }
}

int l2_knn2(L2Workspace &, cudaStream_t, const float *, int, const float *, int, int, int32_t *, float *,
            const mvs_match_params *, mvs_match *, int, int *, int *n_launches, std::string &err)
{
    if (n_launches) *n_launches = 0;
    err = "float L2 matcher not built yet";
    return MVS_E_UNSUPPORTED;
}

}  // namespace mvs
