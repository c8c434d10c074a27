// l2.h — float-descriptor (NORM_L2) brute-force kNN(2) matcher: tensor-core contraction with
// FP32 accumulate followed by an exact FP32 re-rank (BASELINE config 4).  Internal interface.
#pragma once
#include <cuda_runtime.h>
#include <string>

#include "../../include/mvslam_b200.h"

namespace mvs {

struct L2Workspace {
    void *buf[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t cap[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint64_t stats[4] = {0, 0, 0, 0};     // fallbacks fwd, fallbacks rev, gemm us, total us
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // call begin/end, GEMM begin/end per pass
    void release();
};

// idx/dist (knnMatch k=2 output) and/or filtered+sorted matches; host pointers.
int l2_knn2(L2Workspace &ws, cudaStream_t stream, const float *query, int nq, const float *train, int nt, int dim,
            int32_t *idx, float *dist, const mvs_match_params *mp, mvs_match *out, int capacity, int *n_out,
            int *n_launches, std::string &err);

}  // namespace mvs
