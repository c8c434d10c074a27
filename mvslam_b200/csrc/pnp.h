// pnp.h — pnp_solve stage (reference source/vision/pnp-solve.cpp:16-104).  Internal interface between api.cu and pnp.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "../../include/mvslam_b200.h"

namespace mvs {

struct PnpArgs {
    const double *world;       // [total][3]
    const double *image;       // [total][2]
    const int32_t *offsets;    // [problems + 1] first point of each problem
    const uint32_t *table;     // explicit [H][4] sample table shared by all problems, or nullptr -> seeded sampler
    uint64_t seed, problem_id_base;
    int H, tiles, refine_iters, min_inliers;
    double fx, fy, cx, cy, thr2;
    double *poses;             // [problems][H][12] world->camera (R row-major, t)
    uint8_t *valid;            // [problems][H]
    int32_t *part_count;       // [problems][tiles][H]
    uint8_t *mask_ws;          // [total] inliers of the winning hypothesis
    uint8_t *mask;             // optional user-visible copy (device)
    int32_t *all_counts;       // optional [problems][H]
    mvs_pnp_result *results;   // [problems]
};

int pnp_tiles(int max_points);
void launch_pnp(const PnpArgs &a, int n_problems, int max_points, cudaStream_t s);
void pnp_sample_table_host(uint64_t seed, uint64_t problem_id, uint32_t n_points, int H, uint32_t *out);

}  // namespace mvs
