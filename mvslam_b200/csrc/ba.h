// ba.h — bundle adjustment stage (reference source/vision/ba.cpp:26-156).  Internal interface between api.cu and ba.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mvslam_b200.h"

namespace mvs {

struct BaArgs {
    double fx, fy, sk, u0, v0;          // Cal3_S2
    const int32_t *frame_off;           // [problems + 1]
    const int32_t *point_off;           // [problems + 1]
    const double *pose_R, *pose_t;      // guesses = prior means, [frames][9] / [frames][3], camera to world
    const double *pose_prior_cov;       // [frames][36], NaN in [0] = no prior
    const double *points;               // guesses = prior means [points][3]
    const double *point_prior_cov;      // [points][9], NaN in [0] = no prior
    const mvs_ba_observation *obs;      // sorted by (problem, point); frame / point local to the problem
    const int32_t *point_obs_off;       // [points + 1] observation range of every point
    double *ws;                         // [points][48]
    // problems with more than two frames (ba_solve_multi_kernel): per point 12 + 36 F doubles at ws_multi + ws_multi_off[problem],
    // per observation 27 doubles (its camera block and gradient) in ws_obs
    double *ws_multi; const long long *ws_multi_off; double *ws_obs;
    double *pose_R_out, *pose_t_out, *pose_cov_out, *points_out, *point_cov_out;
    mvs_ba_result *results;
    int max_iter;
    double lambda0, rel_tol, abs_tol;   // abs_tol < 0: no absolute test
};

constexpr int BA_MAX_FRAMES = 16;
// max_frames: the largest frame count of the batch (problems of <= 2 frames and of 3..BA_MAX_FRAMES frames run in two kernels)
cudaError_t launch_ba(const BaArgs &a, int n_problems, int min_frames, int max_frames, cudaStream_t s);

}  // namespace mvs
