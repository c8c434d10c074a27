// kernels.h — argument blocks and host-side launchers of the device stages (internal).
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"

#include <utility>

namespace mvs {

// frame table resident in HBM: descriptors as 2 x uint4 (32 B) per keypoint, keypoints as float2
struct KnnArgs {
    const uint4 *desc;
    const int32_t *frame_off;  // first keypoint of each frame in desc / kp
    const int32_t *frame_cnt;
    const int2 *pairs;         // (base, pair) per batch entry; nullptr: frame 0 = train, frame 1 = query
    uint2 *partial;            // [pairs][splits][q_stride] packed (best1, best2) keys
    int q_stride;
    int reverse;               // 1: roles swapped (cross-check pass)
    uint32_t bound;            // bounded search: distances >= bound are "far" (0 = evaluate everything)
};

// tensor-core matcher (match_hamming_tc.cu): same frame table, descriptors expanded to one byte per bit ([rows][256])
struct TcKnnArgs {
    const int32_t *frame_off;
    const int32_t *frame_cnt;
    const int2 *pairs;         // as KnnArgs
    uint2 *partial;            // [pairs][tc_splits(shape)][q_stride]: column halves of the train tiles x train splits
    int q_stride;
    int reverse;
    int t_splits;              // train-dimension splits (tc_train_splits of the launch shape; partial holds 2 x t_splits entries per query)
    int q_tiles, n_pairs;      // filled by the launcher: items = q_tiles x t_splits x n_pairs
    int one;                   // 1 (opaque to the compiler, see the epilogue)
};

struct FinalizeArgs {
    const int32_t *frame_off;
    const int32_t *frame_cnt;
    const int2 *pairs;
    const uint2 *partial; int splits; int q_stride;
    const uint2 *rev_partial; int rev_splits; int rev_stride;
    double ratio, max_dist;
    uint32_t bound;            // bounded search: recorded distances >= bound only mean "far" (0 = exact everywhere)
    const float2 *kp;          // nullptr: no gather
    double Kinv[9];
    mvs_match *matches;        // [pairs][q_stride]
    int32_t *n_matches;        // [pairs]
    double *points;            // [pairs][q_stride][6] = (x1,y1,z1,x2,y2,z2) or nullptr
    PairState *state;          // [pairs] or nullptr
    int32_t *knn_idx, *knn_dist;  // optional raw knnMatch output [pairs][q_stride][2]
    // tensor-core matcher: the second key of a partial is the best of all column streams OTHER than the best's (an upper
    // bound of the true second distance); K2 then evaluates the best's 15 stream-mates from the raw descriptors where the
    // result depends on them (refine_second).  desc = the frame table's 32-byte descriptors, nullptr = partials are exact.
    const uint4 *refine_desc;
};

struct NormArgs { double Kinv[9]; };

struct HypArgs {
    const double *points;      // [pairs][p_stride][6]
    int p_stride;
    const PairState *state;    // n_matches / status per pair
    const uint32_t *table;     // explicit [H][8] sample table shared by all pairs, or nullptr -> seeded sampler
    uint64_t seed;
    uint64_t pair_id_base;
    int H;
    double *F_all;             // [pairs][H][9]
    int solver;                // MVS_SOLVER_*
    int n_pairs;               // filled by the launcher
};

struct ScoreArgs {
    const double *points; int p_stride;
    const PairState *state;
    const double *F_all; int H;
    double max_error_sq;
    double zc1, zc2;           // constant z of image 1 / image 2 points (const-z kernels)
    int tiles;                 // point tiles per pair
    uint32_t *part_count;      // [pairs][tiles][H]
    double *part_res;          // [pairs][tiles][H] residual sums (ALGEBRAIC mode only, else nullptr)
    int solver;
    uint32_t *zero_word;       // optional: a counter this launch clears for the kernels after it (SelectArgs::item_total)
};

struct SelectArgs {
    const double *points; int p_stride;
    PairState *state;
    const double *F_all; int H;
    const uint32_t *part_count; int tiles;
    const double *part_res;    // per-tile residual sums when K4 produced them (ALGEBRAIC), else nullptr
    int32_t *ties;             // scratch [pairs][2][H]: total counts, then the hypotheses sharing the best count
    double max_error_sq;
    double zc1, zc2;
    int min_inliers;
    int decompose;             // 0: stop after the mask (mvs_ransac_fundamental)
    uint8_t *mask;             // [pairs][p_stride]
    int32_t *all_counts;       // optional [pairs][H]
    int solver;
    // work list of the triangulation (decompose only): one entry (pair << 32 | match index) per inlier of every pair
    // that has a pose to recover, appended in any order; valid[] of the other matches is cleared here
    unsigned long long *items; uint32_t *item_total;
    uint8_t *valid;            // [pairs][4][p_stride]
};

struct TriArgs {
    const double *points; int p_stride;
    PairState *state;
    const uint8_t *mask;       // nullptr: all ones
    int n_cand;                // 4 (recover_pose_and_points) or 1 (sfm_triangulate: Rc[0], tc as given)
    uint8_t *valid;            // [pairs][4][p_stride]
    double *tri;               // [pairs][4][p_stride][3]
    int solver;
    const unsigned long long *items; const uint32_t *item_total;   // launch_triangulate_items: K5's work list
};

struct FinishArgs {
    PairState *state; int p_stride;
    int n_cand;
    const uint8_t *valid; const double *tri;
    const mvs_match *matches;  // optional, for match_inlier_ssd
    double *out_points;        // [pairs][p_stride][3]
    uint64_t *out_index;       // [pairs][p_stride]
    mvs_pair_result *results;  // [pairs]
};

void launch_knn2_hamming(const KnnArgs &a, int max_nq, int splits, int n_pairs, cudaStream_t s);
int tc_max_train();            // largest train set of the tensor-core matcher
int tc_train_splits(int max_nq, int max_nt, int n_pairs);   // train-dimension splits for a launch of this shape (1 for large batches)
int tc_splits(int max_nq, int max_nt, int n_pairs);         // partial top-2 pairs per query the kernel writes: 2 x tc_train_splits
void launch_expand_desc(const uint4 *desc, size_t row_begin, size_t n_rows, void *desc8, cudaStream_t s);
cudaError_t launch_knn2_hamming_tc(const void *desc8, size_t total_rows, const TcKnnArgs &a, int max_nq, int n_pairs, cudaStream_t s);
int finalize_sort_capacity(int max_nq);
cudaError_t launch_match_finalize(const FinalizeArgs &a, int max_nq, int n_pairs, cudaStream_t s);
void launch_normalize_points(const double *xy1, const double *xy2, int n, const NormArgs &a, double *pts, PairState *st,
                             cudaStream_t s);

void launch_hypotheses(const HypArgs &a, int n_pairs, cudaStream_t s);
void launch_fundamental_sets(const double *p1s, const double *p2s, int n_sets, double *F_out, int solver, cudaStream_t s);
bool launch_svd_batch(int n, const double *A, int count, int solver, double *U, double *w, double *Vt, cudaStream_t s);
int score_tiles(int max_points);
void launch_score(const ScoreArgs &a, int mode, bool const_z, int n_pairs, cudaStream_t s);
void launch_select(const SelectArgs &a, int mode, bool const_z, int max_points, int n_pairs, cudaStream_t s);
void launch_triangulate(const TriArgs &a, int max_points, int n_pairs, cudaStream_t s);
void launch_finish(const FinishArgs &a, int max_points, int n_pairs, cudaStream_t s);
cudaError_t launch_triangulate_items(const TriArgs &a, size_t max_items, cudaStream_t s);

bool pdl_enabled();   // api.cu: on unless MVS_PDL=0 (an A/B switch for tools/latency_probe, not a documented knob)

// <<<grid, block, smem, s>>> with the programmatic-stream-serialization attribute (see pdl_wait, common.cuh)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_dep(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args &&...args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace mvs
