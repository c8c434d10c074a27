/*
 * mvslam_b200.h — C ABI of libmvslam_b200.so: the B200 (sm_100a) implementation of mvSLAM's
 * two-view front-end hot path (descriptor matching -> RANSAC fundamental/essential matrix ->
 * pose recovery -> linear triangulation) and of the steps on either side of it:
 * feature extraction (cv::ORB), pnp_solve (cv::solvePnPRansac with P3P) and the one-/two-frame bundle
 * adjustment behind sfm_refine / pnp_refine.
 *
 * The reference (lonelycorn/mvSLAM) has no plugin/FFI layer; its boundary for this path is a set
 * of C++ entry points.  Each function below names the reference interface it replaces
 * (file:line relative to the reference tree).  the headers under include/mvslam/ wrap this ABI back into the
 * reference's C++ signatures (namespace mvSLAM); INTEGRATION.md shows how a maintainer binds it.
 *
 * Conventions
 *   - plain C types only; all matrices row-major doubles (the adapters convert from Eigen's
 *     column-major storage); points are AoS.
 *   - the caller owns every host buffer; outputs are written only on success of the stage that
 *     produces them (sfm-solve.cpp:364-366 swaps outputs in only on success).
 *   - every entry point is synchronous w.r.t. the host unless its name ends in _enqueue;
 *     a ctx is bound to one device and one stream and is NOT thread-safe (one ctx per thread).
 *   - there is NO CPU fallback: without a CUDA device mvs_create fails with MVS_E_CUDA.
 *   - host pointers may be pageable or pinned; pinned buffers make the copies asynchronous.
 */
#ifndef MVSLAM_B200_H
#define MVSLAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVS_ABI_VERSION 3

/* status codes (reference: bool / assert / empty vector; SURVEY.md §8b "Errors") */
enum {
    MVS_OK = 0,
    MVS_E_BAD_ARG = 1,          /* precondition the reference asserts on (visual-feature.cpp:56, sfm-solve.cpp:37-41,292) */
    MVS_E_TOO_FEW_POINTS = 2,   /* < 8 correspondences (estimator-RANSAC.cpp:25-29) */
    MVS_E_NO_MODEL = 3,         /* inlier_count_best == 0 (estimator-RANSAC.cpp:89) */
    MVS_E_TOO_FEW_INLIERS = 4,  /* < VF_MATCH_INLIER_MIN (sfm-solve.cpp:326-334) */
    MVS_E_NO_CHEIRALITY = 5,    /* no candidate leaves a point in front of both cameras (sfm-solve.cpp:345-356) */
    MVS_E_CUDA = 6,             /* CUDA runtime/driver failure; see mvs_last_error */
    MVS_E_CAPACITY = 7,         /* caller-provided output capacity too small */
    MVS_E_UNSUPPORTED = 8
};

enum { MVS_SCORE_ALGEBRAIC = 0, /* |x2^T F x1| < max_error_sq, estimator-RANSAC.cpp:114-117 (parity mode) */
       MVS_SCORE_SAMPSON = 1    /* Sampson distance, the score of cv::findEssentialMat used by the default build */ };

/* How the geometry is solved.
 * MVS_SOLVER_REFERENCE (0, the default of a zeroed mvs_ransac_params and of the adapters in include/mvslam/):
 *   the reference's own arithmetic, literally.  The 8-point null vector is vt.row(8) of cv::SVDecomp(A^T A)
 *   (fundamental-matrix.cpp:104-118) and every SVD (3x3 rank-2 / essential projection / decomposition, 4x4 DLT) is
 *   OpenCV's small-matrix routine restated operation for operation, with unfused left-to-right products elsewhere:
 *   F, E, inlier mask, pose and points are bit-identical to the numpy + cv2.SVDecomp restatement of the reference
 *   (oracle/oracle_np.py), ill-conditioned samples included.
 * MVS_SOLVER_FAST (1): same estimator, numerically better and ~50x cheaper per hypothesis: the null vector comes
 *   from a Householder QR of A itself (no squared condition number), SVDs use a round-robin Jacobi with explicit
 *   FMAs.  Agrees with the reference to ~1e-9 on well-conditioned samples; on ill-conditioned ones (sigma_8(A)
 *   ~1e-5) it returns the accurate null vector where the reference returns A^T A round-off (up to 4e-5 apart). */
enum { MVS_SOLVER_REFERENCE = 0, MVS_SOLVER_FAST = 1 };

typedef struct mvs_ctx mvs_ctx;

/* cv::DMatch as the reference uses it (base/image.hpp:31-35): queryIdx indexes the second/pair
 * frame, trainIdx the first/base frame (visual-feature.cpp:59-60), imgIdx is always 0. */
typedef struct {
    int32_t query;
    int32_t train;
    float   distance;
} mvs_match;

/* constants of visual-feature.cpp:21-24 and ImagePair::Params (image-pair.hpp:25-31) */
typedef struct {
    double  ratio;        /* NEAREST_NEIGHBOR_DIST_RATIO, 0.7 */
    double  max_dist;     /* < 0: keep all (match_visual_features default -1; VO default 10) */
    int32_t cross_check;  /* reference: CROSS_CHECK=false */
    int32_t bounded;      /* 1 (needs max_dist >= 0): early-abandon train descriptors that are provably too far to
                             change the filtered result (distance >= B, B = smallest integer with ratio*B > max_dist):
                             the returned matches are bit-identical, only the work shrinks.  0 = evaluate every pair. */
} mvs_match_params;

/* constants of sfm-solve.cpp:18-23,67 made explicit */
typedef struct {
    int32_t  n_hypotheses; /* H >= 1; H == 1 is the reference (single sample {0..7}) */
    int32_t  score_mode;   /* MVS_SCORE_* */
    double   max_error_sq; /* <= 0: reference default 5e-2 / (K00*K11), sfm-solve.cpp:311 */
    uint64_t seed;         /* seeds the rows >= 1 of the sample table (mvs_sample_table) */
    int32_t  min_inliers;  /* <= 0: VF_MATCH_INLIER_MIN = 8 */
    int32_t  solver;       /* MVS_SOLVER_*; 0 = MVS_SOLVER_REFERENCE */
    uint64_t pair_id_base; /* batch entry i samples with pair_id = pair_id_base + i (sharding-invariant results) */
} mvs_ransac_params;

/* fixed-size record of one solved pair == what ImagePair holds after reconstruct()
 * (image-pair.hpp:52-60) plus the intermediate models, all doubles row-major */
typedef struct {
    int32_t status;
    int32_t n_matches;        /* M: ratio/max_dist survivors (== n for mvs_sfm_solve) */
    int32_t n_inliers;        /* inliers of the winning hypothesis */
    int32_t best_hypothesis;  /* row of the sample table that won */
    int32_t n_points;         /* triangulated points that passed cheirality == ImagePair::match_inlier_count */
    int32_t candidate;        /* 0..3 = (Ra,+t),(Ra,-t),(Rb,+t),(Rb,-t), sfm-solve.cpp:259-281 */
    double  residual;         /* sum of residuals over the inliers of the winner */
    double  F[9];             /* winning de-normalised 8-point model */
    double  E[9];             /* after the (s,s,0) projection, sfm-solve.cpp:73-87 */
    double  R1to2[9];
    double  t1to2[3];
    double  R2in1[9];         /* pose2in1 = SE3(SO3(R1to2), t1to2).inverse(), sfm-solve.cpp:364 */
    double  t2in1[3];
    uint64_t match_inlier_ssd; /* sum of squared descriptor distances over the reconstructed points (image-pair.cpp:166) */
} mvs_pair_result;

/* cv::KeyPoint as cv::ORB fills it (class_id is always -1 and is not carried) */
typedef struct {
    float   x, y;      /* pt, level-0 pixel coordinates (level coordinates times the level scale) */
    float   size;     /* patchSize (31) times the level scale */
    float   angle;    /* degrees, [0, 360) */
    float   response; /* Harris response */
    int32_t octave;   /* pyramid level */
} mvs_keypoint;

/* cv::ORB::create(nfeatures) — the only parameter the reference sets (MAX_FEATURE_COUNT = 500,
 * source/vision/visual-feature.cpp:9-17); the rest are OpenCV's defaults: scaleFactor 1.2f, nlevels 8,
 * edgeThreshold 31, firstLevel 0, WTA_K 2, HARRIS_SCORE, patchSize 31, fastThreshold 20. */
typedef struct {
    int32_t n_features;
    int32_t reserved[3];
} mvs_orb_params;

/* constants of pnp-solve.cpp:48-52 made explicit */
typedef struct {
    int32_t  n_hypotheses;       /* <= 0: 100 (iterationsCount) */
    int32_t  refine_iterations;  /* < 0: 10 Gauss-Newton steps on the inliers; 0: return the winning minimal-sample pose */
    double   reprojection_error; /* <= 0: 0.05 pixels (reprojectionError) */
    uint64_t seed;               /* seeds rows >= 1 of the 4-point sample table (mvs_pnp_sample_table) */
    uint64_t problem_id_base;    /* batch entry i samples with problem_id = problem_id_base + i */
    int32_t  min_inliers;        /* <= 0: 4 */
    int32_t  reserved;
} mvs_pnp_params;

typedef struct {
    int32_t status;              /* MVS_OK, MVS_E_TOO_FEW_POINTS (< 4), MVS_E_NO_MODEL */
    int32_t n_points;
    int32_t n_inliers;           /* consensus of the winning minimal-sample pose */
    int32_t best_hypothesis;
    double  R_c2w[9], t_c2w[3];          /* pose: camera to world, SE3(R, t).inverse() of pnp-solve.cpp:99-101 */
    double  R_w2c_p3p[9], t_w2c_p3p[3];  /* the winning P3P pose (world to camera) before refinement */
} mvs_pnp_result;

/* one GenericProjectionFactor of ba.cpp:96-117: image point z with covariance [[cov[0], cov[1]], [cov[1], cov[2]]]
 * of point `point` seen by camera `frame` (both indexes local to their problem) */
typedef struct {
    int32_t frame, point;
    double  uv[2];
    double  cov[3];
} mvs_ba_observation;

/* gtsam::LevenbergMarquardtParams as far as they matter here */
typedef struct {
    int32_t max_iterations;      /* <= 0: 100 (GTSAM default) */
    int32_t reserved;
    double  lambda_initial;      /* <= 0: 1e-5 (GTSAM default) */
    double  relative_tolerance;  /* <= 0: 1e-5 (GTSAM relativeErrorTol): stop when the cost decreases by less than this fraction */
    double  absolute_tolerance;  /* 0: 1e-5 (GTSAM absoluteErrorTol): stop when the cost decreases by less than this; < 0: no such test.
                                    ba.cpp:124 builds LevenbergMarquardtOptimizer with default parameters, so zeros reproduce its
                                    stopping rule; {1e-13, -1} iterates to the minimum itself (what the parity tests use) */
} mvs_ba_params;

typedef struct {
    int32_t status;              /* MVS_OK, MVS_E_UNSUPPORTED (more than 2 frames), MVS_E_BAD_ARG (prior covariance not PD) */
    int32_t iterations;
    double  initial_error;       /* cost at the guesses */
    double  final_error;         /* cost at the result == optimizer.error(), ba.cpp:154 */
} mvs_ba_result;

/* per-stage device time accumulated on the ctx stream while profiling is enabled */
enum { MVS_STAGE_KNN = 0, MVS_STAGE_MATCH_FINALIZE, MVS_STAGE_HYPOTHESES, MVS_STAGE_SCORE,
       MVS_STAGE_SELECT, MVS_STAGE_TRIANGULATE, MVS_STAGE_FINALIZE, MVS_STAGE_L2,
       MVS_STAGE_ORB_PYRAMID, MVS_STAGE_ORB_FAST, MVS_STAGE_ORB_HARRIS, MVS_STAGE_ORB_SELECT,
       MVS_STAGE_ORB_BLUR, MVS_STAGE_ORB_DESCRIBE, MVS_STAGE_PNP, MVS_STAGE_BA, MVS_N_STAGES };
typedef struct {
    double   ms[MVS_N_STAGES];
    uint64_t launches[MVS_N_STAGES];
} mvs_profile;

/* ---- context ---------------------------------------------------------------------------- */
int  mvs_abi_version(void);
const char *mvs_status_string(int status);
/* Creates a context on CUDA device `device` (negative: the calling thread's current device, cudaGetDevice -- what a rank of
 * a multi-GPU job that has already selected its GPU wants) with its own non-blocking stream.  Replaces the
 * reference's hidden globals (static matcher visual-feature.cpp:12-25, global camera
 * camera-manager.cpp:10). */
int  mvs_create(mvs_ctx **out, int device);
void mvs_destroy(mvs_ctx *ctx);
const char *mvs_last_error(const mvs_ctx *ctx);
/* Run all work of this ctx on a caller-owned cudaStream_t (e.g. torch's current stream). */
int  mvs_set_stream(mvs_ctx *ctx, void *cuda_stream);
int  mvs_synchronize(mvs_ctx *ctx);
int  mvs_profile_enable(mvs_ctx *ctx, int on);
int  mvs_profile_read(mvs_ctx *ctx, mvs_profile *out, int reset);
/* number of kernels this ctx has launched so far */
uint64_t mvs_kernel_launches(const mvs_ctx *ctx);
/* Page-locked host memory for callers that do not link the CUDA runtime themselves (the reference is plain C++): buffers
 * from here make the library's host<->device copies asynchronous, and pinned detail outputs of mvs_pair_batch are written
 * by the device directly (only the entries each pair owns).  NULL on failure.  Free with mvs_host_free. */
void *mvs_host_alloc(size_t bytes);
void  mvs_host_free(void *p);
/* Test hook.  With MVS_GUARD=1 in the environment when the library is loaded, every workspace buffer is allocated at
 * exactly the requested size plus a 4 KB guard band; this returns how many buffers had their band overwritten since
 * they were allocated (0 = no kernel wrote past the end of its buffer), -1 when guard mode is off. */
int mvs_debug_guard_check(mvs_ctx *ctx);
/* Guard mode only: writes one byte past the end of the first live buffer, so that a test can see the check fire. */
int mvs_debug_guard_poke(mvs_ctx *ctx);

/* ---- feature extraction: VisualFeature::extract (source/vision/visual-feature.cpp:40-49, decl
 *      visual-feature.hpp:14) = cv::ORB detect + compute, the step FrameManager::add_frame runs per new image
 *      (source/front-end/frame-manager.cpp:107-125) ------------------------------------------------------------ */
/* n_images 8-bit grayscale images of one size (images[i] -> height rows of stride_bytes).  Keypoints come out
 * compact: image i owns the slots [sum(counts[0..i)), +counts[i]) of keypoints[] / descriptors[][32]; `capacity` is the
 * number of slots the caller provided (MVS_E_CAPACITY, with counts[] filled, when too small).  counts, keypoints and
 * descriptors are each optional.  Order within an image (new contract; cv::ORB's order is whatever std::nth_element
 * leaves): pyramid level ascending, then y, then x in level coordinates.  The keypoint SET, responses, angles and
 * descriptor bytes are those of cv::ORB (see oracle/orb_np.py for the pin).
 * append_frames != 0: every image also becomes a new frame of the resident frame table (as mvs_frames_append would
 * make it, but device to device: descriptors never visit the host); *first_frame = index of image 0's frame.
 * A pyramid level keeps at most 4096 keypoints (quota plus ties at the cut-off response): MVS_E_CAPACITY beyond. */
int mvs_orb_extract(mvs_ctx *ctx, const uint8_t *const *images, int n_images, int width, int height, int stride_bytes,
                    const mvs_orb_params *params, int append_frames, int32_t *first_frame,
                    int32_t *counts, mvs_keypoint *keypoints, uint8_t *descriptors, int64_t capacity);
/* Same with the images already in device memory: d_images is [n_images][height][stride_bytes] contiguous. */
int mvs_orb_extract_device(mvs_ctx *ctx, const void *d_images, int n_images, int width, int height, int stride_bytes,
                           const mvs_orb_params *params, int append_frames, int32_t *first_frame,
                           int32_t *counts, mvs_keypoint *keypoints, uint8_t *descriptors, int64_t capacity);

/* ---- matching: VisualFeature::match_visual_features (source/vision/visual-feature.cpp:51-80,
 *      decl visual-feature.hpp:23-26) ------------------------------------------------------------ */
/* cv::BFMatcher(NORM_HAMMING).knnMatch(query, train, k=2) (visual-feature.cpp:59-62):
 * idx/dist are [nq][2]; ties resolve to the lowest train index. desc_bytes must be 32. nt >= 2. */
int mvs_knn2_hamming(mvs_ctx *ctx, const uint8_t *query, int nq, const uint8_t *train, int nt,
                     int desc_bytes, int32_t *idx, int32_t *dist);
/* knnMatch + Lowe ratio + max_dist (+ optional cross-check) + sort.  Output order: distance
 * ascending, then queryIdx ascending (the reference's std::sort leaves ties unspecified). */
int mvs_match_hamming(mvs_ctx *ctx, const uint8_t *query, int nq, const uint8_t *train, int nt,
                      int desc_bytes, const mvs_match_params *params,
                      mvs_match *out, int capacity, int *n_out);
/* float descriptors, NORM_L2 (BASELINE config 4): tensor-core contraction + exact FP32 re-rank.  query / train / idx / dist may
 * be host OR device pointers (unified addressing): with descriptors already resident in HBM the call is the kernels alone
 * (0.45 ms at 32768 x 32768 x 64 against 1.4 ms with the two 8 MB uploads from pageable memory). */
int mvs_knn2_l2(mvs_ctx *ctx, const float *query, int nq, const float *train, int nt, int dim,
                int32_t *idx, float *dist);
int mvs_match_l2(mvs_ctx *ctx, const float *query, int nq, const float *train, int nt, int dim,
                 const mvs_match_params *params, mvs_match *out, int capacity, int *n_out);
/* Diagnostics of the last L2 call: out[0] = queries whose TF32-safety proof failed and were recomputed by the
 * exact brute-force kernel (forward pass), out[1] = same for the cross-check pass, out[2] = device time of the
 * tensor-core kernel(s) in microseconds, out[3] = device time of the whole call in microseconds. */
int mvs_l2_stats(const mvs_ctx *ctx, uint64_t out[4]);

/* ---- geometry --------------------------------------------------------------------------- */
/* find_fundamental_matrix (source/vision/fundamental-matrix.cpp:204-267, decl fundamental-matrix.hpp:16-19)
 * for n_sets independent 8-point samples: p1s/p2s are [n_sets][8][3], F_out is [n_sets][9]. */
int mvs_find_fundamental_matrix(mvs_ctx *ctx, const double *p1s, const double *p2s, int n_sets, int solver /* MVS_SOLVER_* */,
                                double *F_out);

/* SVD<M> (source/math/svd.hpp:13-73: cv::SVDecomp(MODIFY_A | FULL_UV), U, w descending, vt) for `count` square n x n
 * matrices A[count][n*n] (row-major); U, Vt are [count][n*n], w is [count][n].  MVS_SOLVER_REFERENCE returns
 * cv::SVDecomp's own bits for n = 3, 4, 9 (the sizes on the path: F/E, the DLT, A^T A); MVS_SOLVER_FAST is the
 * round-robin Jacobi the fast solver uses and exists for n = 3 only (MVS_E_UNSUPPORTED otherwise). */
int mvs_svd_batch(mvs_ctx *ctx, int n, const double *A, int count, int solver, double *U, double *w, double *Vt);

/* The seeded sample table shared bit-for-bit by the device sampler and the CPU oracle: row 0 is
 * {0..7} (the reference's only sample, estimator-RANSAC.cpp:41-48), rows >= 1 hold 8 distinct
 * indices < n_points.  Pure host function. out is [H][8]. */
void mvs_sample_table(uint64_t seed, uint64_t pair_id, uint32_t n_points, int H, uint32_t *out);

/* FundamentalMatrixEstimatorRANSAC::compute (source/vision/estimator-RANSAC.cpp:16-90, decl
 * estimator-RANSAC.hpp:20-24) over an explicit sample table samples[H][8] (NULL: seeded table of
 * params->seed, pair_id 0).  p1/p2 are [n][3] homogeneous ideal-camera points.
 * all_counts (optional) receives the inlier count of every hypothesis. */
int mvs_ransac_fundamental(mvs_ctx *ctx, const double *p1, const double *p2, int n,
                           const uint32_t *samples, const mvs_ransac_params *params,
                           double F[9], uint8_t *inlier_mask, int *inlier_count, double *residual,
                           int *best_hypothesis, int32_t *all_counts);

/* sfm_solve (source/vision/sfm-solve.cpp:285-368, decl source/vision/sfm.hpp:30-35), own branch of
 * find_essential_matrix (:64-90).  xy1/xy2 are [n][2] pixel coordinates, K row-major.
 * points [capacity][3], indexes [capacity]; inlier_mask [n] optional. */
int mvs_sfm_solve(mvs_ctx *ctx, const double *xy1, const double *xy2, int n, const double K[9],
                  const mvs_ransac_params *params, const uint32_t *samples,
                  mvs_pair_result *result, uint8_t *inlier_mask,
                  double *points, uint64_t *indexes, int capacity);

/* sfm_triangulate (source/vision/sfm-solve.cpp:370-394, decl sfm.hpp:47-53); poses are
 * camera-to-world (R row-major as held by the SO3, t). */
int mvs_sfm_triangulate(mvs_ctx *ctx, const double *xy1, const double *xy2, int n, const double K[9],
                        const double R1[9], const double t1[3], const double R2[9], const double t2[3],
                        int solver /* MVS_SOLVER_*: which 4x4 SVD solves the DLT */,
                        double *points, uint64_t *indexes, int capacity, int *n_out);

/* ---- pnp_solve (source/vision/pnp-solve.cpp:16-104, decl source/vision/pnp.hpp:22-26) = cv::solvePnPRansac with
 *      SOLVEPNP_P3P: seeded 4-point samples, P3P + 4th-point disambiguation per hypothesis, squared reprojection
 *      error <= reprojection_error^2 consensus over all hypotheses x points, refinement on the inliers.
 *      VisualOdometer::track_pnp (source/front-end/visual-odometer.cpp:503-615) is the caller. ---- */
/* row 0 = {0,1,2,3}; rows >= 1 hold 4 distinct indices < n_points.  Pure host function, out is [H][4]. */
void mvs_pnp_sample_table(uint64_t seed, uint64_t problem_id, uint32_t n_points, int H, uint32_t *out);
/* world [n][3], image [n][2] (pixels), K row-major (fx, fy, cx, cy are used, as cv::projectPoints does).
 * samples: optional explicit [H][4] table (H = params->n_hypotheses).  inlier_mask [n] optional.
 * all_counts (optional, [H]) receives the consensus size of every hypothesis. */
int mvs_pnp_solve(mvs_ctx *ctx, const double *world, const double *image, int n, const double K[9],
                  const mvs_pnp_params *params, const uint32_t *samples, mvs_pnp_result *result,
                  uint8_t *inlier_mask, int32_t *all_counts);
/* n_problems independent problems in one pass: problem i owns counts[i] consecutive rows of world / image /
 * inlier_mask.  results[n_problems] always filled (status per problem). */
int mvs_pnp_solve_batch(mvs_ctx *ctx, const double *world, const double *image, const int32_t *counts, int n_problems,
                        const double K[9], const mvs_pnp_params *params, const uint32_t *samples,
                        mvs_pnp_result *results, uint8_t *inlier_mask);

/* ---- ba_frame_pose_and_point (source/vision/ba.cpp:26-156, decl source/vision/ba.hpp:25-36), the optimisation behind
 *      sfm_refine (sfm-refine.cpp:20-139), pnp_refine (pnp-refine.cpp:16-110) and VisualOdometer::track_refine
 *      (visual-odometer.cpp:640-800): Levenberg-Marquardt over 1 .. 16 camera poses (camera to world; the reference's callers
 *      pass one or two, which have their own kernel) and their points
 *      with Gaussian priors (prior mean = the guess) and projection factors; estimates come back with the marginal
 *      covariances gtsam::Marginals would give (blocks of the inverse Gauss-Newton Hessian).
 *      n_problems independent problems per call; problem i owns n_frames[i] / n_points[i] / n_obs[i] consecutive rows
 *      of the pose / point / observation arrays.  K = Cal3_S2(K[0], K[4], K[1], K[2], K[5]).
 *      pose_prior_cov [frames][36] (tangent order: rotation, translation — GTSAM's Pose3) and point_prior_cov
 *      [points][9]: a NaN in the first element means "no prior".  Outputs are optional except results. ---- */
int mvs_ba_solve_batch(mvs_ctx *ctx, int n_problems, const double K[9],
                       const int32_t *n_frames, const int32_t *n_points, const int32_t *n_obs,
                       const double *pose_R, const double *pose_t, const double *pose_prior_cov,
                       const double *points, const double *point_prior_cov, const mvs_ba_observation *obs,
                       const mvs_ba_params *params,
                       double *pose_R_out, double *pose_t_out, double *pose_cov_out,
                       double *points_out, double *point_cov_out, mvs_ba_result *results);

/* ---- batched image pairs: ImagePair::ImagePair + reconstruct (source/front-end/image-pair.cpp:30-71,
 *      115-174) for many (base, pair) frame pairs per call; the natural batch of
 *      VisualOdometer::initialize (visual-odometer.cpp:289-296) and of all-pairs reconstruction ---- */
/* Make n_frames frames resident in HBM: desc[f] is [counts[f]][32] bytes (cv::Mat CV_8U rows),
 * kp[f] is [counts[f]][2] float (cv::KeyPoint::pt).  Replaces the previous frame table. */
int mvs_frames_upload(mvs_ctx *ctx, int n_frames, const uint8_t *const *desc, const float *const *kp,
                      const int32_t *counts, int desc_bytes);
/* The same table from two contiguous host arrays (frame f's rows follow frame f-1's; counts[f] rows each): two copies
 * instead of two per frame, and the call does not wait for them -- with pinned buffers it returns at once, so one
 * context can upload the next chunk of a long sequence while another one matches the previous chunk (bench.py,
 * e2e_distinct).  The buffers must stay valid until the next synchronising call on this ctx. */
int mvs_frames_upload_packed(mvs_ctx *ctx, int n_frames, const uint8_t *desc_all, const float *kp_all,
                             const int32_t *counts, int desc_bytes);
/* Append one frame to the resident table (what FrameManager::add_frame does per new image,
 * source/front-end/frame-manager.cpp:107-125); returns its index through *frame_index.  mvs_frames_upload
 * with n_frames = 0 is not allowed: start a new sequence with mvs_frames_clear. */
int mvs_frames_append(mvs_ctx *ctx, const uint8_t *desc, const float *kp, int32_t count, int desc_bytes, int32_t *frame_index);
int mvs_frames_clear(mvs_ctx *ctx);
/* Solve pairs[i] = (base frame, pair frame) for i < n_pairs against the resident frames.
 * results[n_pairs] always filled (status per pair; one bad pair never aborts the batch).
 * Optional per-pair detail outputs use a common stride `capacity` (>= 1): matches[n_pairs][capacity],
 * inlier_mask[n_pairs][capacity], points[n_pairs][capacity][3], indexes[n_pairs][capacity] (index into that
 * pair's matches).  Entries of a pair beyond its counts (n_matches for matches / inlier_mask, n_points for points /
 * indexes) are unspecified.  Only the first `capacity` entries of a pair are copied back: a pair whose
 * results[i].n_matches exceeds `capacity` is truncated (pass the largest pair-frame keypoint count to rule
 * that out; the VO default max_dist = 10 leaves ~100 matches per 2k-keypoint pair). */
int mvs_pair_batch(mvs_ctx *ctx, const int32_t *pairs, int n_pairs, const double K[9],
                   const mvs_match_params *mparams, const mvs_ransac_params *rparams,
                   mvs_pair_result *results, mvs_match *matches, uint8_t *inlier_mask,
                   double *points, uint64_t *indexes, int capacity);
/* Same, but only enqueues the device work and the device->host copies on the ctx stream;
 * the host buffers are valid after mvs_synchronize(). */
int mvs_pair_batch_enqueue(mvs_ctx *ctx, const int32_t *pairs, int n_pairs, const double K[9],
                           const mvs_match_params *mparams, const mvs_ransac_params *rparams,
                           mvs_pair_result *results, mvs_match *matches, uint8_t *inlier_mask,
                           double *points, uint64_t *indexes, int capacity);

/* ---- multi-GPU: pairs are independent (image-pair.cpp:30-71,143), so the pair list is cut into contiguous slices, one per
 *      rank (one process per GPU), every rank keeps the frame table resident, and the only communication is at the end:
 *      one gather of the fixed-size records and one of the variable-length clouds, placed by the exclusive scan of the
 *      per-pair counts.  NCCL (libnccl.so.2) is bound at run time; nothing here is needed on one GPU. */
typedef struct mvs_comm mvs_comm;
/* ncclGetUniqueId: call on one rank, hand the 128 bytes to the others by any means (file, environment, MPI, ...). */
int  mvs_comm_unique_id(uint8_t id[128]);
/* ncclCommInitRank on ctx's device; collective over the `world` ranks. */
int  mvs_comm_create(mvs_comm **out, mvs_ctx *ctx, const uint8_t id[128], int rank, int world);
void mvs_comm_destroy(mvs_comm *comm);
/* contiguous slice [lo, hi) of n units owned by `rank` (sizes differ by at most one) */
void mvs_shard_bounds(int64_t n, int world, int rank, int64_t *lo, int64_t *hi);
/* Every rank passes the SAME full pair list (frames resident on every rank: mvs_frames_upload / mvs_orb_extract) and solves
 * its slice; pair i samples with pair_id = pair_id_base + i whatever the sharding, so the gathered bytes equal a single-GPU
 * run.  On `root`: results[n_pairs_total]; if point_offsets != NULL (all ranks must agree) also the clouds of all pairs,
 * compacted: pair i owns points[point_offsets[i] .. point_offsets[i+1]) (and indexes, into that pair's matches);
 * likewise matches through match_offsets.  offsets arrays have n_pairs_total + 1 entries.  MVS_E_CAPACITY (with the offsets
 * filled) when a capacity is too small.  Other ranks may pass NULL for every output but must pass non-NULL offsets pointers
 * (any small buffer) to take part in the second gather. */
int  mvs_pair_batch_sharded(mvs_ctx *ctx, mvs_comm *comm, const int32_t *pairs, int64_t n_pairs_total, const double K[9],
                            const mvs_match_params *mparams, const mvs_ransac_params *rparams, int root,
                            mvs_pair_result *results, int64_t *point_offsets, double *points, uint64_t *indexes,
                            int64_t point_capacity, int64_t *match_offsets, mvs_match *matches, int64_t match_capacity);

#ifdef __cplusplus
}
#endif
#endif /* MVSLAM_B200_H */
