/*
 * mvs_oracle.h — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A dependency-free C restatement (double precision, scalar, no FMA contraction)
 * of the two-view front-end hot path of lonelycorn/mvSLAM's *own* geometry branch.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (libmvslam_b200.so) never links or calls it.
 *
 * Parity status: the reference cannot be compiled in this image (needs OpenCV C++,
 * Eigen, GTSAM, scons — all absent), so this restatement is pinned instead against
 *   (1) the known-answer vectors in the reference's own tests (test/test-svd.cpp:10-68,
 *       test/test-sfm.cpp:92-155 sfm_triangulate_cube, test/test-camera.cpp, test/test-lie-group.cpp),
 *   (2) the same third-party routines the reference calls, through Python cv2 4.13
 *       (cv2.batchDistance / BFMatcher.knnMatch, cv2.SVDecomp) — see oracle/oracle_np.py,
 *   (3) committed golden fixtures generated from the reference's bundled Tsukuba frames
 *       (tests/golden/, generator tools/make_golden.py).
 * The matcher and the own-branch 8-point/RANSAC have no direct test in the reference
 * (SURVEY.md §8c): for those two the parity is "pinned to cv2 + known geometry", not to
 * reference-run outputs.
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 */
#ifndef MVS_ORACLE_H
#define MVS_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes mirror include/mvslam_b200.h */
enum {
    ORC_OK = 0,
    ORC_E_BAD_ARG = 1,
    ORC_E_TOO_FEW_POINTS = 2,
    ORC_E_NO_MODEL = 3,
    ORC_E_TOO_FEW_INLIERS = 4,
    ORC_E_NO_CHEIRALITY = 5
};

enum { ORC_SCORE_ALGEBRAIC = 0, ORC_SCORE_SAMPSON = 1 };
/* REFERENCE: literal A^T A + cv::SVDecomp route, unfused arithmetic (what the reference executes).
 * FAST: Householder null vector + round-robin Jacobi + explicit fma (the library's throughput mode). */
enum { ORC_SOLVER_REFERENCE = 0, ORC_SOLVER_FAST = 1 };
void orc_set_solver(int solver);   /* process-wide; default ORC_SOLVER_REFERENCE */
int  orc_get_solver(void);

typedef struct {
    int32_t query;   /* cv::DMatch::queryIdx  (index into frame 2 / pair frame)  */
    int32_t train;   /* cv::DMatch::trainIdx  (index into frame 1 / base frame)  */
    float   distance;
} orc_match;

/* ---- matching (source/vision/visual-feature.cpp:51-80) ---- */
void orc_knn2_hamming(const uint8_t *q, int nq, const uint8_t *t, int nt, int desc_bytes,
                      int32_t *idx /*[nq][2]*/, int32_t *dist /*[nq][2]*/);
void orc_knn2_l2(const float *q, int nq, const float *t, int nt, int dim,
                 int32_t *idx /*[nq][2]*/, float *dist /*[nq][2]*/);
/* ratio test + max_dist filter + canonical sort; dist given as float (DMatch::distance) */
int orc_filter_matches(const int32_t *idx, const float *dist, int nq, double ratio, double max_dist,
                       orc_match *out /*cap nq*/);
int orc_match_hamming(const uint8_t *q, int nq, const uint8_t *t, int nt, int desc_bytes,
                      double ratio, double max_dist, int cross_check, orc_match *out /*cap nq*/);
int orc_match_l2(const float *q, int nq, const float *t, int nt, int dim,
                 double ratio, double max_dist, int cross_check, orc_match *out /*cap nq*/);

/* ---- small dense algebra (source/math/svd.hpp:59-72, cv::SVDecomp restated as one-sided Jacobi) ---- */
/* A is n x n row-major. Outputs: U n x n row-major (may be NULL), w[n] descending, Vt n x n row-major. */
int orc_svd(int n, const double *A, double *U, double *w, double *Vt);
/* cv::SVDecomp(MODIFY_A|FULL_UV) restated bit for bit (OpenCV's own Hestenes Jacobi for small matrices). */
int orc_cv_svd(int n, const double *A, double *U, double *w, double *Vt);

/* ---- lie group pieces on the path (source/math/lie-group.hpp:84-96,75-79,203-234) ---- */
void orc_so3_rectify(const double R[9], double out[9]);
void orc_se3_inverse(const double R[9], const double t[3], double Rout[9], double tout[3]);
void orc_se3_compose(const double Ra[9], const double ta[3], const double Rb[9], const double tb[3],
                     double Rout[9], double tout[3]);

/* ---- camera (source/vision/camera.cpp:14-18,55-79) ---- */
void orc_inverse3(const double K[9], double Kinv[9]);
void orc_normalize_points(const double K[9], const double *xy /*[n][2]*/, int n, double *out /*[n][3]*/);

/* ---- 8-point (source/vision/fundamental-matrix.cpp:18-140,204-267) ---- */
int orc_find_fundamental_matrix(const double *p1s /*[8][3]*/, const double *p2s /*[8][3]*/, double F[9]);

/* ---- RANSAC (source/vision/estimator-RANSAC.cpp:16-129) with an explicit sample table ---- */
void orc_sample_table(uint64_t seed, uint64_t pair_id, uint32_t n_points, int H, uint32_t *out /*[H][8]*/);
int orc_count_inliers(const double *p1, const double *p2, int n, const double F[9], double max_error_sq,
                      int score_mode, uint8_t *mask, double *residual);
int orc_ransac_fundamental(const double *p1, const double *p2, int n, const uint32_t *samples, int H,
                           double max_error_sq, int score_mode, double F[9], uint8_t *mask,
                           int *count, double *residual, int *best_h,
                           int32_t *all_counts /*[H] or NULL*/, double *all_F /*[H][9] or NULL*/);

/* ---- essential matrix / pose / triangulation (source/vision/sfm-solve.cpp:64-90,97-394) ---- */
void orc_project_essential(const double F[9], double E[9]);
void orc_decompose_essential(const double E[9], double Ra[9], double Rb[9], double t[3]);
int orc_triangulate_points(const double R[9], const double t[3], const double *p1, const double *p2,
                           const uint8_t *mask, int n, double *pts /*[n][3]*/, uint64_t *idx /*[n]*/);
int orc_recover_pose_and_points(const double E[9], const double *p1, const double *p2, const uint8_t *mask,
                                int n, double R[9], double t[3], double *pts, uint64_t *idx, int *n_out);

typedef struct {
    int32_t status;
    int32_t n_matches;      /* M  (pair entry only; == n for sfm_solve) */
    int32_t n_inliers;      /* RANSAC inliers of the winning hypothesis */
    int32_t best_hypothesis;
    int32_t n_points;       /* triangulated points that passed cheirality */
    int32_t candidate;      /* 0..3 = (Ra,+t),(Ra,-t),(Rb,+t),(Rb,-t) */
    double  residual;
    double  F[9];           /* winning de-normalised 8-point model (before projection) */
    double  E[9];           /* after (s,s,0) projection */
    double  R1to2[9];
    double  t1to2[3];
    double  R2in1[9];       /* pose2in1 = SE3(SO3(R1to2),t1to2).inverse() */
    double  t2in1[3];
} orc_pair_result;

/* sfm_solve (source/vision/sfm-solve.cpp:285-368). samples==NULL -> H rows from orc_sample_table(seed,pair_id). */
int orc_sfm_solve(const double *xy1, const double *xy2, int n, const double K[9],
                  const uint32_t *samples, int H, uint64_t seed, uint64_t pair_id, int score_mode, double max_error_sq_override,
                  orc_pair_result *res, uint8_t *mask /*[n] or NULL*/,
                  double *pts /*[n][3]*/, uint64_t *idx /*[n]*/);
/* sfm_triangulate (source/vision/sfm-solve.cpp:370-394) */
int orc_sfm_triangulate(const double *xy1, const double *xy2, int n, const double K[9],
                        const double R1[9], const double t1[3], const double R2[9], const double t2[3],
                        double *pts, uint64_t *idx);

/* ImagePair ctor + reconstruct (source/front-end/image-pair.cpp:30-71,115-174):
 * match(base=frame1 train, pair=frame2 query) -> gather keypoints -> sfm_solve. */
int orc_image_pair(const uint8_t *desc1, const float *kp1, int n1,
                   const uint8_t *desc2, const float *kp2, int n2, int desc_bytes,
                   const double K[9], double ratio, double max_dist, int cross_check,
                   int H, uint64_t seed, uint64_t pair_id, int score_mode, double max_error_sq_override,
                   orc_pair_result *res, orc_match *matches /*cap n2*/, uint8_t *mask /*cap n2*/,
                   double *pts /*cap n2*3*/, uint64_t *idx /*cap n2*/);

/* batch driver over pairs, OpenMP over pairs (threads<=0 -> all cores). Only records are returned. */
int orc_pair_batch(const uint8_t *const *desc, const float *const *kp, const int32_t *counts, int n_frames,
                   const int32_t *pairs /*[n_pairs][2] = (base, pair)*/, int n_pairs, int desc_bytes,
                   const double K[9], double ratio, double max_dist, int cross_check,
                   int H, uint64_t seed, int score_mode, double max_error_sq_override, int threads, orc_pair_result *res /*[n_pairs]*/);

int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
