/*
 * pnp_oracle.h — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE) for pnp_solve
 * (reference source/vision/pnp-solve.cpp:16-104 = cv::solvePnPRansac with SOLVEPNP_P3P).  See pnp_oracle.c.
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this; the product never links it.
 */
#ifndef PNP_ORACLE_H
#define PNP_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
void orc_pnp_sample_table(uint64_t seed, uint64_t problem_id, uint32_t n_points, int H, uint32_t *out /*[H][4]*/);
int orc_solve_quartic(const double c[5], double roots[4]);
int orc_p3p(const double f[3][3], const double X[3][3], double R[4][9], double t[4][3]);
int orc_pnp_hypothesis(const double *world, const double *image, const uint32_t idx[4], const double K[9],
                       double R[9], double t[3]);
int orc_pnp_count_inliers(const double *world, const double *image, int n, const double K[9], const double R[9],
                          const double t[3], double thr2, uint8_t *mask);
void orc_pnp_refine(const double *world, const double *image, int n, const uint8_t *mask, const double K[9],
                    double R[9], double t[3], int max_iter);
int orc_pnp_solve(const double *world, const double *image, int n, const double K[9], const uint32_t *samples, int H,
                  uint64_t seed, uint64_t problem_id, double reproj_error, int refine_iters,
                  double R_c2w[9], double t_c2w[3], uint8_t *inlier_mask, int *n_inliers, int *best_h,
                  double R_w2c_p3p[9], double t_w2c_p3p[3], int32_t *all_counts);
#ifdef __cplusplus
}
#endif
#endif
