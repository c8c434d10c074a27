/*
 * mvs_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).  See mvs_oracle.h.
 *
 * Plain C99, double precision, compiled with -ffp-contract=off so that every multiply and
 * add rounds separately — the reference's build (x86-64, g++ without -mfma/-O, SConstruct:70)
 * never fuses either.  All std::cout debugging prints of the reference are dropped.
 *
 * File:line citations are relative to /root/reference.
 */
#include "mvs_oracle.h"
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* source/system-config.hpp:8-14 */
#define ORC_EPSILON   DBL_EPSILON
#define ORC_TOLERANCE (DBL_EPSILON * 1000.0)
#define ORC_INFINITY  (DBL_MAX / 10.0)
/* source/vision/sfm-solve.cpp:18-21 */
#define ORC_MAX_ERROR_SQ 5e-2
#define ORC_VF_MATCH_INLIER_MIN 8

/* ------------------------------------------------------------------------------------------
 * Matching.  cv::BFMatcher(NORM_HAMMING, crossCheck=false).knnMatch(query=vf2, train=vf1, k=2)
 * (source/vision/visual-feature.cpp:21-25,59-62).  OpenCV is an un-vendored dependency
 * ("opencv >= 3.0", README.md:12); its brute-force kNN keeps the k smallest distances with a
 * strict '<' insertion while scanning train rows in ascending order, i.e. ties resolve to the
 * lowest train index for the 1st and the 2nd neighbour alike (pinned against cv2.batchDistance
 * in tests/test_oracle_pinning.py).
 * ------------------------------------------------------------------------------------------ */
static inline int hamming_bytes(const uint8_t *a, const uint8_t *b, int nbytes)
{
    int d = 0, i = 0;
    for (; i + 8 <= nbytes; i += 8) {
        uint64_t x, y;
        memcpy(&x, a + i, 8);
        memcpy(&y, b + i, 8);
        d += __builtin_popcountll(x ^ y);
    }
    for (; i < nbytes; ++i) d += __builtin_popcount((unsigned)(a[i] ^ b[i]));
    return d;
}

void orc_knn2_hamming(const uint8_t *q, int nq, const uint8_t *t, int nt, int desc_bytes,
                      int32_t *idx, int32_t *dist)
{
    for (int i = 0; i < nq; ++i) {
        int32_t b1 = INT32_MAX, b2 = INT32_MAX, i1 = -1, i2 = -1;
        const uint8_t *qi = q + (size_t)i * desc_bytes;
        for (int j = 0; j < nt; ++j) {
            int d = hamming_bytes(qi, t + (size_t)j * desc_bytes, desc_bytes);
            if (d < b1) { b2 = b1; i2 = i1; b1 = d; i1 = j; }
            else if (d < b2) { b2 = d; i2 = j; }
        }
        idx[2 * i] = i1; idx[2 * i + 1] = i2;
        dist[2 * i] = b1; dist[2 * i + 1] = b2;
    }
}

/* NORM_L2 variant of the same knnMatch (float descriptors).  Distances are accumulated in
 * double and rounded to float once (cv2 accumulates in float; parity is tolerance-based,
 * SURVEY.md §8d config 4). */
void orc_knn2_l2(const float *q, int nq, const float *t, int nt, int dim, int32_t *idx, float *dist)
{
    for (int i = 0; i < nq; ++i) {
        double b1 = INFINITY, b2 = INFINITY;
        int32_t i1 = -1, i2 = -1;
        const float *qi = q + (size_t)i * dim;
        for (int j = 0; j < nt; ++j) {
            const float *tj = t + (size_t)j * dim;
            double s = 0.0;
            for (int k = 0; k < dim; ++k) { double e = (double)qi[k] - (double)tj[k]; s += e * e; }
            if (s < b1) { b2 = b1; i2 = i1; b1 = s; i1 = j; }
            else if (s < b2) { b2 = s; i2 = j; }
        }
        idx[2 * i] = i1; idx[2 * i + 1] = i2;
        dist[2 * i] = (float)sqrt(b1); dist[2 * i + 1] = (float)sqrt(b2);
    }
}

static int cmp_match(const void *a, const void *b)
{
    const orc_match *x = (const orc_match *)a, *y = (const orc_match *)b;
    if (x->distance < y->distance) return -1;
    if (x->distance > y->distance) return 1;
    return (x->query > y->query) - (x->query < y->query);
}

/* Lowe ratio + max_dist filter + sort (source/vision/visual-feature.cpp:64-78,29-38).
 * The reference compares `float distance < double 0.7 * float distance` (promoted to double) and
 * `float distance <= double max_dist`.  std::partition/std::sort leave equal-distance matches in an
 * implementation-defined order; the canonical order used everywhere here is (distance, queryIdx). */
int orc_filter_matches(const int32_t *idx, const float *dist, int nq, double ratio, double max_dist,
                       orc_match *out)
{
    int m = 0;
    for (int i = 0; i < nq; ++i) {
        if (idx[2 * i] < 0 || idx[2 * i + 1] < 0) continue;
        float d1 = dist[2 * i], d2 = dist[2 * i + 1];
        int check1 = ((double)d1 < ratio * (double)d2);
        int check2 = (max_dist < 0) || ((double)d1 <= max_dist);
        if (check1 && check2) { out[m].query = i; out[m].train = idx[2 * i]; out[m].distance = d1; ++m; }
    }
    qsort(out, (size_t)m, sizeof(orc_match), cmp_match);
    return m;
}

/* cv::BFMatcher crossCheck=true semantics (a north-star addition; the reference runs with
 * CROSS_CHECK=false, visual-feature.cpp:22): keep (q,t) only if q is also t's nearest query,
 * nearest being decided with the same lowest-index tie-break in the reverse direction. */
static void cross_check_filter(orc_match *m, int *n, const int32_t *rev_idx /*[nt][2]*/)
{
    int k = 0;
    for (int i = 0; i < *n; ++i)
        if (rev_idx[2 * m[i].train] == m[i].query) m[k++] = m[i];
    *n = k;
}

int orc_match_hamming(const uint8_t *q, int nq, const uint8_t *t, int nt, int desc_bytes,
                      double ratio, double max_dist, int cross_check, orc_match *out)
{
    if (nq <= 0 || nt < 2) return 0;
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)nq);
    int32_t *di = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)nq);
    float *df = (float *)malloc(sizeof(float) * 2 * (size_t)nq);
    orc_knn2_hamming(q, nq, t, nt, desc_bytes, idx, di);
    for (int i = 0; i < 2 * nq; ++i) df[i] = (float)di[i];
    int m = orc_filter_matches(idx, df, nq, ratio, max_dist, out);
    if (cross_check && nq >= 2) {
        int32_t *ridx = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)nt);
        int32_t *rd = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)nt);
        orc_knn2_hamming(t, nt, q, nq, desc_bytes, ridx, rd);
        cross_check_filter(out, &m, ridx);
        free(ridx); free(rd);
    }
    free(idx); free(di); free(df);
    return m;
}

int orc_match_l2(const float *q, int nq, const float *t, int nt, int dim,
                 double ratio, double max_dist, int cross_check, orc_match *out)
{
    if (nq <= 0 || nt < 2) return 0;
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)nq);
    float *df = (float *)malloc(sizeof(float) * 2 * (size_t)nq);
    orc_knn2_l2(q, nq, t, nt, dim, idx, df);
    int m = orc_filter_matches(idx, df, nq, ratio, max_dist, out);
    if (cross_check && nq >= 2) {
        int32_t *ridx = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)nt);
        float *rd = (float *)malloc(sizeof(float) * 2 * (size_t)nt);
        orc_knn2_l2(t, nt, q, nq, dim, ridx, rd);
        cross_check_filter(out, &m, ridx);
        free(ridx); free(rd);
    }
    free(idx); free(df);
    return m;
}

/* ------------------------------------------------------------------------------------------
 * SVD.  The reference wraps cv::SVDecomp(MODIFY_A|FULL_UV) (source/math/svd.hpp:59-72) and uses
 * it on 9x9 (A^T A of the 8-point system), 3x3 (F, E) and 4x4 (DLT) matrices.  cv::SVDecomp is
 * restated here as the classical one-sided (Hestenes) Jacobi SVD: columns of W=A are rotated
 * pairwise until mutually orthogonal, V accumulates the rotations, sigma_j = |W_j|, U_j = W_j/sigma_j.
 * Pairs are visited in a round-robin ("chess tournament") order in which the pairs of one step
 * are column-disjoint; the CUDA kernels use the very same order so that both sides perform the
 * same sequence of IEEE operations.  Singular vectors are only defined up to sign (and up to a
 * rotation inside a repeated singular value), so parity against cv2.SVDecomp is checked on
 * sign/rotation-invariant quantities (tests/test_oracle_pinning.py).
 * ------------------------------------------------------------------------------------------ */
#define SVD_MAX_N 9
#define SVD_MAX_SWEEPS 30
#define SVD_EPS2 ((2.0 * DBL_EPSILON) * (2.0 * DBL_EPSILON))
#define SVD_RANK_TOL 1e-12

/* One rotation.  The explicit fma() calls are part of the numerical contract shared bit-for-bit with
 * the CUDA kernels (DFMA there, a hardware FMA here when built with -mfma, glibc's exact fma() otherwise);
 * nothing else is contracted (-ffp-contract=off). */
static void jacobi_rotate(int n, double W[SVD_MAX_N][SVD_MAX_N], double V[SVD_MAX_N][SVD_MAX_N],
                          int p, int q, int *changed)
{
    double a = 0.0, b = 0.0, g = 0.0;
    for (int k = 0; k < n; ++k) {
        a = fma(W[k][p], W[k][p], a);
        b = fma(W[k][q], W[k][q], b);
        g = fma(W[k][p], W[k][q], g);
    }
    /* |g| <= eps * sqrt(a*b), evaluated without the square root */
    if (g * g <= (SVD_EPS2 * a) * b) return;
    *changed = 1;
    double g2 = g * 2.0, beta = a - b;
    double gamma = sqrt(fma(g2, g2, beta * beta));
    double inv = 1.0 / (gamma * 2.0);
    double c, s;
    if (beta < 0) {
        s = sqrt((gamma - beta) * inv);
        c = (g2 * inv) / s;
    } else {
        c = sqrt((gamma + beta) * inv);
        s = (g2 * inv) / c;
    }
    for (int k = 0; k < n; ++k) {
        double wp = W[k][p], wq = W[k][q];
        W[k][p] = fma(c, wp, s * wq);
        W[k][q] = fma(c, wq, -(s * wp));
        double vp = V[k][p], vq = V[k][q];
        V[k][p] = fma(c, vp, s * vq);
        V[k][q] = fma(c, vq, -(s * vp));
    }
}

/* returns number of sweeps used */
static int jacobi_svd_core(int n, double W[SVD_MAX_N][SVD_MAX_N], double V[SVD_MAX_N][SVD_MAX_N])
{
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    /* round-robin: m = n rounded up to even; index m-1 is a bye when n is odd */
    const int m = (n + 1) & ~1;
    const int r = m - 1; /* steps per sweep, also the modulus */
    int sweep = 0;
    for (; sweep < SVD_MAX_SWEEPS; ++sweep) {
        int changed = 0;
        for (int s = 0; s < r; ++s) {
            /* pair (s, m-1) */
            if (m - 1 < n) jacobi_rotate(n, W, V, s, m - 1, &changed);
            for (int k = 1; k < m / 2; ++k) {
                int i = (s + k) % r, j = (s - k + r) % r;
                int p = i < j ? i : j, q = i < j ? j : i;
                jacobi_rotate(n, W, V, p, q, &changed);
            }
        }
        if (!changed) break;
    }
    return sweep;
}


/* ------------------------------------------------------------------------------------------
 * Solver mode.  ORC_SOLVER_REFERENCE (default) is the LITERAL restatement: the 8-point null vector is
 * vt.row(8) of cv::SVDecomp(A^T A) (fundamental-matrix.cpp:104-118), every SVD is OpenCV's own routine
 * restated bit for bit (orc_cv_svd below), and nothing is fused.  ORC_SOLVER_FAST is the library's
 * accuracy/throughput mode (Householder null vector of A, round-robin Jacobi, explicit fma contract).
 * A process-wide switch keeps the many call signatures unchanged; it is read once per entry point.
 * ------------------------------------------------------------------------------------------ */
static int g_solver = ORC_SOLVER_REFERENCE;
void orc_set_solver(int solver) { g_solver = solver; }
int orc_get_solver(void) { return g_solver; }

/* cv::SVDecomp(A, w, u, vt, MODIFY_A | FULL_UV) for a square n x n double matrix (n <= 9), the only way the
 * reference calls it (source/math/svd.hpp:65, fundamental-matrix.cpp:115,131).  OpenCV is an un-vendored
 * dependency ("opencv >= 3.0", README.md:12); for matrices this small it never reaches LAPACK but runs its own
 * one-sided Hestenes Jacobi (modules/core/src/lapack.cpp, JacobiSVDImpl_): the input is transposed, ROWS i<j of
 * A^T are rotated in plain cyclic order until |<Ai,Aj>| <= 10*eps*sqrt(|Ai|^2 |Aj|^2), with the rotation from
 * OpenCV's own scaled hypot; the squared row norms are carried along (W), rows are then sorted by norm
 * (selection sort, swap with the first maximum), normalised rows of A^T are U^T, and rows with a zero norm are
 * replaced by a pseudo-random vector (cv::RNG(0x12345678), +-1/m entries) orthogonalised against the others.
 * Every operation is IEEE +,-,*,/,sqrt in a fixed order and OpenCV's x86-64 baseline build has no FMA, so this
 * restatement is BIT-IDENTICAL to cv2.SVDecomp of this image (4.13.0): tests/test_oracle_pinning.py checks
 * w, u and vt for equality on 3x3, 4x4 and 9x9 inputs, rank-deficient ones included, and tests/golden/
 * cv_svd_golden.npz carries cv2's outputs to the GPU box. */
static double cv_hypot(double a, double b)
{
    a = fabs(a); b = fabs(b);
    if (a > b) { b /= a; return a * sqrt(1 + b * b); }
    if (b > 0) { a /= b; return b * sqrt(1 + a * a); }
    return 0;
}

int orc_cv_svd(int n, const double *A, double *U, double *w, double *Vt)
{
    if (n < 1 || n > SVD_MAX_N) return -1;
    const int m = n;
    const double eps = DBL_EPSILON * 10, minval = DBL_MIN;
    double At[SVD_MAX_N][SVD_MAX_N], V[SVD_MAX_N][SVD_MAX_N], W[SVD_MAX_N];
    for (int i = 0; i < n; ++i) for (int k = 0; k < m; ++k) At[i][k] = A[k * n + i];   /* transpose(src, temp_a) */
    for (int i = 0; i < n; ++i) {
        double sd = 0;
        for (int k = 0; k < m; ++k) { double t = At[i][k]; sd += t * t; }
        W[i] = sd;
        for (int k = 0; k < n; ++k) V[i][k] = 0;
        V[i][i] = 1;
    }
    const int max_iter = m > 30 ? m : 30;
    int iter;
    for (iter = 0; iter < max_iter; ++iter) {
        int changed = 0;
        for (int i = 0; i < n - 1; ++i)
            for (int j = i + 1; j < n; ++j) {
                double *Ai = At[i], *Aj = At[j];
                double a = W[i], p = 0, b = W[j];
                for (int k = 0; k < m; ++k) p += Ai[k] * Aj[k];
                if (fabs(p) <= eps * sqrt(a * b)) continue;
                p *= 2;
                double beta = a - b, gamma = cv_hypot(p, beta), c, s;
                if (beta < 0) {
                    double delta = (gamma - beta) * 0.5;
                    s = sqrt(delta / gamma);
                    c = p / (gamma * s * 2);
                } else {
                    c = sqrt((gamma + beta) / (gamma * 2));
                    s = p / (gamma * c * 2);
                }
                a = b = 0;
                for (int k = 0; k < m; ++k) {
                    double t0 = c * Ai[k] + s * Aj[k];
                    double t1 = -s * Ai[k] + c * Aj[k];
                    Ai[k] = t0; Aj[k] = t1;
                    a += t0 * t0; b += t1 * t1;
                }
                W[i] = a; W[j] = b;
                changed = 1;
                double *Vi = V[i], *Vj = V[j];
                for (int k = 0; k < n; ++k) {
                    double t0 = c * Vi[k] + s * Vj[k];
                    double t1 = -s * Vi[k] + c * Vj[k];
                    Vi[k] = t0; Vj[k] = t1;
                }
            }
        if (!changed) break;
    }
    for (int i = 0; i < n; ++i) {
        double sd = 0;
        for (int k = 0; k < m; ++k) { double t = At[i][k]; sd += t * t; }
        W[i] = sqrt(sd);
    }
    for (int i = 0; i < n - 1; ++i) {
        int j = i;
        for (int k = i + 1; k < n; ++k) if (W[j] < W[k]) j = k;
        if (i != j) {
            double t = W[i]; W[i] = W[j]; W[j] = t;
            for (int k = 0; k < m; ++k) { t = At[i][k]; At[i][k] = At[j][k]; At[j][k] = t; }
            for (int k = 0; k < n; ++k) { t = V[i][k]; V[i][k] = V[j][k]; V[j][k] = t; }
        }
    }
    for (int i = 0; i < n; ++i) w[i] = W[i];
    /* left singular vectors: normalised rows of A^T; zero rows are regenerated (cv::RNG multiply-with-carry) */
    uint64_t rng = 0x12345678;
    for (int i = 0; i < n; ++i) {
        double sd = W[i];
        for (int ii = 0; ii < 100 && sd <= minval; ++ii) {
            const double val0 = 1. / m;
            for (int k = 0; k < m; ++k) {
                rng = (uint64_t)(uint32_t)rng * 4164903690U + (uint32_t)(rng >> 32);
                At[i][k] = ((uint32_t)rng & 256) != 0 ? val0 : -val0;
            }
            for (int it = 0; it < 2; ++it)
                for (int j = 0; j < i; ++j) {
                    sd = 0;
                    for (int k = 0; k < m; ++k) sd += At[i][k] * At[j][k];
                    double asum = 0;
                    for (int k = 0; k < m; ++k) {
                        double t = At[i][k] - sd * At[j][k];
                        At[i][k] = t;
                        asum += fabs(t);
                    }
                    asum = asum > eps * 100 ? 1 / asum : 0;
                    for (int k = 0; k < m; ++k) At[i][k] *= asum;
                }
            sd = 0;
            for (int k = 0; k < m; ++k) { double t = At[i][k]; sd += t * t; }
            sd = sqrt(sd);
        }
        double s = sd > minval ? 1 / sd : 0.;
        for (int k = 0; k < m; ++k) At[i][k] *= s;
    }
    if (U) for (int i = 0; i < n; ++i) for (int k = 0; k < n; ++k) U[i * n + k] = At[k][i];   /* transpose(temp_u, u) */
    if (Vt) for (int i = 0; i < n; ++i) for (int k = 0; k < n; ++k) Vt[i * n + k] = V[i][k];
    return iter;
}

static void cross3(const double a[3], const double b[3], double o[3])
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

int orc_svd(int n, const double *A, double *U, double *w, double *Vt)
{
    if (n < 1 || n > SVD_MAX_N) return -1;
    if (g_solver == ORC_SOLVER_REFERENCE) return orc_cv_svd(n, A, U, w, Vt);
    double W[SVD_MAX_N][SVD_MAX_N], V[SVD_MAX_N][SVD_MAX_N];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) W[i][j] = A[i * n + j];
    int sweeps = jacobi_svd_core(n, W, V);
    double sig[SVD_MAX_N];
    int ord[SVD_MAX_N];
    for (int j = 0; j < n; ++j) {
        double s = 0.0;
        for (int k = 0; k < n; ++k) s += W[k][j] * W[k][j];
        sig[j] = sqrt(s);
        ord[j] = j;
    }
    /* stable insertion sort, descending (cv::SVDecomp returns sigma in descending order) */
    for (int i = 1; i < n; ++i) {
        int o = ord[i], j = i - 1;
        while (j >= 0 && sig[ord[j]] < sig[o]) { ord[j + 1] = ord[j]; --j; }
        ord[j + 1] = o;
    }
    for (int j = 0; j < n; ++j) {
        w[j] = sig[ord[j]];
        if (Vt) for (int k = 0; k < n; ++k) Vt[j * n + k] = V[k][ord[j]];
    }
    if (U) {
        int nvalid = 0;
        double thr = SVD_RANK_TOL * w[0];
        for (int j = 0; j < n; ++j) {
            if (w[j] > thr && w[j] > 0.0) {
                for (int k = 0; k < n; ++k) U[k * n + j] = W[k][ord[j]] / w[j];
                nvalid = j + 1;
            } else break;
        }
        /* FULL_UV completion of the null-space columns with a deterministic orthonormal basis */
        for (int j = nvalid; j < n; ++j) {
            if (n == 3 && j == 2) {
                double u0[3] = {U[0], U[3], U[6]}, u1[3] = {U[1], U[4], U[7]}, u2[3];
                cross3(u0, u1, u2);
                U[2] = u2[0]; U[5] = u2[1]; U[8] = u2[2];
                continue;
            }
            /* Gram-Schmidt of the unit vector least aligned with the existing columns */
            int best = 0; double bestv = INFINITY;
            for (int e = 0; e < n; ++e) {
                double v = 0.0;
                for (int c = 0; c < j; ++c) v += U[e * n + c] * U[e * n + c];
                if (v < bestv) { bestv = v; best = e; }
            }
            double x[SVD_MAX_N];
            for (int k = 0; k < n; ++k) x[k] = (k == best) ? 1.0 : 0.0;
            for (int pass = 0; pass < 2; ++pass)
                for (int c = 0; c < j; ++c) {
                    double d = 0.0;
                    for (int k = 0; k < n; ++k) d += x[k] * U[k * n + c];
                    for (int k = 0; k < n; ++k) x[k] -= d * U[k * n + c];
                }
            double nn = 0.0;
            for (int k = 0; k < n; ++k) nn += x[k] * x[k];
            nn = sqrt(nn);
            for (int k = 0; k < n; ++k) U[k * n + j] = x[k] / nn;
        }
    }
    return sweeps;
}

/* ------------------------------------------------------------------------------------------
 * Lie-group pieces (source/math/lie-group.hpp).
 * ------------------------------------------------------------------------------------------ */
/* SO3::rectify, lie-group.hpp:84-96.  NOTE: row 1 is deliberately NOT normalised (reference quirk). */
void orc_so3_rectify(const double R[9], double out[9])
{
    double n0 = sqrt(R[0] * R[0] + R[1] * R[1] + R[2] * R[2]);
    double u0[3] = {R[0] / n0, R[1] / n0, R[2] / n0};
    double d = R[3] * u0[0] + R[4] * u0[1] + R[5] * u0[2];
    double u1[3] = {R[3] - d * u0[0], R[4] - d * u0[1], R[5] - d * u0[2]};
    double u2[3];
    cross3(u0, u1, u2);
    for (int k = 0; k < 3; ++k) { out[k] = u0[k]; out[3 + k] = u1[k]; out[6 + k] = u2[k]; }
}

static void mat3_mul(const double A[9], const double B[9], double C[9])
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            C[i * 3 + j] = A[i * 3 + 0] * B[0 * 3 + j] + A[i * 3 + 1] * B[1 * 3 + j] + A[i * 3 + 2] * B[2 * 3 + j];
}
static void mat3_vec(const double A[9], const double v[3], double o[3])
{
    for (int i = 0; i < 3; ++i) o[i] = A[i * 3 + 0] * v[0] + A[i * 3 + 1] * v[1] + A[i * 3 + 2] * v[2];
}
static void mat3_transpose(const double A[9], double T[9])
{
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) T[i * 3 + j] = A[j * 3 + i];
}

/* SE3::inverse, lie-group.hpp:203-207 with SO3::inverse :75-79 (transpose, then rectify again).
 * R is the matrix held by the SO3 (i.e. already rectified by its constructor). */
void orc_se3_inverse(const double R[9], const double t[3], double Rout[9], double tout[3])
{
    double Rt[9], v[3];
    mat3_transpose(R, Rt);
    orc_so3_rectify(Rt, Rout);
    mat3_vec(Rout, t, v);
    tout[0] = -v[0]; tout[1] = -v[1]; tout[2] = -v[2];
}

/* SE3::operator*, lie-group.hpp:220-225 with SO3::operator* :118-122 (product, then rectify). */
void orc_se3_compose(const double Ra[9], const double ta[3], const double Rb[9], const double tb[3],
                     double Rout[9], double tout[3])
{
    double P[9], v[3];
    mat3_mul(Ra, Rb, P);
    orc_so3_rectify(P, Rout);
    mat3_vec(Ra, tb, v);
    tout[0] = v[0] + ta[0]; tout[1] = v[1] + ta[1]; tout[2] = v[2] + ta[2];
}

/* ------------------------------------------------------------------------------------------
 * Camera (source/vision/camera.cpp:14-18 K_inv = K.inverse(); :55-79 normalize_point(s)).
 * Eigen's fixed 3x3 inverse is the cofactor formula: inv = cofactor^T * (1/det).
 * ------------------------------------------------------------------------------------------ */
void orc_inverse3(const double K[9], double Ki[9])
{
    double c00 = K[4] * K[8] - K[5] * K[7];
    double c01 = K[5] * K[6] - K[3] * K[8];
    double c02 = K[3] * K[7] - K[4] * K[6];
    double c10 = K[2] * K[7] - K[1] * K[8];
    double c11 = K[0] * K[8] - K[2] * K[6];
    double c12 = K[1] * K[6] - K[0] * K[7];
    double c20 = K[1] * K[5] - K[2] * K[4];
    double c21 = K[2] * K[3] - K[0] * K[5];
    double c22 = K[0] * K[4] - K[1] * K[3];
    double det = K[0] * c00 + K[1] * c01 + K[2] * c02;
    double id = 1.0 / det;
    Ki[0] = c00 * id; Ki[1] = c10 * id; Ki[2] = c20 * id;
    Ki[3] = c01 * id; Ki[4] = c11 * id; Ki[5] = c21 * id;
    Ki[6] = c02 * id; Ki[7] = c12 * id; Ki[8] = c22 * id;
}

void orc_normalize_points(const double K[9], const double *xy, int n, double *out)
{
    double Ki[9];
    orc_inverse3(K, Ki);
    for (int i = 0; i < n; ++i) {
        double x = xy[2 * i], y = xy[2 * i + 1];
        for (int r = 0; r < 3; ++r) out[3 * i + r] = Ki[3 * r] * x + Ki[3 * r + 1] * y + Ki[3 * r + 2] * 1.0;
    }
}

/* ------------------------------------------------------------------------------------------
 * 8-point fundamental matrix (source/vision/fundamental-matrix.cpp).
 * ------------------------------------------------------------------------------------------ */
/* find_normalization_transform, fundamental-matrix.cpp:18-54.  The scale uses the MEAN distance
 * to the centroid (the comment in the reference says RMS; the code is the contract).  The centroid
 * is taken over all three components, so z -> 0 for (x,y,1) inputs. */
static void find_normalization_transform(const double *p /*[8][3]*/, double np[8][3], double T[9])
{
    const double inv = 1.0 / 8.0;
    double mean[3] = {0, 0, 0};
    for (int i = 0; i < 8; ++i) for (int k = 0; k < 3; ++k) mean[k] += p[3 * i + k];
    for (int k = 0; k < 3; ++k) mean[k] *= inv;
    double scale = 0.0;
    for (int i = 0; i < 8; ++i) {
        for (int k = 0; k < 3; ++k) np[i][k] = p[3 * i + k] - mean[k];
        scale += sqrt(np[i][0] * np[i][0] + np[i][1] * np[i][1] + np[i][2] * np[i][2]);
    }
    scale *= inv;
    scale = sqrt(2.0) / scale;
    for (int i = 0; i < 8; ++i) for (int k = 0; k < 3; ++k) np[i][k] *= scale;
    T[0] = scale; T[1] = 0; T[2] = -mean[0] * scale;
    T[3] = 0; T[4] = scale; T[5] = -mean[1] * scale;
    T[6] = 0; T[7] = 0; T[8] = 1;
}

/* The reference obtains f as vt.row(8) of cv::SVDecomp(A^T A) (fundamental-matrix.cpp:104-118), i.e. the
 * right singular vector of the 8x9 matrix A for its zero singular value = the unit vector orthogonal to the
 * 8 rows of A.  It is computed here directly from A: a Householder QR of A^T (9x8) gives an orthogonal Q
 * whose last column spans the null space of A; f = Q e_8 (apply the 8 reflectors to e_8 in reverse order).
 * Working on A instead of A^T A avoids squaring the condition number: measured against the exact null
 * vector this route is ~1000x more accurate than either cv2.SVDecomp(A^T A) or a Jacobi on A^T A, and it
 * differs from cv2.SVDecomp(A^T A) by no more than a literal restatement does (both differences are cv2's
 * own A^T A round-off; numbers in DESIGN.md).  Oracle A (oracle_np.py) keeps the literal A^T A +
 * cv2.SVDecomp route.  The explicit fma() order below is mirrored one-for-one by the CUDA kernel. */
static void null_vector_8x9(double A[8][9], double f[9])
{
    double M[9][8], beta[8];
    for (int i = 0; i < 9; ++i) for (int j = 0; j < 8; ++j) M[i][j] = A[j][i];
    for (int k = 0; k < 8; ++k) {
        double s2 = 0.0;
        for (int i = k; i < 9; ++i) s2 = fma(M[i][k], M[i][k], s2);
        const double nrm = sqrt(s2);
        if (!(nrm > 0.0)) { beta[k] = 0.0; continue; }          /* zero column: H_k = I */
        const double x0 = M[k][k];
        const double alpha = (x0 >= 0.0) ? -nrm : nrm;
        M[k][k] = x0 - alpha;                                    /* v_k stored in place of column k */
        beta[k] = 2.0 / (2.0 * fma(nrm, fabs(x0), s2));          /* 2 / (v^T v), v^T v = 2 (|x|^2 + |x||x0|) */
        for (int j = k + 1; j < 8; ++j) {
            double s = 0.0;
            for (int i = k; i < 9; ++i) s = fma(M[i][k], M[i][j], s);
            s *= beta[k];
            for (int i = k; i < 9; ++i) M[i][j] = fma(-s, M[i][k], M[i][j]);
        }
    }
    for (int i = 0; i < 9; ++i) f[i] = (i == 8) ? 1.0 : 0.0;
    for (int k = 7; k >= 0; --k) {
        if (beta[k] == 0.0) continue;
        double s = 0.0;
        for (int i = k; i < 9; ++i) s = fma(M[i][k], f[i], s);
        s *= beta[k];
        for (int i = k; i < 9; ++i) f[i] = fma(-s, M[i][k], f[i]);
    }
}

/* find_fundamental_matrix_8point, fundamental-matrix.cpp:56-140 */
static void find_fundamental_matrix_8point(double n1[8][3], double n2[8][3], double F[9])
{
    double A[8][9];
    for (int i = 0; i < 8; ++i) {
        double x1 = n1[i][0], y1 = n1[i][1], x2 = n2[i][0], y2 = n2[i][1];
        A[i][0] = x2 * x1; A[i][1] = x2 * y1; A[i][2] = x2;
        A[i][3] = y2 * x1; A[i][4] = y2 * y1; A[i][5] = y2;
        A[i][6] = x1; A[i][7] = y1; A[i][8] = 1.0;
    }
    double Fp[9];
    if (g_solver == ORC_SOLVER_REFERENCE) {
        /* A^T A accumulated element by element over the 8 rows (:104-111), f = vt.row(8) (:114-118) */
        double AtA[81], w9[9], Vt9[81];
        for (int i = 0; i < 9; ++i)
            for (int j = 0; j < 9; ++j) {
                double acc = 0;
                for (int k = 0; k < 8; ++k) acc += A[k][i] * A[k][j];
                AtA[i * 9 + j] = acc;
            }
        orc_cv_svd(9, AtA, NULL, w9, Vt9);
        for (int i = 0; i < 9; ++i) Fp[i] = Vt9[72 + i];
    } else null_vector_8x9(A, Fp);
    /* singular constraint (:128-136): F = u * diag(w0,w1,0) * vt */
    double U[9], w[3], Vt[9];
    orc_svd(3, Fp, U, w, Vt);
    w[2] = 0.0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            F[i * 3 + j] = (U[i * 3 + 0] * w[0]) * Vt[0 * 3 + j] + (U[i * 3 + 1] * w[1]) * Vt[1 * 3 + j];
}

/* find_fundamental_matrix, fundamental-matrix.cpp:204-267: normalise, solve, F = T2^T * F * T1 */
int orc_find_fundamental_matrix(const double *p1s, const double *p2s, double F[9])
{
    double n1[8][3], n2[8][3], T1[9], T2[9], Fh[9], T2t[9], tmp[9];
    find_normalization_transform(p1s, n1, T1);
    find_normalization_transform(p2s, n2, T2);
    find_fundamental_matrix_8point(n1, n2, Fh);
    mat3_transpose(T2, T2t);
    mat3_mul(T2t, Fh, tmp);
    mat3_mul(tmp, T1, F);
    return 1;
}

/* ------------------------------------------------------------------------------------------
 * RANSAC (source/vision/estimator-RANSAC.cpp).  The reference never randomises (shuffle is
 * commented out, :41-42, and the only caller asks for one iteration, sfm-solve.cpp:67): its sample
 * is always index[0..7].  Here the sample set is an explicit uint32[H][8] table whose row 0 is
 * {0..7}; rows >= 1 come from a seeded counter-based generator shared bit-for-bit with the GPU.
 * ------------------------------------------------------------------------------------------ */
static inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

void orc_sample_table(uint64_t seed, uint64_t pair_id, uint32_t n_points, int H, uint32_t *out)
{
    for (int h = 0; h < H; ++h) {
        uint32_t *row = out + 8 * (size_t)h;
        if (h == 0 || n_points < 8) { for (int j = 0; j < 8; ++j) row[j] = (uint32_t)j; continue; }
        uint64_t st = splitmix64(seed ^ splitmix64(pair_id * 0xD1B54A32D192ED03ULL + (uint64_t)h));
        for (int j = 0; j < 8; ++j) {
            for (;;) {
                st = splitmix64(st);
                uint32_t v = (uint32_t)(((st >> 32) * (uint64_t)n_points) >> 32);
                int dup = 0;
                for (int k = 0; k < j; ++k) dup |= (row[k] == v);
                if (!dup) { row[j] = v; break; }
            }
        }
    }
}

/* residual of one correspondence.  ALGEBRAIC = |p2^T F p1| evaluated as (p2^T F) p1
 * (estimator-RANSAC.cpp:114-116).  SAMPSON = r^2 / (|(F p1)_xy|^2 + |(F^T p2)_xy|^2), the score of
 * cv::findEssentialMat's RANSAC (default-build branch, sfm-solve.cpp:42-63; north-star mode). */
/* ALGEBRAIC: e = |r| < thr.  SAMPSON: r^2 / den < thr, decided as r^2 < thr * den (den > 0) so that the
 * division is only needed for the residual of actual inliers.  Explicit fma() = shared contract with the GPU. */
static inline int point_residual_fast(const double *a, const double *b, const double F[9], int mode, double thr, double *res);

/* REFERENCE solver: r = (p2^T F) p1 with every product and sum rounded separately, left to right
 * (Eigen coefficient-based products of the reference build: no FMA, EIGEN_DONT_VECTORIZE, SConstruct:70,86). */
static inline int point_residual(const double *a /*p1*/, const double *b /*p2*/, const double F[9], int mode,
                                 double thr, double *res)
{
    if (g_solver != ORC_SOLVER_REFERENCE) return point_residual_fast(a, b, F, mode, thr, res);
    double v0 = (b[0] * F[0] + b[1] * F[3]) + b[2] * F[6];
    double v1 = (b[0] * F[1] + b[1] * F[4]) + b[2] * F[7];
    double v2 = (b[0] * F[2] + b[1] * F[5]) + b[2] * F[8];
    double r = (v0 * a[0] + v1 * a[1]) + v2 * a[2];
    if (mode == ORC_SCORE_ALGEBRAIC) {
        double e = r < 0 ? -r : r;
        *res = e;
        return e < thr;
    }
    double l0 = (F[0] * a[0] + F[1] * a[1]) + F[2] * a[2];
    double l1 = (F[3] * a[0] + F[4] * a[1]) + F[5] * a[2];
    double den = (l0 * l0 + l1 * l1) + (v0 * v0 + v1 * v1);
    double r2 = r * r;
    if (!(r2 < thr * den)) return 0;
    *res = r2 / den;
    return 1;
}

static inline int point_residual_fast(const double *a /*p1*/, const double *b /*p2*/, const double F[9], int mode,
                                      double thr, double *res)
{
    double v0 = fma(b[0], F[0], fma(b[1], F[3], b[2] * F[6]));
    double v1 = fma(b[0], F[1], fma(b[1], F[4], b[2] * F[7]));
    double v2 = fma(b[0], F[2], fma(b[1], F[5], b[2] * F[8]));
    double r = fma(v0, a[0], fma(v1, a[1], v2 * a[2]));
    if (mode == ORC_SCORE_ALGEBRAIC) {
        double e = r < 0 ? -r : r;
        *res = e;
        return e < thr;
    }
    double l0 = fma(F[0], a[0], fma(F[1], a[1], F[2] * a[2]));
    double l1 = fma(F[3], a[0], fma(F[4], a[1], F[5] * a[2]));
    double den = fma(l0, l0, l1 * l1) + fma(v0, v0, v1 * v1);
    double r2 = r * r;
    if (!(r2 < thr * den)) return 0;
    *res = r2 / den;
    return 1;
}

/* count_inliers, estimator-RANSAC.cpp:100-129 (strict '<', residual summed over inliers in order) */
int orc_count_inliers(const double *p1, const double *p2, int n, const double F[9], double max_error_sq,
                      int score_mode, uint8_t *mask, double *residual)
{
    int cnt = 0;
    double res = 0.0;
    for (int i = 0; i < n; ++i) {
        double r;
        if (point_residual(p1 + 3 * i, p2 + 3 * i, F, score_mode, max_error_sq, &r)) { ++cnt; res += r; if (mask) mask[i] = 1; }
        else if (mask) mask[i] = 0;
    }
    *residual = res;
    return cnt;
}

/* compute(), estimator-RANSAC.cpp:16-90 */
int orc_ransac_fundamental(const double *p1, const double *p2, int n, const uint32_t *samples, int H,
                           double max_error_sq, int score_mode, double F[9], uint8_t *mask,
                           int *count, double *residual, int *best_h, int32_t *all_counts, double *all_F)
{
    *count = 0; *residual = ORC_INFINITY; *best_h = -1;
    if (n < 8) return ORC_E_TOO_FEW_POINTS;
    double res_best = ORC_INFINITY;
    int cnt_best = 0;
    uint8_t *m = (uint8_t *)malloc((size_t)n);
    for (int h = 0; h < H; ++h) {
        double s1[24], s2[24], Fp[9], res;
        for (int j = 0; j < 8; ++j) {
            uint32_t id = samples[8 * (size_t)h + j];
            if (id >= (uint32_t)n) { free(m); return ORC_E_BAD_ARG; }
            for (int k = 0; k < 3; ++k) { s1[3 * j + k] = p1[3 * (size_t)id + k]; s2[3 * j + k] = p2[3 * (size_t)id + k]; }
        }
        orc_find_fundamental_matrix(s1, s2, Fp);
        int cnt = orc_count_inliers(p1, p2, n, Fp, max_error_sq, score_mode, m, &res);
        if (all_counts) all_counts[h] = cnt;
        if (all_F) memcpy(all_F + 9 * (size_t)h, Fp, sizeof(Fp));
        if ((cnt > cnt_best) || ((cnt == cnt_best) && (res < res_best))) {
            cnt_best = cnt; res_best = res; *best_h = h;
            memcpy(F, Fp, sizeof(Fp));
            if (mask) memcpy(mask, m, (size_t)n);
        }
    }
    free(m);
    *count = cnt_best; *residual = res_best;
    return cnt_best > 0 ? ORC_OK : ORC_E_NO_MODEL;
}

/* ------------------------------------------------------------------------------------------
 * Essential matrix, pose recovery, triangulation (source/vision/sfm-solve.cpp).
 * ------------------------------------------------------------------------------------------ */
/* find_essential_matrix own branch, sfm-solve.cpp:73-87: singular values -> (v,v,0), v = sqrt(s0*s1) */
void orc_project_essential(const double F[9], double E[9])
{
    double U[9], w[3], Vt[9];
    orc_svd(3, F, U, w, Vt);
    double v = sqrt(w[0] * w[1]);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            E[i * 3 + j] = (U[i * 3 + 0] * v) * Vt[0 * 3 + j] + (U[i * 3 + 1] * v) * Vt[1 * 3 + j];
}

static double det3(const double M[9])
{
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}

/* decompose_essential_matrix, sfm-solve.cpp:97-127 */
void orc_decompose_essential(const double E[9], double Ra[9], double Rb[9], double t[3])
{
    double U[9], w[3], Vt[9], V[9];
    orc_svd(3, E, U, w, Vt);
    mat3_transpose(Vt, V);
    if (det3(U) < 0) for (int i = 0; i < 9; ++i) U[i] = -U[i];
    if (det3(V) < 0) for (int i = 0; i < 9; ++i) V[i] = -V[i];
    /* U*W: columns (u1, -u0, u2); U*W^T: (-u1, u0, u2); U*Z: (-u1, u0, 0) */
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double u0 = U[i * 3 + 0], u1 = U[i * 3 + 1], u2 = U[i * 3 + 2];
            Ra[i * 3 + j] = (u1 * V[j * 3 + 0] + (-u0) * V[j * 3 + 1]) + u2 * V[j * 3 + 2];
            Rb[i * 3 + j] = ((-u1) * V[j * 3 + 0] + u0 * V[j * 3 + 1]) + u2 * V[j * 3 + 2];
        }
    /* S = U Z U^T, t = (-S12, S02, -S01) */
    double S01 = (-U[0 * 3 + 1]) * U[1 * 3 + 0] + U[0 * 3 + 0] * U[1 * 3 + 1];
    double S02 = (-U[0 * 3 + 1]) * U[2 * 3 + 0] + U[0 * 3 + 0] * U[2 * 3 + 1];
    double S12 = (-U[1 * 3 + 1]) * U[2 * 3 + 0] + U[1 * 3 + 0] * U[2 * 3 + 1];
    t[0] = -S12; t[1] = S02; t[2] = -S01;
}

/* triangulate_points, sfm-solve.cpp:134-227.  P1 = I, P2 = [rectify(R) | t]; cheirality uses the
 * un-rectified R (:218).  Returns the number of surviving points; idx holds their original indexes. */
int orc_triangulate_points(const double R[9], const double t[3], const double *p1, const double *p2,
                           const uint8_t *mask, int n, double *pts, uint64_t *idx)
{
    double Rr[9];
    orc_so3_rectify(R, Rr);
    double P2[3][4];
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) P2[i][j] = Rr[i * 3 + j]; P2[i][3] = t[i]; }
    int m = 0;
    for (int i = 0; i < n; ++i) {
        if (mask && mask[i] == 0) continue;
        const double *x1 = p1 + 3 * (size_t)i, *x2 = p2 + 3 * (size_t)i;
        double A[16];
        /* rows 0,1: x1[k]*P1.row(2) - P1.row(k) with P1 = I4 */
        A[0] = x1[0] * 0.0 - 1.0; A[1] = x1[0] * 0.0 - 0.0; A[2] = x1[0] * 1.0 - 0.0; A[3] = x1[0] * 0.0 - 0.0;
        A[4] = x1[1] * 0.0 - 0.0; A[5] = x1[1] * 0.0 - 1.0; A[6] = x1[1] * 1.0 - 0.0; A[7] = x1[1] * 0.0 - 0.0;
        for (int j = 0; j < 4; ++j) {
            A[8 + j] = x2[0] * P2[2][j] - P2[0][j];
            A[12 + j] = x2[1] * P2[2][j] - P2[1][j];
        }
        double w[4], Vt[16];
        orc_svd(4, A, NULL, w, Vt);
        const double *X = Vt + 12; /* V.col(3); the reference's sign flip (:196-199) cancels in X/X3 */
        if (fabs(X[3]) < ORC_TOLERANCE) continue;
        double scale = 1.0 / X[3];
        double pt[3] = {X[0] * scale, X[1] * scale, X[2] * scale};
        if (pt[2] < ORC_TOLERANCE) continue;
        double z2 = (R[6] * pt[0] + R[7] * pt[1] + R[8] * pt[2]) + t[2];
        if (z2 < ORC_TOLERANCE) continue;
        pts[3 * m] = pt[0]; pts[3 * m + 1] = pt[1]; pts[3 * m + 2] = pt[2];
        idx[m] = (uint64_t)i;
        ++m;
    }
    return m;
}

/* recover_pose_and_points, sfm-solve.cpp:232-284 */
int orc_recover_pose_and_points(const double E[9], const double *p1, const double *p2, const uint8_t *mask,
                                int n, double R[9], double t[3], double *pts, uint64_t *idx, int *n_out)
{
    double Rc[2][9], tc[2][3];
    orc_decompose_essential(E, Rc[0], Rc[1], tc[0]);
    for (int k = 0; k < 3; ++k) tc[1][k] = -tc[0][k];
    double *cp = (double *)malloc(sizeof(double) * 3 * (size_t)(n > 0 ? n : 1));
    uint64_t *ci = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(n > 0 ? n : 1));
    int best = 0, cand = -1;
    for (int r = 0; r < 2; ++r)
        for (int s = 0; s < 2; ++s) {
            int m = orc_triangulate_points(Rc[r], tc[s], p1, p2, mask, n, cp, ci);
            if (m > best) {
                best = m; cand = r * 2 + s;
                memcpy(pts, cp, sizeof(double) * 3 * (size_t)m);
                memcpy(idx, ci, sizeof(uint64_t) * (size_t)m);
                memcpy(R, Rc[r], sizeof(double) * 9);
                memcpy(t, tc[s], sizeof(double) * 3);
            }
        }
    free(cp); free(ci);
    *n_out = best;
    return cand;
}

/* sfm_solve, sfm-solve.cpp:285-368 (own-branch find_essential_matrix :64-90) */
int orc_sfm_solve(const double *xy1, const double *xy2, int n, const double K[9],
                  const uint32_t *samples, int H, uint64_t seed, uint64_t pair_id, int score_mode, double max_error_sq_override,
                  orc_pair_result *res, uint8_t *mask_out, double *pts, uint64_t *idx)
{
    memset(res, 0, sizeof(*res));
    res->n_matches = n; res->best_hypothesis = -1; res->candidate = -1;
    if (n < 8 || H < 1) { res->status = ORC_E_TOO_FEW_POINTS; return res->status; }
    double *p1 = (double *)malloc(sizeof(double) * 3 * (size_t)n);
    double *p2 = (double *)malloc(sizeof(double) * 3 * (size_t)n);
    uint8_t *mask = (uint8_t *)malloc((size_t)n);
    uint32_t *tab = NULL;
    orc_normalize_points(K, xy1, n, p1);
    orc_normalize_points(K, xy2, n, p2);
    if (!samples) {
        tab = (uint32_t *)malloc(sizeof(uint32_t) * 8 * (size_t)H);
        orc_sample_table(seed, pair_id, (uint32_t)n, H, tab);
        samples = tab;
    }
    double max_error_sq = ORC_MAX_ERROR_SQ / K[0] / K[4]; /* :311 */
    if (max_error_sq_override > 0) max_error_sq = max_error_sq_override;
    int st = orc_ransac_fundamental(p1, p2, n, samples, H, max_error_sq, score_mode, res->F, mask,
                                    &res->n_inliers, &res->residual, &res->best_hypothesis, NULL, NULL);
    if (st == ORC_OK) {
        orc_project_essential(res->F, res->E);
        if (res->n_inliers < ORC_VF_MATCH_INLIER_MIN) st = ORC_E_TOO_FEW_INLIERS; /* :326-334 */
    }
    if (st == ORC_OK) {
        res->candidate = orc_recover_pose_and_points(res->E, p1, p2, mask, n, res->R1to2, res->t1to2,
                                                     pts, idx, &res->n_points);
        if (res->candidate < 0) st = ORC_E_NO_CHEIRALITY;
    }
    if (st == ORC_OK) {
        /* pose2in1 = SE3(SO3(R1to2), t1to2).inverse()  (:364) */
        double Rr[9];
        orc_so3_rectify(res->R1to2, Rr);
        orc_se3_inverse(Rr, res->t1to2, res->R2in1, res->t2in1);
    }
    if (mask_out && st != ORC_E_TOO_FEW_POINTS) memcpy(mask_out, mask, (size_t)n);
    free(p1); free(p2); free(mask); free(tab);
    res->status = st;
    return st;
}

/* sfm_triangulate, sfm-solve.cpp:370-394: T_1_to_2 = pose2.inverse() * pose1 */
int orc_sfm_triangulate(const double *xy1, const double *xy2, int n, const double K[9],
                        const double R1[9], const double t1[3], const double R2[9], const double t2[3],
                        double *pts, uint64_t *idx)
{
    double Ri[9], ti[3], R12[9], t12[3];
    orc_se3_inverse(R2, t2, Ri, ti);
    orc_se3_compose(Ri, ti, R1, t1, R12, t12);
    double *p1 = (double *)malloc(sizeof(double) * 3 * (size_t)(n > 0 ? n : 1));
    double *p2 = (double *)malloc(sizeof(double) * 3 * (size_t)(n > 0 ? n : 1));
    orc_normalize_points(K, xy1, n, p1);
    orc_normalize_points(K, xy2, n, p2);
    int m = orc_triangulate_points(R12, t12, p1, p2, NULL, n, pts, idx);
    free(p1); free(p2);
    return m;
}

/* ImagePair::ImagePair + reconstruct, source/front-end/image-pair.cpp:30-71,115-174 */
int orc_image_pair(const uint8_t *desc1, const float *kp1, int n1,
                   const uint8_t *desc2, const float *kp2, int n2, int desc_bytes,
                   const double K[9], double ratio, double max_dist, int cross_check,
                   int H, uint64_t seed, uint64_t pair_id, int score_mode, double max_error_sq_override,
                   orc_pair_result *res, orc_match *matches, uint8_t *mask, double *pts, uint64_t *idx)
{
    memset(res, 0, sizeof(*res));
    res->best_hypothesis = -1; res->candidate = -1;
    if (n1 < 2 || n2 < 1) { res->status = ORC_E_BAD_ARG; return res->status; }
    /* query = pair frame (2), train = base frame (1): visual-feature.cpp:59-60 */
    int m = orc_match_hamming(desc2, n2, desc1, n1, desc_bytes, ratio, max_dist, cross_check, matches);
    double *xy1 = (double *)malloc(sizeof(double) * 2 * (size_t)(m > 0 ? m : 1));
    double *xy2 = (double *)malloc(sizeof(double) * 2 * (size_t)(m > 0 ? m : 1));
    for (int i = 0; i < m; ++i) { /* image-pair.cpp:123-140 */
        xy1[2 * i] = (double)kp1[2 * matches[i].train]; xy1[2 * i + 1] = (double)kp1[2 * matches[i].train + 1];
        xy2[2 * i] = (double)kp2[2 * matches[i].query]; xy2[2 * i + 1] = (double)kp2[2 * matches[i].query + 1];
    }
    int st = orc_sfm_solve(xy1, xy2, m, K, NULL, H, seed, pair_id, score_mode, max_error_sq_override, res, mask, pts, idx);
    res->n_matches = m;
    free(xy1); free(xy2);
    return st;
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int orc_pair_batch(const uint8_t *const *desc, const float *const *kp, const int32_t *counts, int n_frames,
                   const int32_t *pairs, int n_pairs, int desc_bytes,
                   const double K[9], double ratio, double max_dist, int cross_check,
                   int H, uint64_t seed, int score_mode, double max_error_sq_override, int threads, orc_pair_result *res)
{
    (void)n_frames;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
#endif
    for (int p = 0; p < n_pairs; ++p) {
        int a = pairs[2 * p], b = pairs[2 * p + 1];
        int cap = counts[b] > 0 ? counts[b] : 1;
        orc_match *mt = (orc_match *)malloc(sizeof(orc_match) * (size_t)cap);
        uint8_t *mask = (uint8_t *)malloc((size_t)cap);
        double *pts = (double *)malloc(sizeof(double) * 3 * (size_t)cap);
        uint64_t *idx = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)cap);
        orc_image_pair(desc[a], kp[a], counts[a], desc[b], kp[b], counts[b], desc_bytes, K, ratio, max_dist,
                       cross_check, H, seed, (uint64_t)p, score_mode, max_error_sq_override, &res[p], mt, mask, pts, idx);
        free(mt); free(mask); free(pts); free(idx);
    }
    return ORC_OK;
}
