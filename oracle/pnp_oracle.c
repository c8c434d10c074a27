/*
 * pnp_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE) for pnp_solve.
 *
 * The reference's pnp_solve (source/vision/pnp-solve.cpp:16-104) is a thin wrapper over the un-vendored third-party
 * cv::solvePnPRansac(flags = SOLVEPNP_P3P, iterationsCount = 100, reprojectionError = 0.05, confidence = 0.95)
 * (:38-66), followed by pose = SE3(SO3::exp(rvec), tvec).inverse() (:99-101).  OpenCV's source is not under
 * /root/reference, so this file restates the published algorithm of that call:
 *   - minimal samples of 4 correspondences; P3P on the first three (Grunert's quartic, Haralick et al. 1994,
 *     solved in closed form by Ferrari's method), the fourth picks among the up to four poses by reprojection error;
 *   - consensus: squared reprojection error <= reprojectionError^2 (fx, fy, cx, cy of K; no distortion), evaluated
 *     in the division-free form (both sides times z^2);
 *   - the model with most inliers wins (first one on ties);
 *   - final pose on the inliers.  OpenCV runs EPnP there; this restatement runs Gauss-Newton on the reprojection
 *     error from the winning P3P pose (the maximum-likelihood refinement EPnP approximates).
 * Deliberate differences, all stated in DESIGN.md: the sample sets come from the library's seeded table instead of
 * cv::RNG; every one of the H samples is evaluated (no early exit by confidence).
 *
 * PARITY: pinned against cv2.solveP3P (solution sets) and cv2.solvePnPRansac (pose within the reference test's own
 * tolerance 1e-3, test/test-pnp.cpp:16, inlier sets on data with gross outliers) in tests/test_pnp_oracle.py.
 * The RANSAC sample sequence itself cannot be pinned (cv::RNG) -> "parity unpinned" for per-iteration equality.
 *
 * Arithmetic contract shared with the CUDA kernels (pnp.cu): only + - * / sqrt, IEEE double, no contraction, same
 * operation order -> hypothesis poses and inlier counts are bit-identical.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "pnp_oracle.h"

static inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

/* row 0 = {0,1,2,3}; rows >= 1: 4 distinct indices < n_points */
void orc_pnp_sample_table(uint64_t seed, uint64_t problem_id, uint32_t n_points, int H, uint32_t *out)
{
    for (int h = 0; h < H; ++h) {
        uint32_t *row = out + 4 * (size_t)h;
        if (h == 0 || n_points < 4) { for (int j = 0; j < 4; ++j) row[j] = (uint32_t)j; continue; }
        uint64_t st = splitmix64(seed ^ splitmix64(problem_id * 0xD1B54A32D192ED03ULL + (uint64_t)h + 0x504E50ULL));
        for (int j = 0; j < 4; ++j) {
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

/* ---- real roots of c[4] x^4 + c[3] x^3 + c[2] x^2 + c[1] x + c[0] (Ferrari; resolvent cubic by Newton) ---- */
static int quadratic_roots(double b, double c, double *r) /* x^2 + b x + c */
{
    const double disc = b * b - 4.0 * c;
    if (disc < 0.0) return 0;
    const double s = sqrt(disc);
    /* avoid cancellation: larger-magnitude root first */
    const double q = b >= 0.0 ? -0.5 * (b + s) : -0.5 * (b - s);
    r[0] = q;
    r[1] = q != 0.0 ? c / q : 0.0;
    return 2;
}

int orc_solve_quartic(const double c[5], double roots[4])
{
    if (c[4] == 0.0) return 0;
    const double a = c[3] / c[4], b = c[2] / c[4], cc = c[1] / c[4], d = c[0] / c[4];
    const double a2 = a * a;
    const double p = b - 0.375 * a2;
    const double q = cc - 0.5 * a * b + 0.125 * a2 * a;
    const double r = d - 0.25 * a * cc + 0.0625 * a2 * b - (3.0 / 256.0) * a2 * a2;
    double y[4];
    int n = 0;
    const double scale = fabs(p) + sqrt(fabs(r)) + 1e-300;
    if (fabs(q) <= 1e-14 * scale * sqrt(scale)) {          /* biquadratic */
        double z[2];
        const int nz = quadratic_roots(p, r, z);
        for (int i = 0; i < nz; ++i)
            if (z[i] >= 0.0) { const double s = sqrt(z[i]); y[n++] = s; y[n++] = -s; }
    } else {
        /* g(m) = 8 m^3 + 8 p m^2 + (2 p^2 - 8 r) m - q^2: g(0) < 0 < g(bound), so a positive root is bracketed;
           Newton safeguarded by bisection (any positive root of the resolvent factors the quartic) */
        const double g2 = 8.0 * p, g1 = 2.0 * p * p - 8.0 * r, g0 = -q * q;
        double bound = fabs(g2);
        if (fabs(g1) > bound) bound = fabs(g1);
        if (fabs(g0) > bound) bound = fabs(g0);
        double lo = 0.0, hi = 1.0 + bound / 8.0, m = hi;
        for (int it = 0; it < 200; ++it) {
            const double g = ((8.0 * m + g2) * m + g1) * m + g0;
            const double dg = (24.0 * m + 2.0 * g2) * m + g1;
            if (g == 0.0) break;
            if (g < 0.0) lo = m; else hi = m;
            double mn = m - g / dg;
            if (!(mn > lo && mn < hi)) mn = 0.5 * (lo + hi);
            if (mn == m) break;
            m = mn;
        }
        if (!(m > 0.0)) return 0;
        const double s = sqrt(2.0 * m);
        const double h = 0.5 * p + m, k = q / (2.0 * s);
        n += quadratic_roots(-s, h + k, y + n);
        n += quadratic_roots(s, h - k, y + n);
    }
    for (int i = 0; i < n; ++i) {
        double x = y[i] - 0.25 * a;
        for (int it = 0; it < 2; ++it) {                    /* polish on the monic quartic */
            const double f = (((x + a) * x + b) * x + cc) * x + d;
            const double df = ((4.0 * x + 3.0 * a) * x + 2.0 * b) * x + cc;
            if (df != 0.0) x = x - f / df;
        }
        roots[i] = x;
    }
    return n;
}

static inline double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(const double *a, const double *b, double *o)
{
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
static inline void normalize3(double *v)
{
    const double n = sqrt(dot3(v, v));
    v[0] /= n; v[1] /= n; v[2] /= n;
}

/* orthonormal frame of a triangle: e1 along P2 - P1, e3 its normal, e2 = e3 x e1; stored as columns B[r][c] */
static void triangle_frame(const double *P1, const double *P2, const double *P3, double B[9])
{
    double e1[3] = {P2[0] - P1[0], P2[1] - P1[1], P2[2] - P1[2]};
    double w[3] = {P3[0] - P1[0], P3[1] - P1[1], P3[2] - P1[2]};
    double e2[3], e3[3];
    normalize3(e1);
    cross3(e1, w, e3);
    normalize3(e3);
    cross3(e3, e1, e2);
    for (int r = 0; r < 3; ++r) { B[r * 3] = e1[r]; B[r * 3 + 1] = e2[r]; B[r * 3 + 2] = e3[r]; }
}

/* P3P: unit bearings f[3][3] (camera frame), world points X[3][3] -> up to 4 poses (R row-major world->camera, t) */
int orc_p3p(const double f[3][3], const double X[3][3], double R[4][9], double t[4][3])
{
    double d23[3], d13[3], d12[3];
    for (int k = 0; k < 3; ++k) { d23[k] = X[1][k] - X[2][k]; d13[k] = X[0][k] - X[2][k]; d12[k] = X[0][k] - X[1][k]; }
    const double a2 = dot3(d23, d23), b2 = dot3(d13, d13), c2 = dot3(d12, d12);
    if (!(a2 > 0.0) || !(b2 > 0.0) || !(c2 > 0.0)) return 0;
    const double ca = dot3(f[1], f[2]), cb = dot3(f[0], f[2]), cg = dot3(f[0], f[1]);
    const double K1 = (a2 - c2) / b2, K2 = c2 / b2;
    /* u = N(v) / D(v);  s2 = u s1, s3 = v s1 */
    const double n2 = K1 - 1.0, n1 = -2.0 * K1 * cb, n0 = K1 + 1.0;
    const double d1 = -2.0 * ca, d0 = 2.0 * cg;
    /* D^2 + N^2 - 2 cg N D - K2 Q D^2 with Q = v^2 - 2 cb v + 1 */
    const double DD[3] = {d0 * d0, 2.0 * d0 * d1, d1 * d1};
    const double q1 = -2.0 * cb;
    double c[5];
    c[0] = DD[0] + n0 * n0 - 2.0 * cg * (n0 * d0) - K2 * DD[0];
    c[1] = DD[1] + 2.0 * n0 * n1 - 2.0 * cg * (n0 * d1 + n1 * d0) - K2 * (DD[1] + q1 * DD[0]);
    c[2] = DD[2] + (n1 * n1 + 2.0 * n0 * n2) - 2.0 * cg * (n1 * d1 + n2 * d0) - K2 * (DD[2] + q1 * DD[1] + DD[0]);
    c[3] = 2.0 * n1 * n2 - 2.0 * cg * (n2 * d1) - K2 * (q1 * DD[2] + DD[1]);
    c[4] = n2 * n2 - K2 * DD[2];
    double v[4];
    const int nr = orc_solve_quartic(c, v);
    int ns = 0;
    for (int i = 0; i < nr; ++i) {
        const double vv = v[i];
        if (!(vv > 0.0)) continue;
        int dup = 0;                                          /* double roots come out twice */
        for (int j = 0; j < i; ++j) dup |= (v[j] == vv);
        if (dup) continue;
        const double D = d1 * vv + d0;
        if (fabs(D) < 1e-12) continue;
        const double u = ((n2 * vv + n1) * vv + n0) / D;
        if (!(u > 0.0)) continue;
        const double Q = (vv + q1) * vv + 1.0;
        if (!(Q > 0.0)) continue;
        const double s1 = sqrt(b2 / Q), s2 = u * s1, s3 = vv * s1;
        double Y[3][3];
        for (int k = 0; k < 3; ++k) { Y[0][k] = s1 * f[0][k]; Y[1][k] = s2 * f[1][k]; Y[2][k] = s3 * f[2][k]; }
        double Bw[9], Bc[9];
        triangle_frame(X[0], X[1], X[2], Bw);
        triangle_frame(Y[0], Y[1], Y[2], Bc);
        double *Rs = R[ns];
        for (int r = 0; r < 3; ++r)
            for (int cidx = 0; cidx < 3; ++cidx)
                Rs[r * 3 + cidx] = Bc[r * 3] * Bw[cidx * 3] + Bc[r * 3 + 1] * Bw[cidx * 3 + 1] + Bc[r * 3 + 2] * Bw[cidx * 3 + 2];
        for (int r = 0; r < 3; ++r) t[ns][r] = Y[0][r] - (Rs[r * 3] * X[0][0] + Rs[r * 3 + 1] * X[0][1] + Rs[r * 3 + 2] * X[0][2]);
        ++ns;
    }
    return ns;
}

static inline double reproj_err2(const double R[9], const double t[3], const double X[3], const double xy[2],
                                 double fx, double fy, double cx, double cy)
{
    const double x = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
    const double y = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
    const double z = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
    const double du = fx * (x / z) + cx - xy[0], dv = fy * (y / z) + cy - xy[1];
    return du * du + dv * dv;
}

/* one hypothesis: 4 correspondences -> pose (or 0 if none) */
int orc_pnp_hypothesis(const double *world, const double *image, const uint32_t idx[4], const double K[9],
                       double R[9], double t[3])
{
    const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    double f[3][3], X[3][3];
    for (int i = 0; i < 3; ++i) {
        const double *p = image + 2 * (size_t)idx[i];
        f[i][0] = (p[0] - cx) / fx; f[i][1] = (p[1] - cy) / fy; f[i][2] = 1.0;
        normalize3(f[i]);
        for (int k = 0; k < 3; ++k) X[i][k] = world[3 * (size_t)idx[i] + k];
    }
    double Rs[4][9], ts[4][3];
    const int ns = orc_p3p(f, X, Rs, ts);
    int best = -1;
    double best_e = 0.0;
    for (int s = 0; s < ns; ++s) {
        const double e = reproj_err2(Rs[s], ts[s], world + 3 * (size_t)idx[3], image + 2 * (size_t)idx[3], fx, fy, cx, cy);
        if (!(e == e)) continue;
        if (best < 0 || e < best_e) { best = s; best_e = e; }
    }
    if (best < 0) return 0;
    memcpy(R, Rs[best], sizeof(double) * 9);
    memcpy(t, ts[best], sizeof(double) * 3);
    return 1;
}

/* consensus test |K pi(R X + t) - x|^2 <= thr2, multiplied through by z^2 so that no division is needed:
 * (fx x + (cx - u) z)^2 + (fy y + (cy - v) z)^2 <= thr2 z^2 */
static inline int reproj_inlier(const double R[9], const double t[3], const double X[3], const double xy[2],
                                double fx, double fy, double cx, double cy, double thr2)
{
    const double x = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
    const double y = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
    const double z = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
    const double du = fx * x + (cx - xy[0]) * z, dv = fy * y + (cy - xy[1]) * z;
    return du * du + dv * dv <= thr2 * (z * z);
}

int orc_pnp_count_inliers(const double *world, const double *image, int n, const double K[9], const double R[9],
                          const double t[3], double thr2, uint8_t *mask)
{
    int cnt = 0;
    for (int i = 0; i < n; ++i) {
        const int in = reproj_inlier(R, t, world + 3 * (size_t)i, image + 2 * (size_t)i, K[0], K[4], K[2], K[5], thr2);
        if (mask) mask[i] = (uint8_t)in;
        cnt += in;
    }
    return cnt;
}

/* Gauss-Newton on the reprojection error over the masked points; R, t updated in place */
void orc_pnp_refine(const double *world, const double *image, int n, const uint8_t *mask, const double K[9],
                    double R[9], double t[3], int max_iter)
{
    const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    for (int it = 0; it < max_iter; ++it) {
        double H[6][6], g[6];
        memset(H, 0, sizeof(H)); memset(g, 0, sizeof(g));
        for (int i = 0; i < n; ++i) {
            if (mask && !mask[i]) continue;
            const double *X = world + 3 * (size_t)i;
            const double x = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
            const double y = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
            const double z = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
            const double iz = 1.0 / z;
            const double ru = fx * (x * iz) + cx - image[2 * (size_t)i], rv = fy * (y * iz) + cy - image[2 * (size_t)i + 1];
            const double a0 = fx * iz, a2 = -fx * x * iz * iz, b1 = fy * iz, b2 = -fy * y * iz * iz;
            /* d(Xc)/d(omega) = [[0, z, -y], [-z, 0, x], [y, -x, 0]], d(Xc)/d(t) = I */
            const double Ju[6] = {a2 * y, a0 * z - a2 * x, -a0 * y, a0, 0.0, a2};
            const double Jv[6] = {-b1 * z + b2 * y, -b2 * x, b1 * x, 0.0, b1, b2};
            for (int r = 0; r < 6; ++r) {
                g[r] += Ju[r] * ru + Jv[r] * rv;
                for (int c = r; c < 6; ++c) H[r][c] += Ju[r] * Ju[c] + Jv[r] * Jv[c];
            }
        }
        /* Cholesky H = L L^T (upper triangle filled), solve H d = -g */
        double L[6][6];
        int ok = 1;
        for (int r = 0; r < 6 && ok; ++r)
            for (int c = 0; c <= r; ++c) {
                double s = H[c][r];
                for (int k = 0; k < c; ++k) s -= L[r][k] * L[c][k];
                if (r == c) { if (!(s > 0.0)) { ok = 0; break; } L[r][r] = sqrt(s); }
                else L[r][c] = s / L[c][c];
            }
        if (!ok) return;
        double yv[6], d[6];
        for (int r = 0; r < 6; ++r) { double s = -g[r]; for (int k = 0; k < r; ++k) s -= L[r][k] * yv[k]; yv[r] = s / L[r][r]; }
        for (int r = 5; r >= 0; --r) { double s = yv[r]; for (int k = r + 1; k < 6; ++k) s -= L[k][r] * d[k]; d[r] = s / L[r][r]; }
        /* R <- orthonormalise((I + [w]x) R), t <- t + w x t + dt */
        double Rn[9];
        for (int c = 0; c < 3; ++c) {
            Rn[c] = R[c] + (d[1] * R[6 + c] - d[2] * R[3 + c]);
            Rn[3 + c] = R[3 + c] + (d[2] * R[c] - d[0] * R[6 + c]);
            Rn[6 + c] = R[6 + c] + (d[0] * R[3 + c] - d[1] * R[c]);
        }
        double r0[3] = {Rn[0], Rn[1], Rn[2]}, r1[3] = {Rn[3], Rn[4], Rn[5]}, r2[3];
        normalize3(r0);
        const double pr = dot3(r1, r0);
        for (int k = 0; k < 3; ++k) r1[k] -= pr * r0[k];
        normalize3(r1);
        cross3(r0, r1, r2);
        for (int k = 0; k < 3; ++k) { R[k] = r0[k]; R[3 + k] = r1[k]; R[6 + k] = r2[k]; }
        const double tn[3] = {t[0] + (d[1] * t[2] - d[2] * t[1]) + d[3], t[1] + (d[2] * t[0] - d[0] * t[2]) + d[4],
                              t[2] + (d[0] * t[1] - d[1] * t[0]) + d[5]};
        t[0] = tn[0]; t[1] = tn[1]; t[2] = tn[2];
        double mx = 0.0;
        for (int k = 0; k < 6; ++k) if (fabs(d[k]) > mx) mx = fabs(d[k]);
        if (mx < 1e-14) return;
    }
}

/* pnp_solve (source/vision/pnp-solve.cpp:16-104): returns ORC status; pose is camera-to-world like the reference's output */
int orc_pnp_solve(const double *world, const double *image, int n, const double K[9], const uint32_t *samples, int H,
                  uint64_t seed, uint64_t problem_id, double reproj_error, int refine_iters,
                  double R_c2w[9], double t_c2w[3], uint8_t *inlier_mask, int *n_inliers, int *best_h,
                  double R_w2c_p3p[9], double t_w2c_p3p[3], int32_t *all_counts)
{
    if (n < 4 || H < 1) return 2;   /* ORC_E_TOO_FEW_POINTS */
    uint32_t *tab = NULL;
    if (!samples) {
        tab = (uint32_t *)malloc(sizeof(uint32_t) * 4 * (size_t)H);
        orc_pnp_sample_table(seed, problem_id, (uint32_t)n, H, tab);
        samples = tab;
    }
    const double thr2 = reproj_error * reproj_error;
    int best = -1, best_cnt = 0;
    double Rb[9], tb[3];
    for (int h = 0; h < H; ++h) {
        double R[9], t[3];
        int cnt = 0;
        if (orc_pnp_hypothesis(world, image, samples + 4 * (size_t)h, K, R, t))
            cnt = orc_pnp_count_inliers(world, image, n, K, R, t, thr2, NULL);
        if (all_counts) all_counts[h] = cnt;
        if (cnt > best_cnt) { best_cnt = cnt; best = h; memcpy(Rb, R, sizeof(Rb)); memcpy(tb, t, sizeof(tb)); }
    }
    free(tab);
    if (n_inliers) *n_inliers = best_cnt;
    if (best_h) *best_h = best;
    if (best < 0 || best_cnt < 4) return 3;   /* ORC_E_NO_MODEL */
    uint8_t *mask = inlier_mask ? inlier_mask : (uint8_t *)malloc((size_t)n);
    orc_pnp_count_inliers(world, image, n, K, Rb, tb, thr2, mask);
    if (R_w2c_p3p) memcpy(R_w2c_p3p, Rb, sizeof(Rb));
    if (t_w2c_p3p) memcpy(t_w2c_p3p, tb, sizeof(tb));
    orc_pnp_refine(world, image, n, mask, K, Rb, tb, refine_iters);
    if (!inlier_mask) free(mask);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R_c2w[r * 3 + c] = Rb[c * 3 + r];
    for (int r = 0; r < 3; ++r) t_c2w[r] = -(R_c2w[r * 3] * tb[0] + R_c2w[r * 3 + 1] * tb[1] + R_c2w[r * 3 + 2] * tb[2]);
    return 0;
}
