// pnp.cu — pnp_solve on the device (reference source/vision/pnp-solve.cpp:16-104, decl source/vision/pnp.hpp:22-26;
// SURVEY.md §8f rank 3).  The reference forwards to the un-vendored cv::solvePnPRansac(SOLVEPNP_P3P, 100 iterations,
// reprojection error 0.05, confidence 0.95); this file implements the published algorithm of that call as a
// hypothesis x correspondence grid, like the fundamental-matrix RANSAC of ransac.cu:
//
//   Q1 pnp_hypotheses_kernel      thread = hypothesis: seeded 4-point sample, P3P on three points (Grunert's quartic by
//                                 Ferrari's method, resolvent cubic by safeguarded Newton), 4th point picks the pose
//   Q2 pnp_score_kernel           thread = hypothesis (pose in registers), points staged in shared memory tiles,
//                                 squared reprojection error <= threshold^2 (division-free form), counts per tile
//   Q3 pnp_select_refine_kernel   CTA = problem: first best count, inlier mask, Gauss-Newton on the reprojection
//                                 error over the inliers (block-reduced 6x6 normal equations, Cholesky), pose inverse
//
// Only + - * / sqrt in IEEE double without contraction (-fmad=false), in the same order as the CPU checker, so the
// hypothesis poses and inlier counts are bit-identical to it; the refinement differs by summation order only.
#include <cmath>

#include "pnp.h"

namespace mvs {

__host__ __device__ __forceinline__ uint64_t pnp_splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

// row 0 = {0,1,2,3}; rows >= 1: 4 distinct indices < n_points (counter-based, independent of the launch shape)
__host__ __device__ void pnp_sample_row(uint64_t seed, uint64_t problem_id, uint32_t n_points, int h, uint32_t row[4])
{
    if (h == 0 || n_points < 4) { for (int j = 0; j < 4; ++j) row[j] = (uint32_t)j; return; }
    uint64_t st = pnp_splitmix64(seed ^ pnp_splitmix64(problem_id * 0xD1B54A32D192ED03ULL + (uint64_t)h + 0x504E50ULL));
    for (int j = 0; j < 4; ++j) {
        for (;;) {
            st = pnp_splitmix64(st);
            const uint32_t v = (uint32_t)(((st >> 32) * (uint64_t)n_points) >> 32);
            bool dup = false;
            for (int k = 0; k < j; ++k) dup |= (row[k] == v);
            if (!dup) { row[j] = v; break; }
        }
    }
}

void pnp_sample_table_host(uint64_t seed, uint64_t problem_id, uint32_t n_points, int H, uint32_t *out)
{
    for (int h = 0; h < H; ++h) pnp_sample_row(seed, problem_id, n_points, h, out + 4 * (size_t)h);
}

// ------------------------------------------------------------------------------------------ quartic
__device__ __forceinline__ int quadratic_roots(double b, double c, double *r)   // x^2 + b x + c
{
    const double disc = b * b - 4.0 * c;
    if (disc < 0.0) return 0;
    const double s = sqrt(disc);
    const double q = b >= 0.0 ? -0.5 * (b + s) : -0.5 * (b - s);
    r[0] = q;
    r[1] = q != 0.0 ? c / q : 0.0;
    return 2;
}

// real roots of c[4] x^4 + ... + c[0]: depressed quartic, one positive root of the resolvent cubic (bracketed
// Newton), two quadratics, two Newton polishing steps
__device__ int solve_quartic(const double c[5], double roots[4])
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
    if (fabs(q) <= 1e-14 * scale * sqrt(scale)) {
        double z[2];
        const int nz = quadratic_roots(p, r, z);
        for (int i = 0; i < nz; ++i)
            if (z[i] >= 0.0) { const double s = sqrt(z[i]); y[n++] = s; y[n++] = -s; }
    } else {
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
        for (int it = 0; it < 2; ++it) {
            const double f = (((x + a) * x + b) * x + cc) * x + d;
            const double df = ((4.0 * x + 3.0 * a) * x + 2.0 * b) * x + cc;
            if (df != 0.0) x = x - f / df;
        }
        roots[i] = x;
    }
    return n;
}

// ------------------------------------------------------------------------------------------ P3P
__device__ __forceinline__ double pdot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ void pcross3(const double *a, const double *b, double *o)
{
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ void pnormalize3(double *v)
{
    const double n = sqrt(pdot3(v, v));
    v[0] /= n; v[1] /= n; v[2] /= n;
}

// orthonormal frame of a triangle: e1 along P2 - P1, e3 its normal, e2 = e3 x e1, as the columns of B
__device__ void triangle_frame(const double *P1, const double *P2, const double *P3, double B[9])
{
    double e1[3] = {P2[0] - P1[0], P2[1] - P1[1], P2[2] - P1[2]};
    double w[3] = {P3[0] - P1[0], P3[1] - P1[1], P3[2] - P1[2]};
    double e2[3], e3[3];
    pnormalize3(e1);
    pcross3(e1, w, e3);
    pnormalize3(e3);
    pcross3(e3, e1, e2);
    for (int r = 0; r < 3; ++r) { B[r * 3] = e1[r]; B[r * 3 + 1] = e2[r]; B[r * 3 + 2] = e3[r]; }
}

__device__ __forceinline__ double reproj_err2(const double R[9], const double t[3], const double X[3], double u, double v,
                                              double fx, double fy, double cx, double cy)
{
    const double x = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
    const double y = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
    const double z = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
    const double du = fx * (x / z) + cx - u, dv = fy * (y / z) + cy - v;
    return du * du + dv * dv;
}

// consensus test |K pi(R X + t) - x|^2 <= thr2 multiplied through by z^2 (no division):
// (fx x + (cx - u) z)^2 + (fy y + (cy - v) z)^2 <= thr2 z^2
__device__ __forceinline__ bool reproj_inlier(const double R[9], const double t[3], const double X[3], double u, double v,
                                              double fx, double fy, double cx, double cy, double thr2)
{
    const double x = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
    const double y = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
    const double z = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
    const double du = fx * x + (cx - u) * z, dv = fy * y + (cy - v) * z;
    return du * du + dv * dv <= thr2 * (z * z);
}

// 4 correspondences -> pose (world to camera): P3P on the first three, the fourth picks among the real solutions
__device__ bool pnp_hypothesis(const double *world, const double *image, const uint32_t idx[4], double fx, double fy,
                               double cx, double cy, double R[9], double t[3])
{
    double f[3][3], X[3][3];
    for (int i = 0; i < 3; ++i) {
        const double *p = image + 2 * (size_t)idx[i];
        f[i][0] = (p[0] - cx) / fx; f[i][1] = (p[1] - cy) / fy; f[i][2] = 1.0;
        pnormalize3(f[i]);
        for (int k = 0; k < 3; ++k) X[i][k] = world[3 * (size_t)idx[i] + k];
    }
    double d23[3], d13[3], d12[3];
    for (int k = 0; k < 3; ++k) { d23[k] = X[1][k] - X[2][k]; d13[k] = X[0][k] - X[2][k]; d12[k] = X[0][k] - X[1][k]; }
    const double a2 = pdot3(d23, d23), b2 = pdot3(d13, d13), c2 = pdot3(d12, d12);
    if (!(a2 > 0.0) || !(b2 > 0.0) || !(c2 > 0.0)) return false;
    const double ca = pdot3(f[1], f[2]), cb = pdot3(f[0], f[2]), cg = pdot3(f[0], f[1]);
    const double K1 = (a2 - c2) / b2, K2 = c2 / b2;
    // with s2 = u s1, s3 = v s1: u = N(v) / D(v), and D^2 + N^2 - 2 cg N D - K2 Q D^2 = 0 (Q = v^2 - 2 cb v + 1)
    const double n2 = K1 - 1.0, n1 = -2.0 * K1 * cb, n0 = K1 + 1.0;
    const double d1 = -2.0 * ca, d0 = 2.0 * cg;
    const double DD[3] = {d0 * d0, 2.0 * d0 * d1, d1 * d1};
    const double q1 = -2.0 * cb;
    double c[5];
    c[0] = DD[0] + n0 * n0 - 2.0 * cg * (n0 * d0) - K2 * DD[0];
    c[1] = DD[1] + 2.0 * n0 * n1 - 2.0 * cg * (n0 * d1 + n1 * d0) - K2 * (DD[1] + q1 * DD[0]);
    c[2] = DD[2] + (n1 * n1 + 2.0 * n0 * n2) - 2.0 * cg * (n1 * d1 + n2 * d0) - K2 * (DD[2] + q1 * DD[1] + DD[0]);
    c[3] = 2.0 * n1 * n2 - 2.0 * cg * (n2 * d1) - K2 * (q1 * DD[2] + DD[1]);
    c[4] = n2 * n2 - K2 * DD[2];
    double v[4];
    const int nr = solve_quartic(c, v);
    double Bw[9];
    triangle_frame(X[0], X[1], X[2], Bw);
    const double *X4 = world + 3 * (size_t)idx[3];
    const double u4 = image[2 * (size_t)idx[3]], v4 = image[2 * (size_t)idx[3] + 1];
    bool have = false;
    double best_e = 0.0;
    for (int i = 0; i < nr; ++i) {
        const double vv = v[i];
        if (!(vv > 0.0)) continue;
        bool dup = false;
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
        double Bc[9], Rs[9], ts[3];
        triangle_frame(Y[0], Y[1], Y[2], Bc);
        for (int r = 0; r < 3; ++r)
            for (int k = 0; k < 3; ++k)
                Rs[r * 3 + k] = Bc[r * 3] * Bw[k * 3] + Bc[r * 3 + 1] * Bw[k * 3 + 1] + Bc[r * 3 + 2] * Bw[k * 3 + 2];
        for (int r = 0; r < 3; ++r) ts[r] = Y[0][r] - (Rs[r * 3] * X[0][0] + Rs[r * 3 + 1] * X[0][1] + Rs[r * 3 + 2] * X[0][2]);
        const double e = reproj_err2(Rs, ts, X4, u4, v4, fx, fy, cx, cy);
        if (!(e == e)) continue;
        if (!have || e < best_e) {
            have = true; best_e = e;
            for (int k = 0; k < 9; ++k) R[k] = Rs[k];
            for (int k = 0; k < 3; ++k) t[k] = ts[k];
        }
    }
    return have;
}

// ------------------------------------------------------------------------------------------ Q1
__global__ void __launch_bounds__(128)
pnp_hypotheses_kernel(PnpArgs a)
{
    const int prob = blockIdx.y;
    const int h = blockIdx.x * 128 + threadIdx.x;
    if (h >= a.H) return;
    const int off = a.offsets[prob], n = a.offsets[prob + 1] - off;
    double *out = a.poses + ((size_t)prob * a.H + h) * 12;
    bool ok = false;
    double R[9], t[3];
    if (n >= 4) {
        uint32_t row[4];
        if (a.table) { for (int j = 0; j < 4; ++j) row[j] = a.table[4 * h + j]; }
        else pnp_sample_row(a.seed, a.problem_id_base + (uint64_t)prob, (uint32_t)n, h, row);
        bool in_range = true;
        for (int j = 0; j < 4; ++j) in_range &= row[j] < (uint32_t)n;
        if (in_range) ok = pnp_hypothesis(a.world + 3 * (size_t)off, a.image + 2 * (size_t)off, row, a.fx, a.fy, a.cx, a.cy, R, t);
    }
    a.valid[(size_t)prob * a.H + h] = ok ? 1 : 0;
    if (ok) {
        for (int k = 0; k < 9; ++k) out[k] = R[k];
        for (int k = 0; k < 3; ++k) out[9 + k] = t[k];
    }
}

// ------------------------------------------------------------------------------------------ Q2
constexpr int PNP_TILE = 256;

__global__ void __launch_bounds__(128)
pnp_score_kernel(PnpArgs a)
{
    __shared__ double sp[PNP_TILE][5];
    const int prob = blockIdx.z, tile = blockIdx.y;
    const int off = a.offsets[prob], n = a.offsets[prob + 1] - off;
    const int p0 = tile * PNP_TILE;
    if (p0 >= n) return;                                    // uniform per CTA; counts of missing tiles are never read
    const int m = min(PNP_TILE, n - p0);
    for (int i = threadIdx.x; i < m; i += 128) {
        const double *w = a.world + 3 * (size_t)(off + p0 + i), *im = a.image + 2 * (size_t)(off + p0 + i);
        sp[i][0] = w[0]; sp[i][1] = w[1]; sp[i][2] = w[2]; sp[i][3] = im[0]; sp[i][4] = im[1];
    }
    __syncthreads();
    const int h = blockIdx.x * 128 + threadIdx.x;
    if (h >= a.H) return;
    int cnt = 0;
    if (a.valid[(size_t)prob * a.H + h]) {
        const double *P = a.poses + ((size_t)prob * a.H + h) * 12;
        double R[9], t[3];
        for (int k = 0; k < 9; ++k) R[k] = P[k];
        for (int k = 0; k < 3; ++k) t[k] = P[9 + k];
        for (int i = 0; i < m; ++i)
            cnt += reproj_inlier(R, t, sp[i], sp[i][3], sp[i][4], a.fx, a.fy, a.cx, a.cy, a.thr2);
    }
    a.part_count[((size_t)prob * a.tiles + tile) * a.H + h] = cnt;
}

// ------------------------------------------------------------------------------------------ Q3
constexpr int PNP_SEL_THREADS = 256;

__global__ void __launch_bounds__(PNP_SEL_THREADS, 2)
pnp_select_refine_kernel(PnpArgs a)
{
    __shared__ unsigned long long s_best;
    __shared__ double s_pose[12];
    __shared__ double s_part[(PNP_SEL_THREADS / 32) * 27];
    __shared__ double s_sys[27];
    __shared__ int s_stop;
    const int prob = blockIdx.x;
    const int off = a.offsets[prob], n = a.offsets[prob + 1] - off;
    mvs_pnp_result *res = a.results + prob;
    const int tiles_here = (n + PNP_TILE - 1) / PNP_TILE;
    if (threadIdx.x == 0) s_best = 0ull;
    __syncthreads();
    // most inliers, first hypothesis on ties: key = count << 32 | (2^32 - 1 - h)
    unsigned long long mine = 0ull;
    for (int h = threadIdx.x; h < a.H; h += PNP_SEL_THREADS) {
        unsigned cnt = 0;
        for (int tl = 0; tl < tiles_here; ++tl) cnt += (unsigned)a.part_count[((size_t)prob * a.tiles + tl) * a.H + h];
        if (a.all_counts) a.all_counts[(size_t)prob * a.H + h] = (int32_t)cnt;
        const unsigned long long key = ((unsigned long long)cnt << 32) | (unsigned long long)(0xffffffffu - (unsigned)h);
        if (cnt > 0 && key > mine) mine = key;
    }
    if (mine) atomicMax(&s_best, mine);
    __syncthreads();
    const unsigned long long best = s_best;
    const int best_cnt = (int)(best >> 32), best_h = best ? (int)(0xffffffffu - (unsigned)(best & 0xffffffffu)) : -1;
    if (n < 4 || best_h < 0 || best_cnt < a.min_inliers) {
        if (threadIdx.x == 0) {
            res->status = n < 4 ? MVS_E_TOO_FEW_POINTS : MVS_E_NO_MODEL;
            res->n_points = n; res->n_inliers = best_cnt; res->best_hypothesis = best_h;
        }
        if (a.mask) for (int i = threadIdx.x; i < n; i += PNP_SEL_THREADS) a.mask[off + i] = 0;
        return;
    }
    if (threadIdx.x < 12) s_pose[threadIdx.x] = a.poses[((size_t)prob * a.H + best_h) * 12 + threadIdx.x];
    __syncthreads();
    double R[9], t[3];
    for (int k = 0; k < 9; ++k) R[k] = s_pose[k];
    for (int k = 0; k < 3; ++k) t[k] = s_pose[9 + k];
    if (threadIdx.x == 0) {
        for (int k = 0; k < 9; ++k) res->R_w2c_p3p[k] = R[k];
        for (int k = 0; k < 3; ++k) res->t_w2c_p3p[k] = t[k];
    }
    const double *world = a.world + 3 * (size_t)off, *image = a.image + 2 * (size_t)off;
    // inlier flags of the winning minimal-sample pose (what solvePnPRansac returns as inliers)
    for (int i = threadIdx.x; i < n; i += PNP_SEL_THREADS) {
        const bool in = reproj_inlier(R, t, world + 3 * (size_t)i, image[2 * (size_t)i], image[2 * (size_t)i + 1], a.fx, a.fy, a.cx, a.cy, a.thr2);
        a.mask_ws[off + i] = in ? 1 : 0;
        if (a.mask) a.mask[off + i] = in ? 1 : 0;
    }
    __syncthreads();
    // Gauss-Newton on the reprojection error over the inliers: perturbation Xc' = Xc + w x Xc + dt
    for (int it = 0; it < a.refine_iters; ++it) {
        double acc[27];
#pragma unroll
        for (int k = 0; k < 27; ++k) acc[k] = 0.0;
        for (int i = threadIdx.x; i < n; i += PNP_SEL_THREADS) {
            if (!a.mask_ws[off + i]) continue;
            const double *X = world + 3 * (size_t)i;
            const double x = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
            const double y = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
            const double z = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
            const double iz = 1.0 / z;
            const double ru = a.fx * (x * iz) + a.cx - image[2 * (size_t)i], rv = a.fy * (y * iz) + a.cy - image[2 * (size_t)i + 1];
            const double a0 = a.fx * iz, a2 = -a.fx * x * iz * iz, b1 = a.fy * iz, b2 = -a.fy * y * iz * iz;
            const double Ju[6] = {a2 * y, a0 * z - a2 * x, -a0 * y, a0, 0.0, a2};
            const double Jv[6] = {-b1 * z + b2 * y, -b2 * x, b1 * x, 0.0, b1, b2};
            int k = 0;
#pragma unroll
            for (int r = 0; r < 6; ++r) {
#pragma unroll
                for (int c = r; c < 6; ++c) acc[k++] += Ju[r] * Ju[c] + Jv[r] * Jv[c];
            }
#pragma unroll
            for (int r = 0; r < 6; ++r) acc[21 + r] += Ju[r] * ru + Jv[r] * rv;
        }
        // 27 sums: shuffle tree inside each warp, one shared row per warp, summed in warp order by 27 threads
#pragma unroll
        for (int k = 0; k < 27; ++k) {
            double v = acc[k];
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
            if ((threadIdx.x & 31) == 0) s_part[(threadIdx.x >> 5) * 27 + k] = v;
        }
        __syncthreads();
        if (threadIdx.x < 27) {
            double v = 0.0;
            for (int w = 0; w < PNP_SEL_THREADS / 32; ++w) v += s_part[w * 27 + threadIdx.x];
            s_sys[threadIdx.x] = v;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double Hm[6][6], g[6], L[6][6], yv[6], d[6];
            int k = 0;
            for (int r = 0; r < 6; ++r) for (int c = r; c < 6; ++c) Hm[r][c] = s_sys[k++];
            for (int r = 0; r < 6; ++r) g[r] = s_sys[21 + r];
            bool ok = true;
            for (int r = 0; r < 6 && ok; ++r)
                for (int c = 0; c <= r; ++c) {
                    double s = Hm[c][r];
                    for (int q = 0; q < c; ++q) s -= L[r][q] * L[c][q];
                    if (r == c) { if (!(s > 0.0)) { ok = false; break; } L[r][r] = sqrt(s); }
                    else L[r][c] = s / L[c][c];
                }
            int stop = 1;
            if (ok) {
                for (int r = 0; r < 6; ++r) { double s = -g[r]; for (int q = 0; q < r; ++q) s -= L[r][q] * yv[q]; yv[r] = s / L[r][r]; }
                for (int r = 5; r >= 0; --r) { double s = yv[r]; for (int q = r + 1; q < 6; ++q) s -= L[q][r] * d[q]; d[r] = s / L[r][r]; }
                double Rn[9];
                for (int c = 0; c < 3; ++c) {
                    Rn[c] = R[c] + (d[1] * R[6 + c] - d[2] * R[3 + c]);
                    Rn[3 + c] = R[3 + c] + (d[2] * R[c] - d[0] * R[6 + c]);
                    Rn[6 + c] = R[6 + c] + (d[0] * R[3 + c] - d[1] * R[c]);
                }
                double r0[3] = {Rn[0], Rn[1], Rn[2]}, r1[3] = {Rn[3], Rn[4], Rn[5]}, r2[3];
                pnormalize3(r0);
                const double pr = pdot3(r1, r0);
                for (int q = 0; q < 3; ++q) r1[q] -= pr * r0[q];
                pnormalize3(r1);
                pcross3(r0, r1, r2);
                for (int q = 0; q < 3; ++q) { s_pose[q] = r0[q]; s_pose[3 + q] = r1[q]; s_pose[6 + q] = r2[q]; }
                s_pose[9] = t[0] + (d[1] * t[2] - d[2] * t[1]) + d[3];
                s_pose[10] = t[1] + (d[2] * t[0] - d[0] * t[2]) + d[4];
                s_pose[11] = t[2] + (d[0] * t[1] - d[1] * t[0]) + d[5];
                double mx = 0.0;
                for (int q = 0; q < 6; ++q) mx = fmax(mx, fabs(d[q]));
                stop = mx < 1e-14;
            }
            s_stop = stop;
        }
        __syncthreads();
        for (int k = 0; k < 9; ++k) R[k] = s_pose[k];
        for (int k = 0; k < 3; ++k) t[k] = s_pose[9 + k];
        if (s_stop) break;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        res->status = MVS_OK; res->n_points = n; res->n_inliers = best_cnt; res->best_hypothesis = best_h;
        // pose = SE3(R, t).inverse(): camera to world (pnp-solve.cpp:101)
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) res->R_c2w[r * 3 + c] = R[c * 3 + r];
        for (int r = 0; r < 3; ++r)
            res->t_c2w[r] = -(res->R_c2w[r * 3] * t[0] + res->R_c2w[r * 3 + 1] * t[1] + res->R_c2w[r * 3 + 2] * t[2]);
    }
}

int pnp_tiles(int max_points) { return std::max(1, (max_points + PNP_TILE - 1) / PNP_TILE); }

void launch_pnp(const PnpArgs &a, int n_problems, int max_points, cudaStream_t s)
{
    const int hb = (a.H + 127) / 128;
    pnp_hypotheses_kernel<<<dim3(hb, n_problems), 128, 0, s>>>(a);
    pnp_score_kernel<<<dim3(hb, a.tiles, n_problems), 128, 0, s>>>(a);
    pnp_select_refine_kernel<<<n_problems, PNP_SEL_THREADS, 0, s>>>(a);
    (void)max_points;
}

}  // namespace mvs
