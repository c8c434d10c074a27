// common.cuh — shared device helpers for libmvslam_b200 (sm_100a).
//
// Numerical contract: every geometry kernel computes in FP64 in the operation order of the reference's
// Eigen/OpenCV code path.  The translation units are compiled with -fmad=false, so nothing is
// contracted implicitly; the explicit fma() calls (Jacobi inner products/rotations, residuals) are
// part of the contract and are mirrored one-for-one by the CPU oracle, which makes the two sides
// agree to the last bit on almost every input.
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "../../include/mvslam_b200.h"

namespace mvs {

// source/system-config.hpp:8-14
constexpr double kEpsilon = DBL_EPSILON;
constexpr double kTolerance = DBL_EPSILON * 1000.0;
constexpr double kInfinity = DBL_MAX / 10.0;
// source/vision/sfm-solve.cpp:18-21
constexpr double kMaxErrorSq = 5e-2;
constexpr int kMinInliers = 8;

constexpr double kSvdEps2 = (2.0 * DBL_EPSILON) * (2.0 * DBL_EPSILON);
constexpr double kSvdRankTol = 1e-12;
constexpr int kSvdMaxSweeps = 30;

// Hamming top-2 keys pack (distance << 22 | index): lexicographic (distance, index) order in one
// unsigned compare, i.e. OpenCV's lowest-index tie-break for free.  Limits n to 2^22 per side.
constexpr int kIdxBits = 22;
constexpr uint32_t kIdxMask = (1u << kIdxBits) - 1u;
constexpr uint32_t kKeyNone = 0xFFFFFFFFu;

// Programmatic dependent launch: a stage kernel launched with launch_dep() (kernels.h) may become resident while its
// predecessor in the stream is still running; pdl_wait() -- the first statement of every stage kernel, before any read
// of what earlier stages wrote -- blocks until that predecessor has completed and its writes are visible, and
// pdl_launch_dependents() lets the next stage's blocks take their seats likewise.  The launch latency and block
// scheduling of the short geometry kernels then overlap the previous kernel instead of following it (one VO pair is a
// chain of seven launches).  Both are no-ops in a kernel that was launched the ordinary way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// per-pair device state threaded through the stages
struct PairState {
    int32_t status;
    int32_t n_matches;
    int32_t n_inliers;
    int32_t best_h;
    int32_t n_points;
    int32_t candidate;
    int32_t tri_count[4];
    double residual;
    double F[9];
    double E[9];
    double Rc[2][9];
    double tc[3];
    double Rr[2][9];   // SO3(Rc[k]) (so3_rectify), written with Rc by K5
};

// ------------------------------------------------------------------------------------------
// Jacobi rotation of a column pair (p<q): restates the inner step of cv::SVDecomp's one-sided
// Jacobi (un-vendored OpenCV; the reference calls it through source/math/svd.hpp:65).
// Returns false when the pair is already orthogonal to working precision.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool jacobi_cs(double a, double b, double g, double &c, double &s)
{
    // |g| <= eps * sqrt(a*b), evaluated without the square root
    if (g * g <= (kSvdEps2 * a) * b) return false;
    const double g2 = g * 2.0, beta = a - b;
    const double gamma = sqrt(fma(g2, g2, beta * beta));
    const double inv = 1.0 / (gamma * 2.0);
    if (beta < 0) {
        s = sqrt((gamma - beta) * inv);
        c = (g2 * inv) / s;
    } else {
        c = sqrt((gamma + beta) * inv);
        s = (g2 * inv) / c;
    }
    return true;
}

template <int N>
__device__ __forceinline__ bool jacobi_pair(double (&W)[N][N], double (&V)[N][N], const int p, const int q)
{
    double a = 0.0, b = 0.0, g = 0.0;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        a = fma(W[k][p], W[k][p], a);
        b = fma(W[k][q], W[k][q], b);
        g = fma(W[k][p], W[k][q], g);
    }
    double c, s;
    if (!jacobi_cs(a, b, g, c, s)) return false;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double wp = W[k][p], wq = W[k][q];
        W[k][p] = fma(c, wp, s * wq);
        W[k][q] = fma(c, wq, -(s * wp));
        const double vp = V[k][p], vq = V[k][q];
        V[k][p] = fma(c, vp, s * vq);
        V[k][q] = fma(c, vq, -(s * vp));
    }
    return true;
}

// One-sided Jacobi SVD core, one thread per matrix, everything in registers (all indices are
// compile-time after unrolling).  Round-robin pair order identical to the oracle's.
template <int N>
__device__ __forceinline__ void jacobi_svd(double (&W)[N][N], double (&V)[N][N])
{
    constexpr int M = (N + 1) & ~1;
    constexpr int R = M - 1;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < kSvdMaxSweeps; ++sweep) {
        bool changed = false;
#pragma unroll
        for (int s = 0; s < R; ++s) {
            if (M - 1 < N) changed |= jacobi_pair<N>(W, V, s, M - 1);
#pragma unroll
            for (int k = 1; k < M / 2; ++k) {
                const int i = (s + k) % R, j = (s - k + R) % R;
                const int p = i < j ? i : j, q = i < j ? j : i;
                changed |= jacobi_pair<N>(W, V, p, q);
            }
        }
        if (!changed) break;
    }
}

__device__ __forceinline__ void cross3(const double a[3], const double b[3], double o[3])
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

// 3x3 SVD with FULL_UV semantics: A row-major in, U row-major, w descending, Vt row-major.
static __device__ __noinline__ void svd3(const double A[9], double U[9], double w[3], double Vt[9])
{
    double W[3][3], V[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) W[i][j] = A[i * 3 + j];
    jacobi_svd<3>(W, V);
    double sig[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) s += W[k][j] * W[k][j];
        sig[j] = sqrt(s);
    }
    // stable descending order of 3 values
    int o0 = 0, o1 = 1, o2 = 2;
    if (sig[o0] < sig[o1]) { int t = o0; o0 = o1; o1 = t; }
    if (sig[o1] < sig[o2]) {
        int t = o1; o1 = o2; o2 = t;
        if (sig[o0] < sig[o1]) { t = o0; o0 = o1; o1 = t; }
    }
    const int ord[3] = {o0, o1, o2};
    double Wc[3][3], Vc[3][3];  // columns in sorted order (dynamic gather done once through selects)
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int o = ord[j];
            Wc[k][j] = o == 0 ? W[k][0] : (o == 1 ? W[k][1] : W[k][2]);
            Vc[k][j] = o == 0 ? V[k][0] : (o == 1 ? V[k][1] : V[k][2]);
        }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        w[j] = ord[j] == 0 ? sig[0] : (ord[j] == 1 ? sig[1] : sig[2]);
#pragma unroll
        for (int k = 0; k < 3; ++k) Vt[j * 3 + k] = Vc[k][j];
    }
    const double thr = kSvdRankTol * w[0];
    int nvalid = 0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        if (nvalid == j && w[j] > thr && w[j] > 0.0) {
#pragma unroll
            for (int k = 0; k < 3; ++k) U[k * 3 + j] = Wc[k][j] / w[j];
            nvalid = j + 1;
        }
    }
    for (int j = nvalid; j < 3; ++j) {
        if (j == 2) {
            const double u0[3] = {U[0], U[3], U[6]}, u1[3] = {U[1], U[4], U[7]};
            double u2[3];
            cross3(u0, u1, u2);
            U[2] = u2[0]; U[5] = u2[1]; U[8] = u2[2];
            continue;
        }
        // Gram-Schmidt of the unit vector least aligned with the existing columns (rank <= 1 inputs)
        int best = 0;
        double bestv = CUDART_INF;
        for (int e = 0; e < 3; ++e) {
            double v = 0.0;
            for (int c = 0; c < j; ++c) v += U[e * 3 + c] * U[e * 3 + c];
            if (v < bestv) { bestv = v; best = e; }
        }
        double x[3];
        for (int k = 0; k < 3; ++k) x[k] = (k == best) ? 1.0 : 0.0;
        for (int pass = 0; pass < 2; ++pass)
            for (int c = 0; c < j; ++c) {
                double d = 0.0;
                for (int k = 0; k < 3; ++k) d += x[k] * U[k * 3 + c];
                for (int k = 0; k < 3; ++k) x[k] -= d * U[k * 3 + c];
            }
        double nn = 0.0;
        for (int k = 0; k < 3; ++k) nn += x[k] * x[k];
        nn = sqrt(nn);
        for (int k = 0; k < 3; ++k) U[k * 3 + j] = x[k] / nn;
    }
}

// ------------------------------------------------------------------------------------------
// REFERENCE solver: cv::SVDecomp(MODIFY_A | FULL_UV) restated bit for bit.  For the 3x3, 4x4 and 9x9 matrices of
// this path OpenCV (un-vendored dependency, "opencv >= 3.0", README.md:12) never reaches LAPACK: it transposes the
// input and runs its own one-sided Hestenes Jacobi on the ROWS of A^T in plain cyclic (i<j) order, with its own
// scaled hypot, carried squared norms, a selection sort by norm and — for zero singular values — left vectors
// regenerated from cv::RNG(0x12345678).  Everything is IEEE + - * / sqrt in a fixed order (OpenCV's x86-64 baseline
// has no FMA; these translation units are compiled with -fmad=false), so the results equal cv2.SVDecomp's to the
// last bit; tests/test_gpu_reference_solver.py checks that against committed cv2 outputs and live cv2.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double cv_hypot(double a, double b)
{
    a = fabs(a); b = fabs(b);
    if (a > b) { b /= a; return a * sqrt(1 + b * b); }
    if (b > 0) { a /= b; return b * sqrt(1 + a * a); }
    return 0;
}

// At holds A^T on entry (row i = column i of A).  On exit: rows rotated to mutual orthogonality (NOT yet sorted or
// normalised), V the accumulated rotations (row i belongs to row i of At), W[i] = |At row i|.
template <int N>
__device__ __forceinline__ void cv_jacobi(double (&At)[N][N], double (&V)[N][N], double (&W)[N])
{
    constexpr double eps = DBL_EPSILON * 10;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double sd = 0;
#pragma unroll
        for (int k = 0; k < N; ++k) { const double t = At[i][k]; sd += t * t; }
        W[i] = sd;
#pragma unroll
        for (int k = 0; k < N; ++k) V[i][k] = (i == k) ? 1.0 : 0.0;
    }
    for (int iter = 0; iter < 30; ++iter) {   // max_iter = max(m, 30)
        bool changed = false;
#pragma unroll
        for (int i = 0; i < N - 1; ++i)
#pragma unroll
            for (int j = i + 1; j < N; ++j) {
                double a = W[i], p = 0, b = W[j];
#pragma unroll
                for (int k = 0; k < N; ++k) p += At[i][k] * At[j][k];
                if (!(fabs(p) <= eps * sqrt(a * b))) {
                    p *= 2;
                    const double beta = a - b, gamma = cv_hypot(p, beta);
                    double c, s;
                    if (beta < 0) {
                        const double delta = (gamma - beta) * 0.5;
                        s = sqrt(delta / gamma);
                        c = p / (gamma * s * 2);
                    } else {
                        c = sqrt((gamma + beta) / (gamma * 2));
                        s = p / (gamma * c * 2);
                    }
                    a = b = 0;
#pragma unroll
                    for (int k = 0; k < N; ++k) {
                        const double t0 = c * At[i][k] + s * At[j][k];
                        const double t1 = -s * At[i][k] + c * At[j][k];
                        At[i][k] = t0; At[j][k] = t1;
                        a += t0 * t0; b += t1 * t1;
                    }
                    W[i] = a; W[j] = b;
                    changed = true;
#pragma unroll
                    for (int k = 0; k < N; ++k) {
                        const double t0 = c * V[i][k] + s * V[j][k];
                        const double t1 = -s * V[i][k] + c * V[j][k];
                        V[i][k] = t0; V[j][k] = t1;
                    }
                }
            }
        if (!changed) break;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double sd = 0;
#pragma unroll
        for (int k = 0; k < N; ++k) { const double t = At[i][k]; sd += t * t; }
        W[i] = sqrt(sd);
    }
}

// OpenCV's descending selection sort (swap position i with the FIRST maximum of the tail), applied to W and to a
// permutation instead of the rows themselves: afterwards sorted row r is original row perm[r].
template <int N>
__device__ __forceinline__ void cv_sort_perm(double (&W)[N], int (&perm)[N])
{
#pragma unroll
    for (int i = 0; i < N; ++i) perm[i] = i;
#pragma unroll
    for (int i = 0; i < N - 1; ++i) {
        int j = i;
        double wj = W[i];
#pragma unroll
        for (int k = i + 1; k < N; ++k)
            if (wj < W[k]) { j = k; wj = W[k]; }
        // swap(W[i], W[j]), swap(perm[i], perm[j]) with j only known at run time
        const double wi = W[i];
        const int pi = perm[i];
        int pj = pi;
#pragma unroll
        for (int k = i + 1; k < N; ++k)
            if (k == j) { pj = perm[k]; perm[k] = pi; W[k] = wi; }
        W[i] = wj; perm[i] = pj;
    }
}

template <int N>
__device__ __forceinline__ void cv_pick_row(const double (&M)[N][N], int r, double (&out)[N])
{
#pragma unroll
    for (int k = 0; k < N; ++k) {
        double v = M[0][k];
#pragma unroll
        for (int i = 1; i < N; ++i) v = (r == i) ? M[i][k] : v;
        out[k] = v;
    }
}

// Right singular vector of the smallest singular value = vt.row(N-1) of cv::SVDecomp(A) (what the 8-point solve
// and the DLT triangulation take: fundamental-matrix.cpp:114-118, sfm-solve.cpp:193-195).  At = A^T on entry.
template <int N>
__device__ __forceinline__ void cv_svd_last_vt(double (&At)[N][N], double (&x)[N])
{
    double V[N][N], W[N];
    int perm[N];
    cv_jacobi<N>(At, V, W);
    cv_sort_perm<N>(W, perm);
    cv_pick_row<N>(V, perm[N - 1], x);
}

// cv_svd_last_vt with the accumulated rotations V kept in shared memory instead of registers: element (i,k) of this
// thread's V is sV[(i * N + k) * stride].  V is not on the dependent chain of the sweep (the next rotation's angle
// depends on At only), so its loads and stores hide behind the FP64 chain, and the registers it frees double the
// number of decompositions an SM keeps in flight.  Same operations in the same order: bit-identical results.
template <int N>
__device__ __forceinline__ void cv_svd_last_vt_sv(double (&At)[N][N], double *sV, const int stride, double (&x)[N])
{
    constexpr double eps = DBL_EPSILON * 10;
    double W[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double sd = 0;
#pragma unroll
        for (int k = 0; k < N; ++k) { const double t = At[i][k]; sd += t * t; }
        W[i] = sd;
#pragma unroll
        for (int k = 0; k < N; ++k) sV[(i * N + k) * stride] = (i == k) ? 1.0 : 0.0;
    }
    for (int iter = 0; iter < 30; ++iter) {
        bool changed = false;
#pragma unroll
        for (int i = 0; i < N - 1; ++i)
#pragma unroll
            for (int j = i + 1; j < N; ++j) {
                double a = W[i], p = 0, b = W[j];
#pragma unroll
                for (int k = 0; k < N; ++k) p += At[i][k] * At[j][k];
                if (!(fabs(p) <= eps * sqrt(a * b))) {
                    p *= 2;
                    const double beta = a - b, gamma = cv_hypot(p, beta);
                    double c, s;
                    if (beta < 0) {
                        const double delta = (gamma - beta) * 0.5;
                        s = sqrt(delta / gamma);
                        c = p / (gamma * s * 2);
                    } else {
                        c = sqrt((gamma + beta) / (gamma * 2));
                        s = p / (gamma * c * 2);
                    }
                    a = b = 0;
#pragma unroll
                    for (int k = 0; k < N; ++k) {
                        const double t0 = c * At[i][k] + s * At[j][k];
                        const double t1 = -s * At[i][k] + c * At[j][k];
                        At[i][k] = t0; At[j][k] = t1;
                        a += t0 * t0; b += t1 * t1;
                    }
                    W[i] = a; W[j] = b;
                    changed = true;
#pragma unroll
                    for (int k = 0; k < N; ++k) {
                        const double vi = sV[(i * N + k) * stride], vj = sV[(j * N + k) * stride];
                        sV[(i * N + k) * stride] = c * vi + s * vj;
                        sV[(j * N + k) * stride] = -s * vi + c * vj;
                    }
                }
            }
        if (!changed) break;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double sd = 0;
#pragma unroll
        for (int k = 0; k < N; ++k) { const double t = At[i][k]; sd += t * t; }
        W[i] = sqrt(sd);
    }
    int perm[N];
    cv_sort_perm<N>(W, perm);
    const int r = perm[N - 1];
#pragma unroll
    for (int k = 0; k < N; ++k) x[k] = sV[(r * N + k) * stride];
}

// Full cv::SVDecomp of an N x N matrix: A row-major in, U row-major, w descending, Vt row-major (the SVD<> wrapper of
// source/math/svd.hpp:59-72 then takes V = vt^T).
template <int N>
__device__ __forceinline__ void cv_svd_full(const double *A, double *U, double *w, double *Vt)
{
    constexpr double eps = DBL_EPSILON * 10, minval = DBL_MIN;
    double At[N][N], V[N][N], W[N];
    int perm[N];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int k = 0; k < N; ++k) At[i][k] = A[k * N + i];
    cv_jacobi<N>(At, V, W);
    cv_sort_perm<N>(W, perm);
    double As[N][N];
#pragma unroll
    for (int r = 0; r < N; ++r) {
        double a[N], v[N];
        cv_pick_row<N>(At, perm[r], a);
        cv_pick_row<N>(V, perm[r], v);
#pragma unroll
        for (int k = 0; k < N; ++k) { As[r][k] = a[k]; Vt[r * N + k] = v[k]; }
        w[r] = W[r];
    }
    // left singular vectors: normalised rows; zero rows are regenerated from cv::RNG (multiply-with-carry)
    uint64_t rng = 0x12345678ull;
    for (int i = 0; i < N; ++i) {
        double sd = W[i];
        for (int ii = 0; ii < 100 && sd <= minval; ++ii) {
            const double val0 = 1. / N;
            for (int k = 0; k < N; ++k) {
                rng = (uint64_t)(uint32_t)rng * 4164903690u + (uint32_t)(rng >> 32);
                As[i][k] = ((uint32_t)rng & 256u) != 0 ? val0 : -val0;
            }
            for (int it = 0; it < 2; ++it)
                for (int j = 0; j < i; ++j) {
                    sd = 0;
                    for (int k = 0; k < N; ++k) sd += As[i][k] * As[j][k];
                    double asum = 0;
                    for (int k = 0; k < N; ++k) {
                        const double t = As[i][k] - sd * As[j][k];
                        As[i][k] = t;
                        asum += fabs(t);
                    }
                    asum = asum > eps * 100 ? 1 / asum : 0;
                    for (int k = 0; k < N; ++k) As[i][k] *= asum;
                }
            sd = 0;
            for (int k = 0; k < N; ++k) { const double t = As[i][k]; sd += t * t; }
            sd = sqrt(sd);
        }
        const double s = sd > minval ? 1 / sd : 0.;
        for (int k = 0; k < N; ++k) As[i][k] *= s;
    }
    for (int i = 0; i < N; ++i)
        for (int k = 0; k < N; ++k) U[i * N + k] = As[k][i];
}

static __device__ __noinline__ void cv_svd3(const double A[9], double U[9], double w[3], double Vt[9])
{
    cv_svd_full<3>(A, U, w, Vt);
}

__device__ __forceinline__ void mat3_mul(const double A[9], const double B[9], double C[9])
{
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[i * 3 + j] = A[i * 3 + 0] * B[0 * 3 + j] + A[i * 3 + 1] * B[1 * 3 + j] + A[i * 3 + 2] * B[2 * 3 + j];
}

__device__ __forceinline__ void mat3_transpose(const double A[9], double T[9])
{
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) T[i * 3 + j] = A[j * 3 + i];
}

__device__ __forceinline__ void mat3_vec(const double A[9], const double v[3], double o[3])
{
#pragma unroll
    for (int i = 0; i < 3; ++i) o[i] = A[i * 3 + 0] * v[0] + A[i * 3 + 1] * v[1] + A[i * 3 + 2] * v[2];
}

__device__ __forceinline__ double det3(const double M[9])
{
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}

// SO3::rectify (source/math/lie-group.hpp:84-96): row 1 is NOT normalised (reference quirk).
__device__ __forceinline__ void so3_rectify(const double R[9], double out[9])
{
    const double n0 = sqrt(R[0] * R[0] + R[1] * R[1] + R[2] * R[2]);
    const double u0[3] = {R[0] / n0, R[1] / n0, R[2] / n0};
    const double d = R[3] * u0[0] + R[4] * u0[1] + R[5] * u0[2];
    const double u1[3] = {R[3] - d * u0[0], R[4] - d * u0[1], R[5] - d * u0[2]};
    double u2[3];
    cross3(u0, u1, u2);
#pragma unroll
    for (int k = 0; k < 3; ++k) { out[k] = u0[k]; out[3 + k] = u1[k]; out[6 + k] = u2[k]; }
}

// SE3::inverse (lie-group.hpp:203-207) with SO3::inverse (:75-79)
__device__ __forceinline__ void se3_inverse(const double R[9], const double t[3], double Rout[9], double tout[3])
{
    double Rt[9], v[3];
    mat3_transpose(R, Rt);
    so3_rectify(Rt, Rout);
    mat3_vec(Rout, t, v);
    tout[0] = -v[0]; tout[1] = -v[1]; tout[2] = -v[2];
}

// Residual of one correspondence, r = (p2^T F) p1 (estimator-RANSAC.cpp:114-116).
// ALGEBRAIC: inlier iff |r| < thr, residual |r|.  SAMPSON: inlier iff r^2 < thr * den (den > 0, i.e.
// r^2/den < thr without the division), residual r^2/den computed for inliers only.
// CONST_Z: every point of image 1 has the same z (= zc.z1) and every point of image 2 the same z (= zc.z2),
// which is what K^-1 (u,v,1) produces for a pinhole K (z = Kinv[8], usually 1 - 1ulp, not exactly 1).  The
// products z*F are then per-hypothesis constants (FzConst) — the same operands, hence the same bits, as
// multiplying per point.
struct FzConst { double z1, F2z, F5z, F6z, F7z, F8z; };

__device__ __forceinline__ FzConst make_fz(const double (&F)[9], double z1, double z2)
{
    FzConst c;
    c.z1 = z1; c.F2z = F[2] * z1; c.F5z = F[5] * z1; c.F6z = z2 * F[6]; c.F7z = z2 * F[7]; c.F8z = z2 * F[8];
    return c;
}

template <bool CONST_Z, int MODE, bool WANT_RES = true, bool LIT = false>
__device__ __forceinline__ bool point_residual(double x1, double y1, double z1, double x2, double y2, double z2,
                                               const double (&F)[9], const FzConst &zc, double thr, double &res)
{
    double v0, v1, v2, r;
    if (LIT) {
        // REFERENCE solver: (p2^T F) p1 as Eigen's coefficient-based products evaluate it in the reference build
        // (no FMA, EIGEN_DONT_VECTORIZE: SConstruct:70,86): every product rounded, sums left to right
        v0 = (x2 * F[0] + y2 * F[3]) + (CONST_Z ? zc.F6z : z2 * F[6]);
        v1 = (x2 * F[1] + y2 * F[4]) + (CONST_Z ? zc.F7z : z2 * F[7]);
        v2 = (x2 * F[2] + y2 * F[5]) + (CONST_Z ? zc.F8z : z2 * F[8]);
        r = (v0 * x1 + v1 * y1) + v2 * (CONST_Z ? zc.z1 : z1);
    } else if (CONST_Z) {
        v0 = fma(x2, F[0], fma(y2, F[3], zc.F6z));
        v1 = fma(x2, F[1], fma(y2, F[4], zc.F7z));
        v2 = fma(x2, F[2], fma(y2, F[5], zc.F8z));
        r = fma(v0, x1, fma(v1, y1, v2 * zc.z1));
    } else {
        v0 = fma(x2, F[0], fma(y2, F[3], z2 * F[6]));
        v1 = fma(x2, F[1], fma(y2, F[4], z2 * F[7]));
        v2 = fma(x2, F[2], fma(y2, F[5], z2 * F[8]));
        r = fma(v0, x1, fma(v1, y1, v2 * z1));
    }
    if (MODE == MVS_SCORE_ALGEBRAIC) {
        res = fabs(r);
        return res < thr;
    }
    double l0, l1, den;
    if (LIT) {
        l0 = (F[0] * x1 + F[1] * y1) + (CONST_Z ? zc.F2z : F[2] * z1);
        l1 = (F[3] * x1 + F[4] * y1) + (CONST_Z ? zc.F5z : F[5] * z1);
        den = (l0 * l0 + l1 * l1) + (v0 * v0 + v1 * v1);
    } else {
        if (CONST_Z) {
            l0 = fma(F[0], x1, fma(F[1], y1, zc.F2z));
            l1 = fma(F[3], x1, fma(F[4], y1, zc.F5z));
        } else {
            l0 = fma(F[0], x1, fma(F[1], y1, F[2] * z1));
            l1 = fma(F[3], x1, fma(F[4], y1, F[5] * z1));
        }
        den = fma(l0, l0, l1 * l1) + fma(v0, v0, v1 * v1);
    }
    const double r2 = r * r;
    if (!(r2 < thr * den)) return false;
    if (WANT_RES) res = r2 / den;
    return true;
}

// splitmix64 — the seeded sample generator shared with the oracle (integer, bit-exact)
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

__host__ __device__ __forceinline__ void sample_row(uint64_t seed, uint64_t pair_id, uint32_t n_points, int h,
                                                    uint32_t row[8])
{
    if (h == 0 || n_points < 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) row[j] = (uint32_t)j;
        return;
    }
    uint64_t st = splitmix64(seed ^ splitmix64(pair_id * 0xD1B54A32D192ED03ULL + (uint64_t)h));
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        for (;;) {
            st = splitmix64(st);
            const uint32_t v = (uint32_t)(((st >> 32) * (uint64_t)n_points) >> 32);
            bool dup = false;
#pragma unroll
            for (int k = 0; k < 8; ++k) dup |= (k < j) && (row[k] == v);
            if (!dup) { row[j] = v; break; }
        }
    }
}

}  // namespace mvs
