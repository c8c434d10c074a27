// ba.cu — bundle adjustment of camera poses and their points on the device (one or two poses: ba_solve_kernel, the shape
// of every reference caller; three to sixteen: ba_solve_multi_kernel): the optimisation problem that
// ba_frame_pose_and_point states through GTSAM (reference source/vision/ba.cpp:26-156) for its callers sfm_refine
// (source/vision/sfm-refine.cpp:20-139), pnp_refine (source/vision/pnp-refine.cpp:16-110) and
// VisualOdometer::track_refine (source/front-end/visual-odometer.cpp:640-800) — SURVEY.md §8f rank 4.
//
//   cost = sum_f 1/2 |[Log(Rg^T R), Rg^T (t - tg)]|^2_{C_f}        PriorFactor<Pose3>      ba.cpp:57-72
//        + sum_j 1/2 |X_j - Xg_j|^2_{C_j}                          PriorFactor<Point3>     ba.cpp:75-93
//        + sum_o 1/2 |K pi(R_f^T (X_j - t_f)) - z_o|^2_{C_o}       GenericProjectionFactor ba.cpp:96-117
//
// (camera-to-world poses, Cal3_S2 intrinsics with skew, Gaussian::Covariance noise).  GTSAM is not vendored by the
// reference; what is implemented is Levenberg-Marquardt with lambda*I damping on that cost, the value of the cost at
// the result (optimizer.error()) and the marginal covariances as the blocks of the inverse Gauss-Newton Hessian
// (gtsam::Marginals).
//
// One CTA per problem (128 threads, thread = point): per-point 3x3 blocks are eliminated (Schur complement), the reduced
// camera system (6 or 12 unknowns) is accumulated in per-thread shared-memory rows that are summed in a fixed order
// (deterministic), factored by Cholesky in one thread; step acceptance by re-evaluating the cost.
#include <cmath>

#include "ba.h"

namespace mvs {

constexpr int BA_THREADS = 128;    // 93 KB of accumulator rows per CTA: two problems resident per SM (64 threads: +13 % batch throughput, +40 % latency)
constexpr int BA_ACC = 91;        // 90 accumulators per thread (+1 pad): 78 + 12 for the reduced 12x12 system
constexpr int BA_WS = 48;         // doubles of workspace per point: V(6) g(3) W0(18) W1(18) candidate X(3)

__device__ __forceinline__ int sym_idx(int r, int c, int n) { return r * n - r * (r - 1) / 2 + (c - r); }   // r <= c

__device__ __forceinline__ void so3_exp(const double w[3], double E[9])
{
    const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = sqrt(th2);
    double a, b;                  // E = I + a [w]x + b [w]x^2
    if (th < 1e-10) { a = 1.0; b = 0.5; } else { a = sin(th) / th; b = (1.0 - cos(th)) / th2; }
    const double wx = w[0], wy = w[1], wz = w[2];
    E[0] = 1.0 - b * (wy * wy + wz * wz); E[1] = -a * wz + b * wx * wy;       E[2] = a * wy + b * wx * wz;
    E[3] = a * wz + b * wx * wy;          E[4] = 1.0 - b * (wx * wx + wz * wz); E[5] = -a * wx + b * wy * wz;
    E[6] = -a * wy + b * wx * wz;         E[7] = a * wx + b * wy * wz;          E[8] = 1.0 - b * (wx * wx + wy * wy);
}

__device__ __forceinline__ void so3_log(const double R[9], double phi[3])
{
    double c = 0.5 * (R[0] + R[4] + R[8] - 1.0);
    c = fmin(1.0, fmax(-1.0, c));
    const double th = acos(c);
    const double v[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};
    const double k = th < 1e-10 ? 0.5 : th / (2.0 * sin(th));
    phi[0] = k * v[0]; phi[1] = k * v[1]; phi[2] = k * v[2];
}

// inverse right Jacobian of SO(3): Log(R Exp(w)) ~ Log(R) + Jr^-1(Log R) w
__device__ __forceinline__ void so3_jr_inv(const double p[3], double J[9])
{
    const double th2 = p[0] * p[0] + p[1] * p[1] + p[2] * p[2], th = sqrt(th2);
    const double k = th < 1e-6 ? 1.0 / 12.0 : 1.0 / th2 - (1.0 + cos(th)) / (2.0 * th * sin(th));
    const double P[9] = {0, -p[2], p[1], p[2], 0, -p[0], -p[1], p[0], 0};
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            double pp = 0.0;
            for (int q = 0; q < 3; ++q) pp += P[r * 3 + q] * P[q * 3 + c];
            J[r * 3 + c] = (r == c ? 1.0 : 0.0) + 0.5 * P[r * 3 + c] + k * pp;
        }
}

// in-place Cholesky of a symmetric N x N matrix stored dense (upper part read), L in the lower part; false if not PD.
// N is a compile-time constant (6 or 12) so that the loops unroll.
template <int N>
__device__ bool cholesky(double *A)
{
#pragma unroll
    for (int r = 0; r < N; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) {
            double s = A[c * N + r];
#pragma unroll
            for (int k = 0; k < c; ++k) s -= A[r * N + k] * A[c * N + k];
            if (r == c) { if (!(s > 0.0)) return false; A[r * N + r] = sqrt(s); }
            else A[r * N + c] = s / A[c * N + c];
        }
    return true;
}
template <int N>
__device__ void cholesky_solve(const double *L, const double *b, double *x)
{
#pragma unroll
    for (int r = 0; r < N; ++r) { double s = b[r];
#pragma unroll
        for (int k = 0; k < r; ++k) s -= L[r * N + k] * x[k];
        x[r] = s / L[r * N + r]; }
#pragma unroll
    for (int r = N - 1; r >= 0; --r) { double s = x[r];
#pragma unroll
        for (int k = r + 1; k < N; ++k) s -= L[k * N + r] * x[k];
        x[r] = s / L[r * N + r]; }
}

struct Proj {
    double e[2];      // reprojection residual
    double Jc[12];    // 2 x 6 w.r.t. [w, v] of the observing camera (R <- R Exp(w), t <- t + R v)
    double JX[6];     // 2 x 3 w.r.t. the point
};

__device__ __forceinline__ void project(const BaArgs &a, const double *R, const double *t, const double *X, const double *z,
                                        bool jac, Proj &o)
{
    const double d[3] = {X[0] - t[0], X[1] - t[1], X[2] - t[2]};
    const double px = R[0] * d[0] + R[3] * d[1] + R[6] * d[2];       // R^T d
    const double py = R[1] * d[0] + R[4] * d[1] + R[7] * d[2];
    const double pz = R[2] * d[0] + R[5] * d[1] + R[8] * d[2];
    const double iz = 1.0 / pz;
    o.e[0] = a.fx * px * iz + a.sk * py * iz + a.u0 - z[0];
    o.e[1] = a.fy * py * iz + a.v0 - z[1];
    if (!jac) return;
    const double D[6] = {a.fx * iz, a.sk * iz, -(a.fx * px + a.sk * py) * iz * iz, 0.0, a.fy * iz, -a.fy * py * iz * iz};
    for (int r = 0; r < 2; ++r) {
        const double d0 = D[r * 3], d1 = D[r * 3 + 1], d2 = D[r * 3 + 2];
        // D [p]x with [p]x = [[0,-pz,py],[pz,0,-px],[-py,px,0]]
        o.Jc[r * 6] = d1 * pz - d2 * py; o.Jc[r * 6 + 1] = -d0 * pz + d2 * px; o.Jc[r * 6 + 2] = d0 * py - d1 * px;
        o.Jc[r * 6 + 3] = -d0; o.Jc[r * 6 + 4] = -d1; o.Jc[r * 6 + 5] = -d2;
        for (int c = 0; c < 3; ++c) o.JX[r * 3 + c] = d0 * R[c * 3] + d1 * R[c * 3 + 1] + d2 * R[c * 3 + 2];   // D R^T
    }
}

__device__ __forceinline__ void info2(const double cov[3], double I[3])     // inverse of [[xx, xy], [xy, yy]]
{
    const double det = cov[0] * cov[2] - cov[1] * cov[1];
    I[0] = cov[2] / det; I[1] = -cov[1] / det; I[2] = cov[0] / det;
}

__device__ void sym3_inverse(const double V[6], double I[6])                  // xx, xy, xz, yy, yz, zz
{
    const double c00 = V[3] * V[5] - V[4] * V[4], c01 = V[2] * V[4] - V[1] * V[5], c02 = V[1] * V[4] - V[2] * V[3];
    const double det = V[0] * c00 + V[1] * c01 + V[2] * c02;
    const double id = 1.0 / det;
    I[0] = c00 * id; I[1] = c01 * id; I[2] = c02 * id;
    I[3] = (V[0] * V[5] - V[2] * V[2]) * id; I[4] = (V[1] * V[2] - V[0] * V[4]) * id; I[5] = (V[0] * V[3] - V[1] * V[1]) * id;
}

__device__ __forceinline__ double block_sum(double v, double *scratch /*[BA_THREADS / 32]*/)
{
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    for (int w = 0; w < BA_THREADS / 32; ++w) r += scratch[w];
    return r;
}

// sum the first `n` accumulator columns over the per-thread rows (fixed order), result in out[0..n)
__device__ __forceinline__ void reduce_rows(const double *acc, int n, double *out)
{
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += BA_THREADS) {
        double s = 0.0;
        for (int r = 0; r < BA_THREADS; ++r) s += acc[r * BA_ACC + k];
        out[k] = s;
    }
    __syncthreads();
}

// One point's contribution to the reduced camera system, N = 6 (one frame) or 12 (two): (W Vl^-1 W^T) into my[0 .. N(N+1)/2)
// and, when WITH_RHS, W Vl^-1 g into my[N(N+1)/2 ..).  Fully unrolled for a compile-time N so that W Vl^-1 stays in registers
// (with a run-time N the arrays live in local memory and this loop was a fifth of the kernel's instructions).
template <int N, bool WITH_RHS>
__device__ __forceinline__ void schur_accumulate(const double *w, double lam, double *my)
{
    const double Vl[6] = {w[0] + lam, w[1], w[2], w[3] + lam, w[4], w[5] + lam};
    double Vi[6];
    sym3_inverse(Vl, Vi);
    double Wm[N * 3], Y[N * 3];
#pragma unroll
    for (int k = 0; k < N * 3; ++k) Wm[k] = w[9 + k];
#pragma unroll
    for (int r = 0; r < N; ++r) {
        Y[r * 3] = Wm[r * 3] * Vi[0] + Wm[r * 3 + 1] * Vi[1] + Wm[r * 3 + 2] * Vi[2];
        Y[r * 3 + 1] = Wm[r * 3] * Vi[1] + Wm[r * 3 + 1] * Vi[3] + Wm[r * 3 + 2] * Vi[4];
        Y[r * 3 + 2] = Wm[r * 3] * Vi[2] + Wm[r * 3 + 1] * Vi[4] + Wm[r * 3 + 2] * Vi[5];
    }
    int k = 0;
#pragma unroll
    for (int r = 0; r < N; ++r) {
#pragma unroll
        for (int c = r; c < N; ++c) { my[k] += Y[r * 3] * Wm[c * 3] + Y[r * 3 + 1] * Wm[c * 3 + 1] + Y[r * 3 + 2] * Wm[c * 3 + 2]; ++k; }
    }
    if (WITH_RHS) {
#pragma unroll
        for (int r = 0; r < N; ++r) my[N * (N + 1) / 2 + r] += Y[r * 3] * w[6] + Y[r * 3 + 1] * w[7] + Y[r * 3 + 2] * w[8];
    }
}

// marginal covariance of one point: V^-1 + (V^-1 W^T) C (W V^-1), C = inverse of the reduced camera system (N x N, shared)
template <int N>
__device__ __forceinline__ void point_covariance(const double *w, const double *C, double *out)
{
    double Vi[6];
    sym3_inverse(w, Vi);
    const double Vf[9] = {Vi[0], Vi[1], Vi[2], Vi[1], Vi[3], Vi[4], Vi[2], Vi[4], Vi[5]};
    double T[3][N];                                         // V^-1 W^T
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < N; ++c) T[r][c] = Vf[r * 3] * w[9 + c * 3] + Vf[r * 3 + 1] * w[9 + c * 3 + 1] + Vf[r * 3 + 2] * w[9 + c * 3 + 2];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        double m[N];                                        // C T[c]^T
#pragma unroll
        for (int k = 0; k < N; ++k) {
            double tc = 0.0;
#pragma unroll
            for (int q = 0; q < N; ++q) tc += C[k * N + q] * T[c][q];
            m[k] = tc;
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            double v = Vf[r * 3 + c];
#pragma unroll
            for (int k = 0; k < N; ++k) v += T[r][k] * m[k];
            out[r * 3 + c] = v;
        }
    }
}

__global__ void __launch_bounds__(BA_THREADS)
ba_solve_kernel(BaArgs a)
{
    extern __shared__ double smem[];
    double *acc = smem;                               // [BA_THREADS][BA_ACC]
    double *s_red = acc + BA_THREADS * BA_ACC;        // [96] reduced sums
    double *s_pose = s_red + 96;                      // [2][12] current poses
    double *s_cand = s_pose + 24;                     // [2][12] candidate poses
    double *s_pinfo = s_cand + 24;                    // [2][36] pose prior information (or flag)
    double *s_S = s_pinfo + 72;                       // [144] reduced system / its factor / its inverse
    double *s_dc = s_S + 144;                         // [12] camera step
    double *s_scr = s_dc + 12;                        // [4]
    __shared__ int s_flag[4];                         // 0: has prior f0, 1: has prior f1, 2: step ok, 3: stop
    __shared__ double s_lambda;
    __shared__ double s_U[42], s_gc[12];              // camera blocks and gradient of the current linearisation

    const int prob = blockIdx.x, tid = threadIdx.x;
    const int f0 = a.frame_off[prob], F = a.frame_off[prob + 1] - f0;
    const int p0 = a.point_off[prob], P = a.point_off[prob + 1] - p0;
    const int n = 6 * F;
    mvs_ba_result *res = a.results + prob;
    double *my = acc + tid * BA_ACC;

    if (tid < 12 * F) { const int f = tid / 12, k = tid % 12; s_pose[tid] = k < 9 ? a.pose_R[(size_t)(f0 + f) * 9 + k] : a.pose_t[(size_t)(f0 + f) * 3 + k - 9]; }
    if (tid == 0) {
        s_lambda = a.lambda0; s_flag[3] = 0;
        for (int f = 0; f < F && f < 2; ++f) {
            const double *C = a.pose_prior_cov + (size_t)(f0 + f) * 36;
            s_flag[f] = C[0] == C[0];                 // NaN = no prior
            if (!s_flag[f]) continue;
            // information = C^-1 through Cholesky: solve C X = I column by column
            double L[36], col[6], x[6];
            for (int k = 0; k < 36; ++k) L[k] = C[k];
            if (!cholesky<6>(L)) { s_flag[3] = 2; break; }
            for (int c = 0; c < 6; ++c) {
                for (int k = 0; k < 6; ++k) col[k] = k == c ? 1.0 : 0.0;
                cholesky_solve<6>(L, col, x);
                for (int k = 0; k < 6; ++k) s_pinfo[f * 36 + k * 6 + c] = x[k];
            }
        }
    }
    __syncthreads();
    if (F > 2) return;                                // ba_solve_multi_kernel's problem
    if (s_flag[3] == 2 || F < 1) { if (tid == 0) { res->status = F < 1 ? MVS_E_UNSUPPORTED : MVS_E_BAD_ARG; res->iterations = 0; } return; }

    // current points live in a.points_out (initialised from the guesses); the guesses stay in a.points as prior means
    for (int j = tid; j < P; j += BA_THREADS)
        for (int k = 0; k < 3; ++k) a.points_out[(size_t)(p0 + j) * 3 + k] = a.points[(size_t)(p0 + j) * 3 + k];

    // ---- cost of (poses, points); cand = 1 evaluates the candidate state
    auto cost_of = [&](bool cand) -> double {
        const double *poses = cand ? s_cand : s_pose;
        double c = 0.0;
        for (int j = tid; j < P; j += BA_THREADS) {
            const size_t gp = (size_t)(p0 + j);
            const double *X = cand ? a.ws + gp * BA_WS + 45 : a.points_out + gp * 3;
            const double *Cp = a.point_prior_cov + gp * 9;
            if (Cp[0] == Cp[0]) {
                const double V[6] = {Cp[0], Cp[1], Cp[2], Cp[4], Cp[5], Cp[8]};
                double I[6];
                sym3_inverse(V, I);
                const double e[3] = {X[0] - a.points[gp * 3], X[1] - a.points[gp * 3 + 1], X[2] - a.points[gp * 3 + 2]};
                c += 0.5 * (e[0] * (I[0] * e[0] + I[1] * e[1] + I[2] * e[2]) + e[1] * (I[1] * e[0] + I[3] * e[1] + I[4] * e[2]) +
                            e[2] * (I[2] * e[0] + I[4] * e[1] + I[5] * e[2]));
            }
            for (int o = a.point_obs_off[gp]; o < a.point_obs_off[gp + 1]; ++o) {
                const mvs_ba_observation &ob = a.obs[o];
                Proj pr;
                project(a, poses + ob.frame * 12, poses + ob.frame * 12 + 9, X, ob.uv, false, pr);
                double I[3];
                info2(ob.cov, I);
                c += 0.5 * (pr.e[0] * (I[0] * pr.e[0] + I[1] * pr.e[1]) + pr.e[1] * (I[1] * pr.e[0] + I[2] * pr.e[1]));
            }
        }
        if (tid == 0)
            for (int f = 0; f < F; ++f) {
                if (!s_flag[f]) continue;
                const double *Rg = a.pose_R + (size_t)(f0 + f) * 9, *tg = a.pose_t + (size_t)(f0 + f) * 3;
                const double *R = poses + f * 12, *t = R + 9;
                double Rr[9], e[6];
                for (int r = 0; r < 3; ++r) for (int q = 0; q < 3; ++q) Rr[r * 3 + q] = Rg[r] * R[q] + Rg[3 + r] * R[3 + q] + Rg[6 + r] * R[6 + q];
                so3_log(Rr, e);
                for (int r = 0; r < 3; ++r) e[3 + r] = Rg[r] * (t[0] - tg[0]) + Rg[3 + r] * (t[1] - tg[1]) + Rg[6 + r] * (t[2] - tg[2]);
                for (int r = 0; r < 6; ++r) { double s = 0.0; for (int q = 0; q < 6; ++q) s += s_pinfo[f * 36 + r * 6 + q] * e[q]; c += 0.5 * e[r] * s; }
            }
        return block_sum(c, s_scr);
    };

    // ---- linearise at the current state: per point V, g, W in the workspace; camera blocks U (21 per frame) and
    //      gradient (6 per frame) in the per-thread accumulator rows: U_f at f*21, g_f at 42 + f*6
    auto linearise = [&]() {
        for (int k = 0; k < 54; ++k) my[k] = 0.0;
        for (int j = tid; j < P; j += BA_THREADS) {
            const size_t gp = (size_t)(p0 + j);
            const double *X = a.points_out + gp * 3;
            double V[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0}, W[36];
            for (int k = 0; k < 36; ++k) W[k] = 0.0;
            const double *Cp = a.point_prior_cov + gp * 9;
            if (Cp[0] == Cp[0]) {
                const double Vc[6] = {Cp[0], Cp[1], Cp[2], Cp[4], Cp[5], Cp[8]};
                double I[6];
                sym3_inverse(Vc, I);
                const double e[3] = {X[0] - a.points[gp * 3], X[1] - a.points[gp * 3 + 1], X[2] - a.points[gp * 3 + 2]};
                for (int k = 0; k < 6; ++k) V[k] += I[k];
                g[0] += I[0] * e[0] + I[1] * e[1] + I[2] * e[2];
                g[1] += I[1] * e[0] + I[3] * e[1] + I[4] * e[2];
                g[2] += I[2] * e[0] + I[4] * e[1] + I[5] * e[2];
            }
            for (int o = a.point_obs_off[gp]; o < a.point_obs_off[gp + 1]; ++o) {
                const mvs_ba_observation &ob = a.obs[o];
                const int f = ob.frame;
                Proj pr;
                project(a, s_pose + f * 12, s_pose + f * 12 + 9, X, ob.uv, true, pr);
                double I[3];
                info2(ob.cov, I);
                // A = info * J (2 x 6 and 2 x 3), ie = info * e
                double Ac[12], AX[6];
                for (int c = 0; c < 6; ++c) { Ac[c] = I[0] * pr.Jc[c] + I[1] * pr.Jc[6 + c]; Ac[6 + c] = I[1] * pr.Jc[c] + I[2] * pr.Jc[6 + c]; }
                for (int c = 0; c < 3; ++c) { AX[c] = I[0] * pr.JX[c] + I[1] * pr.JX[3 + c]; AX[3 + c] = I[1] * pr.JX[c] + I[2] * pr.JX[3 + c]; }
                const double ie0 = I[0] * pr.e[0] + I[1] * pr.e[1], ie1 = I[1] * pr.e[0] + I[2] * pr.e[1];
                double *U = my + f * 21, *gc = my + 42 + f * 6;
                int k = 0;
                for (int r = 0; r < 6; ++r) {
                    for (int c = r; c < 6; ++c) U[k++] += pr.Jc[r] * Ac[c] + pr.Jc[6 + r] * Ac[6 + c];
                    gc[r] += pr.Jc[r] * ie0 + pr.Jc[6 + r] * ie1;
                    for (int c = 0; c < 3; ++c) W[f * 18 + r * 3 + c] += pr.Jc[r] * AX[c] + pr.Jc[6 + r] * AX[3 + c];
                }
                V[0] += pr.JX[0] * AX[0] + pr.JX[3] * AX[3]; V[1] += pr.JX[0] * AX[1] + pr.JX[3] * AX[4]; V[2] += pr.JX[0] * AX[2] + pr.JX[3] * AX[5];
                V[3] += pr.JX[1] * AX[1] + pr.JX[4] * AX[4]; V[4] += pr.JX[1] * AX[2] + pr.JX[4] * AX[5]; V[5] += pr.JX[2] * AX[2] + pr.JX[5] * AX[5];
                for (int c = 0; c < 3; ++c) g[c] += pr.JX[c] * ie0 + pr.JX[3 + c] * ie1;
            }
            double *w = a.ws + gp * BA_WS;
            for (int k = 0; k < 6; ++k) w[k] = V[k];
            for (int k = 0; k < 3; ++k) w[6 + k] = g[k];
            for (int k = 0; k < 36; ++k) w[9 + k] = W[k];
        }
        if (tid == 0)
            for (int f = 0; f < F; ++f) {
                if (!s_flag[f]) continue;
                const double *Rg = a.pose_R + (size_t)(f0 + f) * 9, *tg = a.pose_t + (size_t)(f0 + f) * 3;
                const double *R = s_pose + f * 12, *t = R + 9;
                double Rr[9], e[6], J[36], Jr[9];
                for (int r = 0; r < 3; ++r) for (int q = 0; q < 3; ++q) Rr[r * 3 + q] = Rg[r] * R[q] + Rg[3 + r] * R[3 + q] + Rg[6 + r] * R[6 + q];
                so3_log(Rr, e);
                for (int r = 0; r < 3; ++r) e[3 + r] = Rg[r] * (t[0] - tg[0]) + Rg[3 + r] * (t[1] - tg[1]) + Rg[6 + r] * (t[2] - tg[2]);
                so3_jr_inv(e, Jr);
                for (int k = 0; k < 36; ++k) J[k] = 0.0;
                for (int r = 0; r < 3; ++r) for (int q = 0; q < 3; ++q) { J[r * 6 + q] = Jr[r * 3 + q]; J[(3 + r) * 6 + 3 + q] = Rr[r * 3 + q]; }
                double IJ[36], Ie[6];                       // info * J, info * e
                for (int r = 0; r < 6; ++r) {
                    double s = 0.0;
                    for (int q = 0; q < 6; ++q) s += s_pinfo[f * 36 + r * 6 + q] * e[q];
                    Ie[r] = s;
                    for (int c = 0; c < 6; ++c) { double v = 0.0; for (int q = 0; q < 6; ++q) v += s_pinfo[f * 36 + r * 6 + q] * J[q * 6 + c]; IJ[r * 6 + c] = v; }
                }
                double *U = my + f * 21, *gc = my + 42 + f * 6;
                int k = 0;
                for (int r = 0; r < 6; ++r) {
                    for (int c = r; c < 6; ++c) { double v = 0.0; for (int q = 0; q < 6; ++q) v += J[q * 6 + r] * IJ[q * 6 + c]; U[k++] += v; }
                    double v = 0.0;
                    for (int q = 0; q < 6; ++q) v += J[q * 6 + r] * Ie[q];
                    gc[r] += v;
                }
            }
        reduce_rows(acc, 54, s_red);                        // s_red[0..41] = U blocks, [42..53] = camera gradient
    };

    // ---- reduced camera system for the current lambda, camera step, point steps (candidates), false if not PD
    auto schur_step = [&]() {
        const double lam = s_lambda;
        const int nu = n * (n + 1) / 2;
        for (int k = 0; k < nu + n; ++k) my[k] = 0.0;
        for (int j = tid; j < P; j += BA_THREADS) {
            const double *w = a.ws + (size_t)(p0 + j) * BA_WS;
            if (n == 12) schur_accumulate<12, true>(w, lam, my); else schur_accumulate<6, true>(w, lam, my);
        }
        reduce_rows(acc, nu + n, s_red);
        if (tid == 0) {
            // S = U + lambda I - sum W Vi W^T,  b = -g_c + sum W Vi g_X
            double b[12];
            for (int r = 0; r < n; ++r) {
                for (int c = r; c < n; ++c) {
                    double u = 0.0;
                    if (r / 6 == c / 6) { const int f = r / 6, rr = r % 6, cc = c % 6; u = s_U[f * 21 + sym_idx(rr, cc, 6)]; }
                    s_S[r * n + c] = u + (r == c ? lam : 0.0) - s_red[sym_idx(r, c, n)];
                }
                b[r] = -s_gc[r] + s_red[nu + r];
            }
            bool ok = n == 12 ? cholesky<12>(s_S) : cholesky<6>(s_S);
            if (ok) {
                if (n == 12) cholesky_solve<12>(s_S, b, s_dc); else cholesky_solve<6>(s_S, b, s_dc);
                for (int f = 0; f < F; ++f) {
                    const double *R = s_pose + f * 12, *t = R + 9, *d = s_dc + f * 6;
                    double E[9];
                    so3_exp(d, E);
                    double *Rn = s_cand + f * 12;
                    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Rn[r * 3 + c] = R[r * 3] * E[c] + R[r * 3 + 1] * E[3 + c] + R[r * 3 + 2] * E[6 + c];
                    for (int r = 0; r < 3; ++r) Rn[9 + r] = t[r] + R[r * 3] * d[3] + R[r * 3 + 1] * d[4] + R[r * 3 + 2] * d[5];
                }
            }
            s_flag[2] = ok;
        }
        __syncthreads();
        if (!s_flag[2]) return;
        for (int j = tid; j < P; j += BA_THREADS) {
            double *w = a.ws + (size_t)(p0 + j) * BA_WS;
            const double Vl[6] = {w[0] + lam, w[1], w[2], w[3] + lam, w[4], w[5] + lam};
            double Vi[6];
            sym3_inverse(Vl, Vi);
            double r3[3] = {w[6], w[7], w[8]};               // g_X + W^T dc
            for (int r = 0; r < n; ++r) { r3[0] += w[9 + r * 3] * s_dc[r]; r3[1] += w[9 + r * 3 + 1] * s_dc[r]; r3[2] += w[9 + r * 3 + 2] * s_dc[r]; }
            const double *X = a.points_out + (size_t)(p0 + j) * 3;
            w[45] = X[0] - (Vi[0] * r3[0] + Vi[1] * r3[1] + Vi[2] * r3[2]);
            w[46] = X[1] - (Vi[1] * r3[0] + Vi[3] * r3[1] + Vi[4] * r3[2]);
            w[47] = X[2] - (Vi[2] * r3[0] + Vi[4] * r3[1] + Vi[5] * r3[2]);
        }
        __syncthreads();
    };

    double cost = cost_of(false);
    if (tid == 0) { res->initial_error = cost; }
    int it = 0;
    for (; it < a.max_iter; ++it) {
        linearise();
        if (tid < 42) s_U[tid] = s_red[tid];
        if (tid < 12) s_gc[tid] = s_red[42 + tid];
        __syncthreads();
        bool improved = false;
        double cn = cost, step = 0.0;
        while (true) {
            schur_step();
            if (s_flag[2]) {
                cn = cost_of(true);
                if (cn <= cost) { improved = true; break; }
            }
            __syncthreads();
            if (tid == 0) s_lambda *= 10.0;
            __syncthreads();
            if (!(s_lambda < 1e12)) break;
        }
        if (!improved) break;
        // accept
        for (int f = 0; f < F; ++f) for (int k = 0; k < 6; ++k) step = fmax(step, fabs(s_dc[f * 6 + k]));
        for (int j = tid; j < P; j += BA_THREADS) {
            const double *w = a.ws + (size_t)(p0 + j) * BA_WS;
            double *X = a.points_out + (size_t)(p0 + j) * 3;
            for (int k = 0; k < 3; ++k) { step = fmax(step, fabs(w[45 + k] - X[k])); X[k] = w[45 + k]; }
        }
        __syncthreads();
        if (tid < 12 * F) s_pose[tid] = s_cand[tid];
        const double rel = (cost - cn) / fmax(cost, 1e-300), absdec = cost - cn;
        cost = cn;
        // every thread needs the same verdict: max step over the block
        __syncthreads();
        {
            double m = step;
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, s));
            if ((tid & 31) == 0) s_scr[tid >> 5] = m;
            __syncthreads();
            step = 0.0;
            for (int w = 0; w < BA_THREADS / 32; ++w) step = fmax(step, s_scr[w]);
            __syncthreads();
        }
        if (tid == 0) s_lambda = fmax(s_lambda / 10.0, 1e-12);
        __syncthreads();
        // gtsam::checkConvergence: relative decrease < relativeErrorTol or absolute decrease < absoluteErrorTol
        if (rel < a.rel_tol || (a.abs_tol >= 0.0 && absdec < a.abs_tol) || step < 1e-14) { ++it; break; }
    }

    // ---- marginal covariances at the result: inverse of the Gauss-Newton Hessian, by blocks
    linearise();
    if (tid < 42) s_U[tid] = s_red[tid];
    if (tid < 12) s_gc[tid] = s_red[42 + tid];
    __syncthreads();
    if (tid == 0) s_lambda = 0.0;
    __syncthreads();
    {
        const int nu = n * (n + 1) / 2;
        for (int k = 0; k < nu; ++k) my[k] = 0.0;
        for (int j = tid; j < P; j += BA_THREADS) {
            const double *w = a.ws + (size_t)(p0 + j) * BA_WS;
            if (n == 12) schur_accumulate<12, false>(w, 0.0, my); else schur_accumulate<6, false>(w, 0.0, my);
        }
        reduce_rows(acc, nu, s_red);
        __shared__ double s_L[144];
        __shared__ int s_ok;
        if (tid == 0) {
            for (int r = 0; r < n; ++r)
                for (int c = r; c < n; ++c) {
                    double u = 0.0;
                    if (r / 6 == c / 6) { const int f = r / 6; u = s_U[f * 21 + sym_idx(r % 6, c % 6, 6)]; }
                    s_L[r * n + c] = u - s_red[sym_idx(r, c, n)];
                }
            s_ok = n == 12 ? cholesky<12>(s_L) : cholesky<6>(s_L);
        }
        __syncthreads();
        if (tid < n) {                                      // column tid of the inverse: solve L L^T x = e_tid
            double col[12], x[12];
            for (int k = 0; k < n; ++k) col[k] = k == tid ? 1.0 : 0.0;
            if (s_ok) { if (n == 12) cholesky_solve<12>(s_L, col, x); else cholesky_solve<6>(s_L, col, x); }
            for (int k = 0; k < n; ++k) s_S[k * n + tid] = s_ok ? x[k] : NAN;
        }
        __syncthreads();
        if (tid == 0) {
            for (int f = 0; f < F; ++f) {
                for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) a.pose_cov_out[(size_t)(f0 + f) * 36 + r * 6 + c] = s_S[(f * 6 + r) * n + f * 6 + c];
                for (int k = 0; k < 9; ++k) a.pose_R_out[(size_t)(f0 + f) * 9 + k] = s_pose[f * 12 + k];
                for (int k = 0; k < 3; ++k) a.pose_t_out[(size_t)(f0 + f) * 3 + k] = s_pose[f * 12 + 9 + k];
            }
            res->status = MVS_OK; res->iterations = it; res->final_error = cost;
        }
        __syncthreads();
        // point covariance = Vi + (Vi W^T) C (W Vi)
        for (int j = tid; j < P; j += BA_THREADS) {
            const double *w = a.ws + (size_t)(p0 + j) * BA_WS;
            double *out = a.point_cov_out + (size_t)(p0 + j) * 9;
            if (n == 12) point_covariance<12>(w, s_S, out); else point_covariance<6>(w, s_S, out);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// The same optimisation for 3 .. BA_MAX_FRAMES camera poses (ba_frame_pose_and_point takes any number of frames,
// ba.cpp:26-156; none of the reference's callers passes more than two, so this path is built for generality, not speed).
// One CTA per problem again, same Levenberg-Marquardt policy, same Schur complement; what changes is where the sums
// live: the reduced camera system has up to 96 unknowns, so it is assembled entry by entry -- thread (r, c) adds the
// points' contributions in point order -- instead of in per-thread accumulator rows, and factored by a block-parallel
// Cholesky.  Every sum has a fixed order: the result does not depend on scheduling.
//   per point (ws_multi): V(6) g(3) candidate X(3), W[6F][3], Y = W (V + lambda I)^-1 [6F][3]
//   per observation (ws_obs): its 6x6 camera block (21) and camera gradient (6)
// ------------------------------------------------------------------------------------------------------------------
constexpr int BAM_THREADS = 128;

__device__ __forceinline__ double block_sum_m(double v, double *scratch)
{
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    for (int w = 0; w < BAM_THREADS / 32; ++w) r += scratch[w];
    __syncthreads();
    return r;
}

// in-place Cholesky of the lower triangle of the n x n matrix S (row-major, shared memory) by the whole block;
// returns false (for every thread) if a pivot is not positive
__device__ bool block_cholesky(double *S, int n, int *s_ok)
{
    const int tid = threadIdx.x;
    if (tid == 0) *s_ok = 1;
    __syncthreads();
    for (int k = 0; k < n; ++k) {
        if (tid == 0) {
            const double d = S[k * n + k];
            if (!(d > 0.0)) *s_ok = 0; else S[k * n + k] = sqrt(d);
        }
        __syncthreads();
        if (!*s_ok) break;
        const double dk = S[k * n + k];
        for (int r = k + 1 + tid; r < n; r += BAM_THREADS) S[r * n + k] /= dk;
        __syncthreads();
        const int m = n - k - 1;
        for (int idx = tid; idx < m * m; idx += BAM_THREADS) {
            const int r = k + 1 + idx / m, c = k + 1 + idx % m;
            if (c <= r) S[r * n + c] -= S[r * n + k] * S[c * n + k];
        }
        __syncthreads();
    }
    const bool ok = *s_ok != 0;
    __syncthreads();
    return ok;
}

// L L^T x = b with L the lower triangle of S; one thread
__device__ void lower_solve(const double *L, int n, const double *b, double *x)
{
    for (int r = 0; r < n; ++r) {
        double s = b[r];
        for (int k = 0; k < r; ++k) s -= L[r * n + k] * x[k];
        x[r] = s / L[r * n + r];
    }
    for (int r = n - 1; r >= 0; --r) {
        double s = x[r];
        for (int k = r + 1; k < n; ++k) s -= L[k * n + r] * x[k];
        x[r] = s / L[r * n + r];
    }
}

__global__ void __launch_bounds__(BAM_THREADS)
ba_solve_multi_kernel(BaArgs a, int max_frames)
{
    extern __shared__ double smem[];
    const int prob = blockIdx.x, tid = threadIdx.x;
    const int f0 = a.frame_off[prob], F = a.frame_off[prob + 1] - f0;
    if (F <= 2) return;                               // ba_solve_kernel's problem
    mvs_ba_result *res = a.results + prob;
    if (F > max_frames || F > BA_MAX_FRAMES) { if (tid == 0) { res->status = MVS_E_UNSUPPORTED; res->iterations = 0; } return; }
    const int p0 = a.point_off[prob], P = a.point_off[prob + 1] - p0;
    const int n = 6 * F, NM = 6 * max_frames;
    const int o0 = a.point_obs_off[p0], o1 = a.point_obs_off[p0 + P];
    const int WS = 12 + 36 * F;
    double *wsp = a.ws_multi + a.ws_multi_off[prob];

    double *s_S = smem;                               // [n][n] reduced system, then its factor
    double *s_C = s_S + NM * NM;                      // [n][n] its inverse (covariance phase)
    double *s_U = s_C + NM * NM;                      // [F][27] camera blocks (21) + gradient (6): priors, then priors + observations
    double *s_b = s_U + max_frames * 27;              // [n]
    double *s_dc = s_b + NM;                          // [n]
    double *s_pose = s_dc + NM;                       // [F][12]
    double *s_cand = s_pose + max_frames * 12;        // [F][12]
    double *s_pinfo = s_cand + max_frames * 12;       // [F][36]
    double *s_scr = s_pinfo + max_frames * 36;        // [8]
    __shared__ int s_has[BA_MAX_FRAMES], s_ok, s_bad;
    __shared__ double s_lambda;

    for (int k = tid; k < 12 * F; k += BAM_THREADS) {
        const int f = k / 12, q = k % 12;
        s_pose[k] = q < 9 ? a.pose_R[(size_t)(f0 + f) * 9 + q] : a.pose_t[(size_t)(f0 + f) * 3 + q - 9];
    }
    if (tid == 0) { s_lambda = a.lambda0; s_bad = 0; }
    __syncthreads();
    if (tid < F) {
        const int f = tid;
        const double *C = a.pose_prior_cov + (size_t)(f0 + f) * 36;
        s_has[f] = C[0] == C[0];                      // NaN = no prior
        if (s_has[f]) {
            double L[36], col[6], x[6];
            for (int k = 0; k < 36; ++k) L[k] = C[k];
            if (!cholesky<6>(L)) s_bad = 1;
            else
                for (int c = 0; c < 6; ++c) {
                    for (int k = 0; k < 6; ++k) col[k] = k == c ? 1.0 : 0.0;
                    cholesky_solve<6>(L, col, x);
                    for (int k = 0; k < 6; ++k) s_pinfo[f * 36 + k * 6 + c] = x[k];
                }
        }
    }
    __syncthreads();
    if (s_bad) { if (tid == 0) { res->status = MVS_E_BAD_ARG; res->iterations = 0; } return; }

    for (int j = tid; j < P; j += BAM_THREADS)
        for (int k = 0; k < 3; ++k) a.points_out[(size_t)(p0 + j) * 3 + k] = a.points[(size_t)(p0 + j) * 3 + k];

    // pose-prior residual e (6) of frame f at `poses`, optionally its Jacobian
    auto prior_residual = [&](const double *poses, int f, double e[6], double *J /*36 or null*/) {
        const double *Rg = a.pose_R + (size_t)(f0 + f) * 9, *tg = a.pose_t + (size_t)(f0 + f) * 3;
        const double *R = poses + f * 12, *t = R + 9;
        double Rr[9];
        for (int r = 0; r < 3; ++r) for (int q = 0; q < 3; ++q) Rr[r * 3 + q] = Rg[r] * R[q] + Rg[3 + r] * R[3 + q] + Rg[6 + r] * R[6 + q];
        so3_log(Rr, e);
        for (int r = 0; r < 3; ++r) e[3 + r] = Rg[r] * (t[0] - tg[0]) + Rg[3 + r] * (t[1] - tg[1]) + Rg[6 + r] * (t[2] - tg[2]);
        if (!J) return;
        double Jr[9];
        so3_jr_inv(e, Jr);
        for (int k = 0; k < 36; ++k) J[k] = 0.0;
        for (int r = 0; r < 3; ++r) for (int q = 0; q < 3; ++q) { J[r * 6 + q] = Jr[r * 3 + q]; J[(3 + r) * 6 + 3 + q] = Rr[r * 3 + q]; }
    };

    auto cost_of = [&](bool cand) -> double {
        const double *poses = cand ? s_cand : s_pose;
        double c = 0.0;
        for (int j = tid; j < P; j += BAM_THREADS) {
            const size_t gp = (size_t)(p0 + j);
            const double *X = cand ? wsp + (size_t)j * WS + 9 : a.points_out + gp * 3;
            const double *Cp = a.point_prior_cov + gp * 9;
            if (Cp[0] == Cp[0]) {
                const double V[6] = {Cp[0], Cp[1], Cp[2], Cp[4], Cp[5], Cp[8]};
                double I[6];
                sym3_inverse(V, I);
                const double e[3] = {X[0] - a.points[gp * 3], X[1] - a.points[gp * 3 + 1], X[2] - a.points[gp * 3 + 2]};
                c += 0.5 * (e[0] * (I[0] * e[0] + I[1] * e[1] + I[2] * e[2]) + e[1] * (I[1] * e[0] + I[3] * e[1] + I[4] * e[2]) +
                            e[2] * (I[2] * e[0] + I[4] * e[1] + I[5] * e[2]));
            }
            for (int o = a.point_obs_off[gp]; o < a.point_obs_off[gp + 1]; ++o) {
                const mvs_ba_observation &ob = a.obs[o];
                Proj pr;
                project(a, poses + ob.frame * 12, poses + ob.frame * 12 + 9, X, ob.uv, false, pr);
                double I[3];
                info2(ob.cov, I);
                c += 0.5 * (pr.e[0] * (I[0] * pr.e[0] + I[1] * pr.e[1]) + pr.e[1] * (I[1] * pr.e[0] + I[2] * pr.e[1]));
            }
        }
        if (tid < F && s_has[tid]) {
            double e[6];
            prior_residual(poses, tid, e, nullptr);
            for (int r = 0; r < 6; ++r) { double s = 0.0; for (int q = 0; q < 6; ++q) s += s_pinfo[tid * 36 + r * 6 + q] * e[q]; c += 0.5 * e[r] * s; }
        }
        return block_sum_m(c, s_scr);
    };

    // linearise at the current state: per point V, g, W; per observation its camera block; s_U = camera blocks and gradients
    auto linearise = [&]() {
        for (int j = tid; j < P; j += BAM_THREADS) {
            const size_t gp = (size_t)(p0 + j);
            const double *X = a.points_out + gp * 3;
            double *w = wsp + (size_t)j * WS, *W = w + 12;
            double V[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
            for (int k = 0; k < 18 * F; ++k) W[k] = 0.0;
            const double *Cp = a.point_prior_cov + gp * 9;
            if (Cp[0] == Cp[0]) {
                const double Vc[6] = {Cp[0], Cp[1], Cp[2], Cp[4], Cp[5], Cp[8]};
                double I[6];
                sym3_inverse(Vc, I);
                const double e[3] = {X[0] - a.points[gp * 3], X[1] - a.points[gp * 3 + 1], X[2] - a.points[gp * 3 + 2]};
                for (int k = 0; k < 6; ++k) V[k] += I[k];
                g[0] += I[0] * e[0] + I[1] * e[1] + I[2] * e[2];
                g[1] += I[1] * e[0] + I[3] * e[1] + I[4] * e[2];
                g[2] += I[2] * e[0] + I[4] * e[1] + I[5] * e[2];
            }
            for (int o = a.point_obs_off[gp]; o < a.point_obs_off[gp + 1]; ++o) {
                const mvs_ba_observation &ob = a.obs[o];
                const int f = ob.frame;
                Proj pr;
                project(a, s_pose + f * 12, s_pose + f * 12 + 9, X, ob.uv, true, pr);
                double I[3];
                info2(ob.cov, I);
                double Ac[12], AX[6];
                for (int c = 0; c < 6; ++c) { Ac[c] = I[0] * pr.Jc[c] + I[1] * pr.Jc[6 + c]; Ac[6 + c] = I[1] * pr.Jc[c] + I[2] * pr.Jc[6 + c]; }
                for (int c = 0; c < 3; ++c) { AX[c] = I[0] * pr.JX[c] + I[1] * pr.JX[3 + c]; AX[3 + c] = I[1] * pr.JX[c] + I[2] * pr.JX[3 + c]; }
                const double ie0 = I[0] * pr.e[0] + I[1] * pr.e[1], ie1 = I[1] * pr.e[0] + I[2] * pr.e[1];
                double *uo = a.ws_obs + (size_t)o * 27;
                int k = 0;
                for (int r = 0; r < 6; ++r) {
                    for (int c = r; c < 6; ++c) uo[k++] = pr.Jc[r] * Ac[c] + pr.Jc[6 + r] * Ac[6 + c];
                    uo[21 + r] = pr.Jc[r] * ie0 + pr.Jc[6 + r] * ie1;
                    for (int c = 0; c < 3; ++c) W[f * 18 + r * 3 + c] += pr.Jc[r] * AX[c] + pr.Jc[6 + r] * AX[3 + c];
                }
                V[0] += pr.JX[0] * AX[0] + pr.JX[3] * AX[3]; V[1] += pr.JX[0] * AX[1] + pr.JX[3] * AX[4]; V[2] += pr.JX[0] * AX[2] + pr.JX[3] * AX[5];
                V[3] += pr.JX[1] * AX[1] + pr.JX[4] * AX[4]; V[4] += pr.JX[1] * AX[2] + pr.JX[4] * AX[5]; V[5] += pr.JX[2] * AX[2] + pr.JX[5] * AX[5];
                for (int c = 0; c < 3; ++c) g[c] += pr.JX[c] * ie0 + pr.JX[3 + c] * ie1;
            }
            for (int k = 0; k < 6; ++k) w[k] = V[k];
            for (int k = 0; k < 3; ++k) w[6 + k] = g[k];
        }
        // pose priors: J^T info J and J^T info e of every frame that has one
        for (int k = tid; k < 27 * F; k += BAM_THREADS) s_U[k] = 0.0;
        __syncthreads();
        if (tid < F && s_has[tid]) {
            const int f = tid;
            double e[6], J[36], IJ[36], Ie[6];
            prior_residual(s_pose, f, e, J);
            for (int r = 0; r < 6; ++r) {
                double s = 0.0;
                for (int q = 0; q < 6; ++q) s += s_pinfo[f * 36 + r * 6 + q] * e[q];
                Ie[r] = s;
                for (int c = 0; c < 6; ++c) { double v = 0.0; for (int q = 0; q < 6; ++q) v += s_pinfo[f * 36 + r * 6 + q] * J[q * 6 + c]; IJ[r * 6 + c] = v; }
            }
            int k = 0;
            for (int r = 0; r < 6; ++r) {
                for (int c = r; c < 6; ++c) { double v = 0.0; for (int q = 0; q < 6; ++q) v += J[q * 6 + r] * IJ[q * 6 + c]; s_U[f * 27 + k++] = v; }
                double v = 0.0;
                for (int q = 0; q < 6; ++q) v += J[q * 6 + r] * Ie[q];
                s_U[f * 27 + 21 + r] = v;
            }
        }
        __syncthreads();
        // + the observations of each frame, in observation order
        for (int k = tid; k < 27 * F; k += BAM_THREADS) {
            const int f = k / 27, q = k % 27;
            double s = s_U[k];
            for (int o = o0; o < o1; ++o)
                if (a.obs[o].frame == f) s += a.ws_obs[(size_t)o * 27 + q];
            s_U[k] = s;
        }
        __syncthreads();
    };

    // S = U + lambda I - sum_j W_j (V_j + lambda I)^-1 W_j^T (lower triangle), b likewise when with_rhs; factor; false if not PD
    auto reduced_system = [&](double lam, bool with_rhs) -> bool {
        for (int j = tid; j < P; j += BAM_THREADS) {
            double *w = wsp + (size_t)j * WS;
            const double *W = w + 12;
            double *Y = w + 12 + 18 * F;
            const double Vl[6] = {w[0] + lam, w[1], w[2], w[3] + lam, w[4], w[5] + lam};
            double Vi[6];
            sym3_inverse(Vl, Vi);
            for (int r = 0; r < n; ++r) {
                Y[r * 3] = W[r * 3] * Vi[0] + W[r * 3 + 1] * Vi[1] + W[r * 3 + 2] * Vi[2];
                Y[r * 3 + 1] = W[r * 3] * Vi[1] + W[r * 3 + 1] * Vi[3] + W[r * 3 + 2] * Vi[4];
                Y[r * 3 + 2] = W[r * 3] * Vi[2] + W[r * 3 + 1] * Vi[4] + W[r * 3 + 2] * Vi[5];
            }
        }
        __syncthreads();
        for (int idx = tid; idx < n * n; idx += BAM_THREADS) {
            const int r = idx / n, c = idx % n;
            if (c > r) continue;
            double s = 0.0;
            for (int j = 0; j < P; ++j) {
                const double *w = wsp + (size_t)j * WS;
                const double *Wc = w + 12 + c * 3, *Yr = w + 12 + 18 * F + r * 3;
                s += Yr[0] * Wc[0] + Yr[1] * Wc[1] + Yr[2] * Wc[2];
            }
            double u = 0.0;
            if (r / 6 == c / 6) { const int f = r / 6; u = s_U[f * 27 + sym_idx(c % 6, r % 6, 6)]; }
            s_S[r * n + c] = u + (r == c ? lam : 0.0) - s;
        }
        if (with_rhs)
            for (int r = tid; r < n; r += BAM_THREADS) {
                double s = 0.0;
                for (int j = 0; j < P; ++j) {
                    const double *w = wsp + (size_t)j * WS;
                    const double *Yr = w + 12 + 18 * F + r * 3;
                    s += Yr[0] * w[6] + Yr[1] * w[7] + Yr[2] * w[8];
                }
                s_b[r] = -s_U[(r / 6) * 27 + 21 + r % 6] + s;
            }
        __syncthreads();
        return block_cholesky(s_S, n, &s_ok);
    };

    auto schur_step = [&]() -> bool {
        const double lam = s_lambda;
        if (!reduced_system(lam, true)) return false;
        if (tid == 0) lower_solve(s_S, n, s_b, s_dc);
        __syncthreads();
        if (tid < F) {
            const int f = tid;
            const double *R = s_pose + f * 12, *t = R + 9, *d = s_dc + f * 6;
            double E[9];
            so3_exp(d, E);
            double *Rn = s_cand + f * 12;
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Rn[r * 3 + c] = R[r * 3] * E[c] + R[r * 3 + 1] * E[3 + c] + R[r * 3 + 2] * E[6 + c];
            for (int r = 0; r < 3; ++r) Rn[9 + r] = t[r] + R[r * 3] * d[3] + R[r * 3 + 1] * d[4] + R[r * 3 + 2] * d[5];
        }
        for (int j = tid; j < P; j += BAM_THREADS) {
            double *w = wsp + (size_t)j * WS;
            const double *W = w + 12;
            const double Vl[6] = {w[0] + lam, w[1], w[2], w[3] + lam, w[4], w[5] + lam};
            double Vi[6];
            sym3_inverse(Vl, Vi);
            double r3[3] = {w[6], w[7], w[8]};               // g_X + W^T dc
            for (int r = 0; r < n; ++r) { r3[0] += W[r * 3] * s_dc[r]; r3[1] += W[r * 3 + 1] * s_dc[r]; r3[2] += W[r * 3 + 2] * s_dc[r]; }
            const double *X = a.points_out + (size_t)(p0 + j) * 3;
            w[9] = X[0] - (Vi[0] * r3[0] + Vi[1] * r3[1] + Vi[2] * r3[2]);
            w[10] = X[1] - (Vi[1] * r3[0] + Vi[3] * r3[1] + Vi[4] * r3[2]);
            w[11] = X[2] - (Vi[2] * r3[0] + Vi[4] * r3[1] + Vi[5] * r3[2]);
        }
        __syncthreads();
        return true;
    };

    double cost = cost_of(false);
    if (tid == 0) res->initial_error = cost;
    int it = 0;
    for (; it < a.max_iter; ++it) {
        linearise();
        bool improved = false;
        double cn = cost, step = 0.0;
        while (true) {
            if (schur_step()) {
                cn = cost_of(true);
                if (cn <= cost) { improved = true; break; }
            }
            __syncthreads();
            if (tid == 0) s_lambda *= 10.0;
            __syncthreads();
            if (!(s_lambda < 1e12)) break;
        }
        if (!improved) break;
        for (int k = 0; k < n; ++k) step = fmax(step, fabs(s_dc[k]));
        for (int j = tid; j < P; j += BAM_THREADS) {
            const double *w = wsp + (size_t)j * WS;
            double *X = a.points_out + (size_t)(p0 + j) * 3;
            for (int k = 0; k < 3; ++k) { step = fmax(step, fabs(w[9 + k] - X[k])); X[k] = w[9 + k]; }
        }
        __syncthreads();
        for (int k = tid; k < 12 * F; k += BAM_THREADS) s_pose[k] = s_cand[k];
        const double rel = (cost - cn) / fmax(cost, 1e-300), absdec = cost - cn;
        cost = cn;
        __syncthreads();
        {
            double m = step;
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, s));
            if ((tid & 31) == 0) s_scr[tid >> 5] = m;
            __syncthreads();
            step = 0.0;
            for (int w = 0; w < BAM_THREADS / 32; ++w) step = fmax(step, s_scr[w]);
            __syncthreads();
        }
        if (tid == 0) s_lambda = fmax(s_lambda / 10.0, 1e-12);
        __syncthreads();
        if (rel < a.rel_tol || (a.abs_tol >= 0.0 && absdec < a.abs_tol) || step < 1e-14) { ++it; break; }
    }

    // ---- marginal covariances at the result
    linearise();
    const bool pd = reduced_system(0.0, false);
    for (int c = tid; c < n; c += BAM_THREADS) {      // column c of the inverse
        double col[6 * BA_MAX_FRAMES], x[6 * BA_MAX_FRAMES];
        for (int k = 0; k < n; ++k) col[k] = k == c ? 1.0 : 0.0;
        if (pd) lower_solve(s_S, n, col, x);
        for (int k = 0; k < n; ++k) s_C[k * n + c] = pd ? x[k] : NAN;
    }
    __syncthreads();
    for (int k = tid; k < 36 * F; k += BAM_THREADS) {
        const int f = k / 36, r = (k % 36) / 6, c = k % 6;
        a.pose_cov_out[(size_t)(f0 + f) * 36 + r * 6 + c] = s_C[(f * 6 + r) * n + f * 6 + c];
    }
    for (int k = tid; k < 12 * F; k += BAM_THREADS) {
        const int f = k / 12, q = k % 12;
        if (q < 9) a.pose_R_out[(size_t)(f0 + f) * 9 + q] = s_pose[k]; else a.pose_t_out[(size_t)(f0 + f) * 3 + q - 9] = s_pose[k];
    }
    if (tid == 0) { res->status = MVS_OK; res->iterations = it; res->final_error = cost; }
    // point covariance = Vi + (Vi W^T) C (W Vi)
    for (int j = tid; j < P; j += BAM_THREADS) {
        const double *w = wsp + (size_t)j * WS;
        const double *Y = w + 12 + 18 * F;            // W Vi at lambda = 0 (left by reduced_system)
        double Vi[6];
        sym3_inverse(w, Vi);
        const double Vf[9] = {Vi[0], Vi[1], Vi[2], Vi[1], Vi[3], Vi[4], Vi[2], Vi[4], Vi[5]};
        double *out = a.point_cov_out + (size_t)(p0 + j) * 9;
        for (int c = 0; c < 3; ++c) {
            double m[6 * BA_MAX_FRAMES];               // C Y[:, c]
            for (int k = 0; k < n; ++k) {
                double tc = 0.0;
                for (int q = 0; q < n; ++q) tc += s_C[k * n + q] * Y[q * 3 + c];
                m[k] = tc;
            }
            for (int r = 0; r < 3; ++r) {
                double v = Vf[r * 3 + c];
                for (int k = 0; k < n; ++k) v += Y[k * 3 + r] * m[k];
                out[r * 3 + c] = v;
            }
        }
    }
}

size_t ba_shared_bytes() { return (size_t)(BA_THREADS * BA_ACC + 96 + 24 + 24 + 72 + 144 + 12 + 4) * sizeof(double); }

static size_t ba_multi_shared_bytes(int mf)
{
    const size_t nm = 6 * (size_t)mf;
    return (2 * nm * nm + (size_t)mf * 27 + 2 * nm + (size_t)mf * (12 + 12 + 36) + 8) * sizeof(double);
}

cudaError_t launch_ba(const BaArgs &a, int n_problems, int min_frames, int max_frames, cudaStream_t s)
{
    cudaError_t e = cudaSuccess;
    if (min_frames <= 2) {
        e = cudaFuncSetAttribute(ba_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ba_shared_bytes());
        if (e != cudaSuccess) return e;
        ba_solve_kernel<<<n_problems, BA_THREADS, ba_shared_bytes(), s>>>(a);
    }
    if (max_frames > 2) {
        const int mf = max_frames < BA_MAX_FRAMES ? max_frames : BA_MAX_FRAMES;
        e = cudaFuncSetAttribute(ba_solve_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ba_multi_shared_bytes(mf));
        if (e != cudaSuccess) return e;
        ba_solve_multi_kernel<<<n_problems, BAM_THREADS, ba_multi_shared_bytes(mf), s>>>(a, mf);
    }
    return cudaGetLastError();
}

}  // namespace mvs
