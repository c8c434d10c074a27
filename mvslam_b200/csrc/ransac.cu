// ransac.cu — seeded RANSAC over the normalised 8-point fundamental-matrix estimator.
//
//   K3 hypotheses_kernel : find_fundamental_matrix (reference source/vision/fundamental-matrix.cpp:18-140,
//                          204-267) for every row of the sample table, one thread per hypothesis.
//   K4 score_kernel      : count_inliers (source/vision/estimator-RANSAC.cpp:100-129) on the
//                          hypothesis x correspondence grid.
//   K5 select_kernel     : best-model rule (estimator-RANSAC.cpp:76-84), inlier mask of the winner,
//                          (s,s,0) projection (source/vision/sfm-solve.cpp:73-87) and
//                          decompose_essential_matrix (sfm-solve.cpp:97-127).
//
// FP64 throughout, compiled with -fmad=false (see common.cuh).
#include <math_constants.h>

#include "common.cuh"
#include "kernels.h"

namespace mvs {

// ------------------------------------------------------------------------------------------
// K3.  f = vt.row(8) of SVD(A^T A) (fundamental-matrix.cpp:104-118) is the right singular vector of the
// 8x9 matrix A for its zero singular value, i.e. the unit vector orthogonal to the 8 rows of A.  It is
// computed from A itself (no squared condition number, ~1000x more accurate than the A^T A route, see
// DESIGN.md): Householder QR of A^T (9x8) entirely in registers, f = Q e_8.  One thread per hypothesis —
// the whole solve is ~450 FMA + 8 sqrt + 8 div, so 32 hypotheses per warp with no communication beats
// any cooperative layout.  Same explicit-fma operation order as the oracle (bit-identical F).
// ------------------------------------------------------------------------------------------
constexpr int HYP_THREADS = 128;
constexpr unsigned FULL = 0xFFFFFFFFu;

// find_normalization_transform (fundamental-matrix.cpp:18-54) of the 8 sampled points of one image:
// returns T = [[s,0,tx],[0,s,ty],[0,0,1]] as (s, tx, ty) and the centroid (mx, my); the normalised
// coordinates of a point are then ((x - mx) * s, (y - my) * s), the same operations the reference applies.
__device__ __forceinline__ void normalize8(const double *pts, const uint32_t (&idx)[8], int off,
                                           double (&T)[3], double &mx, double &my)
{
    double px[8], py[8], pz[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const double *p = pts + (size_t)idx[i] * 6 + off;
        px[i] = p[0]; py[i] = p[1]; pz[i] = p[2];
    }
    double mz = 0.0;
    mx = 0.0; my = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { mx += px[i]; my += py[i]; mz += pz[i]; }
    mx *= 0.125; my *= 0.125; mz *= 0.125;
    double scale = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const double dx = px[i] - mx, dy = py[i] - my, dz = pz[i] - mz;
        scale += sqrt(dx * dx + dy * dy + dz * dz);
    }
    scale *= 0.125;
    scale = 1.4142135623730951 / scale;  // sqrt(2.0) correctly rounded
    T[0] = scale; T[1] = -mx * scale; T[2] = -my * scale;
}

// row of A for one normalised correspondence (fundamental-matrix.cpp:76-87)
__device__ __forceinline__ void epipolar_row(const double *p, const double (&T1)[3], double mx1, double my1,
                                             const double (&T2)[3], double mx2, double my2, double (&a)[9])
{
    const double x1 = (p[0] - mx1) * T1[0], y1 = (p[1] - my1) * T1[0];
    const double x2 = (p[3] - mx2) * T2[0], y2 = (p[4] - my2) * T2[0];
    a[0] = x2 * x1; a[1] = x2 * y1; a[2] = x2; a[3] = y2 * x1; a[4] = y2 * y1; a[5] = y2; a[6] = x1; a[7] = y1; a[8] = 1.0;
}

// 8-point solve of one sample by one thread.  pts: [.][6] correspondences, idx: the 8 sampled rows.
__device__ __forceinline__ void eight_point(const double *pts, const uint32_t (&idx)[8], double (&F)[9])
{
    double T1[3], T2[3], mx1, my1, mx2, my2;
    normalize8(pts, idx, 0, T1, mx1, my1);
    normalize8(pts, idx, 3, T2, mx2, my2);
    // M = A^T (9x8): column r = epipolar row of normalised correspondence r (fundamental-matrix.cpp:76-87)
    double M[9][8], beta[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        double a[9];
        epipolar_row(pts + (size_t)idx[r] * 6, T1, mx1, my1, T2, mx2, my2, a);
#pragma unroll
        for (int i = 0; i < 9; ++i) M[i][r] = a[i];
    }
    // Householder QR of M; reflector v_k overwrites column k
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        double s2 = 0.0;
#pragma unroll
        for (int i = k; i < 9; ++i) s2 = fma(M[i][k], M[i][k], s2);
        const double nrm = sqrt(s2);
        if (!(nrm > 0.0)) { beta[k] = 0.0; continue; }
        const double x0 = M[k][k];
        const double alpha = (x0 >= 0.0) ? -nrm : nrm;
        M[k][k] = x0 - alpha;
        beta[k] = 2.0 / (2.0 * fma(nrm, fabs(x0), s2));
#pragma unroll
        for (int j = k + 1; j < 8; ++j) {
            double sj = 0.0;
#pragma unroll
            for (int i = k; i < 9; ++i) sj = fma(M[i][k], M[i][j], sj);
            sj *= beta[k];
#pragma unroll
            for (int i = k; i < 9; ++i) M[i][j] = fma(-sj, M[i][k], M[i][j]);
        }
    }
    // f = Q e_8
    double Fp[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) Fp[i] = (i == 8) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 7; k >= 0; --k) {
        if (beta[k] == 0.0) continue;
        double sk = 0.0;
#pragma unroll
        for (int i = k; i < 9; ++i) sk = fma(M[i][k], Fp[i], sk);
        sk *= beta[k];
#pragma unroll
        for (int i = k; i < 9; ++i) Fp[i] = fma(-sk, M[i][k], Fp[i]);
    }
    // singular constraint (fundamental-matrix.cpp:128-136), then F = T2^T * F * T1 (:245)
    double U[9], w3[3], Vt[9], Fh[9], T2t[9], tmp[9];
    svd3(Fp, U, w3, Vt);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            Fh[i * 3 + j] = (U[i * 3 + 0] * w3[0]) * Vt[0 * 3 + j] + (U[i * 3 + 1] * w3[1]) * Vt[1 * 3 + j];
    const double T1m[9] = {T1[0], 0.0, T1[1], 0.0, T1[0], T1[2], 0.0, 0.0, 1.0};
    const double T2m[9] = {T2[0], 0.0, T2[1], 0.0, T2[0], T2[2], 0.0, 0.0, 1.0};
    mat3_transpose(T2m, T2t);
    mat3_mul(T2t, Fh, tmp);
    mat3_mul(tmp, T1m, F);
}

// REFERENCE solver of the same sample: find_normalization_transform (fundamental-matrix.cpp:18-54),
// find_fundamental_matrix_8point (:56-140) and the de-normalisation (:245) exactly as written there — normalised
// points as (p - mean) * scale, A^T A accumulated element by element over the 8 rows (:104-111), f = vt.row(8) of
// cv::SVDecomp (:114-118), singular constraint u * diag(w0, w1, 0) * vt through a second cv::SVDecomp (:128-136).
__device__ __forceinline__ void eight_point_reference(const double *pts, const uint32_t (&idx)[8], double (&F)[9])
{
    double T1[3], T2[3], mx1, my1, mx2, my2;
    normalize8(pts, idx, 0, T1, mx1, my1);
    normalize8(pts, idx, 3, T2, mx2, my2);
    double At[9][9];
    {
        double A[8][9];
#pragma unroll
        for (int r = 0; r < 8; ++r) epipolar_row(pts + (size_t)idx[r] * 6, T1, mx1, my1, T2, mx2, my2, A[r]);
#pragma unroll
        for (int i = 0; i < 9; ++i)
#pragma unroll
            for (int j = i; j < 9; ++j) {
                double acc = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) acc += A[k][i] * A[k][j];
                At[i][j] = acc; At[j][i] = acc;   // products commute bitwise: A^T A is exactly symmetric, (A^T A)^T = A^T A
            }
    }
    double Fp[9];
    cv_svd_last_vt<9>(At, Fp);
    double U[9], w3[3], Vt[9], Fh[9], T2t[9], tmp[9];
    cv_svd3(Fp, U, w3, Vt);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            Fh[i * 3 + j] = (U[i * 3 + 0] * w3[0]) * Vt[0 * 3 + j] + (U[i * 3 + 1] * w3[1]) * Vt[1 * 3 + j];
    const double T1m[9] = {T1[0], 0.0, T1[1], 0.0, T1[0], T1[2], 0.0, 0.0, 1.0};
    const double T2m[9] = {T2[0], 0.0, T2[1], 0.0, T2[0], T2[2], 0.0, 0.0, 1.0};
    mat3_transpose(T2m, T2t);
    mat3_mul(T2t, Fh, tmp);
    mat3_mul(tmp, T1m, F);
}

// FAST solver: one thread per (pair, hypothesis), flattened over the batch.
__device__ __forceinline__ void hypothesis_body(const HypArgs &a)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int pair = (int)(gid / a.H), h = (int)(gid % a.H);
    if (pair >= a.n_pairs) return;
    if (a.state[pair].status != MVS_OK) return;
    const int n = a.state[pair].n_matches;
    uint32_t idx[8];
    if (a.table) {
#pragma unroll
        for (int j = 0; j < 8; ++j) idx[j] = a.table[(size_t)h * 8 + j];
    } else {
        sample_row(a.seed, a.pair_id_base + (uint64_t)pair, (uint32_t)n, h, idx);
    }
    double F[9];
    eight_point(a.points + (size_t)pair * a.p_stride * 6, idx, F);
    double *o = a.F_all + ((size_t)pair * a.H + h) * 9;
#pragma unroll
    for (int i = 0; i < 9; ++i) o[i] = F[i];
}

// ------------------------------------------------------------------------------------------
// The REFERENCE solve of the batch (eight_point_reference above is the one-thread form, kept for explicit 8-point sets),
// with the 9x9 Jacobi spread over 4 lanes (H = 1 is the reference's configuration: one thread per pair would leave a
// 300-clock div/sqrt chain per rotation, ~290 rotations, on the critical path of every call; for many hypotheses the
// shared-memory state also beats a thread per hypothesis, whose 162 doubles spill: 145 M against 55 M hypotheses/s).  OpenCV visits the row pairs (i, j) of a sweep in lexicographic order; rotation
// (i, j) only depends on the latest earlier rotations that touched row i or row j, i.e. on (i, j-1), (i-1, i) and
// (i-1, j): it can run at "time" i + j, rotations with equal i + j touch disjoint rows, and the next sweep may start 9
// time steps after the current one (row 0 is free after (0, 8)).  Four lanes therefore execute the exact sequential
// algorithm -- every rotation sees bit for bit the operands it would see in program order -- in 9 steps per sweep
// instead of 36.  A sweep that rotates nothing ends the iteration (OpenCV's `if (!changed) break`); the rotations of the
// following sweep that were started early have then, by construction, rotated nothing either.
// Rows live in shared memory: At[9][9], V[9][9], W[9] per hypothesis.
// ------------------------------------------------------------------------------------------
constexpr int WAVE_LANES = 4;
constexpr int WAVE_HYP_PER_BLOCK = 16;                       // 64 threads
constexpr int WAVE_STATE = 9 * 9 * 2 + 9 + 1;                // doubles per hypothesis (+1 keeps consecutive blocks off one bank)

struct WaveSlot { int8_t i, j, older; };                    // older = 1: the rotation belongs to the previous sweep
// phase f = T mod 9 of global step T: rotations with i + j = f (sweep started later) and i + j = f + 9 (started earlier)
__constant__ WaveSlot kWave[9][WAVE_LANES] = {
    {{1, 8, 1}, {2, 7, 1}, {3, 6, 1}, {4, 5, 1}},            // f = 0: t = 9
    {{0, 1, 0}, {2, 8, 1}, {3, 7, 1}, {4, 6, 1}},            // f = 1: t = 1, 10
    {{0, 2, 0}, {3, 8, 1}, {4, 7, 1}, {5, 6, 1}},            // f = 2: t = 2, 11
    {{0, 3, 0}, {1, 2, 0}, {4, 8, 1}, {5, 7, 1}},            // f = 3: t = 3, 12
    {{0, 4, 0}, {1, 3, 0}, {5, 8, 1}, {6, 7, 1}},            // f = 4: t = 4, 13
    {{0, 5, 0}, {1, 4, 0}, {2, 3, 0}, {6, 8, 1}},            // f = 5: t = 5, 14
    {{0, 6, 0}, {1, 5, 0}, {2, 4, 0}, {7, 8, 1}},            // f = 6: t = 6, 15
    {{0, 7, 0}, {1, 6, 0}, {2, 5, 0}, {3, 4, 0}},            // f = 7: t = 7
    {{0, 8, 0}, {1, 7, 0}, {2, 6, 0}, {3, 5, 0}},            // f = 8: t = 8
};

// one rotation of cv_jacobi<9> on rows i, j held in shared memory; returns whether the pair was rotated
__device__ __forceinline__ bool wave_rotate(double *At, double *V, double *W, int i, int j)
{
    constexpr double eps = DBL_EPSILON * 10;
    double ai[9], aj[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) { ai[k] = At[i * 9 + k]; aj[k] = At[j * 9 + k]; }
    double a = W[i], p = 0, b = W[j];
#pragma unroll
    for (int k = 0; k < 9; ++k) p += ai[k] * aj[k];
    if (fabs(p) <= eps * sqrt(a * b)) return false;
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
    for (int k = 0; k < 9; ++k) {
        const double t0 = c * ai[k] + s * aj[k];
        const double t1 = -s * ai[k] + c * aj[k];
        At[i * 9 + k] = t0; At[j * 9 + k] = t1;
        a += t0 * t0; b += t1 * t1;
    }
    W[i] = a; W[j] = b;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const double vi = V[i * 9 + k], vj = V[j * 9 + k];
        V[i * 9 + k] = c * vi + s * vj;
        V[j * 9 + k] = -s * vi + c * vj;
    }
    return true;
}

__global__ void __launch_bounds__(WAVE_HYP_PER_BLOCK * WAVE_LANES, 8)
hypotheses_reference_wave_kernel(HypArgs a)
{
    pdl_wait();
    pdl_launch_dependents();
    __shared__ double s_state[WAVE_HYP_PER_BLOCK][WAVE_STATE];
    __shared__ double s_T[WAVE_HYP_PER_BLOCK][10];            // T1 (s, tx, ty), T2, mx1, my1, mx2, my2 of the sample
    const int slot = threadIdx.x / WAVE_LANES, sub = threadIdx.x % WAVE_LANES;
    const long long gid = (long long)blockIdx.x * WAVE_HYP_PER_BLOCK + slot;
    const int pair = (int)(gid / a.H), h = (int)(gid % a.H);
    const bool live = pair < a.n_pairs && a.state[pair].status == MVS_OK;
    double *At = s_state[slot], *V = At + 81, *W = V + 81;
    const double *pts = a.points + (size_t)(live ? pair : 0) * a.p_stride * 6;
    if (live && sub == 0) {
        uint32_t idx[8];
        if (a.table) {
#pragma unroll
            for (int j = 0; j < 8; ++j) idx[j] = a.table[(size_t)h * 8 + j];
        } else {
            sample_row(a.seed, a.pair_id_base + (uint64_t)pair, (uint32_t)a.state[pair].n_matches, h, idx);
        }
        double T1[3], T2[3], mx1, my1, mx2, my2;
        normalize8(pts, idx, 0, T1, mx1, my1);
        normalize8(pts, idx, 3, T2, mx2, my2);
        double A[8][9];
#pragma unroll
        for (int r = 0; r < 8; ++r) epipolar_row(pts + (size_t)idx[r] * 6, T1, mx1, my1, T2, mx2, my2, A[r]);
#pragma unroll
        for (int i = 0; i < 9; ++i)
#pragma unroll
            for (int j = i; j < 9; ++j) {
                double acc = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) acc += A[k][i] * A[k][j];
                At[i * 9 + j] = acc; At[j * 9 + i] = acc;
            }
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            double sd = 0;
#pragma unroll
            for (int k = 0; k < 9; ++k) { const double t = At[i * 9 + k]; sd += t * t; }
            W[i] = sd;
#pragma unroll
            for (int k = 0; k < 9; ++k) V[i * 9 + k] = (i == k) ? 1.0 : 0.0;
        }
        s_T[slot][0] = T1[0]; s_T[slot][1] = T1[1]; s_T[slot][2] = T1[2]; s_T[slot][3] = T2[0]; s_T[slot][4] = T2[1]; s_T[slot][5] = T2[2];
    }
    __syncwarp();
    // pipelined sweeps: global step T, sweep s covers steps 9 s + 1 .. 9 s + 15
    if (live) {
        unsigned changed_new = 0, changed_old = 0;              // per lane: did one of MY rotations of that sweep rotate
        const unsigned gmask = 0xFu << ((threadIdx.x & 31) / WAVE_LANES * WAVE_LANES);
        for (int T = 1; T <= 9 * 30 + 15; ++T) {
            const int f = T % 9, s_new = T / 9;                 // f >= 1: sweep s_new runs its step f, sweep s_new - 1 its step f + 9
            const WaveSlot ws = kWave[f][sub];
            const int sweep = (f == 0) ? s_new - 1 : (ws.older ? s_new - 1 : s_new);
            bool did = false;
            if (sweep >= 0 && sweep < 30) did = wave_rotate(At, V, W, ws.i, ws.j);
            const bool is_old = (f == 0) || ws.older;
            if (did) { if (is_old) changed_old = 1; else changed_new = 1; }
            __syncwarp(gmask);
            if (f == 6) {                                      // local step 15 of the older sweep just ran: that sweep is complete
                if (T >= 15) {
                    const unsigned any = __ballot_sync(gmask, changed_old != 0) & gmask;
                    if (!any) break;                           // `if (!changed) break`
                    if (s_new - 1 >= 29) break;                // max_iter = 30 sweeps
                }
                changed_old = 0;
            } else if (f == 8) {                               // after its local step 8 the newer sweep becomes the older one
                changed_old = changed_new; changed_new = 0;
            }
        }
    }
    __syncwarp();
    if (live && sub == 0) {
        double Wn[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            double sd = 0;
#pragma unroll
            for (int k = 0; k < 9; ++k) { const double t = At[i * 9 + k]; sd += t * t; }
            Wn[i] = sqrt(sd);
        }
        int perm[9];
        cv_sort_perm<9>(Wn, perm);
        double Fp[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) Fp[k] = V[perm[8] * 9 + k];
        double U[9], w3[3], Vt[9], Fh[9], T2t[9], tmp[9], F[9];
        cv_svd3(Fp, U, w3, Vt);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                Fh[i * 3 + j] = (U[i * 3 + 0] * w3[0]) * Vt[0 * 3 + j] + (U[i * 3 + 1] * w3[1]) * Vt[1 * 3 + j];
        const double T1m[9] = {s_T[slot][0], 0.0, s_T[slot][1], 0.0, s_T[slot][0], s_T[slot][2], 0.0, 0.0, 1.0};
        const double T2m[9] = {s_T[slot][3], 0.0, s_T[slot][4], 0.0, s_T[slot][3], s_T[slot][5], 0.0, 0.0, 1.0};
        mat3_transpose(T2m, T2t);
        mat3_mul(T2t, Fh, tmp);
        mat3_mul(tmp, T1m, F);
        double *o = a.F_all + ((size_t)pair * a.H + h) * 9;
#pragma unroll
        for (int i = 0; i < 9; ++i) o[i] = F[i];
    }
}

__global__ void __launch_bounds__(HYP_THREADS)
hypotheses_kernel(HypArgs a) { pdl_wait(); pdl_launch_dependents(); hypothesis_body(a); }

// a9 entry for explicit 8-point sets: p1s/p2s [n_sets][8][3]
__global__ void __launch_bounds__(HYP_THREADS)
fundamental_sets_kernel(const double *p1s, const double *p2s, int n_sets, double *pts6, double *F_out, int solver)
{
    const int h = blockIdx.x * HYP_THREADS + threadIdx.x;
    if (h >= n_sets) return;
    double *blk = pts6 + (size_t)h * 48;   // this set in the interleaved [8][6] layout
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            blk[r * 6 + k] = p1s[((size_t)h * 8 + r) * 3 + k];
            blk[r * 6 + 3 + k] = p2s[((size_t)h * 8 + r) * 3 + k];
        }
    const uint32_t idx[8] = {0, 1, 2, 3, 4, 5, 6, 7};
    double F[9];
    if (solver == MVS_SOLVER_REFERENCE) eight_point_reference(blk, idx, F);
    else eight_point(blk, idx, F);
#pragma unroll
    for (int i = 0; i < 9; ++i) F_out[(size_t)h * 9 + i] = F[i];
}

// ------------------------------------------------------------------------------------------
// K4.  thread = hypothesis (F in registers), correspondences streamed through shared memory as
// broadcast 128-bit reads; grid = (H/128, point tiles, pairs).  Each thread keeps a private inlier
// count for its tile.  The residual sum only breaks ties between hypotheses with equal counts
// (estimator-RANSAC.cpp:76-84): in ALGEBRAIC mode it costs one predicated add and is accumulated here
// (per tile, in point order); in SAMPSON mode it needs a division, so K5 evaluates it for the tied leaders only.
// ------------------------------------------------------------------------------------------
constexpr int SC_THREADS = 128;
constexpr int SC_TILE = 512;

template <bool UNIT_Z, int MODE, bool LIT>
__global__ void __launch_bounds__(SC_THREADS)
score_kernel(ScoreArgs a)
{
    pdl_wait();
    pdl_launch_dependents();
    constexpr int W = UNIT_Z ? 4 : 6;
    __shared__ __align__(16) double sp[SC_TILE * W];
    const int pair = blockIdx.z, tile = blockIdx.y;
    if (a.zero_word && (blockIdx.x | blockIdx.y | blockIdx.z | threadIdx.x) == 0) *a.zero_word = 0;
    if (a.state[pair].status != MVS_OK) return;
    const int n = a.state[pair].n_matches;
    const int p0 = tile * SC_TILE;
    if (p0 >= n) return;
    const int cnt = min(SC_TILE, n - p0);
    const double *src = a.points + ((size_t)pair * a.p_stride + p0) * 6;
    for (int i = threadIdx.x; i < cnt; i += SC_THREADS) {
        const double *p = src + (size_t)i * 6;
        if (UNIT_Z) { sp[4 * i] = p[0]; sp[4 * i + 1] = p[1]; sp[4 * i + 2] = p[3]; sp[4 * i + 3] = p[4]; }
        else {
#pragma unroll
            for (int k = 0; k < 6; ++k) sp[6 * i + k] = p[k];
        }
    }
    __syncthreads();
    const int h = blockIdx.x * SC_THREADS + threadIdx.x;
    if (h >= a.H) return;
    double F[9];
    const double *Fg = a.F_all + ((size_t)pair * a.H + h) * 9;
#pragma unroll
    for (int i = 0; i < 9; ++i) F[i] = Fg[i];
    const FzConst zc = make_fz(F, a.zc1, a.zc2);
    const double thr = a.max_error_sq;
    uint32_t c = 0;
    double res = 0.0;
    constexpr bool kRes = (MODE == MVS_SCORE_ALGEBRAIC);
#pragma unroll 4
    for (int i = 0; i < cnt; ++i) {
        double r;
        bool in;
        if (UNIT_Z) {
            const double2 u = *reinterpret_cast<const double2 *>(sp + 4 * i);
            const double2 v = *reinterpret_cast<const double2 *>(sp + 4 * i + 2);
            in = point_residual<true, MODE, kRes, LIT>(u.x, u.y, 0.0, v.x, v.y, 0.0, F, zc, thr, r);
        } else {
            const double2 u = *reinterpret_cast<const double2 *>(sp + 6 * i);
            const double2 v = *reinterpret_cast<const double2 *>(sp + 6 * i + 2);
            const double2 w = *reinterpret_cast<const double2 *>(sp + 6 * i + 4);
            in = point_residual<false, MODE, kRes, LIT>(u.x, u.y, v.x, v.y, w.x, w.y, F, zc, thr, r);
        }
        c += in ? 1u : 0u;
        if (kRes && in) res += r;
    }
    a.part_count[((size_t)pair * a.tiles + tile) * a.H + h] = c;
    if (kRes) a.part_res[((size_t)pair * a.tiles + tile) * a.H + h] = res;
}

// ------------------------------------------------------------------------------------------
// K5.  One CTA per pair.
// ------------------------------------------------------------------------------------------
// Block size: 256 threads for long match lists; 64 when a pair has at most a few thousand matches, so that every pair
// of a 1024-pair batch is resident at once (the kernel's time is thread 0's two 3x3 SVDs, a latency every block pays).
constexpr int SEL_THREADS_BIG = 256, SEL_THREADS_SMALL = 64, SEL_SMALL_MAX_POINTS = 4096;

struct Best { uint32_t cnt; double res; int h; };

__device__ __forceinline__ bool better(const Best &x, const Best &y)
{   // x beats y under the sequential rule of estimator-RANSAC.cpp:76-84 (earlier hypothesis wins ties)
    if (x.cnt != y.cnt) return x.cnt > y.cnt;
    if (x.res != y.res) return x.res < y.res;
    return x.h < y.h;
}

// find_essential_matrix own branch (sfm-solve.cpp:73-87)
static __device__ __noinline__ void project_essential(const double F[9], double E[9], bool ref)
{
    double U[9], w[3], Vt[9];
    if (ref) cv_svd3(F, U, w, Vt); else svd3(F, U, w, Vt);
    const double v = sqrt(w[0] * w[1]);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            E[i * 3 + j] = (U[i * 3 + 0] * v) * Vt[0 * 3 + j] + (U[i * 3 + 1] * v) * Vt[1 * 3 + j];
}

// decompose_essential_matrix (sfm-solve.cpp:97-127)
static __device__ __noinline__ void decompose_essential(const double E[9], double Ra[9], double Rb[9], double t[3], bool ref)
{
    double U[9], w[3], Vt[9], V[9];
    if (ref) cv_svd3(E, U, w, Vt); else svd3(E, U, w, Vt);
    mat3_transpose(Vt, V);
    if (det3(U) < 0) {
#pragma unroll
        for (int i = 0; i < 9; ++i) U[i] = -U[i];
    }
    if (det3(V) < 0) {
#pragma unroll
        for (int i = 0; i < 9; ++i) V[i] = -V[i];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double u0 = U[i * 3 + 0], u1 = U[i * 3 + 1], u2 = U[i * 3 + 2];
            Ra[i * 3 + j] = (u1 * V[j * 3 + 0] + (-u0) * V[j * 3 + 1]) + u2 * V[j * 3 + 2];
            Rb[i * 3 + j] = ((-u1) * V[j * 3 + 0] + u0 * V[j * 3 + 1]) + u2 * V[j * 3 + 2];
        }
    const double S01 = (-U[0 * 3 + 1]) * U[1 * 3 + 0] + U[0 * 3 + 0] * U[1 * 3 + 1];
    const double S02 = (-U[0 * 3 + 1]) * U[2 * 3 + 0] + U[0 * 3 + 0] * U[2 * 3 + 1];
    const double S12 = (-U[1 * 3 + 1]) * U[2 * 3 + 0] + U[1 * 3 + 0] * U[2 * 3 + 1];
    t[0] = -S12; t[1] = S02; t[2] = -S01;
}

template <bool UNIT_Z, int MODE, bool LIT, int SEL_THREADS>
__global__ void __launch_bounds__(SEL_THREADS)
select_kernel(SelectArgs a)
{
    pdl_wait();
    pdl_launch_dependents();
    __shared__ uint32_t s_cnt[SEL_THREADS / 32];
    __shared__ int s_status, s_nin, s_k;
    __shared__ uint32_t s_item_base;
    __shared__ Best s_best[SEL_THREADS / 32];
    __shared__ double s_F[9];
    __shared__ uint32_t s_max;
    __shared__ int s_nties;
    const int pair = blockIdx.x;
    PairState *st = a.state + pair;
    if (st->status != MVS_OK) {   // no model for this pair (fewer than 8 matches): its mask row is all outliers, not stale workspace
        uint8_t *mk = a.mask + (size_t)pair * a.p_stride;
        for (int i = threadIdx.x; i < st->n_matches && i < a.p_stride; i += SEL_THREADS) mk[i] = 0;
        return;
    }
    const int n = st->n_matches;
    const int tiles_used = (n + SC_TILE - 1) / SC_TILE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double *pts = a.points + (size_t)pair * a.p_stride * 6;
    int32_t *ties = a.ties + (size_t)pair * a.H * 2;   // [0,H): total counts, [H,2H): tie list

    // ---- total inlier count of every hypothesis, block-wide maximum
    uint32_t cmax = 0;
    for (int h = threadIdx.x; h < a.H; h += SEL_THREADS) {
        uint32_t c = 0;
        for (int t = 0; t < tiles_used; ++t) c += a.part_count[((size_t)pair * a.tiles + t) * a.H + h];
        ties[h] = (int32_t)c;   // parked here until the tie list is built
        if (a.all_counts) a.all_counts[(size_t)pair * a.H + h] = (int32_t)c;
        cmax = max(cmax, c);
    }
    cmax = __reduce_max_sync(FULL, cmax);
    if (lane == 0) s_cnt[warp] = cmax;
    if (threadIdx.x == 0) s_nties = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t m = 0;
        for (int w = 0; w < SEL_THREADS / 32; ++w) m = max(m, s_cnt[w]);
        s_max = m;
    }
    __syncthreads();
    cmax = s_max;

    // ---- best-model rule (estimator-RANSAC.cpp:76-84): most inliers, then the smaller residual sum,
    //      then the earlier hypothesis.  Only the hypotheses tied at the top need their residual sum.
    Best b; b.cnt = cmax; b.res = kInfinity; b.h = 0x7FFFFFFF;
    if (cmax == 0) {
        if (threadIdx.x == 0) { b.res = 0.0; b.h = 0; s_best[0] = b; }   // every residual sum is 0: the first one wins
        __syncthreads();
    } else {
        // list the tied hypotheses (any order: the comparator below is a total order)
        int32_t *list = ties + a.H;
        for (int h = threadIdx.x; h < a.H; h += SEL_THREADS)
            if ((uint32_t)ties[h] == cmax) list[atomicAdd(&s_nties, 1)] = h;
        __syncthreads();
        const int nties = s_nties;
        if (a.part_res) {   // K4 already summed the residuals per tile: add the tiles in order
            for (int t = threadIdx.x; t < nties; t += SEL_THREADS) {
                const int h = list[t];
                double res = 0.0;
                for (int tl = 0; tl < tiles_used; ++tl) res += a.part_res[((size_t)pair * a.tiles + tl) * a.H + h];
                Best x; x.cnt = cmax; x.res = res; x.h = h;
                if (better(x, b)) b = x;
            }
        } else {
        // L lanes cooperate on one tied hypothesis: L = 32 when ties are rare (noisy data), down to one
        // thread per hypothesis when (almost) every hypothesis ties (noise-free data)
        int L = 32;
        while (L > 1 && nties * L > SEL_THREADS) L >>= 1;
        const int sub = threadIdx.x & (L - 1);
        double F[9];
        for (int t = threadIdx.x / L; t < nties; t += SEL_THREADS / L) {
            const int h = list[t];
            const double *Fg = a.F_all + ((size_t)pair * a.H + h) * 9;
#pragma unroll
            for (int i = 0; i < 9; ++i) F[i] = Fg[i];
            const FzConst zc = make_fz(F, a.zc1, a.zc2);
            double res = 0.0;
            for (int i = sub; i < n; i += L) {
                const double *p = pts + (size_t)i * 6;
                double r;
                if (point_residual<UNIT_Z, MODE, true, LIT>(p[0], p[1], p[2], p[3], p[4], p[5], F, zc, a.max_error_sq, r)) res += r;
            }
            for (int off = L >> 1; off > 0; off >>= 1) res += __shfl_xor_sync(__activemask(), res, off);
            Best x; x.cnt = cmax; x.res = res; x.h = h;
            if (better(x, b)) b = x;
        }
        }
        // fold the per-thread leaders of the warp
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            Best o;
            o.cnt = cmax;
            o.res = __shfl_xor_sync(FULL, b.res, off);
            o.h = __shfl_xor_sync(FULL, b.h, off);
            if (better(o, b)) b = o;
        }
        if (lane == 0) s_best[warp] = b;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (cmax != 0) {
            b = s_best[0];
            for (int w = 1; w < SEL_THREADS / 32; ++w)
                if (better(s_best[w], b)) b = s_best[w];
        } else b = s_best[0];
        st->best_h = b.h;
        st->n_inliers = (int)b.cnt;
        st->residual = b.res;
        double F[9];
        const double *Fg = a.F_all + ((size_t)pair * a.H + b.h) * 9;
#pragma unroll
        for (int i = 0; i < 9; ++i) { F[i] = Fg[i]; s_F[i] = F[i]; st->F[i] = F[i]; }
        int status = MVS_OK;
        if (b.cnt == 0) status = MVS_E_NO_MODEL;
        if (status == MVS_OK && a.decompose) {
            double E[9];
            project_essential(F, E, LIT);
#pragma unroll
            for (int i = 0; i < 9; ++i) st->E[i] = E[i];
            if ((int)b.cnt < a.min_inliers) status = MVS_E_TOO_FEW_INLIERS;
            else {
                double Ra[9], Rb[9], t[3];
                decompose_essential(E, Ra, Rb, t, LIT);
#pragma unroll
                for (int i = 0; i < 9; ++i) { st->Rc[0][i] = Ra[i]; st->Rc[1][i] = Rb[i]; }
                st->tc[0] = t[0]; st->tc[1] = t[1]; st->tc[2] = t[2];
                double Rr[9];
                so3_rectify(Ra, Rr);
#pragma unroll
                for (int i = 0; i < 9; ++i) st->Rr[0][i] = Rr[i];
                so3_rectify(Rb, Rr);
#pragma unroll
                for (int i = 0; i < 9; ++i) st->Rr[1][i] = Rr[i];
            }
        }
        st->tri_count[0] = st->tri_count[1] = st->tri_count[2] = st->tri_count[3] = 0;
        st->status = status;
        s_status = status; s_nin = 0; s_k = 0;
    }
    __syncthreads();
    // inlier mask of the winner (estimator-RANSAC.cpp:112-127), same residual expression as K4
    double F[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) F[i] = s_F[i];
    const FzConst zc = make_fz(F, a.zc1, a.zc2);
    uint8_t *mask = a.mask + (size_t)pair * a.p_stride;
    const bool list = a.items != nullptr && a.decompose;
    int mine = 0;
    for (int i = threadIdx.x; i < n; i += SEL_THREADS) {
        const double *p = pts + (size_t)i * 6;
        double r;
        const bool in = point_residual<UNIT_Z, MODE, false, LIT>(p[0], p[1], p[2], p[3], p[4], p[5], F, zc, a.max_error_sq, r);
        mask[i] = in ? 1 : 0;
        mine += in ? 1 : 0;
    }
    if (!list) return;
    // ---- the triangulation's work list: reserve this pair's share, then append its inliers
    const bool ok = s_status == MVS_OK;
    if (ok) {
        mine = __reduce_add_sync(FULL, mine);
        if (lane == 0 && mine) atomicAdd(&s_nin, mine);
    }
    __syncthreads();
    if (threadIdx.x == 0 && ok) s_item_base = atomicAdd(a.item_total, (uint32_t)s_nin);
    __syncthreads();
    const uint32_t base = s_item_base;
    for (int i0 = 0; i0 < n; i0 += SEL_THREADS) {
        const int i = i0 + threadIdx.x;
        const bool in = ok && i < n && mask[i] != 0;
        const unsigned bal = __ballot_sync(FULL, in);
        int at = 0;
        if (lane == 0 && bal) at = atomicAdd(&s_k, __popc(bal));
        at = __shfl_sync(FULL, at, 0);
        if (in) a.items[base + at + __popc(bal & ((1u << lane) - 1u))] = ((unsigned long long)pair << 32) | (unsigned)i;
        else if (i < n)
#pragma unroll
            for (int c = 0; c < 4; ++c) a.valid[((size_t)pair * 4 + c) * a.p_stride + i] = 0;
    }
}

// SVD<M> (source/math/svd.hpp:13-73) for a batch of N x N matrices, one thread per matrix.
template <int N>
__global__ void __launch_bounds__(64)
svd_batch_kernel(const double *A, int count, int solver, double *U, double *w, double *Vt)
{
    const int i = blockIdx.x * 64 + threadIdx.x;
    if (i >= count) return;
    double Ul[N * N], wl[N], Vl[N * N];
    if (N == 3 && solver == MVS_SOLVER_FAST) svd3(A + (size_t)i * 9, Ul, wl, Vl);
    else cv_svd_full<N>(A + (size_t)i * N * N, Ul, wl, Vl);
    for (int k = 0; k < N * N; ++k) { U[(size_t)i * N * N + k] = Ul[k]; Vt[(size_t)i * N * N + k] = Vl[k]; }
    for (int k = 0; k < N; ++k) w[(size_t)i * N + k] = wl[k];
}

bool launch_svd_batch(int n, const double *A, int count, int solver, double *U, double *w, double *Vt, cudaStream_t s)
{
    const int blocks = (count + 63) / 64;
    if (n == 3) svd_batch_kernel<3><<<blocks, 64, 0, s>>>(A, count, solver, U, w, Vt);
    else if (n == 4 && solver == MVS_SOLVER_REFERENCE) svd_batch_kernel<4><<<blocks, 64, 0, s>>>(A, count, solver, U, w, Vt);
    else if (n == 9 && solver == MVS_SOLVER_REFERENCE) svd_batch_kernel<9><<<blocks, 64, 0, s>>>(A, count, solver, U, w, Vt);
    else return false;
    return true;
}

// ------------------------------------------------------------------------------------------ launchers
void launch_hypotheses(const HypArgs &a, int n_pairs, cudaStream_t s)
{
    HypArgs b = a;
    b.n_pairs = n_pairs;
    const long long total = (long long)n_pairs * a.H;
    if (a.solver == MVS_SOLVER_REFERENCE)   // 4 lanes per Jacobi: shortest chain for H = 1, and 2.6x the rate of a thread per hypothesis at H = 1024
        launch_dep(hypotheses_reference_wave_kernel, dim3((unsigned)((total + WAVE_HYP_PER_BLOCK - 1) / WAVE_HYP_PER_BLOCK)),
                   dim3(WAVE_HYP_PER_BLOCK * WAVE_LANES), 0, s, b);
    else
        launch_dep(hypotheses_kernel, dim3((unsigned)((total + HYP_THREADS - 1) / HYP_THREADS)), dim3(HYP_THREADS), 0, s, b);
}

void launch_fundamental_sets(const double *p1s, const double *p2s, int n_sets, double *F_out, int solver, cudaStream_t s)
{
    // scratch for the interleaved [n_sets][8][6] view lives right behind F_out's device buffer: the
    // caller passes F_out with room for n_sets*9 + n_sets*48 doubles
    fundamental_sets_kernel<<<(n_sets + HYP_THREADS - 1) / HYP_THREADS, HYP_THREADS, 0, s>>>(
        p1s, p2s, n_sets, F_out + (size_t)n_sets * 9, F_out, solver);
}

int score_tiles(int max_points) { return max_points > 0 ? (max_points + SC_TILE - 1) / SC_TILE : 1; }

template <bool UNIT_Z, int MODE>
static void launch_score_t(const ScoreArgs &a, bool lit, dim3 grid, cudaStream_t s)
{
    if (lit) launch_dep(score_kernel<UNIT_Z, MODE, true>, grid, dim3(SC_THREADS), 0, s, a);
    else launch_dep(score_kernel<UNIT_Z, MODE, false>, grid, dim3(SC_THREADS), 0, s, a);
}

void launch_score(const ScoreArgs &a, int mode, bool unit_z, int n_pairs, cudaStream_t s)
{
    dim3 grid((a.H + SC_THREADS - 1) / SC_THREADS, a.tiles, n_pairs);
    const bool lit = a.solver == MVS_SOLVER_REFERENCE;
    if (unit_z) {
        if (mode == MVS_SCORE_ALGEBRAIC) launch_score_t<true, MVS_SCORE_ALGEBRAIC>(a, lit, grid, s);
        else launch_score_t<true, MVS_SCORE_SAMPSON>(a, lit, grid, s);
    } else {
        if (mode == MVS_SCORE_ALGEBRAIC) launch_score_t<false, MVS_SCORE_ALGEBRAIC>(a, lit, grid, s);
        else launch_score_t<false, MVS_SCORE_SAMPSON>(a, lit, grid, s);
    }
}

template <bool UNIT_Z, int MODE>
static void launch_select_t(const SelectArgs &a, bool lit, bool small, int n_pairs, cudaStream_t s)
{
    if (small) {
        if (lit) launch_dep(select_kernel<UNIT_Z, MODE, true, SEL_THREADS_SMALL>, dim3(n_pairs), dim3(SEL_THREADS_SMALL), 0, s, a);
        else launch_dep(select_kernel<UNIT_Z, MODE, false, SEL_THREADS_SMALL>, dim3(n_pairs), dim3(SEL_THREADS_SMALL), 0, s, a);
    } else {
        if (lit) launch_dep(select_kernel<UNIT_Z, MODE, true, SEL_THREADS_BIG>, dim3(n_pairs), dim3(SEL_THREADS_BIG), 0, s, a);
        else launch_dep(select_kernel<UNIT_Z, MODE, false, SEL_THREADS_BIG>, dim3(n_pairs), dim3(SEL_THREADS_BIG), 0, s, a);
    }
}

void launch_select(const SelectArgs &a, int mode, bool unit_z, int max_points, int n_pairs, cudaStream_t s)
{
    const bool lit = a.solver == MVS_SOLVER_REFERENCE;
    // (a batch that fits one wave of the big blocks anyway keeps them: 256 threads walk H hypotheses 4x faster)
    const bool small = max_points <= SEL_SMALL_MAX_POINTS && a.H <= 2048 && n_pairs > 148 * 3;
    if (unit_z) {
        if (mode == MVS_SCORE_ALGEBRAIC) launch_select_t<true, MVS_SCORE_ALGEBRAIC>(a, lit, small, n_pairs, s);
        else launch_select_t<true, MVS_SCORE_SAMPSON>(a, lit, small, n_pairs, s);
    } else {
        if (mode == MVS_SCORE_ALGEBRAIC) launch_select_t<false, MVS_SCORE_ALGEBRAIC>(a, lit, small, n_pairs, s);
        else launch_select_t<false, MVS_SCORE_SAMPSON>(a, lit, small, n_pairs, s);
    }
}

}  // namespace mvs
