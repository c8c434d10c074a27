// ransac.cu — seeded RANSAC over the normalised 8-point fundamental-matrix estimator.
//
//   K3 hypotheses_kernel : find_fundamental_matrix (reference source/vision/fundamental-matrix.cpp:18-140,
//                          204-267) for every row of the sample table, 3 hypotheses per warp.
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
// K3.  The 9x9 problem SVD(A^T A) is solved by a one-sided Jacobi laid out as a systolic "chess
// tournament": a hypothesis is owned by 5 lanes (seats); every seat holds the two columns of W (= A^T A)
// and of V that meet in the current step, so the three inner products, the rotation parameters and the
// rotation itself are lane-local (one c,s computation per column pair).  Between steps the columns
// move one seat along the ring  top1<-top2<-top3<-top4<-bot4<-bot3<-bot2<-bot1<-bot0<-top1  with two
// warp shuffles per element.  Seat 0 pairs the bye (9 columns, 10 players) with player s.  In step s
// seat k>=1 holds players ((s+k) mod 9, (s-k) mod 9): exactly the oracle's round-robin order.
// 6 hypotheses (30 lanes) per warp; lanes 30,31 shadow lanes 0,1.
// ------------------------------------------------------------------------------------------
constexpr int HYP_WARPS = 4;
constexpr int HYP_PER_WARP = 6;
constexpr unsigned FULL = 0xFFFFFFFFu;

// find_normalization_transform (fundamental-matrix.cpp:18-54) of the 8 sampled points of one image;
// T = [[s,0,tx],[0,s,ty],[0,0,1]] returned as (s, tx, ty)
__device__ __forceinline__ void normalize8(const double *pts, const uint32_t (&idx)[8], int off,
                                           double (&nx)[8], double (&ny)[8], double (&T)[3])
{
    double px[8], py[8], pz[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const double *p = pts + (size_t)idx[i] * 6 + off;
        px[i] = p[0]; py[i] = p[1]; pz[i] = p[2];
    }
    double mx = 0.0, my = 0.0, mz = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { mx += px[i]; my += py[i]; mz += pz[i]; }
    mx *= 0.125; my *= 0.125; mz *= 0.125;
    double scale = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const double dx = px[i] - mx, dy = py[i] - my, dz = pz[i] - mz;
        scale += sqrt(dx * dx + dy * dy + dz * dz);
        nx[i] = dx; ny[i] = dy;
    }
    scale *= 0.125;
    scale = 1.4142135623730951 / scale;  // sqrt(2.0) correctly rounded
#pragma unroll
    for (int i = 0; i < 8; ++i) { nx[i] *= scale; ny[i] *= scale; }
    T[0] = scale; T[1] = -mx * scale; T[2] = -my * scale;
}

__device__ __forceinline__ double sel9(const double (&a)[9], int c)
{
    double r = a[0];
#pragma unroll
    for (int k = 1; k < 9; ++k) r = (c == k) ? a[k] : r;
    return r;
}

// Cooperative 8-point solve by the 5 lanes [gbase, gbase+5) of a warp (all 32 lanes must call).
// pts: [.][6] correspondences, idx: the 8 sampled rows. Every lane of the group returns the full F.
__device__ __forceinline__ void eight_point_group(const double *pts, const uint32_t (&idx)[8], int k, int gbase,
                                                  double (&F)[9])
{
    double T1[3], T2[3];
    double wt[9], vt[9], wb[9], vb[9];
    int pt = k, pb = (k == 0) ? 0 : 9 - k;  // players (columns) held at step 0; seat 0's top is the bye
    {
        double x1[8], y1[8], x2[8], y2[8];
        normalize8(pts, idx, 0, x1, y1, T1);
        normalize8(pts, idx, 3, x2, y2, T2);
        // columns pt, pb of A^T A, accumulated over the 8 rows in order (fundamental-matrix.cpp:76-111)
#pragma unroll
        for (int i = 0; i < 9; ++i) { wt[i] = 0.0; wb[i] = 0.0; }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const double a[9] = {x2[r] * x1[r], x2[r] * y1[r], x2[r], y2[r] * x1[r], y2[r] * y1[r], y2[r], x1[r], y1[r], 1.0};
            const double at = sel9(a, pt), ab = sel9(a, pb);
#pragma unroll
            for (int i = 0; i < 9; ++i) { wt[i] += a[i] * at; wb[i] += a[i] * ab; }
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) { vt[i] = (i == pt) ? 1.0 : 0.0; vb[i] = (i == pb) ? 1.0 : 0.0; }
    }
    const int src_next = gbase + min(k + 1, 4), src_prev = gbase + max(k - 1, 0);
    int step = 0;
    for (int sweep = 0; sweep < kSvdMaxSweeps; ++sweep) {
        bool changed = false;
        for (int s9 = 0; s9 < 9; ++s9) {
            double at = 0.0, ab = 0.0, g = 0.0;
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                at = fma(wt[i], wt[i], at);
                ab = fma(wb[i], wb[i], ab);
                g = fma(wt[i], wb[i], g);   // products commute exactly
            }
            const bool top_first = pt < pb;  // the column with the smaller index is "p" of the pair (p<q)
            double cs, sn;
            const bool rot = (k != 0) && jacobi_cs(top_first ? at : ab, top_first ? ab : at, g, cs, sn);
            if (rot) {
                // p' = c*p + s*q ; q' = c*q - s*p
                const double st = top_first ? sn : -sn;   // top' = c*top + st*bot ; bot' = c*bot - st*top
#pragma unroll
                for (int i = 0; i < 9; ++i) {
                    const double a0 = wt[i], b0 = wb[i];
                    wt[i] = fma(cs, a0, st * b0);
                    wb[i] = fma(cs, b0, -(st * a0));
                    const double a1 = vt[i], b1 = vb[i];
                    vt[i] = fma(cs, a1, st * b1);
                    vb[i] = fma(cs, b1, -(st * a1));
                }
            }
            changed |= rot;
            // ring move to the seating of the next step
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                const double tn = __shfl_sync(FULL, wt[i], src_next), bp = __shfl_sync(FULL, wb[i], src_prev);
                wt[i] = (k == 4) ? wb[i] : tn;
                wb[i] = (k == 0) ? tn : bp;
                const double un = __shfl_sync(FULL, vt[i], src_next), up = __shfl_sync(FULL, vb[i], src_prev);
                vt[i] = (k == 4) ? vb[i] : un;
                vb[i] = (k == 0) ? un : up;
            }
            ++step;
            const int sm = step % 9;
            pt = (sm + k) % 9;
            pb = (k == 0) ? sm : (sm - k + 9) % 9;
        }
        if (__ballot_sync(FULL, changed) == 0u) break;
    }
    // f = V column of the smallest singular value (vt.row(8), fundamental-matrix.cpp:114-118); among
    // equal values the column with the larger index, as a stable descending sort would leave it last
    double sg_t = 0.0, sg_b = 0.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) { sg_t = fma(wt[i], wt[i], sg_t); sg_b = fma(wb[i], wb[i], sg_b); }
    sg_t = (k == 0) ? CUDART_INF : sqrt(sg_t);
    sg_b = sqrt(sg_b);
    double best = CUDART_INF;
    int bj = -1, blane = 0, btop = 0;
#pragma unroll
    for (int l = 0; l < 5; ++l) {
        const double vt_ = __shfl_sync(FULL, sg_t, gbase + l), vb_ = __shfl_sync(FULL, sg_b, gbase + l);
        const int jt = __shfl_sync(FULL, pt, gbase + l), jb = __shfl_sync(FULL, pb, gbase + l);
        if (l != 0 && (vt_ < best || (vt_ == best && jt > bj))) { best = vt_; bj = jt; blane = l; btop = 1; }
        if (vb_ < best || (vb_ == best && jb > bj)) { best = vb_; bj = jb; blane = l; btop = 0; }
    }
    double Fp[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) Fp[i] = __shfl_sync(FULL, btop ? vt[i] : vb[i], gbase + blane);
    // singular constraint (fundamental-matrix.cpp:128-136), then F = T2^T * F * T1 (:245)
    double U[9], w[3], Vt[9], Fh[9], T2t[9], tmp[9];
    svd3(Fp, U, w, Vt);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            Fh[i * 3 + j] = (U[i * 3 + 0] * w[0]) * Vt[0 * 3 + j] + (U[i * 3 + 1] * w[1]) * Vt[1 * 3 + j];
    const double T1m[9] = {T1[0], 0.0, T1[1], 0.0, T1[0], T1[2], 0.0, 0.0, 1.0};
    const double T2m[9] = {T2[0], 0.0, T2[1], 0.0, T2[0], T2[2], 0.0, 0.0, 1.0};
    mat3_transpose(T2m, T2t);
    mat3_mul(T2t, Fh, tmp);
    mat3_mul(tmp, T1m, F);
}

__global__ void __launch_bounds__(HYP_WARPS * 32)
hypotheses_kernel(HypArgs a)
{
    const int pair = blockIdx.y;
    int n = a.n_fixed;
    if (a.state) {
        if (a.state[pair].status != MVS_OK) return;
        n = a.state[pair].n_matches;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane < 30 ? lane / 5 : 0;
    const int c = lane < 30 ? lane % 5 : lane - 30;
    const int gbase = g * 5;
    const int h0 = (blockIdx.x * HYP_WARPS + warp) * HYP_PER_WARP;
    if (h0 >= a.H) return;  // warp-uniform
    const int h = min(h0 + g, a.H - 1);
    uint32_t idx[8];
    if (a.table) {
#pragma unroll
        for (int j = 0; j < 8; ++j) idx[j] = a.table[(size_t)h * 8 + j];
    } else {
        sample_row(a.seed, a.pair_id_base + (uint64_t)pair, (uint32_t)n, h, idx);
    }
    double F[9];
    eight_point_group(a.points + (size_t)pair * a.p_stride * 6, idx, c, gbase, F);
    if (lane < 30 && c == 0 && h0 + g < a.H) {
        double *o = a.F_all + ((size_t)pair * a.H + h) * 9;
#pragma unroll
        for (int i = 0; i < 9; ++i) o[i] = F[i];
    }
}

// a9 entry for explicit 8-point sets: p1s/p2s [n_sets][8][3] -> interleave into the [.][6] layout
__global__ void __launch_bounds__(HYP_WARPS * 32)
fundamental_sets_kernel(const double *p1s, const double *p2s, int n_sets, double *pts6, double *F_out)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane < 30 ? lane / 5 : 0;
    const int c = lane < 30 ? lane % 5 : lane - 30;
    const int gbase = g * 5;
    const int h0 = (blockIdx.x * HYP_WARPS + warp) * HYP_PER_WARP;
    if (h0 >= n_sets) return;
    const int h = min(h0 + g, n_sets - 1);
    (void)pts6;
    // gather this set into a per-lane view through a tiny index table over a virtual [8][6] block
    // stored in global scratch (written by the same lanes, then re-read; volume is negligible)
    double *blk = pts6 + (size_t)h * 48;
    if (lane < 30) {
        for (int r = c; r < 8; r += 5)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                blk[r * 6 + k] = p1s[((size_t)h * 8 + r) * 3 + k];
                blk[r * 6 + 3 + k] = p2s[((size_t)h * 8 + r) * 3 + k];
            }
    }
    __syncwarp();
    __threadfence_block();
    const uint32_t idx[8] = {0, 1, 2, 3, 4, 5, 6, 7};
    double F[9];
    eight_point_group(blk, idx, c, gbase, F);
    if (lane < 30 && c == 0 && h0 + g < n_sets) {
#pragma unroll
        for (int i = 0; i < 9; ++i) F_out[(size_t)h * 9 + i] = F[i];
    }
}

// ------------------------------------------------------------------------------------------
// K4.  thread = hypothesis (F in registers), correspondences streamed through shared memory as
// broadcast 128-bit reads; grid = (H/128, point tiles, pairs).  Each thread keeps a private
// (count, residual) for its tile, summed over the tile's points in index order.
// ------------------------------------------------------------------------------------------
constexpr int SC_THREADS = 128;
constexpr int SC_TILE = 512;

template <bool UNIT_Z, int MODE>
__global__ void __launch_bounds__(SC_THREADS)
score_kernel(ScoreArgs a)
{
    constexpr int W = UNIT_Z ? 4 : 6;
    __shared__ __align__(16) double sp[SC_TILE * W];
    const int pair = blockIdx.z, tile = blockIdx.y;
    int n = a.n_fixed;
    if (a.state) {
        if (a.state[pair].status != MVS_OK) return;
        n = a.state[pair].n_matches;
    }
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
    const double thr = a.max_error_sq;
    uint32_t c = 0;
    double res = 0.0;
#pragma unroll 4
    for (int i = 0; i < cnt; ++i) {
        double r;
        bool in;
        if (UNIT_Z) {
            const double2 u = *reinterpret_cast<const double2 *>(sp + 4 * i);
            const double2 v = *reinterpret_cast<const double2 *>(sp + 4 * i + 2);
            in = point_residual<true, MODE>(u.x, u.y, 1.0, v.x, v.y, 1.0, F, thr, r);
        } else {
            const double2 u = *reinterpret_cast<const double2 *>(sp + 6 * i);
            const double2 v = *reinterpret_cast<const double2 *>(sp + 6 * i + 2);
            const double2 w = *reinterpret_cast<const double2 *>(sp + 6 * i + 4);
            in = point_residual<false, MODE>(u.x, u.y, v.x, v.y, w.x, w.y, F, thr, r);
        }
        if (in) { ++c; res += r; }
    }
    const size_t o = ((size_t)pair * a.tiles + tile) * a.H + h;
    a.part_count[o] = c;
    a.part_res[o] = res;
}

// ------------------------------------------------------------------------------------------
// K5.  One CTA per pair.
// ------------------------------------------------------------------------------------------
constexpr int SEL_THREADS = 256;

struct Best { uint32_t cnt; double res; int h; };

__device__ __forceinline__ bool better(const Best &x, const Best &y)
{   // x beats y under the sequential rule of estimator-RANSAC.cpp:76-84 (earlier hypothesis wins ties)
    if (x.cnt != y.cnt) return x.cnt > y.cnt;
    if (x.res != y.res) return x.res < y.res;
    return x.h < y.h;
}

// find_essential_matrix own branch (sfm-solve.cpp:73-87)
static __device__ __noinline__ void project_essential(const double F[9], double E[9])
{
    double U[9], w[3], Vt[9];
    svd3(F, U, w, Vt);
    const double v = sqrt(w[0] * w[1]);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            E[i * 3 + j] = (U[i * 3 + 0] * v) * Vt[0 * 3 + j] + (U[i * 3 + 1] * v) * Vt[1 * 3 + j];
}

// decompose_essential_matrix (sfm-solve.cpp:97-127)
static __device__ __noinline__ void decompose_essential(const double E[9], double Ra[9], double Rb[9], double t[3])
{
    double U[9], w[3], Vt[9], V[9];
    svd3(E, U, w, Vt);
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

template <bool UNIT_Z, int MODE>
__global__ void __launch_bounds__(SEL_THREADS)
select_kernel(SelectArgs a)
{
    __shared__ Best s_best[SEL_THREADS / 32];
    __shared__ double s_F[9];
    __shared__ int s_go;
    const int pair = blockIdx.x;
    PairState *st = a.state + pair;
    if (st->status != MVS_OK) return;
    const int n = st->n_matches;
    const int tiles_used = (n + SC_TILE - 1) / SC_TILE;

    Best b; b.cnt = 0; b.res = kInfinity; b.h = 0x7FFFFFFF;
    for (int h = threadIdx.x; h < a.H; h += SEL_THREADS) {
        Best x; x.cnt = 0; x.res = 0.0; x.h = h;
        for (int t = 0; t < tiles_used; ++t) {
            const size_t o = ((size_t)pair * a.tiles + t) * a.H + h;
            x.cnt += a.part_count[o];
            x.res += a.part_res[o];
        }
        if (a.all_counts) a.all_counts[(size_t)pair * a.H + h] = (int32_t)x.cnt;
        if (better(x, b)) b = x;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        Best o;
        o.cnt = __shfl_xor_sync(FULL, b.cnt, off);
        o.res = __shfl_xor_sync(FULL, b.res, off);
        o.h = __shfl_xor_sync(FULL, b.h, off);
        if (better(o, b)) b = o;
    }
    if ((threadIdx.x & 31) == 0) s_best[threadIdx.x >> 5] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < SEL_THREADS / 32; ++w)
            if (better(s_best[w], b)) b = s_best[w];
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
            project_essential(F, E);
#pragma unroll
            for (int i = 0; i < 9; ++i) st->E[i] = E[i];
            if ((int)b.cnt < a.min_inliers) status = MVS_E_TOO_FEW_INLIERS;
            else {
                double Ra[9], Rb[9], t[3];
                decompose_essential(E, Ra, Rb, t);
#pragma unroll
                for (int i = 0; i < 9; ++i) { st->Rc[0][i] = Ra[i]; st->Rc[1][i] = Rb[i]; }
                st->tc[0] = t[0]; st->tc[1] = t[1]; st->tc[2] = t[2];
            }
        }
        st->tri_count[0] = st->tri_count[1] = st->tri_count[2] = st->tri_count[3] = 0;
        st->status = status;
        s_go = 1;
    }
    __syncthreads();
    (void)s_go;
    // inlier mask of the winner (estimator-RANSAC.cpp:112-127), same residual expression as K4
    double F[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) F[i] = s_F[i];
    const double *pts = a.points + (size_t)pair * a.p_stride * 6;
    uint8_t *mask = a.mask + (size_t)pair * a.p_stride;
    for (int i = threadIdx.x; i < n; i += SEL_THREADS) {
        const double *p = pts + (size_t)i * 6;
        double r;
        mask[i] = point_residual<UNIT_Z, MODE>(p[0], p[1], p[2], p[3], p[4], p[5], F, a.max_error_sq, r) ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------------ launchers
void launch_hypotheses(const HypArgs &a, int n_pairs, cudaStream_t s)
{
    const int per_block = HYP_WARPS * HYP_PER_WARP;
    dim3 grid((a.H + per_block - 1) / per_block, n_pairs);
    hypotheses_kernel<<<grid, HYP_WARPS * 32, 0, s>>>(a);
}

void launch_fundamental_sets(const double *p1s, const double *p2s, int n_sets, double *F_out, cudaStream_t s)
{
    // scratch for the interleaved [n_sets][8][6] view lives right behind F_out's device buffer: the
    // caller passes F_out with room for n_sets*9 + n_sets*48 doubles
    const int per_block = HYP_WARPS * HYP_PER_WARP;
    fundamental_sets_kernel<<<(n_sets + per_block - 1) / per_block, HYP_WARPS * 32, 0, s>>>(
        p1s, p2s, n_sets, F_out + (size_t)n_sets * 9, F_out);
}

int score_tiles(int max_points) { return max_points > 0 ? (max_points + SC_TILE - 1) / SC_TILE : 1; }

void launch_score(const ScoreArgs &a, int mode, bool unit_z, int n_pairs, cudaStream_t s)
{
    dim3 grid((a.H + SC_THREADS - 1) / SC_THREADS, a.tiles, n_pairs);
    if (unit_z) {
        if (mode == MVS_SCORE_ALGEBRAIC) score_kernel<true, MVS_SCORE_ALGEBRAIC><<<grid, SC_THREADS, 0, s>>>(a);
        else score_kernel<true, MVS_SCORE_SAMPSON><<<grid, SC_THREADS, 0, s>>>(a);
    } else {
        if (mode == MVS_SCORE_ALGEBRAIC) score_kernel<false, MVS_SCORE_ALGEBRAIC><<<grid, SC_THREADS, 0, s>>>(a);
        else score_kernel<false, MVS_SCORE_SAMPSON><<<grid, SC_THREADS, 0, s>>>(a);
    }
}

void launch_select(const SelectArgs &a, int mode, bool unit_z, int n_pairs, cudaStream_t s)
{
    if (unit_z) {
        if (mode == MVS_SCORE_ALGEBRAIC) select_kernel<true, MVS_SCORE_ALGEBRAIC><<<n_pairs, SEL_THREADS, 0, s>>>(a);
        else select_kernel<true, MVS_SCORE_SAMPSON><<<n_pairs, SEL_THREADS, 0, s>>>(a);
    } else {
        if (mode == MVS_SCORE_ALGEBRAIC) select_kernel<false, MVS_SCORE_ALGEBRAIC><<<n_pairs, SEL_THREADS, 0, s>>>(a);
        else select_kernel<false, MVS_SCORE_SAMPSON><<<n_pairs, SEL_THREADS, 0, s>>>(a);
    }
}

}  // namespace mvs
