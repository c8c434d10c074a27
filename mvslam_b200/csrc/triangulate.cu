// triangulate.cu — batched linear (DLT) triangulation with cheirality, candidate selection and
// the final pose record.
//
//   K6 triangulate_kernel : triangulate_points (reference source/vision/sfm-solve.cpp:134-227) for
//                           the 4 (R,t) candidates of recover_pose_and_points (:232-284); one thread
//                           per (point, candidate), the 4x4 SVD solved in registers.
//   K7 finish_kernel      : strict-'>' candidate choice (:259-281), order-preserving compaction of
//                           the surviving points + original indexes, pose2in1 = SE3(SO3(R),t).inverse()
//                           (:364), ImagePair bookkeeping (source/front-end/image-pair.cpp:158-167).
#include <math_constants.h>

#include <algorithm>
#include <mutex>

#include "common.cuh"
#include "kernels.h"

namespace mvs {

constexpr int TRI_THREADS = 128;

// X = V.col(3) of SVD<4x4>(A) (source/vision/sfm-solve.cpp:193-195).  W is overwritten.
template <bool REF>
__device__ __forceinline__ void dlt_null_vector(double (&W)[4][4], double (&X)[4])
{
    if (REF) {
        // vt.row(3) of cv::SVDecomp(A) (svd.hpp:65-67), bit for bit
        double At[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) At[c][r] = W[r][c];
        cv_svd_last_vt<4>(At, X);
    } else {
        double V[4][4];
        jacobi_svd<4>(W, V);
        // the V column of the smallest singular value
        double best = CUDART_INF;
        int bj = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) s += W[k][j] * W[k][j];
            s = sqrt(s);
            if (s <= best) { best = s; bj = j; }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) X[k] = bj == 0 ? V[k][0] : (bj == 1 ? V[k][1] : (bj == 2 ? V[k][2] : V[k][3]));
    }
}


template <bool REF>
__global__ void __launch_bounds__(TRI_THREADS, REF ? 4 : 5)
triangulate_kernel(TriArgs a)
{
    __shared__ int s_cnt;
    const int pair = blockIdx.z, cand = blockIdx.y;
    PairState *st = a.state + pair;
    if (st->status != MVS_OK) return;
    const int n = st->n_matches;
    if ((int)(blockIdx.x * TRI_THREADS) >= n) return;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();

    // candidate order (Ra,+t),(Ra,-t),(Rb,+t),(Rb,-t)
    double R[9], t[3], Rr[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = st->Rc[cand >> 1][i];
#pragma unroll
    for (int i = 0; i < 3; ++i) t[i] = (cand & 1) ? -st->tc[i] : st->tc[i];
    so3_rectify(R, Rr);  // P2 = SE3(SO3(R), t).get_matrix() (:155)

    const int i = blockIdx.x * TRI_THREADS + threadIdx.x;
    bool ok = false;
    double pt[3] = {0.0, 0.0, 0.0};
    if (i < n && (!a.mask || a.mask[(size_t)pair * a.p_stride + i] != 0)) {
        const double *p = a.points + ((size_t)pair * a.p_stride + i) * 6;
        const double x1 = p[0], y1 = p[1], x2 = p[3], y2 = p[4];
        double W[4][4];
        // rows 0,1: x1[k]*P1.row(2) - P1.row(k) with P1 = I4 (:185-188)
        W[0][0] = x1 * 0.0 - 1.0; W[0][1] = x1 * 0.0 - 0.0; W[0][2] = x1 * 1.0 - 0.0; W[0][3] = x1 * 0.0 - 0.0;
        W[1][0] = y1 * 0.0 - 0.0; W[1][1] = y1 * 0.0 - 1.0; W[1][2] = y1 * 1.0 - 0.0; W[1][3] = y1 * 0.0 - 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double p2j = j < 3 ? Rr[6 + j] : t[2];
            const double p0j = j < 3 ? Rr[0 + j] : t[0];
            const double p1j = j < 3 ? Rr[3 + j] : t[1];
            W[2][j] = x2 * p2j - p0j;
            W[3][j] = y2 * p2j - p1j;
        }
        double X[4];
        dlt_null_vector<REF>(W, X);
        if (!(fabs(X[3]) < kTolerance)) {
            const double scale = 1.0 / X[3];
            pt[0] = X[0] * scale; pt[1] = X[1] * scale; pt[2] = X[2] * scale;
            if (!(pt[2] < kTolerance)) {
                const double z2 = (R[6] * pt[0] + R[7] * pt[1] + R[8] * pt[2]) + t[2];  // un-rectified R (:218)
                ok = !(z2 < kTolerance);
            }
        }
    }
    if (i < n) {
        const size_t o = ((size_t)pair * 4 + cand) * a.p_stride + i;
        a.valid[o] = ok ? 1 : 0;
        double *tp = a.tri + o * 3;
        tp[0] = pt[0]; tp[1] = pt[1]; tp[2] = pt[2];
    }
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, ok);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(&s_cnt, __popc(bal));
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt) atomicAdd(&st->tri_count[cand], s_cnt);
}

// 256 threads for long match lists, 128 for ordinary ones (twice the pairs resident: the kernel's time is thread 0's
// candidate choice and record, a latency every block pays).
constexpr int FIN2_THREADS = 256, FIN2_THREADS_SMALL = 128, FIN2_SMALL_MAX_POINTS = 4096;

// The per-pair tail shared by finish_kernel and triangulate_finish_kernel: one block of THREADS threads per pair.
template <int THREADS>
__device__ __forceinline__ void finish_body(const FinishArgs &a, const int pair)
{
    __shared__ int s_cand, s_base, s_warp[THREADS / 32];
    __shared__ unsigned long long s_ssd;
    PairState *st = a.state + pair;
    mvs_pair_result *res = a.results ? a.results + pair : nullptr;
    if (threadIdx.x == 0) {
        int cand = -1;
        if (st->status == MVS_OK) {
            int best = 0;
            for (int c = 0; c < a.n_cand; ++c)
                if (st->tri_count[c] > best) { best = st->tri_count[c]; cand = c; }
            if (cand < 0) st->status = MVS_E_NO_CHEIRALITY;
            st->candidate = cand;
            st->n_points = best;
        }
        s_cand = cand; s_base = 0; s_ssd = 0ull;
    }
    __syncthreads();
    const int cand = s_cand;
    const int n = st->n_matches;
    if (cand >= 0) {
        const uint8_t *valid = a.valid + ((size_t)pair * 4 + cand) * a.p_stride;
        const double *tri = a.tri + ((size_t)pair * 4 + cand) * a.p_stride * 3;
        double *op = a.out_points + (size_t)pair * a.p_stride * 3;
        uint64_t *oi = a.out_index + (size_t)pair * a.p_stride;
        const mvs_match *mt = a.matches ? a.matches + (size_t)pair * a.p_stride : nullptr;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        unsigned long long ssd = 0ull;
        for (int i0 = 0; i0 < n; i0 += THREADS) {
            const int i = i0 + threadIdx.x;
            const bool f = (i < n) && valid[i];
            const unsigned bal = __ballot_sync(0xFFFFFFFFu, f);
            if (lane == 0) s_warp[warp] = __popc(bal);
            __syncthreads();
            int off = s_base;
            for (int w = 0; w < warp; ++w) off += s_warp[w];
            if (f) {
                const int r = off + __popc(bal & ((1u << lane) - 1u));
                op[3 * (size_t)r] = tri[3 * (size_t)i]; op[3 * (size_t)r + 1] = tri[3 * (size_t)i + 1];
                op[3 * (size_t)r + 2] = tri[3 * (size_t)i + 2];
                oi[r] = (uint64_t)i;
                if (mt) { const unsigned long long d = (unsigned long long)mt[i].distance; ssd += d * d; }
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                int tot = 0;
                for (int w = 0; w < THREADS / 32; ++w) tot += s_warp[w];
                s_base += tot;
            }
            __syncthreads();
        }
        if (ssd) atomicAdd(&s_ssd, ssd);
        __syncthreads();
    }
    if (threadIdx.x == 0 && res) {
        res->status = st->status;
        res->n_matches = st->n_matches;
        res->n_inliers = st->n_inliers;
        res->best_hypothesis = st->best_h;
        res->n_points = st->n_points;
        res->candidate = st->candidate;
        res->residual = st->residual;
        const bool have_model = (st->status == MVS_OK || st->status == MVS_E_TOO_FEW_INLIERS ||
                                 st->status == MVS_E_NO_CHEIRALITY || st->status == MVS_E_NO_MODEL);
        for (int i = 0; i < 9; ++i) {
            res->F[i] = have_model ? st->F[i] : 0.0;
            res->E[i] = (have_model && st->status != MVS_E_NO_MODEL) ? st->E[i] : 0.0;
            res->R1to2[i] = 0.0; res->R2in1[i] = 0.0;
        }
        for (int i = 0; i < 3; ++i) { res->t1to2[i] = 0.0; res->t2in1[i] = 0.0; }
        res->match_inlier_ssd = 0;
        if (cand >= 0) {
            double R[9], t[3], Rr[9], Ri[9], ti[3];
            for (int i = 0; i < 9; ++i) R[i] = st->Rc[cand >> 1][i];
            for (int i = 0; i < 3; ++i) t[i] = (cand & 1) ? -st->tc[i] : st->tc[i];
            so3_rectify(R, Rr);
            se3_inverse(Rr, t, Ri, ti);
            for (int i = 0; i < 9; ++i) { res->R1to2[i] = R[i]; res->R2in1[i] = Ri[i]; }
            for (int i = 0; i < 3; ++i) { res->t1to2[i] = t[i]; res->t2in1[i] = ti[i]; }
            res->match_inlier_ssd = s_ssd;
        }
    }
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
finish_kernel(FinishArgs a)
{
    pdl_wait();
    pdl_launch_dependents();
    finish_body<THREADS>(a, blockIdx.x);
}

// K6 over K5's work list (SelectArgs::items): the (inlier, candidate) pairs of the whole batch, flattened, walked by a
// grid that just fills the GPU.  Every lane of every warp that enters the 4x4 SVD has a point, whatever the pairs'
// inlier counts are, and no block waits for its slowest decomposition; the rectified candidates come from K5
// (PairState::Rr).  REF on a big job keeps V in shared memory (cv_svd_last_vt_sv): 64 registers, two 512-thread blocks per SM.
// Results go to the same per-(pair, candidate, match) scratch as triangulate_kernel, so K7 is unchanged.
// A job that one wave of small blocks covers is a latency problem, not a throughput one: 64-thread blocks spread its
// FP64 chains over as many SMs as there are warps (V in registers: the shorter chain).
constexpr int TI_THREADS_BIG = 512, TI_THREADS_SMALL = 64;

template <bool REF, bool SMEM_V, int TI_THREADS>
__global__ void __launch_bounds__(TI_THREADS, SMEM_V ? 2 : 1)
triangulate_items_kernel(TriArgs a)
{
    pdl_wait();
    pdl_launch_dependents();
    extern __shared__ double s_V[];   // SMEM_V: [16][TI_THREADS]
    const size_t total = (size_t)*a.item_total * 4;
    for (size_t it = (size_t)blockIdx.x * TI_THREADS + threadIdx.x; it < total; it += (size_t)gridDim.x * TI_THREADS) {
        const unsigned long long e = a.items[it >> 2];
        const int cand = (int)(it & 3), pair = (int)(e >> 32), i = (int)(e & 0xFFFFFFFFu);   // (Ra,+t),(Ra,-t),(Rb,+t),(Rb,-t)
        PairState *st = a.state + pair;
        const double *R = st->Rc[cand >> 1], *Rr = st->Rr[cand >> 1];   // P2 = SE3(SO3(R), t).get_matrix() (:155)
        double t[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) t[k] = (cand & 1) ? -st->tc[k] : st->tc[k];
        const double *p = a.points + ((size_t)pair * a.p_stride + i) * 6;
        const double x1 = p[0], y1 = p[1], x2 = p[3], y2 = p[4];
        double W[4][4];
        // rows 0,1: x1[k]*P1.row(2) - P1.row(k) with P1 = I4 (:185-188)
        W[0][0] = x1 * 0.0 - 1.0; W[0][1] = x1 * 0.0 - 0.0; W[0][2] = x1 * 1.0 - 0.0; W[0][3] = x1 * 0.0 - 0.0;
        W[1][0] = y1 * 0.0 - 0.0; W[1][1] = y1 * 0.0 - 1.0; W[1][2] = y1 * 1.0 - 0.0; W[1][3] = y1 * 0.0 - 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double p2j = j < 3 ? Rr[6 + j] : t[2];
            const double p0j = j < 3 ? Rr[0 + j] : t[0];
            const double p1j = j < 3 ? Rr[3 + j] : t[1];
            W[2][j] = x2 * p2j - p0j;
            W[3][j] = y2 * p2j - p1j;
        }
        double X[4];
        if (REF) {
            // X = vt.row(3) of cv::SVDecomp(A) (svd.hpp:65-67), bit for bit
            double At[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) At[c][r] = W[r][c];
            if (SMEM_V) cv_svd_last_vt_sv<4>(At, s_V + threadIdx.x, TI_THREADS, X);
            else cv_svd_last_vt<4>(At, X);
        } else dlt_null_vector<false>(W, X);
        bool ok = false;
        double pt[3] = {0.0, 0.0, 0.0};
        if (!(fabs(X[3]) < kTolerance)) {
            const double scale = 1.0 / X[3];
            pt[0] = X[0] * scale; pt[1] = X[1] * scale; pt[2] = X[2] * scale;
            if (!(pt[2] < kTolerance)) {
                const double z2 = (R[6] * pt[0] + R[7] * pt[1] + R[8] * pt[2]) + t[2];  // un-rectified R (:218)
                ok = !(z2 < kTolerance);
            }
        }
        const size_t o = ((size_t)pair * 4 + cand) * a.p_stride + i;
        a.valid[o] = ok ? 1 : 0;
        double *tp = a.tri + o * 3;
        tp[0] = pt[0]; tp[1] = pt[1]; tp[2] = pt[2];
        if (ok) atomicAdd(&st->tri_count[cand], 1);
    }
}

void launch_triangulate(const TriArgs &a, int max_points, int n_pairs, cudaStream_t s)
{
    dim3 grid(max_points > 0 ? (max_points + TRI_THREADS - 1) / TRI_THREADS : 1, a.n_cand, n_pairs);
    if (a.solver == MVS_SOLVER_REFERENCE) triangulate_kernel<true><<<grid, TRI_THREADS, 0, s>>>(a);
    else triangulate_kernel<false><<<grid, TRI_THREADS, 0, s>>>(a);
}

cudaError_t launch_triangulate_items(const TriArgs &a, size_t max_items, cudaStream_t s)
{
    const bool ref = a.solver == MVS_SOLVER_REFERENCE;
    const size_t lanes = std::max<size_t>(1, max_items * 4);
    if (lanes <= (size_t)148 * TI_THREADS_BIG) {
        const unsigned grid = (unsigned)((lanes + TI_THREADS_SMALL - 1) / TI_THREADS_SMALL);
        if (ref) return launch_dep(triangulate_items_kernel<true, false, TI_THREADS_SMALL>, dim3(grid), dim3(TI_THREADS_SMALL), 0, s, a);
        return launch_dep(triangulate_items_kernel<false, false, TI_THREADS_SMALL>, dim3(grid), dim3(TI_THREADS_SMALL), 0, s, a);
    }
    if (!ref) {
        return launch_dep(triangulate_items_kernel<false, false, TI_THREADS_BIG>, dim3(148), dim3(TI_THREADS_BIG), 0, s, a);
    }
    constexpr int smem = 16 * TI_THREADS_BIG * (int)sizeof(double);
    // per-device opt-in above 48 KB, as in launch_match_finalize
    static std::mutex mu;
    static bool configured[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (dev < 0 || dev >= 64 || !configured[dev]) {
            e = cudaFuncSetAttribute(triangulate_items_kernel<true, true, TI_THREADS_BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return e;
            if (dev >= 0 && dev < 64) configured[dev] = true;
        }
    }
    return launch_dep(triangulate_items_kernel<true, true, TI_THREADS_BIG>, dim3(148 * 2), dim3(TI_THREADS_BIG), smem, s, a);
}

void launch_finish(const FinishArgs &a, int max_points, int n_pairs, cudaStream_t s)
{
    if (max_points <= FIN2_SMALL_MAX_POINTS) launch_dep(finish_kernel<FIN2_THREADS_SMALL>, dim3(n_pairs), dim3(FIN2_THREADS_SMALL), 0, s, a);
    else launch_dep(finish_kernel<FIN2_THREADS>, dim3(n_pairs), dim3(FIN2_THREADS), 0, s, a);
}

}  // namespace mvs
