// match_hamming.cu — brute-force 256-bit Hamming kNN(2) + Lowe ratio / max-dist / cross-check
// filter + canonical sort + keypoint gather/normalisation, for batches of frame pairs.
//
// Replaces cv::BFMatcher(NORM_HAMMING).knnMatch(k=2) and the filtering around it in
// VisualFeature::match_visual_features (reference source/vision/visual-feature.cpp:51-80) and the
// keypoint gather + PinholeCamera::normalize_points of ImagePair::reconstruct / sfm_solve
// (source/front-end/image-pair.cpp:123-140, source/vision/sfm-solve.cpp:300-302, camera.cpp:55-79).
#include <math_constants.h>

#include <mutex>

#include "common.cuh"
#include "kernels.h"

namespace mvs {

// ------------------------------------------------------------------------------------------
// K1: knn2.  One thread owns one query descriptor (8 x u32 in registers) and scans a slice of the
// train set that the CTA stages through shared memory in tiles (128-bit loads, broadcast reads).
// Per (query, train) pair: 8 XOR + 6 LOP3 (carry-save adders) + 5 POPC + adds, then a 3-instruction top-2 update on packed
// (distance << 22 | trainIdx) keys — unsigned min/max gives the (distance, index) lexicographic
// order, i.e. OpenCV's strict-'<' ascending scan, independent of the order tiles are visited.
// grid = (query tiles, train splits, pairs); partial top-2 per split are merged in K2.
// ------------------------------------------------------------------------------------------
constexpr int KNN_THREADS = 256;
constexpr int KNN_TILE = 256;  // train descriptors per smem tile (8 KB)

// 256-bit Hamming distance.  POPC issues on the XU pipe at 16 lanes/clk/SM (measured 4.46e12/s on B200,
// profiles/ubench_peaks.json) — 4x scarcer than LOP3 — so three carry-save adders (sum = a^b^c,
// carry = maj(a,b,c), one LOP3 each) first compress 7 of the 8 XOR words into 2 weight-1 and 3 weight-2
// words: 5 POPC instead of 8, XU and ALU pipes roughly balanced (measured +44% over the 8-POPC form).
__device__ __forceinline__ uint32_t hamming256(const uint4 &qa, const uint4 &qb, const uint4 &ta, const uint4 &tb)
{
    const uint32_t x0 = qa.x ^ ta.x, x1 = qa.y ^ ta.y, x2 = qa.z ^ ta.z, x3 = qa.w ^ ta.w;
    const uint32_t x4 = qb.x ^ tb.x, x5 = qb.y ^ tb.y, x6 = qb.z ^ tb.z, x7 = qb.w ^ tb.w;
    const uint32_t s0 = x0 ^ x1 ^ x2, c0 = (x0 & x1) | (x2 & (x0 | x1));
    const uint32_t s1 = x3 ^ x4 ^ x5, c1 = (x3 & x4) | (x5 & (x3 | x4));
    const uint32_t s2 = s0 ^ s1 ^ x6, c2 = (s0 & s1) | (x6 & (s0 | s1));
#ifdef MVS_CSA4
    const uint32_t s3 = c0 ^ c1 ^ c2, c3 = (c0 & c1) | (c2 & (c0 | c1));
    return (__popc(s2) + __popc(x7)) + 2u * __popc(s3) + 4u * __popc(c3);
#else
    return (__popc(s2) + __popc(x7)) + 2u * (__popc(c0) + __popc(c1) + __popc(c2));
#endif
}

__device__ __forceinline__ void top2_insert(uint32_t &b1, uint32_t &b2, uint32_t key)
{
    const uint32_t hi = max(b1, key);
    b1 = min(b1, key);
    b2 = min(b2, hi);
}

template <bool BOUNDED>
__global__ void __launch_bounds__(KNN_THREADS)
knn2_hamming_kernel(KnnArgs a)
{
    __shared__ uint4 tile[KNN_TILE * 2];
    const int pair = blockIdx.z;
    int fq, ft;
    if (a.pairs) {  // query = pair frame (second), train = base frame (first): visual-feature.cpp:59-60
        const int2 pr = a.pairs[pair];
        fq = a.reverse ? pr.x : pr.y;
        ft = a.reverse ? pr.y : pr.x;
    } else { fq = a.reverse ? 0 : 1; ft = a.reverse ? 1 : 0; }
    const int nq = a.frame_cnt[fq], nt = a.frame_cnt[ft];
    const uint4 *Q = a.desc + 2 * (size_t)a.frame_off[fq];
    const uint4 *T = a.desc + 2 * (size_t)a.frame_off[ft];

    const int q = blockIdx.x * KNN_THREADS + threadIdx.x;
    if (blockIdx.x * KNN_THREADS >= nq) return;
    // contiguous train slice of this split, aligned to tiles
    const int tiles_total = (nt + KNN_TILE - 1) / KNN_TILE;
    const int tiles_per = (tiles_total + gridDim.y - 1) / gridDim.y;
    const int t_begin = blockIdx.y * tiles_per * KNN_TILE;
    const int t_end = min(nt, t_begin + tiles_per * KNN_TILE);

    uint4 qa = make_uint4(0, 0, 0, 0), qb = qa;
    if (q < nq) { qa = __ldg(Q + 2 * (size_t)q); qb = __ldg(Q + 2 * (size_t)q + 1); }
    uint32_t b1 = kKeyNone, b2 = kKeyNone;

    for (int t0 = t_begin; t0 < t_end; t0 += KNN_TILE) {
        const int cnt = min(KNN_TILE, t_end - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * cnt; i += KNN_THREADS) tile[i] = __ldg(T + 2 * (size_t)t0 + i);
        __syncthreads();
        const uint32_t kbase = (uint32_t)t0;
        int j = 0;
        if (BOUNDED) {
            // early abandon: the first 96 bits already put this train descriptor at distance >= bound for
            // every query of the warp -> it cannot change any filtered result (match_finalize_kernel)
#pragma unroll 4
            for (; j + 1 <= cnt; ++j) {
                const uint4 ta = tile[2 * j];
                const uint32_t x0 = qa.x ^ ta.x, x1 = qa.y ^ ta.y, x2 = qa.z ^ ta.z;
                const uint32_t s0 = x0 ^ x1 ^ x2, c0 = (x0 & x1) | (x2 & (x0 | x1));
                const uint32_t pc0 = __popc(c0);
                if (__all_sync(0xFFFFFFFFu, __popc(s0) + 2u * pc0 >= a.bound)) continue;
                const uint4 tb = tile[2 * j + 1];
                const uint32_t x3 = qa.w ^ ta.w, x4 = qb.x ^ tb.x, x5 = qb.y ^ tb.y, x6 = qb.z ^ tb.z, x7 = qb.w ^ tb.w;
                const uint32_t s1 = x3 ^ x4 ^ x5, c1 = (x3 & x4) | (x5 & (x3 | x4));
                const uint32_t s2 = s0 ^ s1 ^ x6, c2 = (s0 & s1) | (x6 & (s0 | s1));
                const uint32_t d = (__popc(s2) + __popc(x7)) + 2u * (pc0 + __popc(c1) + __popc(c2));
                top2_insert(b1, b2, (d << kIdxBits) + (kbase + (uint32_t)j));
            }
        } else {
#pragma unroll 4
            for (; j + 1 <= cnt; ++j) {
                const uint4 ta = tile[2 * j], tb = tile[2 * j + 1];
                const uint32_t d = hamming256(qa, qb, ta, tb);
                top2_insert(b1, b2, (d << kIdxBits) + (kbase + (uint32_t)j));
            }
        }
    }
    if (q < nq) a.partial[((size_t)pair * gridDim.y + blockIdx.y) * a.q_stride + q] = make_uint2(b1, b2);
}

// ------------------------------------------------------------------------------------------
// K2: finalize one pair per CTA: merge split partials, Lowe ratio + max_dist (+ cross-check),
// compact the survivors, sort them by (distance, queryIdx) with an in-smem bitonic network, emit
// DMatch records and (optionally) gather + K^-1-normalise the matched keypoints.
// ------------------------------------------------------------------------------------------
constexpr int FIN_THREADS = 1024;

__device__ __forceinline__ uint2 merge_partials(const uint2 *partial, int splits, size_t q_stride, int q)
{
    uint2 b = partial[q];
    for (int s = 1; s < splits; ++s) {
        const uint2 o = partial[(size_t)s * q_stride + q];
        top2_insert(b.x, b.y, o.x);
        top2_insert(b.x, b.y, o.y);
    }
    return b;
}

// Exact second neighbour from the tensor-core matcher's partials (match_hamming_tc.cu).  Its epilogue tracks, per 64-column
// block of train rows, the maximum of the four streams t = 0..3 (mod 4); `second` is the best key over every stream except
// the one the best came from.  The only train rows that can beat it are therefore the best's stream-mates: the rows of the
// same 64-row block with the same index mod 4 (15 of them).  Warp-cooperative: for every lane that asks (`want`), lanes
// 0..15 evaluate one stream slot each with popcounts and a shuffle-minimum brings the result back to the asking lane.
__device__ __forceinline__ uint32_t refine_second_warp(const uint4 *desc, bool want, int row_q, int row_t0, int nt, uint32_t best,
                                                       uint32_t second)
{
    const int lane = threadIdx.x & 31, sub = lane & 15, hi = lane >> 4;
    unsigned bal = __ballot_sync(0xFFFFFFFFu, want);
    while (bal) {                       // two asking lanes per round: lanes 0-15 serve the first, lanes 16-31 the second
        const int L0 = __ffs(bal) - 1;
        bal &= bal - 1;
        const int L1 = bal ? __ffs(bal) - 1 : -1;
        if (bal) bal &= bal - 1;
        const int L = hi ? L1 : L0;
        const uint32_t bL = __shfl_sync(0xFFFFFFFFu, best, L < 0 ? 0 : L);
        const int rq = __shfl_sync(0xFFFFFFFFu, row_q, L < 0 ? 0 : L);
        const int t1 = (int)(bL & kIdxMask);
        const int t = (t1 & ~63) + (t1 & 3) + 4 * sub;
        uint32_t key = kKeyNone;
        if (L >= 0 && t != t1 && t < nt) {
            const uint4 qa = desc[2 * (size_t)rq], qb = desc[2 * (size_t)rq + 1];
            const uint4 ta = desc[2 * (size_t)(row_t0 + t)], tb = desc[2 * (size_t)(row_t0 + t) + 1];
            key = (hamming256(qa, qb, ta, tb) << kIdxBits) | (uint32_t)t;
        }
#pragma unroll
        for (int off = 8; off > 0; off >>= 1) key = min(key, __shfl_xor_sync(0xFFFFFFFFu, key, off));   // stays inside a half
        const uint32_t k0 = __shfl_sync(0xFFFFFFFFu, key, 0), k1 = __shfl_sync(0xFFFFFFFFu, key, 16);
        if (lane == L0) second = min(second, k0);
        if (lane == L1) second = min(second, k1);
    }
    return second;
}

// FIN_T threads per CTA (pair): 1024 (two CTAs per SM) for long frames; 512 (four per SM) for VO-sized ones, where the kernel is a
// chain of barriers and dependent loads per pair and more resident pairs hide it better.  32 registers either way.
template <int FIN_T>
__global__ void __launch_bounds__(FIN_T, 2048 / FIN_T)
match_finalize_kernel(FinalizeArgs a)
{
    pdl_wait();
    pdl_launch_dependents();
    extern __shared__ uint32_t keys[];  // sort_cap entries
    __shared__ int s_count;
    const int pair = blockIdx.x;
    int fq, ft;
    if (a.pairs) { const int2 pr = a.pairs[pair]; fq = pr.y; ft = pr.x; } else { fq = 1; ft = 0; }
    const int nq = a.frame_cnt[fq], nt = a.frame_cnt[ft];
    const uint2 *part = a.partial + (size_t)pair * a.splits * a.q_stride;
    const uint2 *rpart = a.rev_partial ? a.rev_partial + (size_t)pair * a.rev_splits * a.rev_stride : nullptr;
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();

    for (int q0 = 0; q0 < nq; q0 += FIN_T) {    // warp-uniform trip count (refine_second_warp shuffles)
        const int q = q0 + threadIdx.x;
        const bool valid = q < nq;
        uint2 b = valid ? merge_partials(part, a.splits, a.q_stride, q) : make_uint2(kKeyNone, kKeyNone);
        if (a.refine_desc && a.knn_idx)    // raw output: every query needs its exact second neighbour
            b.y = refine_second_warp(a.refine_desc, valid && b.x != kKeyNone, a.frame_off[fq] + q, a.frame_off[ft], nt, b.x, b.y);
        if (valid && a.knn_idx) {  // raw knnMatch(k=2) output
            int32_t *ki = a.knn_idx + ((size_t)pair * a.q_stride + q) * 2;
            int32_t *kd = a.knn_dist + ((size_t)pair * a.q_stride + q) * 2;
            ki[0] = b.x == kKeyNone ? -1 : (int)(b.x & kIdxMask); kd[0] = b.x == kKeyNone ? -1 : (int)(b.x >> kIdxBits);
            ki[1] = b.y == kKeyNone ? -1 : (int)(b.y & kIdxMask); kd[1] = b.y == kKeyNone ? -1 : (int)(b.y >> kIdxBits);
        }
        const bool far2 = a.bound && (b.y == kKeyNone || (b.y >> kIdxBits) >= a.bound);   // bounded search: d2 >= bound
        // the reference compares float distances promoted to double (visual-feature.cpp:66-68);
        // a "far" second neighbour passes the ratio test for every d1 <= max_dist by construction of the bound
        const double d1 = (double)(float)(b.x >> kIdxBits);
        double d2 = (double)(float)(b.y >> kIdxBits);
        bool keep = valid && nt >= 2 && b.x != kKeyNone && (b.y != kKeyNone || far2) &&
                    (far2 || d1 < a.ratio * d2) && ((a.max_dist < 0) || (d1 <= a.max_dist));
        if (a.refine_desc && !a.knn_idx) {   // passed with the upper bound of d2: decide with the exact one
            b.y = refine_second_warp(a.refine_desc, keep, a.frame_off[fq] + q, a.frame_off[ft], nt, b.x, b.y);
            d2 = (double)(float)(b.y >> kIdxBits);
            keep = keep && d1 < a.ratio * d2;
        }
        if (keep && rpart) {  // cross-check: q must be the nearest query of its train descriptor
            const uint2 rb = merge_partials(rpart, a.rev_splits, a.rev_stride, (int)(b.x & kIdxMask));
            keep = ((int)(rb.x & kIdxMask) == q);
        }
        if (keep) {
            const int slot = atomicAdd(&s_count, 1);
            keys[slot] = ((b.x >> kIdxBits) << kIdxBits) | (uint32_t)q;
        }
    }
    __syncthreads();
    const int m = s_count;
    if (m <= FIN_T / 4) {
        // few survivors (the VO threshold leaves ~100): rank sort, four lanes per key, two barriers instead of the
        // log^2 of a bitonic network.  Keys are distinct (they end in the query index).
        const int i = threadIdx.x >> 2, part = threadIdx.x & 3;
        uint32_t mine = kKeyNone;
        int rank = 0;
        if (i < m) {
            mine = keys[i];
            for (int j = part; j < m; j += 4) rank += keys[j] < mine ? 1 : 0;
        }
        rank += __shfl_xor_sync(0xFFFFFFFFu, rank, 1);
        rank += __shfl_xor_sync(0xFFFFFFFFu, rank, 2);
        __syncthreads();
        if (i < m && part == 0) keys[rank] = mine;
        __syncthreads();
    } else {
    int n2 = 1;
    while (n2 < m) n2 <<= 1;
    for (int i = m + threadIdx.x; i < n2; i += FIN_T) keys[i] = kKeyNone;
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n2; i += FIN_T) {
                const int l = i ^ j;
                if (l > i) {
                    const uint32_t x = keys[i], y = keys[l];
                    const bool up = ((i & k) == 0);
                    if ((x > y) == up) { keys[i] = y; keys[l] = x; }
                }
            }
            __syncthreads();
        }
    }

    mvs_match *out = a.matches + (size_t)pair * a.q_stride;
    double *pts = a.points ? a.points + (size_t)pair * a.q_stride * 6 : nullptr;
    const float2 *kpq = a.kp ? a.kp + a.frame_off[fq] : nullptr;
    const float2 *kpt = a.kp ? a.kp + a.frame_off[ft] : nullptr;
    for (int i = threadIdx.x; i < m; i += FIN_T) {
        const uint32_t key = keys[i];
        const int q = (int)(key & kIdxMask);
        const uint2 b = merge_partials(part, a.splits, a.q_stride, q);
        const int t = (int)(b.x & kIdxMask);
        mvs_match mm;
        mm.query = q; mm.train = t; mm.distance = (float)(key >> kIdxBits);
        out[i] = mm;
        if (pts) {
            // base frame <- trainIdx, pair frame <- queryIdx (image-pair.cpp:129,138);
            // normalize_point: K^-1 (u, v, 1) (camera.cpp:55-64)
            const float2 k1 = kpt[t], k2 = kpq[q];
            const double u1 = (double)k1.x, v1 = (double)k1.y, u2 = (double)k2.x, v2 = (double)k2.y;
            double *p = pts + (size_t)i * 6;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                p[r] = (a.Kinv[3 * r] * u1 + a.Kinv[3 * r + 1] * v1) + a.Kinv[3 * r + 2] * 1.0;
                p[3 + r] = (a.Kinv[3 * r] * u2 + a.Kinv[3 * r + 1] * v2) + a.Kinv[3 * r + 2] * 1.0;
            }
        }
    }
    if (threadIdx.x == 0) {
        a.n_matches[pair] = m;
        if (a.state) {
            PairState *st = a.state + pair;
            st->n_matches = m;
            st->status = (m < 8) ? MVS_E_TOO_FEW_POINTS : MVS_OK;
            st->n_inliers = 0; st->best_h = -1; st->n_points = 0; st->candidate = -1; st->residual = 0.0;
        }
    }
}

// normalise caller-provided pixel correspondences (sfm_solve / sfm_triangulate entry, sfm-solve.cpp:300-302)
__global__ void normalize_points_kernel(const double *xy1, const double *xy2, int n, NormArgs a, double *pts, PairState *st)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && st) {
        st->n_matches = n;
        st->status = (n < 8) ? MVS_E_TOO_FEW_POINTS : MVS_OK;
        st->n_inliers = 0; st->best_h = -1; st->n_points = 0; st->candidate = -1; st->residual = 0.0;
    }
    if (i >= n) return;
    const double u1 = xy1[2 * i], v1 = xy1[2 * i + 1], u2 = xy2[2 * i], v2 = xy2[2 * i + 1];
    double *p = pts + (size_t)i * 6;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        p[r] = (a.Kinv[3 * r] * u1 + a.Kinv[3 * r + 1] * v1) + a.Kinv[3 * r + 2] * 1.0;
        p[3 + r] = (a.Kinv[3 * r] * u2 + a.Kinv[3 * r + 1] * v2) + a.Kinv[3 * r + 2] * 1.0;
    }
}

// ------------------------------------------------------------------------------------------ launchers
void launch_knn2_hamming(const KnnArgs &a, int max_nq, int splits, int n_pairs, cudaStream_t s)
{
    dim3 grid((max_nq + KNN_THREADS - 1) / KNN_THREADS, splits, n_pairs);
    if (a.bound) knn2_hamming_kernel<true><<<grid, KNN_THREADS, 0, s>>>(a);
    else knn2_hamming_kernel<false><<<grid, KNN_THREADS, 0, s>>>(a);
}

int finalize_sort_capacity(int max_nq)
{
    int n2 = 1;
    while (n2 < max_nq) n2 <<= 1;
    return n2;
}

cudaError_t launch_match_finalize(const FinalizeArgs &a, int max_nq, int n_pairs, cudaStream_t s)
{
    const size_t smem = (size_t)finalize_sort_capacity(max_nq) * sizeof(uint32_t);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    if (smem > 48 * 1024) {
        // the opt-in is a per-device function attribute: set it (to the kernel's maximum, once) on whichever device
        // the calling context is bound to; contexts of several threads / GPUs may get here concurrently
        static std::mutex mu;
        static bool configured[64] = {};
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        std::lock_guard<std::mutex> lock(mu);
        if (dev < 0 || dev >= 64 || !configured[dev]) {
            e = cudaFuncSetAttribute(match_finalize_kernel<FIN_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (e != cudaSuccess) return e;
            if (dev >= 0 && dev < 64) configured[dev] = true;
        }
    }
    // measured on 1764-keypoint VO frames: 1024 pairs 0.042 -> 0.035 ms with 512-thread CTAs; ten pairs 13.7 -> 16.6 us (a lone CTA
    // walks its queries in twice as many rounds), so the small CTA is for batches that fill the GPU more than once
    if (max_nq <= 2048 && n_pairs >= 296)
        return launch_dep(match_finalize_kernel<FIN_THREADS / 2>, dim3(n_pairs), dim3(FIN_THREADS / 2), smem, s, a);
    return launch_dep(match_finalize_kernel<FIN_THREADS>, dim3(n_pairs), dim3(FIN_THREADS), smem, s, a);
}

void launch_normalize_points(const double *xy1, const double *xy2, int n, const NormArgs &a, double *pts, PairState *st,
                             cudaStream_t s)
{
    const int blocks = n > 0 ? (n + 255) / 256 : 1;
    normalize_points_kernel<<<blocks, 256, 0, s>>>(xy1, xy2, n, a, pts, st);
}

}  // namespace mvs
