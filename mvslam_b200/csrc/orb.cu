// orb.cu — feature extraction on the device: the cv::ORB pipeline that VisualFeature::extract runs
// (reference source/vision/visual-feature.cpp:9-17,40-49; cv::ORB::create(MAX_FEATURE_COUNT), every other
// parameter at its OpenCV default).  OpenCV itself is not vendored by the reference; the stages below follow the
// published algorithm (8-level 1.2x pyramid by bit-exact fixed-point bilinear resize, FAST-9/16 with 3x3 non-max
// suppression, Harris re-scoring, intensity-centroid orientation, 7x7 sigma-2 Gaussian, steered BRIEF) and are
// pinned bit-for-bit against cv2 by the test suite (tests/test_orb_*.py).
//
// Everything is integer or explicitly-rounded FP32 arithmetic: this file is compiled with -fmad=false and uses
// fmaf() exactly where OpenCV's FMA-contracted separable filter does.
//
//   O1 orb_pyramid_kernel   one thread-block cluster per image: level 0 copy + 7 chained fixed-point resizes
//   O2 orb_fast_kernel      FAST score + 3x3 NMS per 32x16 tile -> candidate list + score histogram
//   O3 orb_harris_kernel    first retainBest (2 x quota by FAST score, histogram threshold) + Harris response
//   O4 orb_select_kernel    second retainBest (quota by Harris response, radix select), raster-order sort
//   O5 orb_blur_kernel      7x7 Gaussian, reflect-101
//   O6 orb_describe_kernel  warp per keypoint: IC angle, fastAtan2, 256 steered tests -> 32 bytes
#include <cmath>
#include <vector>

#include "orb.h"
#include "orb_pattern.h"

namespace mvs {

// ------------------------------------------------------------------------------------------ host geometry
static inline int cv_round_f(float v) { return (int)lrintf(v); }

bool orb_make_geometry(int w, int h, int nfeatures, OrbGeom &g, std::vector<int32_t> &tabs)
{
    const double scale_factor = (double)1.2f;   // ORB::create(float scaleFactor = 1.2f) stored as double
    tabs.clear();
    int off = 0, cand = 0, ftiles = 0, btiles = 0;
    // nfeaturesPerLevel (orb.cpp computeKeyPoints)
    const float factor = (float)(1.0 / scale_factor);
    float nd = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)kOrbLevels));
    int sum = 0;
    for (int l = 0; l < kOrbLevels; ++l) {
        OrbLevel &L = g.lv[l];
        L.scale = (float)std::pow(scale_factor, (double)l);
        const float inv = 1.0f / L.scale;
        L.w = cv_round_f(w * inv);
        L.h = cv_round_f(h * inv);
        if (L.w < 1 || L.h < 1) return false;
        L.pitch = (L.w + 15) & ~15;
        L.off = off;
        off += (L.pitch * L.h + 127) & ~127;      // levels start on their own 128-byte line
        if (l < kOrbLevels - 1) {
            L.quota = cv_round_f(nd);
            sum += L.quota;
            nd *= factor;
        } else {
            L.quota = std::max(nfeatures - sum, 0);
        }
        const int rw = L.w - 2 * kOrbEdge, rh = L.h - 2 * kOrbEdge;   // KeyPointsFilter::runByImageBorder region
        L.cand_off = cand;
        L.cand_cap = (rw > 0 && rh > 0) ? ((rw + 1) / 2) * ((rh + 1) / 2) : 0;   // strict 3x3 maxima: one per 2x2 block
        cand += L.cand_cap;
        L.tile_off = ftiles;
        if (rw > 0 && rh > 0) ftiles += ((rw + 31) / 32) * ((rh + 15) / 16);
        L.blur_tile_off = btiles;
        btiles += ((L.w + 31) / 32) * ((L.h + 31) / 32);
        // resize tables (imgproc resize.cpp, INTER_LINEAR_EXACT 8-bit: interpolationLinear<ufixedpoint16>), one packed
        // int per destination index: source offset | 8.8 weight of the next source pixel << 16; x table padded to 4
        L.tab_off = (int)tabs.size();
        if (l > 0) {
            const OrbLevel &P = g.lv[l - 1];
            for (int axis = 0; axis < 2; ++axis) {
                const int src = axis ? P.h : P.w, dst = axis ? L.h : L.w;
                const double inv_scale = (double)dst / (double)src, sc = 1.0 / inv_scale;
                for (int d = 0; d < dst; ++d) {
                    double f = (d + 0.5) * sc - 0.5;
                    int s0 = (int)std::floor(f), al = 0;
                    f -= s0;
                    if (s0 < 0) s0 = 0;
                    else if (s0 >= src - 1) s0 = src - 1;
                    else al = (int)std::nearbyint(f * 256.0);
                    tabs.push_back(s0 | (al << 16));
                }
                while (tabs.size() & 3) tabs.push_back(0);
            }
        }
    }
    g.slab = (off + 255) & ~255;
    g.cand_total = std::max(cand, 1);
    g.fast_tiles = ftiles;
    g.blur_tiles = btiles;
    const float hs = 1.f / ((1 << 2) * 7 * 255.f);
    g.harris_scale4 = hs * hs * hs * hs;
    // umax: end of each row of the circular patch (orb.cpp computeKeyPoints)
    const int hp = kOrbHalfPatch;
    int umax[hp + 2];
    const int vmax = (int)std::floor(hp * std::sqrt(2.f) / 2 + 1);
    const int vmin = (int)std::ceil(hp * std::sqrt(2.f) / 2);
    for (int v = 0; v <= vmax; ++v) umax[v] = (int)std::nearbyint(std::sqrt((double)hp * hp - v * v));
    for (int v = hp, v0 = 0; v >= vmin; --v) {
        while (umax[v0] == umax[v0 + 1]) ++v0;
        umax[v] = v0;
        ++v0;
    }
    for (int v = 0; v <= hp; ++v) g.umax[v] = umax[v];
    return true;
}

// ------------------------------------------------------------------------------------------ O1 pyramid
// One thread-block cluster (8 CTAs x 512 threads) per image builds the whole pyramid: level 0 is copied from the
// staging buffer, level l is resized from level l-1 (cv::resize INTER_LINEAR_EXACT, 8-bit: horizontal pass in 8.8 fixed
// point, vertical pass rounded at 16 fractional bits), with a cluster barrier (release/acquire) between levels instead
// of a kernel boundary (the acquire side makes the other SMs' global writes visible to ordinary loads; every level starts
// on its own 128-byte line).  A thread produces 4 adjacent pixels (one packed x-table int4 load, one 32-bit store).
constexpr int kPyrCluster = 8, kPyrThreads = 512;

__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(kPyrCluster, 1, 1) __launch_bounds__(kPyrThreads)
orb_pyramid_kernel(const __grid_constant__ OrbGeom g, OrbBuffers b, const uint8_t *stage, int stride)
{
    const int img = blockIdx.x / kPyrCluster;
    const int tid = (blockIdx.x % kPyrCluster) * kPyrThreads + threadIdx.x, nthr = kPyrCluster * kPyrThreads;
    uint8_t *slab = b.pyr + (size_t)img * g.slab;
    const uint8_t *img0 = stage + (size_t)img * g.lv[0].h * stride;
    {   // level 0
        const OrbLevel &L = g.lv[0];
        const int wq = (L.w + 3) >> 2;
        const bool words = (stride & 3) == 0 && ((uintptr_t)stage & 3) == 0;
        for (int i = tid; i < wq * L.h; i += nthr) {
            const int y = i / wq, x = (i - y * wq) * 4;
            const uint8_t *sp = img0 + (size_t)y * stride + x;
            uint32_t v = 0;
            if (words && x + 4 <= L.w) v = *reinterpret_cast<const uint32_t *>(sp);
            else for (int k = 0; k < 4 && x + k < L.w; ++k) v |= (uint32_t)sp[k] << (8 * k);
            *reinterpret_cast<uint32_t *>(slab + L.off + (size_t)y * L.pitch + x) = v;
        }
    }
    for (int l = 1; l < kOrbLevels; ++l) {
        const OrbLevel &D = g.lv[l];
        const OrbLevel &S = g.lv[l - 1];
        const uint8_t *src = l == 1 ? img0 : slab + S.off;      // level 1 reads the input itself: no wait for the copy
        const int sp = l == 1 ? stride : S.pitch;
        const int wq = (D.w + 3) >> 2;
        const int4 *tx = reinterpret_cast<const int4 *>(b.tabs + D.tab_off);
        const int32_t *ty = b.tabs + D.tab_off + wq * 4;
        uint8_t *dst = slab + D.off;
        for (int i = tid; i < wq * D.h; i += nthr) {
            const int y = i / wq, xq = i - y * wq;
            const int4 t4 = tx[xq];
            const int tyv = ty[y];
            const int y0 = tyv & 0xffff, ay = tyv >> 16, y1 = min(y0 + 1, S.h - 1);
            const uint8_t *r0 = src + (size_t)y0 * sp, *r1 = src + (size_t)y1 * sp;
            const int te[4] = {t4.x, t4.y, t4.z, t4.w};
            uint32_t out = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int x0 = te[k] & 0xffff, ax = te[k] >> 16, x1 = min(x0 + 1, S.w - 1);
                const uint32_t h0 = r0[x0] * (256 - ax) + r0[x1] * ax;     // 8.8 fixed point
                const uint32_t h1 = r1[x0] * (256 - ax) + r1[x1] * ax;
                const uint32_t v = h0 * (256 - ay) + h1 * ay;                                // 16.16
                out |= ((v + (1u << 15)) >> 16) << (8 * k);
            }
            *reinterpret_cast<uint32_t *>(dst + (size_t)y * D.pitch + 4 * xq) = out;   // row padding absorbs x >= w
        }
        if (l + 1 < kOrbLevels) cluster_sync_all();
    }
}

void launch_orb_pyramid(const OrbGeom &g, const OrbBuffers &b, const uint8_t *stage, int stride, int n_images, cudaStream_t s)
{
    orb_pyramid_kernel<<<n_images * kPyrCluster, kPyrThreads, 0, s>>>(g, b, stage, stride);
}

// ------------------------------------------------------------------------------------------ O2 FAST + NMS
// cornerScore<16> of fast_score.cpp in closed form: the largest threshold at which the pixel is still a FAST-9
// corner = max over the 16 arcs of 9 contiguous ring pixels of min(v - ring) (darker arc) or min(ring - v)
// (brighter arc), minus 1; 0 when that does not exceed the detector threshold.
__device__ __forceinline__ int fast_corner_score(const uint8_t *c, int p)
{
    const int v = c[0];
    int d[16];
    d[0] = v - c[3 * p];       d[1] = v - c[3 * p + 1];   d[2] = v - c[2 * p + 2];   d[3] = v - c[p + 3];
    d[4] = v - c[3];           d[5] = v - c[-p + 3];      d[6] = v - c[-2 * p + 2];  d[7] = v - c[-3 * p + 1];
    d[8] = v - c[-3 * p];      d[9] = v - c[-3 * p - 1];  d[10] = v - c[-2 * p - 2]; d[11] = v - c[-p - 3];
    d[12] = v - c[-3];         d[13] = v - c[p - 3];      d[14] = v - c[2 * p - 2];  d[15] = v - c[3 * p - 1];
    const int t = kOrbFastThreshold;
    // an arc of 9 contains one pixel of every antipodal pair
    const bool pb = (max(d[0], d[8]) > t) & (max(d[4], d[12]) > t) & (max(d[2], d[10]) > t) & (max(d[6], d[14]) > t);
    const bool pd = (min(d[0], d[8]) < -t) & (min(d[4], d[12]) < -t) & (min(d[2], d[10]) < -t) & (min(d[6], d[14]) < -t);
    if (!(pb | pd)) return 0;
    int lo2[16], hi2[16], lo4[16], hi4[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { lo2[i] = min(d[i], d[(i + 1) & 15]); hi2[i] = max(d[i], d[(i + 1) & 15]); }
#pragma unroll
    for (int i = 0; i < 16; ++i) { lo4[i] = min(lo2[i], lo2[(i + 2) & 15]); hi4[i] = max(hi2[i], hi2[(i + 2) & 15]); }
    int a = -255, bmin = 255;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int lo9 = min(min(lo4[i], lo4[(i + 4) & 15]), d[(i + 8) & 15]);
        const int hi9 = max(max(hi4[i], hi4[(i + 4) & 15]), d[(i + 8) & 15]);
        a = max(a, lo9);
        bmin = min(bmin, hi9);
    }
    const int m = max(a, -bmin);
    return m > t ? m - 1 : 0;
}

constexpr int FT_W = 32, FT_H = 16;
constexpr int FT_PW = FT_W + 8, FT_PH = FT_H + 8;      // pixel tile (halo 4: 1 for the NMS ring + 3 for the FAST ring)
constexpr int FT_SW = FT_W + 2, FT_SH = FT_H + 2;      // score tile

constexpr int FT_PITCH = 44;       // bytes per pixel-tile row: 11 aligned words starting at x0 - 7

// 4 adjacent bytes starting at byte `off` of a word-aligned shared-memory row
__device__ __forceinline__ uint32_t ld4_unaligned(const uint32_t *row, int off)
{
    const uint32_t *w = row + (off >> 2);
    return __funnelshift_r(w[0], w[1], (off & 3) * 8);
}
// per byte: bit 7 set iff the byte exceeds the FAST threshold (a > 20 <=> a >= 128 or (a & 127) + 107 >= 128)
__device__ __forceinline__ uint32_t over_threshold4(uint32_t a)
{
    static_assert(kOrbFastThreshold == 20, "0x6b = 127 - threshold");
    return (((a & 0x7f7f7f7fu) + 0x6b6b6b6bu) | a) & 0x80808080u;
}
// Necessary condition for 4 horizontally adjacent centres at once (packed bytes, VABSDIFF4): an arc of 9 contains one
// pixel of every antipodal pair, so every pair must hold a pixel that differs from the centre by more than the threshold.
// Sign-agnostic and on 4 of the 8 pairs only: a superset of the corners (11.5 % of the Tsukuba pixels pass, 4.8 % are corners).
__device__ __forceinline__ uint32_t fast_maybe4(const uint32_t *pxw, int r, int cb)   // cb = byte column of the first centre
{
    const uint32_t *row = pxw + r * (FT_PITCH / 4);
    const uint32_t c = ld4_unaligned(row, cb);
    uint32_t f = over_threshold4(__vabsdiffu4(c, ld4_unaligned(row + 3 * (FT_PITCH / 4), cb))) |
                 over_threshold4(__vabsdiffu4(c, ld4_unaligned(row - 3 * (FT_PITCH / 4), cb)));
    f &= over_threshold4(__vabsdiffu4(c, ld4_unaligned(row, cb + 3))) | over_threshold4(__vabsdiffu4(c, ld4_unaligned(row, cb - 3)));
    if (!f) return 0;
    f &= over_threshold4(__vabsdiffu4(c, ld4_unaligned(row + 2 * (FT_PITCH / 4), cb + 2))) |
         over_threshold4(__vabsdiffu4(c, ld4_unaligned(row - 2 * (FT_PITCH / 4), cb - 2)));
    f &= over_threshold4(__vabsdiffu4(c, ld4_unaligned(row - 2 * (FT_PITCH / 4), cb + 2))) |
         over_threshold4(__vabsdiffu4(c, ld4_unaligned(row + 2 * (FT_PITCH / 4), cb - 2)));
    return f;
}

__global__ void __launch_bounds__(256, 8)
orb_fast_kernel(const __grid_constant__ OrbGeom g, OrbBuffers b)
{
    __shared__ uint32_t pxw[FT_PH + 1][FT_PITCH / 4];       // +1 row: the unaligned 4-byte window of the last group may touch it
    __shared__ __align__(4) uint8_t sc[FT_SH][FT_SW + 2];
    __shared__ uint16_t todo[FT_SH * FT_SW];       // phase 1 survivors, then the 3x3 maxima: r << 6 | c
    __shared__ uint16_t corners[FT_H * FT_W];      // positive scores inside the output tile
    __shared__ int n_todo, n_corner, n_keep, keep_base;
    int l = 0;
    while (l + 1 < kOrbLevels && (int)blockIdx.x >= g.lv[l + 1].tile_off) ++l;
    const OrbLevel &L = g.lv[l];
    const int img = blockIdx.y, tid = threadIdx.x;
    const int tiles_x = (L.w - 2 * kOrbEdge + 31) / 32;
    const int t = blockIdx.x - L.tile_off;
    const int x0 = kOrbEdge + (t % tiles_x) * FT_W, y0 = kOrbEdge + (t / tiles_x) * FT_H;
    const uint8_t *src = b.pyr + (size_t)img * g.slab + L.off;
    const uint8_t *px = reinterpret_cast<const uint8_t *>(pxw);
    if (tid == 0) { n_todo = 0; n_corner = 0; n_keep = 0; }
    if (tid < FT_SH * (FT_SW + 2) / 4) reinterpret_cast<uint32_t *>(sc)[tid] = 0u;
    // pixel tile: rows [y0-4, y0+20), 11 aligned words per row from x0 - 7 (x0 = 31 + 32k, so x0 - 7 is a multiple of 4);
    // byte column of image x is x - (x0 - 7).  Rows beyond the image repeat the last row (no tested pixel reads them).
    for (int i = tid; i < (FT_PH + 1) * 11; i += 256) {
        const int r = i / 11, wd = i - r * 11;
        const int y = min(y0 - 4 + r, L.h - 1), xw = x0 - 7 + 4 * wd;
        pxw[r][wd] = xw < L.pitch ? *reinterpret_cast<const uint32_t *>(src + (size_t)y * L.pitch + xw) : 0u;
    }
    __syncthreads();
    // phase 1: cheap rejection over the 18 x 34 score positions (origin (x0-1, y0-1)), 4 adjacent positions per thread in
    // packed bytes; survivors are queued so that phase 2 runs the full score on dense warps
    if (tid < FT_SH * 9) {
        const int r = tid / 9, q = tid - r * 9;
        const uint32_t f = fast_maybe4(&pxw[0][0], r + 3, 4 * q + 6);
        if (f) {
            const int y = y0 - 1 + r;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c = 4 * q + k, x = x0 - 1 + c;
                if (((f >> (8 * k + 7)) & 1u) && c < FT_SW && x < L.w - 3 && y < L.h - 3)       // fast.cpp tests [3, n-3)
                    todo[atomicAdd(&n_todo, 1)] = (uint16_t)(r << 6 | c);
            }
        }
    }
    __syncthreads();
    const int n = n_todo;
    for (int j = tid; j < n; j += 256) {
        const int v = todo[j], r = v >> 6, c = v & 63;
        const int s = fast_corner_score(px + (r + 3) * FT_PITCH + c + 6, FT_PITCH);
        sc[r][c] = (uint8_t)s;
        // candidates for the 3x3 maximum test: corners inside the output tile and inside the detector's border
        if (s > 0 && r >= 1 && r <= FT_H && c >= 1 && c <= FT_W && x0 - 1 + c < L.w - kOrbEdge && y0 - 1 + r < L.h - kOrbEdge)
            corners[atomicAdd(&n_corner, 1)] = (uint16_t)v;
    }
    __syncthreads();
    const int nc = n_corner;
    for (int j = tid; j < nc; j += 256) {
        const int v = corners[j], r = v >> 6, c = v & 63;
        const int s = sc[r][c];
        if (s > sc[r - 1][c - 1] && s > sc[r - 1][c] && s > sc[r - 1][c + 1] && s > sc[r][c - 1] && s > sc[r][c + 1] &&
            s > sc[r + 1][c - 1] && s > sc[r + 1][c] && s > sc[r + 1][c + 1])
            todo[atomicAdd(&n_keep, 1)] = (uint16_t)v;      // todo is free again
    }
    __syncthreads();
    const int nk = n_keep;
    if (nk == 0) return;
    if (tid == 0) keep_base = atomicAdd(&b.cand_cnt[img * kOrbLevels + l], nk);
    __syncthreads();
    for (int j = tid; j < nk; j += 256) {
        const int v = todo[j], r = v >> 6, c = v & 63;
        const int s = sc[r][c];
        const size_t o = (size_t)img * g.cand_total + L.cand_off + keep_base + j;
        b.cand_xy[o] = (uint32_t)(x0 - 1 + c) | ((uint32_t)(y0 - 1 + r) << 16);
        b.cand_val[o] = (float)s;
        atomicAdd(&b.hist[(img * kOrbLevels + l) * 256 + s], 1);
    }
}

void launch_orb_fast(const OrbGeom &g, const OrbBuffers &b, int n_images, cudaStream_t s)
{
    if (!g.fast_tiles) return;
    orb_fast_kernel<<<dim3(g.fast_tiles, n_images), 256, 0, s>>>(g, b);
}

// 256 threads, one histogram bin each: the largest bin d whose suffix count sum_{j >= d} h[j] reaches `need`
// (the bin that holds the need-th largest element); out[0] = d, out[1] = how many of bin d are still needed.
// Requires sum(h) >= need >= 1.  Ends with a barrier.
__device__ __forceinline__ void suffix_pick_256(int h, int need, int *warp_tot /*[8]*/, int *out /*[2]*/)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int v = h;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_down_sync(0xffffffffu, v, o);
        if (lane + o < 32) v += u;
    }
    if (lane == 0) warp_tot[w] = v;
    __syncthreads();
    for (int k = w + 1; k < 8; ++k) v += warp_tot[k];
    if (v >= need && v - h < need) { out[0] = threadIdx.x; out[1] = need - (v - h); }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------ O3 first cut + Harris
// KeyPointsFilter::retainBest(keypoints, 2 * featuresNum) on the FAST score keeps everything >= the n-th largest
// score (keypoint.cpp): with integer scores that threshold comes from the 256-bin histogram.
__global__ void __launch_bounds__(256)
orb_harris_kernel(const __grid_constant__ OrbGeom g, OrbBuffers b)
{
    __shared__ int s_wt[8], s_out[2];
    const int l = blockIdx.y, img = blockIdx.z;
    const OrbLevel &L = g.lv[l];
    const int n = b.cand_cnt[img * kOrbLevels + l];
    if ((int)(blockIdx.x * 32) >= n) return;
    const int want = 2 * L.quota;
    int thr = 0;
    if (n > want) {
        thr = 256;                           // want == 0: drop everything
        if (want > 0) {
            suffix_pick_256(b.hist[(img * kOrbLevels + l) * 256 + threadIdx.x], want, s_wt, s_out);
            thr = s_out[0];
        }
    }
    const uint8_t *im = b.pyr + (size_t)img * g.slab + L.off;
    const int p = L.pitch;
    // 8 lanes per candidate: lane r < 7 sums row r of the 7x7 block, then a 3-step shuffle reduction
    const int sub = threadIdx.x & 7;
    for (int i0 = blockIdx.x * 32; i0 < n; i0 += gridDim.x * 32) {     // uniform trip count: full-warp shuffles below
        const int i = i0 + (threadIdx.x >> 3);
        const size_t o = (size_t)img * g.cand_total + L.cand_off + min(i, n - 1);
        const bool live = i < n && b.cand_val[o] >= (float)thr;
        int a = 0, bb = 0, c = 0;
        if (live && sub < 7) {
            const uint32_t xy = b.cand_xy[o];
            const uint8_t *q = im + (size_t)((int)(xy >> 16) - 3 + sub) * p + (int)(xy & 0xffff) - 3;
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                const int ix = (q[k + 1] - q[k - 1]) * 2 + (q[k - p + 1] - q[k - p - 1]) + (q[k + p + 1] - q[k + p - 1]);
                const int iy = (q[k + p] - q[k - p]) * 2 + (q[k + p - 1] - q[k - p - 1]) + (q[k + p + 1] - q[k - p + 1]);
                a += ix * ix; bb += iy * iy; c += ix * iy;
            }
        }
#pragma unroll
        for (int s = 4; s > 0; s >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, s);
            bb += __shfl_xor_sync(0xffffffffu, bb, s);
            c += __shfl_xor_sync(0xffffffffu, c, s);
        }
        if (sub == 0 && i < n) {
            float r = -INFINITY;
            if (live) {
                // ((float)a * b - (float)c * c - harris_k * ((float)a + b) * ((float)a + b)) * scale^4, one rounding per operation
                const float fa = (float)a, fb = (float)bb, fc = (float)c;
                const float tr = __fadd_rn(fa, fb);
                r = __fmul_rn(__fsub_rn(__fsub_rn(__fmul_rn(fa, fb), __fmul_rn(fc, fc)), __fmul_rn(__fmul_rn(0.04f, tr), tr)),
                              g.harris_scale4);
            }
            b.cand_val[o] = r;
        }
    }
}

void launch_orb_harris(const OrbGeom &g, const OrbBuffers &b, int n_images, cudaStream_t s)
{
    int most = 1;
    for (int l = 0; l < kOrbLevels; ++l) most = std::max(most, g.lv[l].cand_cap);
    const int gx = std::min((most + 31) / 32, n_images >= 32 ? 16 : 128);
    orb_harris_kernel<<<dim3(gx, kOrbLevels, n_images), 256, 0, s>>>(g, b);
}

// ------------------------------------------------------------------------------------------ O4 second cut + order
__device__ __forceinline__ uint32_t float_order(float f)
{
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(256)
orb_select_kernel(const __grid_constant__ OrbGeom g, OrbBuffers b)
{
    __shared__ unsigned long long keys[kOrbSortCap];
    __shared__ int s_hist[256];
    __shared__ int s_a, s_wt[8], s_out[2];
    const int l = blockIdx.x, img = blockIdx.y;
    const OrbLevel &L = g.lv[l];
    const int n = b.cand_cnt[img * kOrbLevels + l];
    const float *val = b.cand_val + (size_t)img * g.cand_total + L.cand_off;
    const uint32_t *xy = b.cand_xy + (size_t)img * g.cand_total + L.cand_off;
    int32_t *out_cnt = b.kept_cnt + img * kOrbLevels + l;
    if (n == 0 || L.quota == 0) { if (threadIdx.x == 0) *out_cnt = 0; return; }
    // survivors of the first cut
    if (threadIdx.x == 0) s_a = 0;
    __syncthreads();
    int mine = 0;
    for (int i = threadIdx.x; i < n; i += 256) mine += val[i] != -INFINITY;
    if (mine) atomicAdd(&s_a, mine);
    __syncthreads();
    const int survivors = s_a;
    __syncthreads();
    uint32_t thr = 0;       // ordered-uint threshold: keep float_order(val) >= thr
    if (survivors > L.quota) {
        // retainBest(featuresNum): the quota-th largest response, by 4 x 8-bit radix select
        uint32_t prefix = 0, mask = 0;
        int need = L.quota;
        for (int pass = 3; pass >= 0; --pass) {
            s_hist[threadIdx.x] = 0;
            __syncthreads();
            for (int i = threadIdx.x; i < n; i += 256) {
                const float v = val[i];
                const uint32_t u = float_order(v);
                if (v != -INFINITY && (u & mask) == prefix) atomicAdd(&s_hist[(u >> (8 * pass)) & 255], 1);
            }
            __syncthreads();
            suffix_pick_256(s_hist[threadIdx.x], need, s_wt, s_out);
            prefix |= (uint32_t)s_out[0] << (8 * pass);
            mask |= 0xffu << (8 * pass);
            need = s_out[1];
            __syncthreads();
        }
        thr = prefix;
    }
    if (threadIdx.x == 0) s_a = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += 256) {
        const float v = val[i];
        if (v != -INFINITY && float_order(v) >= thr) {
            const int slot = atomicAdd(&s_a, 1);
            if (slot < kOrbSortCap) keys[slot] = ((unsigned long long)xy[i] << 32) | (uint32_t)i;   // (y, x) raster key
        }
    }
    __syncthreads();
    const int kept = s_a;
    if (kept > kOrbSortCap) { if (threadIdx.x == 0) *out_cnt = -1; return; }
    int cap = 1;
    while (cap < kept) cap <<= 1;
    for (int i = kept + threadIdx.x; i < cap; i += 256) keys[i] = ~0ull;
    __syncthreads();
    for (int k = 2; k <= cap; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < cap; i += 256) {
                const int p = i ^ j;
                if (p > i) {
                    const unsigned long long x = keys[i], y = keys[p];
                    if ((x > y) == ((i & k) == 0)) { keys[i] = y; keys[p] = x; }
                }
            }
            __syncthreads();
        }
    uint32_t *out = b.kept_idx + (size_t)(img * kOrbLevels + l) * kOrbSortCap;
    for (int i = threadIdx.x; i < kept; i += 256) out[i] = (uint32_t)keys[i];
    if (threadIdx.x == 0) *out_cnt = kept;
}

void launch_orb_select(const OrbGeom &g, const OrbBuffers &b, int n_images, cudaStream_t s)
{
    orb_select_kernel<<<dim3(kOrbLevels, n_images), 256, 0, s>>>(g, b);
}

// ------------------------------------------------------------------------------------------ O5 Gaussian 7x7
// GaussianBlur(level, 7x7, sigma 2, BORDER_REFLECT_101) of orb.cpp runs OpenCV's float separable filter on the 8-bit
// level (the level is a sub-matrix, so the fixed-point path is not taken): row pass left to right, column pass centre
// first then symmetric pairs, every multiply-add fused, result rounded half-to-even.  getGaussianKernel(7, 2, CV_32F):
__constant__ float c_gauss[4] = {0x1.ba95c0p-3f, 0x1.869472p-3f, 0x1.0c70fcp-3f, 0x1.1f5f62p-4f};   // centre, +-1, +-2, +-3

__device__ __forceinline__ int reflect101(int i, int n)
{
    if (n >= 8) {                       // one reflection is enough for |overhang| <= n - 2; indices further out feed no output
        if (i < 0) i = -i;
        if (i >= n) i = 2 * n - 2 - i;
        return min(max(i, 0), n - 1);
    }
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

__device__ __forceinline__ float gauss_row7(const float *p, float k0, float k1, float k2, float k3)
{
    float s = __fmul_rn(k3, p[0]);
    s = fmaf(k2, p[1], s); s = fmaf(k1, p[2], s); s = fmaf(k0, p[3], s);
    s = fmaf(k1, p[4], s); s = fmaf(k2, p[5], s); s = fmaf(k3, p[6], s);
    return s;
}

__device__ __forceinline__ uint32_t gauss_col7(float c, float a1, float b1, float a2, float b2, float a3, float b3,
                                               float k0, float k1, float k2, float k3)
{
    float s = __fmul_rn(k0, c);
    s = fmaf(k1, __fadd_rn(a1, b1), s);
    s = fmaf(k2, __fadd_rn(a2, b2), s);
    s = fmaf(k3, __fadd_rn(a3, b3), s);
    return (uint32_t)min(max(__float2int_rn(s), 0), 255);
}

// 32x32 output tile: 38 x 40 input bytes as 10 aligned words per row, row pass 4 outputs per thread (float4 to shared
// memory), column pass 4 adjacent columns per thread (7 LDS.128, one 32-bit store).
__global__ void __launch_bounds__(256, 8)
orb_blur_kernel(const __grid_constant__ OrbGeom g, OrbBuffers b)
{
    __shared__ uint32_t in[38][10];
    __shared__ __align__(16) float row[38][32];
    int l = 0;
    while (l + 1 < kOrbLevels && (int)blockIdx.x >= g.lv[l + 1].blur_tile_off) ++l;
    const OrbLevel &L = g.lv[l];
    const int img = blockIdx.y;
    const int tiles_x = (L.w + 31) / 32;
    const int t = blockIdx.x - L.blur_tile_off;
    const int x0 = (t % tiles_x) * 32, y0 = (t / tiles_x) * 32;
    const uint8_t *src = b.pyr + (size_t)img * g.slab + L.off;
    const bool interior = x0 >= 4 && x0 + 36 <= L.w && y0 >= 3 && y0 + 35 <= L.h;
    for (int i = threadIdx.x; i < 380; i += 256) {
        const int r = i / 10, wd = i % 10;
        uint32_t v = 0;
        if (interior) {
            v = *reinterpret_cast<const uint32_t *>(src + (size_t)(y0 - 3 + r) * L.pitch + x0 - 4 + 4 * wd);
        } else {
            const uint8_t *rp = src + (size_t)reflect101(y0 - 3 + r, L.h) * L.pitch;
#pragma unroll
            for (int k = 0; k < 4; ++k) v |= (uint32_t)rp[reflect101(x0 - 4 + 4 * wd + k, L.w)] << (8 * k);
        }
        in[r][wd] = v;
    }
    __syncthreads();
    const float k0 = c_gauss[0], k1 = c_gauss[1], k2 = c_gauss[2], k3 = c_gauss[3];
    for (int i = threadIdx.x; i < 38 * 8; i += 256) {
        const int r = i >> 3, gq = i & 7;
        const uint32_t w0 = in[r][gq], w1 = in[r][gq + 1], w2 = in[r][gq + 2];
        float p[10];                    // tile columns 4gq+1 .. 4gq+10 = outputs 4gq..4gq+3 with their +-3 neighbours
        p[0] = (float)((w0 >> 8) & 255); p[1] = (float)((w0 >> 16) & 255); p[2] = (float)(w0 >> 24);
        p[3] = (float)(w1 & 255); p[4] = (float)((w1 >> 8) & 255); p[5] = (float)((w1 >> 16) & 255); p[6] = (float)(w1 >> 24);
        p[7] = (float)(w2 & 255); p[8] = (float)((w2 >> 8) & 255); p[9] = (float)((w2 >> 16) & 255);
        float4 o;
        o.x = gauss_row7(p, k0, k1, k2, k3); o.y = gauss_row7(p + 1, k0, k1, k2, k3);
        o.z = gauss_row7(p + 2, k0, k1, k2, k3); o.w = gauss_row7(p + 3, k0, k1, k2, k3);
        *reinterpret_cast<float4 *>(&row[r][4 * gq]) = o;
    }
    __syncthreads();
    const int r = threadIdx.x >> 3, gq = threadIdx.x & 7;
    const int x = x0 + 4 * gq, y = y0 + r;
    if (x < L.w && y < L.h) {
        const float4 c = *reinterpret_cast<const float4 *>(&row[r + 3][4 * gq]);
        const float4 a1 = *reinterpret_cast<const float4 *>(&row[r + 4][4 * gq]), b1 = *reinterpret_cast<const float4 *>(&row[r + 2][4 * gq]);
        const float4 a2 = *reinterpret_cast<const float4 *>(&row[r + 5][4 * gq]), b2 = *reinterpret_cast<const float4 *>(&row[r + 1][4 * gq]);
        const float4 a3 = *reinterpret_cast<const float4 *>(&row[r + 6][4 * gq]), b3 = *reinterpret_cast<const float4 *>(&row[r][4 * gq]);
        const uint32_t v = gauss_col7(c.x, a1.x, b1.x, a2.x, b2.x, a3.x, b3.x, k0, k1, k2, k3) |
                           gauss_col7(c.y, a1.y, b1.y, a2.y, b2.y, a3.y, b3.y, k0, k1, k2, k3) << 8 |
                           gauss_col7(c.z, a1.z, b1.z, a2.z, b2.z, a3.z, b3.z, k0, k1, k2, k3) << 16 |
                           gauss_col7(c.w, a1.w, b1.w, a2.w, b2.w, a3.w, b3.w, k0, k1, k2, k3) << 24;
        // the row padding (pitch is a multiple of 16) absorbs the bytes beyond the image width
        *reinterpret_cast<uint32_t *>(b.blur + (size_t)img * g.slab + L.off + (size_t)y * L.pitch + x) = v;
    }
}

void launch_orb_blur(const OrbGeom &g, const OrbBuffers &b, int n_images, cudaStream_t s)
{
    orb_blur_kernel<<<dim3(g.blur_tiles, n_images), 256, 0, s>>>(g, b);
}

// ------------------------------------------------------------------------------------------ O6 orientation + rBRIEF
// cv::fastAtan2 (degrees, 7th-order odd polynomial, FP32, one rounding per operation)
__device__ __forceinline__ float fast_atan2_deg(float y, float x)
{
    const float s = (float)(180.0 / 3.14159265358979323846);
    const float p1 = __fmul_rn(0.9997878412794807f, s), p3 = __fmul_rn(-0.3258083974640975f, s);
    const float p5 = __fmul_rn(0.1555786518463281f, s), p7 = __fmul_rn(-0.04432655554792128f, s);
    const float ax = fabsf(x), ay = fabsf(y);
    const float eps = 2.220446049250313e-16f;
    float a;
    if (ax >= ay) {
        const float c = __fdiv_rn(ay, __fadd_rn(ax, eps)), c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        const float c = __fdiv_rn(ax, __fadd_rn(ay, eps)), c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

__constant__ int8_t c_pattern[1024];

// cvRound for |x| < 2^22 without the XU pipe: adding 1.5 * 2^23 leaves the integer, rounded half to even by the FP32 add,
// in the low mantissa bits
__device__ __forceinline__ int round_half_even(float x)
{
    return __float_as_int(__fadd_rn(x, 12582912.0f)) - 0x4B400000;
}

// A warp takes G keypoints per pass: moments and descriptor bits are computed by the whole warp for one keypoint at a
// time, while the scalar part in between (fastAtan2, double-precision cos/sin, the keypoint record) runs once with
// lane j working on keypoint j.  G = 32 amortises the trigonometry best; small G keeps more warps busy when a call
// holds few keypoints (single-frame latency).
template <int G>
__global__ void __launch_bounds__(256)
orb_describe_kernel(const __grid_constant__ OrbGeom g, OrbBuffers b, OrbDescribeArgs d)
{
    // pattern transposed for conflict-free access: s_pat[k][lane] = point k (0..15) of the byte `lane`, already as floats
    // (int -> float conversions and float -> int roundings run on the quarter-rate XU pipe: none is left in the loop)
    __shared__ float2 s_pat[16][32];
    __shared__ int s_start[kOrbLevels + 1];
    const int img = blockIdx.y;
    for (int i = threadIdx.x; i < 512; i += 256)
        s_pat[i & 15][i >> 4] = make_float2((float)c_pattern[2 * i], (float)c_pattern[2 * i + 1]);
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int l = 0; l < kOrbLevels; ++l) { s_start[l] = acc; acc += max(b.kept_cnt[img * kOrbLevels + l], 0); }
        s_start[kOrbLevels] = acc;
    }
    __syncthreads();
    const int total = s_start[kOrbLevels];
    const int lane = threadIdx.x & 31;
    const int warps = gridDim.x * 8;
    // circular patch: lane = column u + 15; bit v + 15 of rowmask says whether (u, v) lies inside (|u| <= umax[|v|])
    const int u = lane - kOrbHalfPatch;
    uint32_t rowmask = 0;
    if (lane < 31)
        for (int v = -kOrbHalfPatch; v <= kOrbHalfPatch; ++v)
            if (abs(u) <= g.umax[abs(v)]) rowmask |= 1u << (v + kOrbHalfPatch);
    const size_t out0 = (size_t)d.img_off[img];
    const uint8_t *pyr = b.pyr + (size_t)img * g.slab, *blur = b.blur + (size_t)img * g.slab;
    for (int base = (blockIdx.x * 8 + (threadIdx.x >> 5)) * G; base < total; base += warps * G) {
        // lane j < G owns keypoint base + j
        const int k = base + lane;
        const bool own = lane < G && k < total;
        int l = 0, cx = 0, cy = 0;
        float resp = 0.f;
        if (own) {
            while (k >= s_start[l + 1]) ++l;
            const uint32_t ci = b.kept_idx[(size_t)(img * kOrbLevels + l) * kOrbSortCap + (k - s_start[l])];
            const size_t o = (size_t)img * g.cand_total + g.lv[l].cand_off + ci;
            const uint32_t xy = b.cand_xy[o];
            cx = xy & 0xffff; cy = xy >> 16;
            resp = b.cand_val[o];
        }
        const int n_here = min(G, total - base);
        // ICAngles: moments over the circular patch of the un-blurred level; lane = u + 15
        int m10_own = 0, m01_own = 0;
        for (int j = 0; j < n_here; ++j) {
            const int jl = __shfl_sync(0xffffffffu, l, j), jx = __shfl_sync(0xffffffffu, cx, j), jy = __shfl_sync(0xffffffffu, cy, j);
            const int pitch = g.lv[jl].pitch;
            const uint8_t *im = pyr + g.lv[jl].off + (size_t)jy * pitch + jx;
            int m10 = 0, m01 = 0;
            const uint8_t *row = im - kOrbHalfPatch * pitch + u;
#pragma unroll
            for (int v = -kOrbHalfPatch; v <= kOrbHalfPatch; ++v, row += pitch) {
                if ((rowmask >> (v + kOrbHalfPatch)) & 1u) {
                    const int p = *row;
                    m10 += u * p; m01 += v * p;
                }
            }
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) {
                m10 += __shfl_xor_sync(0xffffffffu, m10, s);
                m01 += __shfl_xor_sync(0xffffffffu, m01, s);
            }
            if (lane == j) { m10_own = m10; m01_own = m01; }
        }
        // per-keypoint scalars, one keypoint per lane
        float angle = 0.f, ca = 1.f, sa = 0.f;
        if (own) {
            angle = fast_atan2_deg((float)m01_own, (float)m10_own);
            // computeOrbDescriptors: angle in radians (float), cos/sin evaluated in double and rounded to float
            const float rad = __fmul_rn(angle, (float)(3.14159265358979323846 / 180.f));
            double sd, cd;
            sincos((double)rad, &sd, &cd);
            ca = (float)cd; sa = (float)sd;
            const float scale = g.lv[l].scale;
            const float px = __fmul_rn((float)cx, scale), py = __fmul_rn((float)cy, scale);
            if (d.kp) {
                mvs_keypoint kp;
                kp.x = px; kp.y = py;
                kp.size = __fmul_rn(31.f, scale);
                kp.angle = angle;
                kp.response = resp;
                kp.octave = l;
                d.kp[out0 + k] = kp;
            }
            if (d.frame_kp) d.frame_kp[out0 + k] = make_float2(px, py);
        }
        // steered BRIEF: lane = descriptor byte
        for (int j = 0; j < n_here; ++j) {
            const int jl = __shfl_sync(0xffffffffu, l, j), jx = __shfl_sync(0xffffffffu, cx, j), jy = __shfl_sync(0xffffffffu, cy, j);
            const float jc = __shfl_sync(0xffffffffu, ca, j), js = __shfl_sync(0xffffffffu, sa, j);
            const int pitch = g.lv[jl].pitch;
            const uint8_t *bl = blur + g.lv[jl].off + (size_t)jy * pitch + jx;
            int byte = 0;
#pragma unroll
            for (int bit = 0; bit < 8; ++bit) {
                const float2 q0 = s_pat[2 * bit][lane], q1 = s_pat[2 * bit + 1][lane];
                const int x0 = round_half_even(__fsub_rn(__fmul_rn(q0.x, jc), __fmul_rn(q0.y, js)));
                const int y0 = round_half_even(__fadd_rn(__fmul_rn(q0.x, js), __fmul_rn(q0.y, jc)));
                const int x1 = round_half_even(__fsub_rn(__fmul_rn(q1.x, jc), __fmul_rn(q1.y, js)));
                const int y1 = round_half_even(__fadd_rn(__fmul_rn(q1.x, js), __fmul_rn(q1.y, jc)));
                byte |= (int)(bl[y0 * pitch + x0] < bl[y1 * pitch + x1]) << bit;
            }
            d.desc[(out0 + base + j) * 32 + lane] = (uint8_t)byte;
        }
    }
}

void launch_orb_describe(const OrbGeom &g, const OrbBuffers &b, const OrbDescribeArgs &d, int n_images, int max_keypoints,
                         cudaStream_t s)
{
    static bool pattern_loaded[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 64 || !pattern_loaded[dev]) {   // synchronous copy: complete before any kernel that reads it is enqueued
        cudaMemcpyToSymbol(c_pattern, MVS_ORB_PATTERN, sizeof(MVS_ORB_PATTERN));
        if (dev < 64) pattern_loaded[dev] = true;
    }
    // enough warps to fill the GPU (148 SMs x 8 resident CTAs) before the group size grows
    const int kp = std::max(max_keypoints, 1);
    auto ctas = [&](int G) { return (kp + 8 * G - 1) / (8 * G); };
    if ((long)ctas(32) * n_images >= 1184) orb_describe_kernel<32><<<dim3(ctas(32), n_images), 256, 0, s>>>(g, b, d);
    else if ((long)ctas(8) * n_images >= 592) orb_describe_kernel<8><<<dim3(ctas(8), n_images), 256, 0, s>>>(g, b, d);
    else orb_describe_kernel<2><<<dim3(ctas(2), n_images), 256, 0, s>>>(g, b, d);
}

}  // namespace mvs
