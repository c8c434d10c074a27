// api.cu — context, HBM workspace and the extern "C" entry points of libmvslam_b200.so.
// Everything here is host-side plumbing: argument checks, host<->device copies, stage launches.
// There is deliberately no CPU implementation of any stage.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "l2.h"
#include "orb.h"
#include "pnp.h"
#include "ba.h"

using namespace mvs;

static_assert(sizeof(mvs_pair_result) == 376, "mvs_pair_result layout is part of the ABI");
static_assert(sizeof(mvs_match) == 12, "mvs_match layout is part of the ABI");
static_assert(sizeof(mvs_keypoint) == 24 && sizeof(mvs_orb_params) == 16, "extraction structs are part of the ABI");
static_assert(sizeof(mvs_pnp_result) == 208 && sizeof(mvs_pnp_params) == 40, "pnp structs are part of the ABI");
static_assert(sizeof(mvs_ba_observation) == 48 && sizeof(mvs_ba_result) == 24 && sizeof(mvs_ba_params) == 32, "BA structs are part of the ABI");
static_assert(MVS_N_STAGES == 16, "mvs_profile layout is part of the ABI (capi.py STAGES)");

namespace {

// Guard mode (MVS_GUARD=1 in the environment when the library is loaded; a test hook, see mvs_debug_guard_check): every
// workspace buffer is allocated at exactly the size that was asked for, followed by a 4 KB band of a known byte, so a kernel
// that writes past the end of its buffer is caught by the check instead of landing in the allocation slack.
constexpr size_t kGuardBytes = 4096;
constexpr int kGuardByte = 0xA5;
inline bool guard_mode()
{
    static const bool on = [] { const char *e = std::getenv("MVS_GUARD"); return e && e[0] == '1'; }();
    return on;
}

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        if (guard_mode()) {
            cudaError_t e = cudaMalloc(&p, bytes + kGuardBytes);
            if (e != cudaSuccess) return e;
            cap = bytes;
            return cudaMemset(static_cast<uint8_t *>(p) + bytes, kGuardByte, kGuardBytes);
        }
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct Pending { int stage; cudaEvent_t a, b; };

}  // namespace

struct mvs_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    uint64_t launches = 0;
    // resident frame table
    DevBuf d_desc, d_kp, d_foff, d_fcnt;
    std::vector<int32_t> h_off, h_cnt;
    // scratch frame table of the two-set matcher entry points
    DevBuf t_desc, t_foff, t_fcnt;
    // tensor-core matcher (match_hamming_tc.cu): the descriptor tables expanded to one +-1 byte per bit ([rows][256]);
    // rows [0, desc8_rows) of d_desc8 mirror d_desc
    DevBuf d_desc8, t_desc8;
    size_t desc8_rows = 0;
    bool use_tc = true;          // MVS_MATCHER=popc selects the integer-pipe kernel (A/B measurements, > 32768 train descriptors)
    // workspace
    DevBuf d_pairs, d_partial, d_rev, d_matches, d_nmatch, d_points, d_state, d_Fall, d_pc, d_pr, d_mask,
        d_valid, d_tri, d_items, d_item_total, d_opts, d_oidx, d_results, d_table, d_in1, d_in2, d_knn_i, d_knn_d, d_counts, d_pres;
    mvs::L2Workspace l2;
    // feature extraction (orb.cu): pyramid geometry cached per (width, height, nfeatures), workspace, last results
    int orb_w = 0, orb_h = 0, orb_nf = -1;
    mvs::OrbGeom orb_geom;
    DevBuf o_tabs, o_stage, o_pyr, o_blur, o_cxy, o_cval, o_cnt, o_kidx, o_kcnt, o_off, o_kp, o_desc;
    DevBuf p_world, p_image, p_off, p_table, p_poses, p_valid, p_pc, p_maskws, p_mask, p_counts, p_results;
    DevBuf b_foff, b_poff, b_R, b_t, b_pc, b_X, b_xc, b_obs, b_ooff, b_ws, b_wsm, b_wsm_off, b_wsobs, b_Ro, b_to, b_pco, b_Xo, b_xco, b_res;
    int32_t *o_pinned = nullptr;
    int32_t *h_counts = nullptr;         // pinned: per-pair match counts read back before the detail copies (synchronous calls)
    size_t h_counts_cap = 0;
    // host images are fetched one chunk ahead on a copy stream (orb_extract_impl)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t copy_done[2] = {nullptr, nullptr};
    size_t o_pinned_cap = 0;
    // pinned staging for small device->host results that the caller wants in pageable memory: the copies are enqueued
    // asynchronously into the staging area and scattered to the caller's buffers after the one synchronisation
    uint8_t *h_stage = nullptr;
    size_t h_stage_cap = 0, h_stage_used = 0;
    // small host->device arguments (pair lists, frame offset tables) go through a pinned ring: cudaMemcpyAsync from pageable
    // memory first waits for the stream's earlier work, which serialises contexts that are meant to overlap
    static constexpr int kArgSlots = 32;
    static constexpr size_t kArgSlotBytes = 64 << 10;
    uint8_t *h_args = nullptr;
    cudaEvent_t arg_ev[kArgSlots] = {};
    int arg_next = 0;
    bool skip_d2h = false;      // pair_batch_chunk leaves every output on the device (mvs_pair_batch_device_only)
    int last_stride = 0;        // detail stride (largest pair frame) of the last pair_batch chunk: sharded.cu reads the device outputs
    bool allow_stage = false;   // set by the synchronous entry points only: _enqueue callers may synchronise the stream themselves
    struct StagedCopy { void *dst; size_t dpitch; size_t src_off; size_t width; size_t rows; };
    std::vector<StagedCopy> staged;
    // details of a small synchronous batch exported by the device into the staging area (export_details_kernel): scattered to
    // the caller's pageable buffers -- only the entries each pair owns -- by flush_staged()
    struct StagedExport {
        mvs_pair_result *results; mvs_match *matches; uint8_t *mask; double *points; uint64_t *indexes;
        int n_pairs, w, capacity;
        size_t off_res, off_m, off_k, off_p, off_i;
    };
    std::vector<StagedExport> staged_exports;
    // profiling
    bool prof = false;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<Pending> pending;
    mvs_profile acc;
};

namespace {

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);                  \
            return MVS_E_CUDA;                                                                \
        }                                                                                     \
    } while (0)

int fail(mvs_ctx *ctx, int code, const char *msg)
{
    if (ctx) ctx->err = msg;
    return code;
}

// host -> device copy of a small argument block on the ctx stream, staged through the pinned ring when it fits
cudaError_t h2d_args(mvs_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    if (!bytes) return cudaSuccess;
    if (bytes > mvs_ctx::kArgSlotBytes) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (!ctx->h_args) {
        cudaError_t e = cudaMallocHost((void **)&ctx->h_args, mvs_ctx::kArgSlots * mvs_ctx::kArgSlotBytes);
        if (e != cudaSuccess) { ctx->h_args = nullptr; (void)cudaGetLastError(); return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream); }
        for (int i = 0; i < mvs_ctx::kArgSlots; ++i) cudaEventCreateWithFlags(&ctx->arg_ev[i], cudaEventDisableTiming);
    }
    const int slot = ctx->arg_next;
    ctx->arg_next = (slot + 1) % mvs_ctx::kArgSlots;
    cudaEventSynchronize(ctx->arg_ev[slot]);                 // the copy that last used this slot (32 calls ago) has left it
    uint8_t *at = ctx->h_args + (size_t)slot * mvs_ctx::kArgSlotBytes;
    std::memcpy(at, src, bytes);
    cudaError_t e = cudaMemcpyAsync(dst, at, bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->arg_ev[slot], ctx->stream);
    return e;
}

cudaEvent_t get_event(mvs_ctx *ctx)
{
    if (!ctx->ev_pool.empty()) { cudaEvent_t e = ctx->ev_pool.back(); ctx->ev_pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

struct StageTimer {
    mvs_ctx *ctx; int stage; cudaEvent_t a = nullptr;
    StageTimer(mvs_ctx *c, int s, int n_launch = 1) : ctx(c), stage(s)
    {
        ctx->launches += (uint64_t)n_launch;
        ctx->acc.launches[stage] += (uint64_t)n_launch;
        if (ctx->prof) { a = get_event(ctx); cudaEventRecord(a, ctx->stream); }
    }
    ~StageTimer()
    {
        if (a) { cudaEvent_t b = get_event(ctx); cudaEventRecord(b, ctx->stream); ctx->pending.push_back({stage, a, b}); }
    }
};

// ---- host copies of the tiny SE3/SO3 algebra needed to prepare inputs (same operation order as
//      the device versions; reference source/math/lie-group.hpp:75-96,203-225, camera.cpp:14-18)
void h_cross3(const double a[3], const double b[3], double o[3])
{
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
void h_rectify(const double R[9], double out[9])
{
    double n0 = std::sqrt(R[0] * R[0] + R[1] * R[1] + R[2] * R[2]);
    double u0[3] = {R[0] / n0, R[1] / n0, R[2] / n0};
    double d = R[3] * u0[0] + R[4] * u0[1] + R[5] * u0[2];
    double u1[3] = {R[3] - d * u0[0], R[4] - d * u0[1], R[5] - d * u0[2]}, u2[3];
    h_cross3(u0, u1, u2);
    for (int k = 0; k < 3; ++k) { out[k] = u0[k]; out[3 + k] = u1[k]; out[6 + k] = u2[k]; }
}
void h_mat3_vec(const double A[9], const double v[3], double o[3])
{
    for (int i = 0; i < 3; ++i) o[i] = A[i * 3] * v[0] + A[i * 3 + 1] * v[1] + A[i * 3 + 2] * v[2];
}
void h_se3_inverse(const double R[9], const double t[3], double Ro[9], double to[3])
{
    double Rt[9], v[3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Rt[i * 3 + j] = R[j * 3 + i];
    h_rectify(Rt, Ro);
    h_mat3_vec(Ro, t, v);
    to[0] = -v[0]; to[1] = -v[1]; to[2] = -v[2];
}
void h_se3_compose(const double Ra[9], const double ta[3], const double Rb[9], const double tb[3], double Ro[9], double to[3])
{
    double P[9], v[3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) P[i * 3 + j] = Ra[i * 3] * Rb[j] + Ra[i * 3 + 1] * Rb[3 + j] + Ra[i * 3 + 2] * Rb[6 + j];
    h_rectify(P, Ro);
    h_mat3_vec(Ra, tb, v);
    to[0] = v[0] + ta[0]; to[1] = v[1] + ta[1]; to[2] = v[2] + ta[2];
}
void h_inverse3(const double K[9], double Ki[9])
{   // Eigen's fixed 3x3 inverse: cofactor^T * (1/det)
    double c00 = K[4] * K[8] - K[5] * K[7], c01 = K[5] * K[6] - K[3] * K[8], c02 = K[3] * K[7] - K[4] * K[6];
    double c10 = K[2] * K[7] - K[1] * K[8], c11 = K[0] * K[8] - K[2] * K[6], c12 = K[1] * K[6] - K[0] * K[7];
    double c20 = K[1] * K[5] - K[2] * K[4], c21 = K[2] * K[3] - K[0] * K[5], c22 = K[0] * K[4] - K[1] * K[3];
    double det = K[0] * c00 + K[1] * c01 + K[2] * c02, id = 1.0 / det;
    Ki[0] = c00 * id; Ki[1] = c10 * id; Ki[2] = c20 * id;
    Ki[3] = c01 * id; Ki[4] = c11 * id; Ki[5] = c21 * id;
    Ki[6] = c02 * id; Ki[7] = c12 * id; Ki[8] = c22 * id;
}

struct RansacCfg { int H; int mode; double thr; uint64_t seed; int min_inl; uint64_t pair_base; int solver; };

int resolve_ransac(mvs_ctx *ctx, const mvs_ransac_params *rp, const double *K, RansacCfg &c)
{
    c.H = rp ? rp->n_hypotheses : 1;
    c.mode = rp ? rp->score_mode : MVS_SCORE_ALGEBRAIC;
    c.thr = rp ? rp->max_error_sq : 0.0;
    c.seed = rp ? rp->seed : 0;
    c.pair_base = rp ? rp->pair_id_base : 0;
    c.min_inl = (rp && rp->min_inliers > 0) ? rp->min_inliers : kMinInliers;
    c.solver = rp ? rp->solver : MVS_SOLVER_REFERENCE;
    if (c.solver != MVS_SOLVER_REFERENCE && c.solver != MVS_SOLVER_FAST) return fail(ctx, MVS_E_BAD_ARG, "bad solver");
    if (c.H < 1) return fail(ctx, MVS_E_BAD_ARG, "n_hypotheses must be >= 1");
    if (c.mode != MVS_SCORE_ALGEBRAIC && c.mode != MVS_SCORE_SAMPSON) return fail(ctx, MVS_E_BAD_ARG, "bad score_mode");
    if (!(c.thr > 0.0)) {
        if (!K) return fail(ctx, MVS_E_BAD_ARG, "max_error_sq must be > 0");
        c.thr = kMaxErrorSq / K[0] / K[4];  // sfm-solve.cpp:311
    }
    if (!(c.thr > kEpsilon)) return fail(ctx, MVS_E_BAD_ARG, "max_error_sq must be > epsilon");  // sfm-solve.cpp:39
    return MVS_OK;
}

// K3..K5 (+K6,K7) over n_pairs pairs whose correspondences sit in d_points[pairs][p_stride][6]
int run_geometry(mvs_ctx *ctx, int n_pairs, int p_stride, const RansacCfg &rc, bool unit_z, double zc1, double zc2, const uint32_t *d_table,
                 uint64_t pair_id_base, bool decompose, bool want_all_counts, const mvs_match *d_matches)
{
    const int tiles = score_tiles(p_stride);
    CK(ctx->d_Fall.ensure((size_t)n_pairs * rc.H * 9 * sizeof(double)));
    CK(ctx->d_pc.ensure((size_t)n_pairs * tiles * rc.H * sizeof(uint32_t)));
    CK(ctx->d_pr.ensure((size_t)n_pairs * rc.H * 2 * sizeof(int32_t)));
    const bool k4_res = (rc.mode == MVS_SCORE_ALGEBRAIC);
    if (k4_res) CK(ctx->d_pres.ensure((size_t)n_pairs * tiles * rc.H * sizeof(double)));
    CK(ctx->d_mask.ensure((size_t)n_pairs * p_stride));
    if (want_all_counts) CK(ctx->d_counts.ensure((size_t)n_pairs * rc.H * sizeof(int32_t)));
    if (decompose) {
        CK(ctx->d_valid.ensure((size_t)n_pairs * 4 * p_stride));
        CK(ctx->d_items.ensure((size_t)n_pairs * p_stride * sizeof(unsigned long long)));
        CK(ctx->d_item_total.ensure(sizeof(uint32_t)));
    }
    PairState *state = ctx->d_state.as<PairState>();
    {
        StageTimer t(ctx, MVS_STAGE_HYPOTHESES);
        HypArgs a{};
        a.points = ctx->d_points.as<double>(); a.p_stride = p_stride; a.state = state;
        a.table = d_table; a.seed = rc.seed; a.pair_id_base = pair_id_base; a.H = rc.H; a.F_all = ctx->d_Fall.as<double>();
        a.solver = rc.solver;
        launch_hypotheses(a, n_pairs, ctx->stream);
    }
    {
        StageTimer t(ctx, MVS_STAGE_SCORE);
        ScoreArgs a{};
        a.points = ctx->d_points.as<double>(); a.p_stride = p_stride; a.state = state; a.F_all = ctx->d_Fall.as<double>();
        a.H = rc.H; a.max_error_sq = rc.thr; a.zc1 = zc1; a.zc2 = zc2; a.tiles = tiles; a.part_count = ctx->d_pc.as<uint32_t>();
        a.part_res = k4_res ? ctx->d_pres.as<double>() : nullptr; a.solver = rc.solver;
        a.zero_word = decompose ? ctx->d_item_total.as<uint32_t>() : nullptr;
        launch_score(a, rc.mode, unit_z, n_pairs, ctx->stream);
    }
    {
        StageTimer t(ctx, MVS_STAGE_SELECT);
        SelectArgs a{};
        a.points = ctx->d_points.as<double>(); a.p_stride = p_stride; a.state = state; a.F_all = ctx->d_Fall.as<double>();
        a.H = rc.H; a.part_count = ctx->d_pc.as<uint32_t>(); a.ties = ctx->d_pr.as<int32_t>(); a.tiles = tiles;
        a.part_res = k4_res ? ctx->d_pres.as<double>() : nullptr;
        a.max_error_sq = rc.thr; a.zc1 = zc1; a.zc2 = zc2; a.min_inliers = rc.min_inl; a.decompose = decompose ? 1 : 0;
        a.mask = ctx->d_mask.as<uint8_t>(); a.all_counts = want_all_counts ? ctx->d_counts.as<int32_t>() : nullptr;
        a.solver = rc.solver;
        if (decompose) {
            a.items = ctx->d_items.as<unsigned long long>(); a.item_total = ctx->d_item_total.as<uint32_t>();
            a.valid = ctx->d_valid.as<uint8_t>();
        }
        launch_select(a, rc.mode, unit_z, p_stride, n_pairs, ctx->stream);
    }
    if (!decompose) return MVS_OK;
    CK(ctx->d_tri.ensure((size_t)n_pairs * 4 * p_stride * 3 * sizeof(double)));
    CK(ctx->d_opts.ensure((size_t)n_pairs * p_stride * 3 * sizeof(double)));
    CK(ctx->d_oidx.ensure((size_t)n_pairs * p_stride * sizeof(uint64_t)));
    CK(ctx->d_results.ensure((size_t)n_pairs * sizeof(mvs_pair_result)));
    {
        StageTimer t(ctx, MVS_STAGE_TRIANGULATE);
        TriArgs a{};
        a.points = ctx->d_points.as<double>(); a.p_stride = p_stride; a.state = state; a.mask = ctx->d_mask.as<uint8_t>();
        a.n_cand = 4; a.valid = ctx->d_valid.as<uint8_t>(); a.tri = ctx->d_tri.as<double>(); a.solver = rc.solver;
        a.items = ctx->d_items.as<unsigned long long>(); a.item_total = ctx->d_item_total.as<uint32_t>();
        CK(launch_triangulate_items(a, (size_t)n_pairs * p_stride, ctx->stream));
    }
    {
        StageTimer t(ctx, MVS_STAGE_FINALIZE);
        FinishArgs a{};
        a.state = state; a.p_stride = p_stride; a.n_cand = 4; a.valid = ctx->d_valid.as<uint8_t>(); a.tri = ctx->d_tri.as<double>();
        a.matches = d_matches; a.out_points = ctx->d_opts.as<double>(); a.out_index = ctx->d_oidx.as<uint64_t>();
        a.results = ctx->d_results.as<mvs_pair_result>();
        launch_finish(a, p_stride, n_pairs, ctx->stream);
    }
    CK(cudaGetLastError());
    return MVS_OK;
}

// Tensor-core matcher eligibility: train index must fit the epilogue key, the expanded table must fit comfortably.
bool tc_eligible(const mvs_ctx *ctx, int max_nq, int max_nt, bool cross, size_t table_rows)
{
    return ctx->use_tc && max_nt <= tc_max_train() && (!cross || max_nq <= tc_max_train()) && table_rows * 256 <= ((size_t)16 << 30);
}

// bring d_desc8 up to date with the first `total` rows of d_desc (lazy: only rows added since the last call are expanded)
int sync_desc8(mvs_ctx *ctx, size_t total)
{
    if (ctx->d_desc8.cap < total * 256) { CK(ctx->d_desc8.ensure(total * 256)); ctx->desc8_rows = 0; }
    if (ctx->desc8_rows < total) {
        launch_expand_desc(ctx->d_desc.as<uint4>(), ctx->desc8_rows, total - ctx->desc8_rows, ctx->d_desc8.p, ctx->stream);
        ctx->launches += 1; ctx->acc.launches[MVS_STAGE_KNN] += 1;
        ctx->desc8_rows = total;
    }
    return MVS_OK;
}

int choose_splits(int q_tiles, int n_pairs, int nt_max)
{
    const int train_tiles = std::max(1, (nt_max + 255) / 256);
    const long ctas = (long)q_tiles * n_pairs;
    const long target = 148L * 6;  // a few CTAs of 256 threads per SM
    long s = (target + ctas - 1) / ctas;
    s = std::max(1L, std::min<long>(s, train_tiles));
    return (int)std::min<long>(s, 64);
}

// K^-1 (u,v,1) has the constant z = Kinv[8] for every point when the last row of K^-1 is (0, 0, c)
// bounded search (mvs_match_params.bounded): smallest integer B > max_dist with ratio * B > max_dist, evaluated
// with the same double arithmetic as the filter; 0 disables (max_dist < 0, or nothing to gain)
uint32_t search_bound(const mvs_match_params *mp)
{
    if (!mp || !mp->bounded || mp->max_dist < 0 || !(mp->ratio > 0)) return 0;
    uint32_t b = (uint32_t)std::min(300.0, std::floor(mp->max_dist)) + 1;
    while (b <= 256 && !(mp->ratio * (double)(float)b > mp->max_dist)) ++b;
    return b > 256 ? 0 : b;      // beyond the largest possible distance: plain evaluation
}

// mvs_frames_upload from pinned host memory: one kernel reads every frame's descriptors and keypoints over the host interface
// and writes the offset / count tables, instead of two DMA copies per frame that the stream runs one after the other (five VO
// frames: 12 copies, ~60 us of a 90 us upload; the kernel: ~10 us).  Up to kGatherFrames frames per launch (kernel parameters).
constexpr int kGatherFrames = 64;
struct GatherArgs {
    const uint4 *desc[kGatherFrames];
    const float2 *kp[kGatherFrames];
    int32_t off[kGatherFrames], cnt[kGatherFrames];
    uint4 *d_desc; float2 *d_kp; int32_t *d_foff, *d_fcnt;
    int f0;
};

__global__ void __launch_bounds__(256) gather_frames_kernel(const __grid_constant__ GatherArgs a)
{
    const int f = blockIdx.y;
    const int cnt = a.cnt[f], off = a.off[f];
    if (blockIdx.x == 0 && threadIdx.x == 0) { a.d_foff[a.f0 + f] = off; a.d_fcnt[a.f0 + f] = cnt; }
    const uint4 *sd = a.desc[f];
    uint4 *dd = a.d_desc + 2 * (size_t)off;
    const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = t0; i < 2 * cnt; i += stride) dd[i] = sd[i];
    const float2 *sk = a.kp[f];
    float2 *dk = a.d_kp + off;
    for (int i = t0; i < cnt; i += stride) dk[i] = sk[i];
}

// Detail outputs straight into the caller's buffers when those are pinned (device-accessible) host memory: one kernel writes
// exactly the entries every pair owns (n_matches matches / mask bytes, n_points points / indexes) over the host interface,
// instead of four strided copies of max-count rows after a round trip for the counts.  The records say how many.
struct ExportArgs {
    const mvs_pair_result *res;
    int stride, capacity;
    const mvs_match *m; mvs_match *mo;
    const uint8_t *k; uint8_t *ko;
    const double *p; double *po;
    const uint64_t *i; uint64_t *io;
};

__global__ void __launch_bounds__(128) export_details_kernel(ExportArgs a)
{
    mvs::pdl_wait();
    const int pair = blockIdx.x;
    const int lim = min(a.capacity, a.stride);
    const int nm = max(0, min(a.res[pair].n_matches, lim)), np = max(0, min(a.res[pair].n_points, lim));
    if (a.mo) {     // 12-byte structs as 32-bit words
        const uint32_t *s = reinterpret_cast<const uint32_t *>(a.m + (size_t)pair * a.stride);
        uint32_t *d = reinterpret_cast<uint32_t *>(a.mo + (size_t)pair * a.capacity);
        for (int i = threadIdx.x; i < nm * 3; i += blockDim.x) d[i] = s[i];
    }
    if (a.ko) {
        const uint8_t *s = a.k + (size_t)pair * a.stride;
        uint8_t *d = a.ko + (size_t)pair * a.capacity;
        for (int i = threadIdx.x; i < nm; i += blockDim.x) d[i] = s[i];
    }
    if (a.po) {
        const double *s = a.p + (size_t)pair * a.stride * 3;
        double *d = a.po + (size_t)pair * a.capacity * 3;
        for (int i = threadIdx.x; i < np * 3; i += blockDim.x) d[i] = s[i];
    }
    if (a.io) {
        const uint64_t *s = a.i + (size_t)pair * a.stride;
        uint64_t *d = a.io + (size_t)pair * a.capacity;
        for (int i = threadIdx.x; i < np; i += blockDim.x) d[i] = s[i];
    }
}

// the device-side address of a pinned host buffer (null for pageable or device memory, and for a null pointer)
void *pinned_device_ptr(const void *host)
{
    if (!host) return nullptr;
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}

// Device -> host copy of `rows` rows of `width` bytes.  With `stage` the data lands in the ctx's pinned staging area
// (asynchronous even when `dst` is pageable) and is scattered by flush_staged() after the stream has been synchronised.
cudaError_t d2h_rows(mvs_ctx *ctx, bool stage, void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t rows)
{
    if (!width || !rows) return cudaSuccess;
    if (!stage || ctx->h_stage_used + width * rows > ctx->h_stage_cap)
        return cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, cudaMemcpyDeviceToHost, ctx->stream);
    uint8_t *at = ctx->h_stage + ctx->h_stage_used;
    cudaError_t e = cudaMemcpy2DAsync(at, width, src, spitch, width, rows, cudaMemcpyDeviceToHost, ctx->stream);
    if (e != cudaSuccess) return e;
    ctx->staged.push_back({dst, dpitch, ctx->h_stage_used, width, rows});
    ctx->h_stage_used += width * rows;
    return cudaSuccess;
}

void flush_staged(mvs_ctx *ctx)     // call only after the stream has been synchronised
{
    for (const auto &c : ctx->staged)
        for (size_t r = 0; r < c.rows; ++r)
            std::memcpy(static_cast<uint8_t *>(c.dst) + r * c.dpitch, ctx->h_stage + c.src_off + r * c.width, c.width);
    ctx->staged.clear();
    for (const auto &e : ctx->staged_exports) {
        const mvs_pair_result *res = reinterpret_cast<const mvs_pair_result *>(ctx->h_stage + e.off_res);
        std::memcpy(e.results, res, (size_t)e.n_pairs * sizeof(mvs_pair_result));
        for (int i = 0; i < e.n_pairs; ++i) {
            const size_t nm = (size_t)std::max(0, std::min(res[i].n_matches, e.w)), np = (size_t)std::max(0, std::min(res[i].n_points, e.w));
            const size_t row = (size_t)i * e.w, urow = (size_t)i * e.capacity;
            if (e.matches && nm) std::memcpy(e.matches + urow, ctx->h_stage + e.off_m + row * sizeof(mvs_match), nm * sizeof(mvs_match));
            if (e.mask && nm) std::memcpy(e.mask + urow, ctx->h_stage + e.off_k + row, nm);
            if (e.points && np) std::memcpy(e.points + urow * 3, ctx->h_stage + e.off_p + row * 24, np * 24);
            if (e.indexes && np) std::memcpy(e.indexes + urow, ctx->h_stage + e.off_i + row * 8, np * 8);
        }
    }
    ctx->staged_exports.clear();
    ctx->h_stage_used = 0;
}

}  // namespace

namespace mvs {
bool pdl_enabled()
{
    static const bool on = [] { const char *e = std::getenv("MVS_PDL"); return !(e && e[0] == '0'); }();
    return on;
}
}  // namespace mvs

namespace {

bool unit_z_intrinsics(const double Ki[9]) { return Ki[6] == 0.0 && Ki[7] == 0.0; }

}  // namespace

// ---- accessors for the multi-GPU driver (sharded.cu), which lives in its own translation unit
int mvs_ctx_device(const mvs_ctx *ctx) { return ctx->device; }
cudaStream_t mvs_ctx_stream(const mvs_ctx *ctx) { return ctx->stream; }
void mvs_ctx_set_error(mvs_ctx *ctx, const std::string &msg) { if (ctx) ctx->err = msg; }
void mvs_ctx_last_outputs(mvs_ctx *ctx, const mvs_pair_result **res, const mvs_match **matches, const double **points,
                          const uint64_t **indexes, int *stride)
{
    *res = ctx->d_results.as<mvs_pair_result>(); *matches = ctx->d_matches.as<mvs_match>(); *points = ctx->d_opts.as<double>();
    *indexes = ctx->d_oidx.as<uint64_t>(); *stride = ctx->last_stride;
}

// ============================================================================================ C ABI
extern "C" {

int mvs_abi_version(void) { return MVS_ABI_VERSION; }

const char *mvs_status_string(int s)
{
    switch (s) {
    case MVS_OK: return "ok";
    case MVS_E_BAD_ARG: return "bad argument";
    case MVS_E_TOO_FEW_POINTS: return "fewer than 8 correspondences";
    case MVS_E_NO_MODEL: return "no model with any inlier";
    case MVS_E_TOO_FEW_INLIERS: return "fewer inliers than the minimum";
    case MVS_E_NO_CHEIRALITY: return "no candidate pose passes the cheirality test";
    case MVS_E_CUDA: return "CUDA error";
    case MVS_E_CAPACITY: return "output capacity too small";
    case MVS_E_UNSUPPORTED: return "unsupported";
    default: return "unknown status";
    }
}

int mvs_create(mvs_ctx **out, int device)
{
    if (!out) return MVS_E_BAD_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return MVS_E_CUDA;  // no CPU fallback
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return MVS_E_CUDA;   // negative: the caller's current device
    if (device < 0 || device >= count) return MVS_E_BAD_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return MVS_E_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return MVS_E_CUDA;
    if (prop.major != 10) return MVS_E_UNSUPPORTED;  // sm_100a binary only
    mvs_ctx *ctx = new mvs_ctx();
    ctx->device = device;
    { const char *e = std::getenv("MVS_MATCHER"); ctx->use_tc = !(e && e[0] == 'p'); }
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return MVS_E_CUDA; }
    ctx->own_stream = true;
    std::memset(&ctx->acc, 0, sizeof(ctx->acc));
    *out = ctx;
    return MVS_OK;
}

static std::vector<DevBuf *> all_buffers(mvs_ctx *ctx)
{
    return {&ctx->d_desc, &ctx->d_kp, &ctx->d_foff, &ctx->d_fcnt, &ctx->t_desc, &ctx->t_foff, &ctx->t_fcnt, &ctx->d_desc8, &ctx->t_desc8,
                      &ctx->d_pairs, &ctx->d_partial, &ctx->d_rev, &ctx->d_matches, &ctx->d_nmatch, &ctx->d_points,
                      &ctx->d_state, &ctx->d_Fall, &ctx->d_pc, &ctx->d_pr, &ctx->d_mask, &ctx->d_valid, &ctx->d_tri, &ctx->d_items, &ctx->d_item_total,
                      &ctx->d_opts, &ctx->d_oidx, &ctx->d_results, &ctx->d_table, &ctx->d_in1, &ctx->d_in2,
                      &ctx->d_knn_i, &ctx->d_knn_d, &ctx->d_counts, &ctx->d_pres, &ctx->o_tabs, &ctx->o_stage,
                      &ctx->o_pyr, &ctx->o_blur, &ctx->o_cxy, &ctx->o_cval, &ctx->o_cnt, &ctx->o_kidx, &ctx->o_kcnt,
                      &ctx->o_off, &ctx->o_kp, &ctx->o_desc, &ctx->p_world, &ctx->p_image, &ctx->p_off, &ctx->p_table,
                      &ctx->p_poses, &ctx->p_valid, &ctx->p_pc, &ctx->p_maskws, &ctx->p_mask, &ctx->p_counts, &ctx->p_results,
                      &ctx->b_foff, &ctx->b_poff, &ctx->b_R, &ctx->b_t, &ctx->b_pc, &ctx->b_X, &ctx->b_xc, &ctx->b_obs, &ctx->b_ooff,
                      &ctx->b_ws, &ctx->b_wsm, &ctx->b_wsm_off, &ctx->b_wsobs, &ctx->b_Ro, &ctx->b_to, &ctx->b_pco, &ctx->b_Xo, &ctx->b_xco, &ctx->b_res};
}

// Guard mode only (else -1): the number of workspace buffers whose guard band no longer holds the fill byte.
int mvs_debug_guard_check(mvs_ctx *ctx)
{
    if (!ctx || !guard_mode()) return -1;
    cudaSetDevice(ctx->device);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return -2;
    std::vector<uint8_t> h(kGuardBytes);
    int bad = 0;
    for (DevBuf *b : all_buffers(ctx)) {
        if (!b->p) continue;
        if (cudaMemcpy(h.data(), static_cast<uint8_t *>(b->p) + b->cap, kGuardBytes, cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
        for (uint8_t v : h)
            if (v != kGuardByte) { ++bad; break; }
    }
    return bad;
}

// Guard mode only: overwrite one byte just past the end of the first live buffer (the self-test of the check above).
int mvs_debug_guard_poke(mvs_ctx *ctx)
{
    if (!ctx || !guard_mode()) return -1;
    cudaSetDevice(ctx->device);
    for (DevBuf *b : all_buffers(ctx))
        if (b->p) return cudaMemset(static_cast<uint8_t *>(b->p) + b->cap, 0, 1) == cudaSuccess ? 0 : -2;
    return -3;
}

void mvs_destroy(mvs_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    const std::vector<DevBuf *> bufs = all_buffers(ctx);
    if (ctx->o_pinned) cudaFreeHost(ctx->o_pinned);
    if (ctx->h_counts) cudaFreeHost(ctx->h_counts);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->h_args) { cudaFreeHost(ctx->h_args); for (auto &e : ctx->arg_ev) if (e) cudaEventDestroy(e); }
    for (cudaEvent_t e : ctx->copy_done) if (e) cudaEventDestroy(e);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (DevBuf *b : bufs) b->release();
    ctx->l2.release();
    for (auto &p : ctx->pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *mvs_last_error(const mvs_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int mvs_set_stream(mvs_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return MVS_E_BAD_ARG;
    cudaStreamSynchronize(ctx->stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    ctx->own_stream = false;
    return MVS_OK;
}

int mvs_synchronize(mvs_ctx *ctx)
{
    if (!ctx) return MVS_E_BAD_ARG;
    CK(cudaStreamSynchronize(ctx->stream));
    flush_staged(ctx);
    return MVS_OK;
}

void *mvs_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    return p;
}

void mvs_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

int mvs_profile_enable(mvs_ctx *ctx, int on)
{
    if (!ctx) return MVS_E_BAD_ARG;
    ctx->prof = on != 0;
    return MVS_OK;
}

int mvs_profile_read(mvs_ctx *ctx, mvs_profile *out, int reset)
{
    if (!ctx || !out) return MVS_E_BAD_ARG;
    CK(cudaStreamSynchronize(ctx->stream));
    for (auto &p : ctx->pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) ctx->acc.ms[p.stage] += (double)ms;
        ctx->ev_pool.push_back(p.a); ctx->ev_pool.push_back(p.b);
    }
    ctx->pending.clear();
    *out = ctx->acc;
    if (reset) std::memset(&ctx->acc, 0, sizeof(ctx->acc));
    return MVS_OK;
}

uint64_t mvs_kernel_launches(const mvs_ctx *ctx) { return ctx ? ctx->launches : 0; }

void mvs_sample_table(uint64_t seed, uint64_t pair_id, uint32_t n_points, int H, uint32_t *out)
{
    for (int h = 0; h < H; ++h) {
        uint32_t row[8];
        sample_row(seed, pair_id, n_points, h, row);
        for (int j = 0; j < 8; ++j) out[(size_t)h * 8 + j] = row[j];
    }
}

// ------------------------------------------------------------------------------------------ matching
static int match_two_sets(mvs_ctx *ctx, const uint8_t *query, int nq, const uint8_t *train, int nt, int desc_bytes,
                          const mvs_match_params *mp, bool want_knn)
{
    if (!ctx) return MVS_E_BAD_ARG;
    if (!query || !train || nq < 1) return fail(ctx, MVS_E_BAD_ARG, "null descriptors or nq < 1");
    if (nt < 2) return fail(ctx, MVS_E_BAD_ARG, "knnMatch(k=2) needs at least 2 train descriptors");
    if (desc_bytes != 32) return fail(ctx, MVS_E_UNSUPPORTED, "only 256-bit (32-byte) descriptors are supported");
    if (nq > (int)kIdxMask || nt > (int)kIdxMask) return fail(ctx, MVS_E_UNSUPPORTED, "more than 2^22-1 descriptors per side");
    CK(cudaSetDevice(ctx->device));
    // scratch frame table: frame 0 = train, frame 1 = query
    CK(ctx->t_desc.ensure(((size_t)nq + nt) * 32));
    CK(ctx->t_foff.ensure(2 * sizeof(int32_t)));
    CK(ctx->t_fcnt.ensure(2 * sizeof(int32_t)));
    const int32_t off[2] = {0, nt}, cnt[2] = {nt, nq};
    CK(cudaMemcpyAsync(ctx->t_desc.p, train, (size_t)nt * 32, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->t_desc.as<uint8_t>() + (size_t)nt * 32, query, (size_t)nq * 32, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->t_foff.p, off, sizeof(off), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->t_fcnt.p, cnt, sizeof(cnt), cudaMemcpyHostToDevice, ctx->stream));
    const bool cross = mp && mp->cross_check;
    const bool tc = tc_eligible(ctx, nq, nt, cross, (size_t)nq + nt);
    const int splits = tc ? tc_splits(nq, nt, 1) : choose_splits((nq + 255) / 256, 1, nt);
    const int rsplits = cross ? (tc ? tc_splits(nt, nq, 1) : choose_splits((nt + 255) / 256, 1, nq)) : 0;
    CK(ctx->d_partial.ensure((size_t)splits * nq * sizeof(uint2)));
    if (cross) CK(ctx->d_rev.ensure((size_t)rsplits * nt * sizeof(uint2)));
    CK(ctx->d_matches.ensure((size_t)nq * sizeof(mvs_match)));
    CK(ctx->d_nmatch.ensure(sizeof(int32_t)));
    if (want_knn) { CK(ctx->d_knn_i.ensure((size_t)nq * 2 * sizeof(int32_t))); CK(ctx->d_knn_d.ensure((size_t)nq * 2 * sizeof(int32_t))); }
    KnnArgs ka{};
    ka.desc = ctx->t_desc.as<uint4>(); ka.frame_off = ctx->t_foff.as<int32_t>(); ka.frame_cnt = ctx->t_fcnt.as<int32_t>();
    ka.pairs = nullptr; ka.partial = ctx->d_partial.as<uint2>(); ka.q_stride = nq; ka.reverse = 0;
    ka.bound = (want_knn || tc) ? 0 : search_bound(mp);
    if (tc) {
        CK(ctx->t_desc8.ensure(((size_t)nq + nt) * 256));
        StageTimer t(ctx, MVS_STAGE_KNN, cross ? 3 : 2);
        launch_expand_desc(ctx->t_desc.as<uint4>(), 0, (size_t)nq + nt, ctx->t_desc8.p, ctx->stream);
        TcKnnArgs ta{};
        ta.frame_off = ka.frame_off; ta.frame_cnt = ka.frame_cnt; ta.pairs = nullptr; ta.partial = ka.partial; ta.q_stride = nq; ta.reverse = 0;
        ta.t_splits = tc_train_splits(nq, nt, 1);
        CK(launch_knn2_hamming_tc(ctx->t_desc8.p, (size_t)nq + nt, ta, nq, 1, ctx->stream));
        if (cross) {
            ta.partial = ctx->d_rev.as<uint2>(); ta.q_stride = nt; ta.reverse = 1; ta.t_splits = tc_train_splits(nt, nq, 1);
            CK(launch_knn2_hamming_tc(ctx->t_desc8.p, (size_t)nq + nt, ta, nt, 1, ctx->stream));
        }
    } else {
        StageTimer t(ctx, MVS_STAGE_KNN, cross ? 2 : 1);
        launch_knn2_hamming(ka, nq, splits, 1, ctx->stream);
        if (cross) {
            KnnArgs kr = ka;
            kr.partial = ctx->d_rev.as<uint2>(); kr.q_stride = nt; kr.reverse = 1;
            launch_knn2_hamming(kr, nt, rsplits, 1, ctx->stream);
        }
    }
    FinalizeArgs fa{};
    fa.frame_off = ka.frame_off; fa.frame_cnt = ka.frame_cnt; fa.pairs = nullptr;
    fa.partial = ka.partial; fa.splits = splits; fa.q_stride = nq;
    fa.rev_partial = cross ? ctx->d_rev.as<uint2>() : nullptr; fa.rev_splits = rsplits; fa.rev_stride = nt;
    fa.ratio = mp ? mp->ratio : 0.7; fa.max_dist = mp ? mp->max_dist : -1.0; fa.bound = ka.bound;
    fa.kp = nullptr; fa.matches = ctx->d_matches.as<mvs_match>(); fa.n_matches = ctx->d_nmatch.as<int32_t>();
    fa.points = nullptr; fa.state = nullptr;
    fa.knn_idx = want_knn ? ctx->d_knn_i.as<int32_t>() : nullptr; fa.knn_dist = want_knn ? ctx->d_knn_d.as<int32_t>() : nullptr;
    fa.refine_desc = tc ? ctx->t_desc.as<uint4>() : nullptr;
    {
        StageTimer t(ctx, MVS_STAGE_MATCH_FINALIZE);
        CK(launch_match_finalize(fa, nq, 1, ctx->stream));
    }
    CK(cudaGetLastError());
    return MVS_OK;
}

int mvs_knn2_hamming(mvs_ctx *ctx, const uint8_t *query, int nq, const uint8_t *train, int nt, int desc_bytes,
                     int32_t *idx, int32_t *dist)
{
    if (!idx || !dist) return fail(ctx, MVS_E_BAD_ARG, "null output");
    mvs_match_params mp{0.7, -1.0, 0, 0};
    int st = match_two_sets(ctx, query, nq, train, nt, desc_bytes, &mp, true);
    if (st != MVS_OK) return st;
    CK(cudaMemcpyAsync(idx, ctx->d_knn_i.p, (size_t)nq * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(dist, ctx->d_knn_d.p, (size_t)nq * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return MVS_OK;
}

int mvs_match_hamming(mvs_ctx *ctx, const uint8_t *query, int nq, const uint8_t *train, int nt, int desc_bytes,
                      const mvs_match_params *params, mvs_match *out, int capacity, int *n_out)
{
    if (!n_out) return fail(ctx, MVS_E_BAD_ARG, "null n_out");
    *n_out = 0;
    int st = match_two_sets(ctx, query, nq, train, nt, desc_bytes, params, false);
    if (st != MVS_OK) return st;
    int32_t m = 0;
    CK(cudaMemcpyAsync(&m, ctx->d_nmatch.p, sizeof(m), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (m > capacity) { *n_out = m; return fail(ctx, MVS_E_CAPACITY, "match capacity too small"); }
    if (m > 0) {
        if (!out) return fail(ctx, MVS_E_BAD_ARG, "null output");
        CK(cudaMemcpyAsync(out, ctx->d_matches.p, (size_t)m * sizeof(mvs_match), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    *n_out = m;
    return MVS_OK;
}

int mvs_knn2_l2(mvs_ctx *ctx, const float *query, int nq, const float *train, int nt, int dim, int32_t *idx, float *dist)
{
    if (!ctx) return MVS_E_BAD_ARG;
    if (!query || !train || !idx || !dist || nq < 1) return fail(ctx, MVS_E_BAD_ARG, "null argument or nq < 1");
    if (nt < 2) return fail(ctx, MVS_E_BAD_ARG, "knnMatch(k=2) needs at least 2 train descriptors");
    CK(cudaSetDevice(ctx->device));
    StageTimer t(ctx, MVS_STAGE_L2, 0);
    std::string err;
    int nl = 0;
    int st = mvs::l2_knn2(ctx->l2, ctx->stream, query, nq, train, nt, dim, idx, dist, nullptr, nullptr, 0, nullptr, &nl, err);
    ctx->launches += nl; ctx->acc.launches[MVS_STAGE_L2] += nl;
    if (st != MVS_OK) ctx->err = err;
    return st;
}

int mvs_match_l2(mvs_ctx *ctx, const float *query, int nq, const float *train, int nt, int dim,
                 const mvs_match_params *params, mvs_match *out, int capacity, int *n_out)
{
    if (!ctx) return MVS_E_BAD_ARG;
    if (!query || !train || !n_out || nq < 1) return fail(ctx, MVS_E_BAD_ARG, "null argument or nq < 1");
    if (nt < 2) return fail(ctx, MVS_E_BAD_ARG, "knnMatch(k=2) needs at least 2 train descriptors");
    CK(cudaSetDevice(ctx->device));
    StageTimer t(ctx, MVS_STAGE_L2, 0);
    std::string err;
    int nl = 0;
    mvs_match_params mp = params ? *params : mvs_match_params{0.7, -1.0, 0, 0};
    int st = mvs::l2_knn2(ctx->l2, ctx->stream, query, nq, train, nt, dim, nullptr, nullptr, &mp, out, capacity, n_out, &nl, err);
    ctx->launches += nl; ctx->acc.launches[MVS_STAGE_L2] += nl;
    if (st != MVS_OK) ctx->err = err;
    return st;
}

int mvs_l2_stats(const mvs_ctx *ctx, uint64_t out[4])
{
    if (!ctx || !out) return MVS_E_BAD_ARG;
    for (int i = 0; i < 4; ++i) out[i] = ctx->l2.stats[i];
    return MVS_OK;
}

// ------------------------------------------------------------------------------------------ geometry
int mvs_find_fundamental_matrix(mvs_ctx *ctx, const double *p1s, const double *p2s, int n_sets, int solver, double *F_out)
{
    if (ctx && solver != MVS_SOLVER_REFERENCE && solver != MVS_SOLVER_FAST) return fail(ctx, MVS_E_BAD_ARG, "bad solver");
    if (!ctx) return MVS_E_BAD_ARG;
    if (!p1s || !p2s || !F_out || n_sets < 1) return fail(ctx, MVS_E_BAD_ARG, "null argument or n_sets < 1");
    CK(cudaSetDevice(ctx->device));
    const size_t in_b = (size_t)n_sets * 24 * sizeof(double);
    CK(ctx->d_in1.ensure(in_b)); CK(ctx->d_in2.ensure(in_b));
    CK(ctx->d_Fall.ensure((size_t)n_sets * (9 + 48) * sizeof(double)));
    CK(cudaMemcpyAsync(ctx->d_in1.p, p1s, in_b, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_in2.p, p2s, in_b, cudaMemcpyHostToDevice, ctx->stream));
    {
        StageTimer t(ctx, MVS_STAGE_HYPOTHESES);
        launch_fundamental_sets(ctx->d_in1.as<double>(), ctx->d_in2.as<double>(), n_sets, ctx->d_Fall.as<double>(), solver, ctx->stream);
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(F_out, ctx->d_Fall.p, (size_t)n_sets * 9 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return MVS_OK;
}

int mvs_svd_batch(mvs_ctx *ctx, int n, const double *A, int count, int solver, double *U, double *w, double *Vt)
{
    if (!ctx) return MVS_E_BAD_ARG;
    if (!A || !U || !w || !Vt || count < 1) return fail(ctx, MVS_E_BAD_ARG, "null argument or count < 1");
    if (solver != MVS_SOLVER_REFERENCE && solver != MVS_SOLVER_FAST) return fail(ctx, MVS_E_BAD_ARG, "bad solver");
    if (!(n == 3 || ((n == 4 || n == 9) && solver == MVS_SOLVER_REFERENCE)))
        return fail(ctx, MVS_E_UNSUPPORTED, "n must be 3, or 4 / 9 with MVS_SOLVER_REFERENCE");
    CK(cudaSetDevice(ctx->device));
    const size_t mb = (size_t)count * n * n * sizeof(double), wb = (size_t)count * n * sizeof(double);
    CK(ctx->d_in1.ensure(mb)); CK(ctx->d_in2.ensure(mb)); CK(ctx->d_Fall.ensure(mb + wb));
    CK(cudaMemcpyAsync(ctx->d_in1.p, A, mb, cudaMemcpyHostToDevice, ctx->stream));
    double *dVt = ctx->d_Fall.as<double>(), *dw = dVt + (size_t)count * n * n;
    {
        StageTimer t(ctx, MVS_STAGE_HYPOTHESES);
        launch_svd_batch(n, ctx->d_in1.as<double>(), count, solver, ctx->d_in2.as<double>(), dw, dVt, ctx->stream);
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(U, ctx->d_in2.p, mb, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(Vt, dVt, mb, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(w, dw, wb, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return MVS_OK;
}

static int upload_table(mvs_ctx *ctx, const uint32_t *samples, int H, int n, const uint32_t **d_table)
{
    *d_table = nullptr;
    if (!samples) return MVS_OK;
    for (size_t i = 0; i < (size_t)H * 8; ++i)
        if (samples[i] >= (uint32_t)n) return fail(ctx, MVS_E_BAD_ARG, "sample index out of range");
    CK(ctx->d_table.ensure((size_t)H * 8 * sizeof(uint32_t)));
    CK(cudaMemcpyAsync(ctx->d_table.p, samples, (size_t)H * 8 * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    *d_table = ctx->d_table.as<uint32_t>();
    return MVS_OK;
}

static int init_state(mvs_ctx *ctx, int n)
{
    PairState st;
    std::memset(&st, 0, sizeof(st));
    st.status = n < 8 ? MVS_E_TOO_FEW_POINTS : MVS_OK;
    st.n_matches = n; st.best_h = -1; st.candidate = -1;
    CK(ctx->d_state.ensure(sizeof(PairState)));
    CK(cudaMemcpyAsync(ctx->d_state.p, &st, sizeof(st), cudaMemcpyHostToDevice, ctx->stream));
    return MVS_OK;
}

int mvs_ransac_fundamental(mvs_ctx *ctx, const double *p1, const double *p2, int n, const uint32_t *samples,
                           const mvs_ransac_params *params, double F[9], uint8_t *inlier_mask, int *inlier_count,
                           double *residual, int *best_hypothesis, int32_t *all_counts)
{
    if (!ctx) return MVS_E_BAD_ARG;
    if (!p1 || !p2 || n < 0) return fail(ctx, MVS_E_BAD_ARG, "null points");
    if (inlier_count) *inlier_count = 0;
    if (n < 8) return fail(ctx, MVS_E_TOO_FEW_POINTS, "fewer than 8 correspondences");  // estimator-RANSAC.cpp:25-29
    RansacCfg rc;
    int st = resolve_ransac(ctx, params, nullptr, rc);
    if (st != MVS_OK) return st;
    CK(cudaSetDevice(ctx->device));
    std::vector<double> inter((size_t)n * 6);
    bool unit_z = true;   // "constant z per image": the points come from K^-1 (u,v,1)
    const double zc1 = p1[2], zc2 = p2[2];
    for (int i = 0; i < n; ++i) {
        for (int k = 0; k < 3; ++k) { inter[(size_t)i * 6 + k] = p1[(size_t)i * 3 + k]; inter[(size_t)i * 6 + 3 + k] = p2[(size_t)i * 3 + k]; }
        unit_z = unit_z && p1[(size_t)i * 3 + 2] == zc1 && p2[(size_t)i * 3 + 2] == zc2;
    }
    CK(ctx->d_points.ensure(inter.size() * sizeof(double)));
    CK(cudaMemcpyAsync(ctx->d_points.p, inter.data(), inter.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    const uint32_t *d_table = nullptr;
    if ((st = upload_table(ctx, samples, rc.H, n, &d_table)) != MVS_OK) return st;
    if ((st = init_state(ctx, n)) != MVS_OK) return st;
    if ((st = run_geometry(ctx, 1, n, rc, unit_z, zc1, zc2, d_table, rc.pair_base, false, all_counts != nullptr, nullptr)) != MVS_OK) return st;
    PairState hs;
    CK(cudaMemcpyAsync(&hs, ctx->d_state.p, sizeof(hs), cudaMemcpyDeviceToHost, ctx->stream));
    if (inlier_mask) CK(cudaMemcpyAsync(inlier_mask, ctx->d_mask.p, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (all_counts) CK(cudaMemcpyAsync(all_counts, ctx->d_counts.p, (size_t)rc.H * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (F) std::memcpy(F, hs.F, sizeof(hs.F));
    if (inlier_count) *inlier_count = hs.n_inliers;
    if (residual) *residual = hs.residual;
    if (best_hypothesis) *best_hypothesis = hs.best_h;
    return hs.status;
}

int mvs_sfm_solve(mvs_ctx *ctx, const double *xy1, const double *xy2, int n, const double K[9],
                  const mvs_ransac_params *params, const uint32_t *samples, mvs_pair_result *result,
                  uint8_t *inlier_mask, double *points, uint64_t *indexes, int capacity)
{
    if (!ctx) return MVS_E_BAD_ARG;
    if (!xy1 || !xy2 || !K || !result || n < 0) return fail(ctx, MVS_E_BAD_ARG, "null argument");
    std::memset(result, 0, sizeof(*result));
    result->n_matches = n; result->best_hypothesis = -1; result->candidate = -1;
    if (n < 8) { result->status = MVS_E_TOO_FEW_POINTS; return fail(ctx, MVS_E_TOO_FEW_POINTS, "fewer than 8 correspondences"); }
    RansacCfg rc;
    int st = resolve_ransac(ctx, params, K, rc);
    if (st != MVS_OK) return st;
    CK(cudaSetDevice(ctx->device));
    NormArgs na;
    h_inverse3(K, na.Kinv);
    const bool unit_z = unit_z_intrinsics(na.Kinv);
    const size_t xb = (size_t)n * 2 * sizeof(double);
    CK(ctx->d_in1.ensure(xb)); CK(ctx->d_in2.ensure(xb));
    CK(ctx->d_points.ensure((size_t)n * 6 * sizeof(double)));
    CK(ctx->d_state.ensure(sizeof(PairState)));
    CK(cudaMemcpyAsync(ctx->d_in1.p, xy1, xb, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_in2.p, xy2, xb, cudaMemcpyHostToDevice, ctx->stream));
    const uint32_t *d_table = nullptr;
    if ((st = upload_table(ctx, samples, rc.H, n, &d_table)) != MVS_OK) return st;
    {
        StageTimer t(ctx, MVS_STAGE_MATCH_FINALIZE);
        launch_normalize_points(ctx->d_in1.as<double>(), ctx->d_in2.as<double>(), n, na, ctx->d_points.as<double>(),
                                ctx->d_state.as<PairState>(), ctx->stream);
    }
    if ((st = run_geometry(ctx, 1, n, rc, unit_z, na.Kinv[8], na.Kinv[8], d_table, rc.pair_base, true, false, nullptr)) != MVS_OK) return st;
    CK(cudaMemcpyAsync(result, ctx->d_results.p, sizeof(*result), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (inlier_mask && result->status != MVS_E_TOO_FEW_POINTS)
        CK(cudaMemcpyAsync(inlier_mask, ctx->d_mask.p, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (result->status == MVS_OK) {
        if (result->n_points > capacity) return fail(ctx, MVS_E_CAPACITY, "point capacity too small");
        if (points) CK(cudaMemcpyAsync(points, ctx->d_opts.p, (size_t)result->n_points * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (indexes) CK(cudaMemcpyAsync(indexes, ctx->d_oidx.p, (size_t)result->n_points * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return result->status;
}

int mvs_sfm_triangulate(mvs_ctx *ctx, const double *xy1, const double *xy2, int n, const double K[9],
                        const double R1[9], const double t1[3], const double R2[9], const double t2[3], int solver,
                        double *points, uint64_t *indexes, int capacity, int *n_out)
{
    if (!ctx) return MVS_E_BAD_ARG;
    if (solver != MVS_SOLVER_REFERENCE && solver != MVS_SOLVER_FAST) return fail(ctx, MVS_E_BAD_ARG, "bad solver");
    if (!xy1 || !xy2 || !K || !R1 || !t1 || !R2 || !t2 || !n_out || n < 0) return fail(ctx, MVS_E_BAD_ARG, "null argument");
    *n_out = 0;
    if (n == 0) return MVS_OK;
    CK(cudaSetDevice(ctx->device));
    // T_1_to_2 = pose2.inverse() * pose1 (sfm-solve.cpp:382)
    PairState st;
    std::memset(&st, 0, sizeof(st));
    double Ri[9], ti[3];
    h_se3_inverse(R2, t2, Ri, ti);
    h_se3_compose(Ri, ti, R1, t1, st.Rc[0], st.tc);
    st.status = MVS_OK; st.n_matches = n; st.best_h = -1; st.candidate = -1;
    NormArgs na;
    h_inverse3(K, na.Kinv);
    const size_t xb = (size_t)n * 2 * sizeof(double);
    CK(ctx->d_in1.ensure(xb)); CK(ctx->d_in2.ensure(xb));
    CK(ctx->d_points.ensure((size_t)n * 6 * sizeof(double)));
    CK(ctx->d_state.ensure(sizeof(PairState)));
    CK(ctx->d_valid.ensure((size_t)4 * n));
    CK(ctx->d_tri.ensure((size_t)4 * n * 3 * sizeof(double)));
    CK(ctx->d_opts.ensure((size_t)n * 3 * sizeof(double)));
    CK(ctx->d_oidx.ensure((size_t)n * sizeof(uint64_t)));
    CK(ctx->d_results.ensure(sizeof(mvs_pair_result)));
    CK(cudaMemcpyAsync(ctx->d_in1.p, xy1, xb, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_in2.p, xy2, xb, cudaMemcpyHostToDevice, ctx->stream));
    {
        StageTimer t(ctx, MVS_STAGE_MATCH_FINALIZE);
        launch_normalize_points(ctx->d_in1.as<double>(), ctx->d_in2.as<double>(), n, na, ctx->d_points.as<double>(), nullptr, ctx->stream);
    }
    CK(cudaMemcpyAsync(ctx->d_state.p, &st, sizeof(st), cudaMemcpyHostToDevice, ctx->stream));
    {
        StageTimer t(ctx, MVS_STAGE_TRIANGULATE);
        TriArgs a{};
        a.points = ctx->d_points.as<double>(); a.p_stride = n; a.state = ctx->d_state.as<PairState>(); a.mask = nullptr;
        a.n_cand = 1; a.valid = ctx->d_valid.as<uint8_t>(); a.tri = ctx->d_tri.as<double>(); a.solver = solver;
        launch_triangulate(a, n, 1, ctx->stream);
    }
    {
        StageTimer t(ctx, MVS_STAGE_FINALIZE);
        FinishArgs f{};
        f.state = ctx->d_state.as<PairState>(); f.p_stride = n; f.n_cand = 1; f.valid = ctx->d_valid.as<uint8_t>();
        f.tri = ctx->d_tri.as<double>(); f.matches = nullptr; f.out_points = ctx->d_opts.as<double>();
        f.out_index = ctx->d_oidx.as<uint64_t>(); f.results = ctx->d_results.as<mvs_pair_result>();
        launch_finish(f, n, 1, ctx->stream);
    }
    CK(cudaGetLastError());
    mvs_pair_result r;
    CK(cudaMemcpyAsync(&r, ctx->d_results.p, sizeof(r), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const int m = r.n_points;
    if (m > capacity) { *n_out = m; return fail(ctx, MVS_E_CAPACITY, "point capacity too small"); }
    if (m > 0) {
        if (points) CK(cudaMemcpyAsync(points, ctx->d_opts.p, (size_t)m * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (indexes) CK(cudaMemcpyAsync(indexes, ctx->d_oidx.p, (size_t)m * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    *n_out = m;
    return MVS_OK;
}

// ------------------------------------------------------------------------------------------ batch
int mvs_frames_upload(mvs_ctx *ctx, int n_frames, const uint8_t *const *desc, const float *const *kp,
                      const int32_t *counts, int desc_bytes)
{
    if (!ctx) return MVS_E_BAD_ARG;
    if (n_frames < 1 || !desc || !kp || !counts) return fail(ctx, MVS_E_BAD_ARG, "null argument or n_frames < 1");
    if (desc_bytes != 32) return fail(ctx, MVS_E_UNSUPPORTED, "only 256-bit (32-byte) descriptors are supported");
    CK(cudaSetDevice(ctx->device));
    std::vector<int32_t> off(n_frames), cnt(n_frames);
    size_t total = 0;
    for (int f = 0; f < n_frames; ++f) {
        if (counts[f] < 0 || counts[f] > (int)kIdxMask) return fail(ctx, MVS_E_BAD_ARG, "bad keypoint count");
        if (counts[f] > 0 && (!desc[f] || !kp[f])) return fail(ctx, MVS_E_BAD_ARG, "null frame buffer");
        off[f] = (int32_t)total; cnt[f] = counts[f];
        total += (size_t)counts[f];
        if (total > 0x7FFFFFFFull) return fail(ctx, MVS_E_UNSUPPORTED, "more than 2^31 keypoints in the frame table");
    }
    CK(ctx->d_desc.ensure(std::max<size_t>(total, 1) * 32));
    CK(ctx->d_kp.ensure(std::max<size_t>(total, 1) * sizeof(float2)));
    CK(ctx->d_foff.ensure((size_t)n_frames * sizeof(int32_t)));
    CK(ctx->d_fcnt.ensure((size_t)n_frames * sizeof(int32_t)));
    // every frame in pinned, suitably aligned host memory (and a window small enough for a few launches): gathered by a kernel
    bool gather = n_frames <= 16 * kGatherFrames;
    std::vector<const void *> dev_desc(gather ? n_frames : 0), dev_kp(gather ? n_frames : 0);
    for (int f = 0; gather && f < n_frames; ++f) {
        if (!cnt[f]) continue;
        dev_desc[f] = pinned_device_ptr(desc[f]); dev_kp[f] = pinned_device_ptr(kp[f]);
        gather = dev_desc[f] && dev_kp[f] && !((uintptr_t)dev_desc[f] & 15u) && !((uintptr_t)dev_kp[f] & 7u);
    }
    if (gather) {
        for (int f0 = 0; f0 < n_frames; f0 += kGatherFrames) {
            const int nf = std::min(kGatherFrames, n_frames - f0);
            GatherArgs ga{};
            int most = 1;
            for (int f = 0; f < nf; ++f) {
                ga.desc[f] = static_cast<const uint4 *>(dev_desc[f0 + f]); ga.kp[f] = static_cast<const float2 *>(dev_kp[f0 + f]);
                ga.off[f] = off[f0 + f]; ga.cnt[f] = cnt[f0 + f];
                most = std::max(most, cnt[f0 + f]);
            }
            ga.d_desc = ctx->d_desc.as<uint4>(); ga.d_kp = ctx->d_kp.as<float2>();
            ga.d_foff = ctx->d_foff.as<int32_t>(); ga.d_fcnt = ctx->d_fcnt.as<int32_t>(); ga.f0 = f0;
            const unsigned bx = (unsigned)std::max(1, std::min(16, (2 * most + 511) / 512));     // two 16-byte loads per thread
            gather_frames_kernel<<<dim3(bx, (unsigned)nf), 256, 0, ctx->stream>>>(ga);
            ctx->launches += 1;
        }
        CK(cudaGetLastError());
    } else {
        for (int f = 0; f < n_frames; ++f) {
            if (!cnt[f]) continue;
            CK(cudaMemcpyAsync(ctx->d_desc.as<uint8_t>() + (size_t)off[f] * 32, desc[f], (size_t)cnt[f] * 32, cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaMemcpyAsync(ctx->d_kp.as<float2>() + off[f], kp[f], (size_t)cnt[f] * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
        }
        CK(cudaMemcpyAsync(ctx->d_foff.p, off.data(), (size_t)n_frames * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_fcnt.p, cnt.data(), (size_t)n_frames * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));  // the caller may reuse its buffers on return; off/cnt are stack/vector memory
    ctx->h_off.swap(off); ctx->h_cnt.swap(cnt);
    ctx->desc8_rows = 0;
    return MVS_OK;
}

int mvs_frames_upload_packed(mvs_ctx *ctx, int n_frames, const uint8_t *desc_all, const float *kp_all, const int32_t *counts,
                             int desc_bytes)
{
    if (!ctx) return MVS_E_BAD_ARG;
    if (n_frames < 1 || !desc_all || !kp_all || !counts) return fail(ctx, MVS_E_BAD_ARG, "null argument or n_frames < 1");
    if (desc_bytes != 32) return fail(ctx, MVS_E_UNSUPPORTED, "only 256-bit (32-byte) descriptors are supported");
    CK(cudaSetDevice(ctx->device));
    std::vector<int32_t> off(n_frames), cnt(n_frames);
    size_t total = 0;
    for (int f = 0; f < n_frames; ++f) {
        if (counts[f] < 0 || counts[f] > (int)kIdxMask) return fail(ctx, MVS_E_BAD_ARG, "bad keypoint count");
        off[f] = (int32_t)total; cnt[f] = counts[f];
        total += (size_t)counts[f];
        if (total > 0x7FFFFFFFull) return fail(ctx, MVS_E_UNSUPPORTED, "more than 2^31 keypoints in the frame table");
    }
    CK(ctx->d_desc.ensure(std::max<size_t>(total, 1) * 32));
    CK(ctx->d_kp.ensure(std::max<size_t>(total, 1) * sizeof(float2)));
    CK(ctx->d_foff.ensure((size_t)n_frames * sizeof(int32_t)));
    CK(ctx->d_fcnt.ensure((size_t)n_frames * sizeof(int32_t)));
    // the offset / count tables are copied from the ctx's own vectors, which live until the next upload: no need to wait
    ctx->h_off.swap(off); ctx->h_cnt.swap(cnt);
    if (total) {
        CK(cudaMemcpyAsync(ctx->d_desc.p, desc_all, total * 32, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_kp.p, kp_all, total * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
    }
    CK(h2d_args(ctx, ctx->d_foff.p, ctx->h_off.data(), (size_t)n_frames * sizeof(int32_t)));
    CK(h2d_args(ctx, ctx->d_fcnt.p, ctx->h_cnt.data(), (size_t)n_frames * sizeof(int32_t)));
    ctx->desc8_rows = 0;
    return MVS_OK;
}

static int pair_batch_chunk(mvs_ctx *ctx, const int32_t *pairs, int n_pairs, const double K[9],
                            const mvs_match_params *mparams, const mvs_ransac_params *rparams,
                            mvs_pair_result *results, mvs_match *matches, uint8_t *inlier_mask,
                            double *points, uint64_t *indexes, int capacity)
{
    if (ctx->h_cnt.empty()) return fail(ctx, MVS_E_BAD_ARG, "no frames uploaded (mvs_frames_upload)");
    const int nf = (int)ctx->h_cnt.size();
    int max_nq = 0, max_nt = 0;
    for (int i = 0; i < n_pairs; ++i) {
        const int a = pairs[2 * i], b = pairs[2 * i + 1];
        if (a < 0 || a >= nf || b < 0 || b >= nf || a == b) return fail(ctx, MVS_E_BAD_ARG, "bad frame index in pairs");  // image-pair.cpp:49
        // a frame with fewer than 2 (base) / 1 (pair) keypoints (a dark image in a window) gives that pair no matches:
        // it comes back with MVS_E_TOO_FEW_POINTS and n_matches = 0; the other pairs of the batch are unaffected
        max_nq = std::max(max_nq, ctx->h_cnt[b]); max_nt = std::max(max_nt, ctx->h_cnt[a]);
    }
    max_nq = std::max(max_nq, 1); max_nt = std::max(max_nt, 1);
    const bool details = matches || inlier_mask || points || indexes;
    if (details && capacity < 1) return fail(ctx, MVS_E_CAPACITY, "detail capacity must be >= 1");
    if ((size_t)finalize_sort_capacity(max_nq) * sizeof(uint32_t) > 200 * 1024)
        return fail(ctx, MVS_E_UNSUPPORTED, "more than 32768 keypoints in a pair frame");
    RansacCfg rc;
    int st = resolve_ransac(ctx, rparams, K, rc);
    if (st != MVS_OK) return st;
    CK(cudaSetDevice(ctx->device));
    const int qs = max_nq;
    ctx->last_stride = qs;
    const bool cross = mparams && mparams->cross_check;
    const size_t table_rows = (size_t)ctx->h_off[nf - 1] + (size_t)ctx->h_cnt[nf - 1];
    const bool tc = tc_eligible(ctx, max_nq, max_nt, cross, table_rows);
    const int splits = tc ? tc_splits(max_nq, max_nt, n_pairs) : choose_splits((max_nq + 255) / 256, n_pairs, max_nt);
    const int rsplits = cross ? (tc ? tc_splits(max_nt, max_nq, n_pairs) : choose_splits((max_nt + 255) / 256, n_pairs, max_nq)) : 0;
    CK(ctx->d_pairs.ensure((size_t)n_pairs * sizeof(int2)));
    CK(ctx->d_partial.ensure((size_t)n_pairs * splits * qs * sizeof(uint2)));
    if (cross) CK(ctx->d_rev.ensure((size_t)n_pairs * rsplits * max_nt * sizeof(uint2)));
    CK(ctx->d_matches.ensure((size_t)n_pairs * qs * sizeof(mvs_match)));
    CK(ctx->d_nmatch.ensure((size_t)n_pairs * sizeof(int32_t)));
    CK(ctx->d_points.ensure((size_t)n_pairs * qs * 6 * sizeof(double)));
    CK(ctx->d_state.ensure((size_t)n_pairs * sizeof(PairState)));
    CK(h2d_args(ctx, ctx->d_pairs.p, pairs, (size_t)n_pairs * sizeof(int2)));

    KnnArgs ka{};
    ka.desc = ctx->d_desc.as<uint4>(); ka.frame_off = ctx->d_foff.as<int32_t>(); ka.frame_cnt = ctx->d_fcnt.as<int32_t>();
    ka.pairs = ctx->d_pairs.as<int2>(); ka.partial = ctx->d_partial.as<uint2>(); ka.q_stride = qs; ka.reverse = 0;
    ka.bound = tc ? 0 : search_bound(mparams);     // the tensor-core kernel evaluates every pair: nothing to bound
    if (tc) {
        StageTimer t(ctx, MVS_STAGE_KNN, cross ? 2 : 1);
        if ((st = sync_desc8(ctx, table_rows)) != MVS_OK) return st;
        TcKnnArgs ta{};
        ta.frame_off = ka.frame_off; ta.frame_cnt = ka.frame_cnt; ta.pairs = ka.pairs; ta.partial = ka.partial; ta.q_stride = qs; ta.reverse = 0;
        ta.t_splits = tc_train_splits(max_nq, max_nt, n_pairs);
        CK(launch_knn2_hamming_tc(ctx->d_desc8.p, table_rows, ta, max_nq, n_pairs, ctx->stream));
        if (cross) {
            ta.partial = ctx->d_rev.as<uint2>(); ta.q_stride = max_nt; ta.reverse = 1; ta.t_splits = tc_train_splits(max_nt, max_nq, n_pairs);
            CK(launch_knn2_hamming_tc(ctx->d_desc8.p, table_rows, ta, max_nt, n_pairs, ctx->stream));
        }
    } else {
        StageTimer t(ctx, MVS_STAGE_KNN, cross ? 2 : 1);
        launch_knn2_hamming(ka, max_nq, splits, n_pairs, ctx->stream);
        if (cross) {
            KnnArgs kr = ka;
            kr.partial = ctx->d_rev.as<uint2>(); kr.q_stride = max_nt; kr.reverse = 1;
            launch_knn2_hamming(kr, max_nt, rsplits, n_pairs, ctx->stream);
        }
    }
    FinalizeArgs fa{};
    fa.frame_off = ka.frame_off; fa.frame_cnt = ka.frame_cnt; fa.pairs = ka.pairs;
    fa.partial = ka.partial; fa.splits = splits; fa.q_stride = qs;
    fa.rev_partial = cross ? ctx->d_rev.as<uint2>() : nullptr; fa.rev_splits = rsplits; fa.rev_stride = max_nt;
    fa.ratio = mparams ? mparams->ratio : 0.7; fa.max_dist = mparams ? mparams->max_dist : -1.0; fa.bound = ka.bound;
    fa.kp = ctx->d_kp.as<float2>();
    h_inverse3(K, fa.Kinv);
    const bool unit_z = unit_z_intrinsics(fa.Kinv);
    fa.matches = ctx->d_matches.as<mvs_match>(); fa.n_matches = ctx->d_nmatch.as<int32_t>();
    fa.points = ctx->d_points.as<double>(); fa.state = ctx->d_state.as<PairState>();
    fa.knn_idx = nullptr; fa.knn_dist = nullptr;
    fa.refine_desc = tc ? ctx->d_desc.as<uint4>() : nullptr;
    {
        StageTimer t(ctx, MVS_STAGE_MATCH_FINALIZE);
        CK(launch_match_finalize(fa, max_nq, n_pairs, ctx->stream));
    }
    if ((st = run_geometry(ctx, n_pairs, qs, rc, unit_z, fa.Kinv[8], fa.Kinv[8], nullptr, rc.pair_base, true, false, ctx->d_matches.as<mvs_match>())) != MVS_OK) return st;

    if (ctx->skip_d2h) return MVS_OK;
    // Small batches whose outputs go to pageable memory are staged through pinned memory: one synchronisation for all
    // copies instead of one blocking copy each (a single VO pair drops from ~280 us to ~170 us per call).
    if (details) {
        // every requested detail buffer pinned: the device writes the used entries itself
        ExportArgs ea{};
        ea.mo = static_cast<mvs_match *>(pinned_device_ptr(matches)); ea.ko = static_cast<uint8_t *>(pinned_device_ptr(inlier_mask));
        ea.po = static_cast<double *>(pinned_device_ptr(points)); ea.io = static_cast<uint64_t *>(pinned_device_ptr(indexes));
        if ((!matches || ea.mo) && (!inlier_mask || ea.ko) && (!points || ea.po) && (!indexes || ea.io)) {
            ea.res = ctx->d_results.as<mvs_pair_result>(); ea.stride = qs; ea.capacity = capacity;
            ea.m = ctx->d_matches.as<mvs_match>(); ea.k = ctx->d_mask.as<uint8_t>();
            ea.p = ctx->d_opts.as<double>(); ea.i = ctx->d_oidx.as<uint64_t>();
            CK(mvs::launch_dep(export_details_kernel, dim3((unsigned)n_pairs), dim3(128), 0, ctx->stream, ea));
            ctx->launches += 1;
            CK(cudaMemcpyAsync(results, ctx->d_results.p, (size_t)n_pairs * sizeof(mvs_pair_result), cudaMemcpyDeviceToHost, ctx->stream));
            return MVS_OK;
        }
    }
    const size_t w = (size_t)std::min(capacity, qs);
    const size_t detail_bytes = (size_t)n_pairs * w * ((matches ? sizeof(mvs_match) : 0) + (inlier_mask ? 1 : 0) + (points ? 24 : 0) + (indexes ? 8 : 0));
    const size_t stage_bytes = detail_bytes + (size_t)n_pairs * sizeof(mvs_pair_result);
    bool stage = ctx->allow_stage && stage_bytes <= ((size_t)4 << 20);
    if (stage) {
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, results) == cudaSuccess && at.type == cudaMemoryTypeHost) stage = false;   // pinned already
        else (void)cudaGetLastError();
    }
    if (stage && !ctx->h_stage) {
        if (cudaMallocHost((void **)&ctx->h_stage, (size_t)8 << 20) == cudaSuccess) ctx->h_stage_cap = (size_t)8 << 20;
        else { (void)cudaGetLastError(); stage = false; }
    }
    if (stage && details) {
        // the device writes the used entries of every pair into the staging area (one kernel, no count round trip, no strided
        // copies); flush_staged() hands exactly those entries to the caller's buffers after the one synchronisation
        auto al = [](size_t x) { return (x + 15) & ~(size_t)15; };
        mvs_ctx::StagedExport e{results, matches, inlier_mask, points, indexes, n_pairs, (int)w, capacity, 0, 0, 0, 0, 0};
        size_t off = al(ctx->h_stage_used);
        e.off_res = off; off = al(off + (size_t)n_pairs * sizeof(mvs_pair_result));
        e.off_m = off; if (matches) off = al(off + (size_t)n_pairs * w * sizeof(mvs_match));
        e.off_k = off; if (inlier_mask) off = al(off + (size_t)n_pairs * w);
        e.off_p = off; if (points) off = al(off + (size_t)n_pairs * w * 24);
        e.off_i = off; if (indexes) off = al(off + (size_t)n_pairs * w * 8);
        void *stage_dev = pinned_device_ptr(ctx->h_stage);
        if (off <= ctx->h_stage_cap && stage_dev) {
            uint8_t *sd = static_cast<uint8_t *>(stage_dev);
            ExportArgs ea{};
            ea.res = ctx->d_results.as<mvs_pair_result>(); ea.stride = qs; ea.capacity = (int)w;
            ea.m = ctx->d_matches.as<mvs_match>(); ea.k = ctx->d_mask.as<uint8_t>();
            ea.p = ctx->d_opts.as<double>(); ea.i = ctx->d_oidx.as<uint64_t>();
            ea.mo = matches ? reinterpret_cast<mvs_match *>(sd + e.off_m) : nullptr; ea.ko = inlier_mask ? sd + e.off_k : nullptr;
            ea.po = points ? reinterpret_cast<double *>(sd + e.off_p) : nullptr; ea.io = indexes ? reinterpret_cast<uint64_t *>(sd + e.off_i) : nullptr;
            CK(mvs::launch_dep(export_details_kernel, dim3((unsigned)n_pairs), dim3(128), 0, ctx->stream, ea));
            ctx->launches += 1;
            CK(cudaMemcpyAsync(ctx->h_stage + e.off_res, ctx->d_results.p, (size_t)n_pairs * sizeof(mvs_pair_result), cudaMemcpyDeviceToHost, ctx->stream));
            ctx->staged_exports.push_back(e);
            ctx->h_stage_used = off;
            return MVS_OK;
        }
    }
    CK(d2h_rows(ctx, stage, results, sizeof(mvs_pair_result), ctx->d_results.p, sizeof(mvs_pair_result), sizeof(mvs_pair_result), (size_t)n_pairs));
    // details: the first min(capacity, stride) entries of every pair (a pair with n_matches > capacity is truncated).
    // A synchronous call with more than 1 MB of details (128 KB when they are staged and scattered by the host as well)
    // first reads the match counts back (one small round trip after the kernels) and copies only as many entries per pair
    // as the fullest pair holds: with the VO threshold a pair keeps ~100 of its slots, and the device-to-host copy is a
    // fifth of an end-to-end step (ten VO pairs: 0.7 MB of slots, 100 us of a 280 us call).
    size_t wc = w;
    if (ctx->allow_stage && detail_bytes > (stage ? (size_t)128 << 10 : (size_t)1 << 20)) {
        if (ctx->h_counts_cap < (size_t)n_pairs) {
            if (ctx->h_counts) cudaFreeHost(ctx->h_counts);
            ctx->h_counts = nullptr; ctx->h_counts_cap = 0;
            CK(cudaMallocHost((void **)&ctx->h_counts, (size_t)n_pairs * sizeof(int32_t)));
            ctx->h_counts_cap = (size_t)n_pairs;
        }
        CK(cudaMemcpyAsync(ctx->h_counts, ctx->d_nmatch.p, (size_t)n_pairs * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        int32_t most = 1;
        for (int i = 0; i < n_pairs; ++i) most = std::max(most, ctx->h_counts[i]);
        wc = std::min(w, (size_t)most);       // inliers and points are subsets of the matches
    }
    if (matches) CK(d2h_rows(ctx, stage, matches, (size_t)capacity * sizeof(mvs_match), ctx->d_matches.p, (size_t)qs * sizeof(mvs_match), wc * sizeof(mvs_match), (size_t)n_pairs));
    if (inlier_mask) CK(d2h_rows(ctx, stage, inlier_mask, (size_t)capacity, ctx->d_mask.p, (size_t)qs, wc, (size_t)n_pairs));
    if (points) CK(d2h_rows(ctx, stage, points, (size_t)capacity * 24, ctx->d_opts.p, (size_t)qs * 24, wc * 24, (size_t)n_pairs));
    if (indexes) CK(d2h_rows(ctx, stage, indexes, (size_t)capacity * 8, ctx->d_oidx.p, (size_t)qs * 8, wc * 8, (size_t)n_pairs));
    return MVS_OK;
}

// one chunk (<= 8192 pairs), every output left in the ctx's device workspace (read back through mvs_ctx_last_outputs)
int mvs_pair_batch_device_only(mvs_ctx *ctx, const int32_t *pairs, int n_pairs, const double K[9], const mvs_match_params *mparams,
                               const mvs_ransac_params *rparams)
{
    if (!ctx || !pairs || n_pairs < 1 || n_pairs > 8192 || !K) return MVS_E_BAD_ARG;
    ctx->skip_d2h = true;
    mvs_pair_result dummy;
    const int st = pair_batch_chunk(ctx, pairs, n_pairs, K, mparams, rparams, &dummy, nullptr, nullptr, nullptr, nullptr, 0);
    ctx->skip_d2h = false;
    return st;
}

// grow a device buffer geometrically, keeping the first n_used elements (device-to-device copy on the ctx stream)
static cudaError_t grow_keep(mvs_ctx *ctx, DevBuf &b, size_t elem, size_t n_used, size_t n_need)
{
    if (n_need * elem <= b.cap) return cudaSuccess;
    DevBuf nb;
    cudaError_t e = nb.ensure(std::max(n_need * 2, (size_t)4096) * elem);
    if (e != cudaSuccess) return e;
    if (n_used) e = cudaMemcpyAsync(nb.p, b.p, n_used * elem, cudaMemcpyDeviceToDevice, ctx->stream);
    if (e != cudaSuccess) { nb.release(); return e; }
    cudaStreamSynchronize(ctx->stream);
    b.release();
    b = nb;
    return cudaSuccess;
}

int mvs_frames_clear(mvs_ctx *ctx)
{
    if (!ctx) return MVS_E_BAD_ARG;
    ctx->h_off.clear(); ctx->h_cnt.clear();
    ctx->desc8_rows = 0;
    return MVS_OK;
}

int mvs_frames_append(mvs_ctx *ctx, const uint8_t *desc, const float *kp, int32_t count, int desc_bytes, int32_t *frame_index)
{
    if (!ctx) return MVS_E_BAD_ARG;
    if (count < 0 || count > (int)kIdxMask || (count > 0 && (!desc || !kp))) return fail(ctx, MVS_E_BAD_ARG, "bad frame");
    if (desc_bytes != 32) return fail(ctx, MVS_E_UNSUPPORTED, "only 256-bit (32-byte) descriptors are supported");
    CK(cudaSetDevice(ctx->device));
    const int nf = (int)ctx->h_cnt.size();
    const size_t used = nf ? (size_t)ctx->h_off[nf - 1] + (size_t)ctx->h_cnt[nf - 1] : 0;
    const size_t total = used + (size_t)count;
    if (total > 0x7FFFFFFFull) return fail(ctx, MVS_E_UNSUPPORTED, "more than 2^31 keypoints in the frame table");
    CK(grow_keep(ctx, ctx->d_desc, 32, used, std::max<size_t>(total, 1)));
    CK(grow_keep(ctx, ctx->d_kp, sizeof(float2), used, std::max<size_t>(total, 1)));
    CK(grow_keep(ctx, ctx->d_foff, sizeof(int32_t), (size_t)nf, (size_t)nf + 1));
    CK(grow_keep(ctx, ctx->d_fcnt, sizeof(int32_t), (size_t)nf, (size_t)nf + 1));
    const int32_t off = (int32_t)used;
    if (count) {
        CK(cudaMemcpyAsync(ctx->d_desc.as<uint8_t>() + used * 32, desc, (size_t)count * 32, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_kp.as<float2>() + used, kp, (size_t)count * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
    }
    CK(cudaMemcpyAsync(ctx->d_foff.as<int32_t>() + nf, &off, sizeof(off), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_fcnt.as<int32_t>() + nf, &count, sizeof(count), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->h_off.push_back(off); ctx->h_cnt.push_back(count);
    if (frame_index) *frame_index = nf;
    return MVS_OK;
}

int mvs_pair_batch_enqueue(mvs_ctx *ctx, const int32_t *pairs, int n_pairs, const double K[9],
                           const mvs_match_params *mparams, const mvs_ransac_params *rparams,
                           mvs_pair_result *results, mvs_match *matches, uint8_t *inlier_mask,
                           double *points, uint64_t *indexes, int capacity)
{
    if (!ctx) return MVS_E_BAD_ARG;
    if (!pairs || n_pairs < 1 || !K || !results) return fail(ctx, MVS_E_BAD_ARG, "null argument or n_pairs < 1");
    // grid.z carries the pair index (<= 65535) and the workspace scales with the batch: large batches run
    // as stream-ordered chunks; sampling stays invariant through pair_id_base
    constexpr int kChunk = 8192;
    for (int c0 = 0; c0 < n_pairs; c0 += kChunk) {
        const int n = std::min(kChunk, n_pairs - c0);
        mvs_ransac_params rp = rparams ? *rparams : mvs_ransac_params{1, MVS_SCORE_ALGEBRAIC, 0.0, 0, 0, 0, 0};
        rp.pair_id_base += (uint64_t)c0;
        const size_t o = (size_t)c0 * (size_t)std::max(capacity, 0);
        int st = pair_batch_chunk(ctx, pairs + 2 * (size_t)c0, n, K, mparams, &rp, results + c0,
                                  matches ? matches + o : nullptr, inlier_mask ? inlier_mask + o : nullptr,
                                  points ? points + 3 * o : nullptr, indexes ? indexes + o : nullptr, capacity);
        if (st != MVS_OK) return st;
    }
    return MVS_OK;
}

int mvs_pair_batch(mvs_ctx *ctx, const int32_t *pairs, int n_pairs, const double K[9],
                   const mvs_match_params *mparams, const mvs_ransac_params *rparams,
                   mvs_pair_result *results, mvs_match *matches, uint8_t *inlier_mask,
                   double *points, uint64_t *indexes, int capacity)
{
    if (!ctx) return MVS_E_BAD_ARG;
    ctx->allow_stage = true;
    int st = mvs_pair_batch_enqueue(ctx, pairs, n_pairs, K, mparams, rparams, results, matches, inlier_mask, points, indexes, capacity);
    ctx->allow_stage = false;
    if (st != MVS_OK) { cudaStreamSynchronize(ctx->stream); flush_staged(ctx); return st; }
    CK(cudaStreamSynchronize(ctx->stream));
    flush_staged(ctx);
    return MVS_OK;
}

// ------------------------------------------------------------------------------------------ feature extraction
// VisualFeature::extract (visual-feature.cpp:40-49) for a batch of same-size images.  Two device passes per chunk of
// images with one small device->host read between them (the per-level keypoint counts size the compact outputs).
static int orb_extract_impl(mvs_ctx *ctx, const uint8_t *const *h_images, const uint8_t *d_images, int n_images, int width,
                            int height, int stride, const mvs_orb_params *params, int append_frames, int32_t *first_frame,
                            int32_t *counts, mvs_keypoint *keypoints, uint8_t *descriptors, int64_t capacity)
{
    if (!ctx) return MVS_E_BAD_ARG;
    if (n_images < 1 || (!h_images && !d_images) || width < 1 || height < 1 || stride < width)
        return fail(ctx, MVS_E_BAD_ARG, "orb_extract: null images, n_images < 1 or bad size/stride");
    const int nf = params ? params->n_features : 500;   // visual-feature.cpp:9 MAX_FEATURE_COUNT
    if (nf < 0) return fail(ctx, MVS_E_BAD_ARG, "orb_extract: n_features < 0");
    if (width > 65535 || height > 65535) return fail(ctx, MVS_E_UNSUPPORTED, "orb_extract: image side > 65535");
    if ((keypoints || descriptors) && capacity < 0) return fail(ctx, MVS_E_CAPACITY, "orb_extract: negative capacity");
    CK(cudaSetDevice(ctx->device));
    if (ctx->orb_w != width || ctx->orb_h != height || ctx->orb_nf != nf) {
        std::vector<int32_t> tabs;
        if (!orb_make_geometry(width, height, nf, ctx->orb_geom, tabs))
            return fail(ctx, MVS_E_UNSUPPORTED, "orb_extract: image too small for an 8-level pyramid");
        if (ctx->orb_geom.lv[0].quota > kOrbSortCap) {      // level 0 has the largest quota (about 0.217 * n_features)
            ctx->orb_w = 0;
            return fail(ctx, MVS_E_UNSUPPORTED, "orb_extract: n_features too large (a level keeps at most 4096 keypoints: n_features <= 18800)");
        }
        CK(cudaStreamSynchronize(ctx->stream));   // earlier work may still read the old tables
        CK(ctx->o_tabs.ensure(std::max<size_t>(tabs.size(), 1) * sizeof(int32_t)));
        CK(cudaMemcpy(ctx->o_tabs.p, tabs.data(), tabs.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        ctx->orb_w = width; ctx->orb_h = height; ctx->orb_nf = nf;
    }
    const OrbGeom &g = ctx->orb_geom;
    // chunk size: keep the workspace (2 pyramids + candidate lists + kept lists per image) near 2 GB
    const size_t per_image = 2 * (size_t)g.slab + 8 * (size_t)g.cand_total + (size_t)kOrbLevels * kOrbSortCap * 4 + (h_images ? (size_t)height * stride : 0);
    int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_images, ((size_t)2 << 30) / per_image));
    // Host images: chunks of at most kOrbHostChunk frames, fetched one chunk ahead on a copy stream into a double-buffered
    // staging area, so that the host-to-device copy of chunk k+1 (for pageable memory: the driver's staged copy, which
    // occupies the calling thread) runs under the kernels of chunk k.
    constexpr int kOrbHostChunk = 64;
    const bool prefetch = h_images && n_images > kOrbHostChunk;
    if (prefetch) chunk = std::min(chunk, kOrbHostChunk);
    const size_t isz = (size_t)height * stride;
    if (h_images) CK(ctx->o_stage.ensure((size_t)(prefetch ? 2 : 1) * chunk * isz));
    if (prefetch && !ctx->copy_stream) {
        CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        for (cudaEvent_t &e : ctx->copy_done) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    // one copy per run of images that are contiguous in host memory
    auto copy_chunk = [&](int k, cudaStream_t st) -> cudaError_t {
        const int c0 = k * chunk, n = std::min(chunk, n_images - c0);
        uint8_t *dst = ctx->o_stage.as<uint8_t>() + (prefetch ? (size_t)(k & 1) * chunk * isz : 0);
        for (int i = 0; i < n;) {
            int j = i + 1;
            while (j < n && h_images[c0 + j] == h_images[c0 + j - 1] + isz) ++j;
            cudaError_t e = cudaMemcpyAsync(dst + (size_t)i * isz, h_images[c0 + i], (size_t)(j - i) * isz, cudaMemcpyHostToDevice, st);
            if (e != cudaSuccess) return e;
            i = j;
        }
        return cudaSuccess;
    };
    if (h_images)
        for (int i = 0; i < n_images; ++i)
            if (!h_images[i]) return fail(ctx, MVS_E_BAD_ARG, "orb_extract: null image");
    struct CopyGuard {      // never return while the copy stream may still read the caller's images
        cudaStream_t s = nullptr;
        ~CopyGuard() { if (s) cudaStreamSynchronize(s); }
    } copy_guard;
    if (prefetch) {
        copy_guard.s = ctx->copy_stream;
        CK(cudaStreamSynchronize(ctx->stream));     // earlier work of this context may still read the staging area
        CK(copy_chunk(0, ctx->copy_stream));
        CK(cudaEventRecord(ctx->copy_done[0], ctx->copy_stream));
    }
    CK(ctx->o_pyr.ensure((size_t)chunk * g.slab));
    CK(ctx->o_blur.ensure((size_t)chunk * g.slab));
    CK(ctx->o_cxy.ensure((size_t)chunk * g.cand_total * 4));
    CK(ctx->o_cval.ensure((size_t)chunk * g.cand_total * 4));
    CK(ctx->o_cnt.ensure((size_t)chunk * kOrbLevels * 257 * sizeof(int32_t)));
    CK(ctx->o_kidx.ensure((size_t)chunk * kOrbLevels * kOrbSortCap * 4));
    CK(ctx->o_kcnt.ensure((size_t)chunk * kOrbLevels * sizeof(int32_t)));
    CK(ctx->o_off.ensure((size_t)chunk * sizeof(int32_t)));
    const size_t pin_need = (size_t)chunk * (kOrbLevels + 1);
    if (ctx->o_pinned_cap < pin_need) {
        if (ctx->o_pinned) cudaFreeHost(ctx->o_pinned);
        ctx->o_pinned = nullptr; ctx->o_pinned_cap = 0;
        CK(cudaMallocHost((void **)&ctx->o_pinned, pin_need * sizeof(int32_t)));
        ctx->o_pinned_cap = pin_need;
    }
    OrbBuffers b{};
    b.pyr = ctx->o_pyr.as<uint8_t>(); b.blur = ctx->o_blur.as<uint8_t>(); b.tabs = ctx->o_tabs.as<int32_t>();
    b.cand_xy = ctx->o_cxy.as<uint32_t>(); b.cand_val = ctx->o_cval.as<float>();
    b.cand_cnt = ctx->o_cnt.as<int32_t>(); b.hist = b.cand_cnt + (size_t)chunk * kOrbLevels;
    b.kept_idx = ctx->o_kidx.as<uint32_t>(); b.kept_cnt = ctx->o_kcnt.as<int32_t>();

    // where the results go: the resident frame table (append) or the ctx's own result buffers
    const int frames0 = (int)ctx->h_cnt.size();
    const size_t used0 = (append_frames && frames0) ? (size_t)ctx->h_off[frames0 - 1] + (size_t)ctx->h_cnt[frames0 - 1] : 0;
    size_t total = 0;                       // keypoints produced so far in this call
    std::vector<int32_t> img_count(n_images);
    for (int c0 = 0; c0 < n_images; c0 += chunk) {
        const int n = std::min(chunk, n_images - c0);
        const uint8_t *stage = nullptr;
        const int ck = c0 / chunk;
        if (d_images) {
            stage = d_images + (size_t)c0 * height * stride;
        } else if (prefetch) {
            CK(cudaStreamWaitEvent(ctx->stream, ctx->copy_done[ck & 1], 0));
            stage = ctx->o_stage.as<uint8_t>() + (size_t)(ck & 1) * chunk * isz;
        } else {
            CK(copy_chunk(ck, ctx->stream));
            stage = ctx->o_stage.as<uint8_t>();
        }
        CK(cudaMemsetAsync(b.cand_cnt, 0, (size_t)chunk * kOrbLevels * 257 * sizeof(int32_t), ctx->stream));
        {
            StageTimer t(ctx, MVS_STAGE_ORB_PYRAMID);
            launch_orb_pyramid(g, b, stage, stride, n, ctx->stream);
            CK(cudaGetLastError());
        }
        { StageTimer t(ctx, MVS_STAGE_ORB_FAST, g.fast_tiles ? 1 : 0); launch_orb_fast(g, b, n, ctx->stream); CK(cudaGetLastError()); }
        { StageTimer t(ctx, MVS_STAGE_ORB_HARRIS); launch_orb_harris(g, b, n, ctx->stream); CK(cudaGetLastError()); }
        { StageTimer t(ctx, MVS_STAGE_ORB_SELECT); launch_orb_select(g, b, n, ctx->stream); CK(cudaGetLastError()); }
        CK(cudaMemcpyAsync(ctx->o_pinned, b.kept_cnt, (size_t)n * kOrbLevels * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        { StageTimer t(ctx, MVS_STAGE_ORB_BLUR); launch_orb_blur(g, b, n, ctx->stream); CK(cudaGetLastError()); }
        if (prefetch && c0 + chunk < n_images) {    // the other staging half was last read by chunk k-1, which has completed
            CK(copy_chunk(ck + 1, ctx->copy_stream));
            CK(cudaEventRecord(ctx->copy_done[(ck + 1) & 1], ctx->copy_stream));
        }
        CK(cudaStreamSynchronize(ctx->stream));
        int32_t *off = ctx->o_pinned + (size_t)chunk * kOrbLevels;
        size_t chunk_total = 0;
        int32_t most = 0;
        for (int i = 0; i < n; ++i) {
            int32_t c = 0;
            for (int l = 0; l < kOrbLevels; ++l) {
                const int32_t k = ctx->o_pinned[i * kOrbLevels + l];
                if (k < 0) return fail(ctx, MVS_E_CAPACITY, "orb_extract: more than 4096 keypoints tie at a level's cut-off response");
                c += k;
            }
            img_count[c0 + i] = c;
            most = std::max(most, c);
            off[i] = (int32_t)(total + chunk_total);
            chunk_total += (size_t)c;
        }
        if (used0 + total + chunk_total > 0x7FFFFFFFull) return fail(ctx, MVS_E_UNSUPPORTED, "more than 2^31 keypoints");
        CK(grow_keep(ctx, ctx->o_kp, sizeof(mvs_keypoint), total, std::max<size_t>(total + chunk_total, 1)));
        OrbDescribeArgs d{};
        d.img_off = ctx->o_off.as<int32_t>();
        d.kp = ctx->o_kp.as<mvs_keypoint>();
        if (append_frames) {
            CK(grow_keep(ctx, ctx->d_desc, 32, used0 + total, std::max<size_t>(used0 + total + chunk_total, 1)));
            CK(grow_keep(ctx, ctx->d_kp, sizeof(float2), used0 + total, std::max<size_t>(used0 + total + chunk_total, 1)));
            d.desc = ctx->d_desc.as<uint8_t>() + used0 * 32;
            d.frame_kp = ctx->d_kp.as<float2>() + used0;
        } else {
            CK(grow_keep(ctx, ctx->o_desc, 32, total, std::max<size_t>(total + chunk_total, 1)));
            d.desc = ctx->o_desc.as<uint8_t>();
        }
        CK(cudaMemcpyAsync(ctx->o_off.p, off, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        {
            StageTimer t(ctx, MVS_STAGE_ORB_DESCRIBE);
            launch_orb_describe(g, b, d, n, most, ctx->stream);
            CK(cudaGetLastError());
        }
        CK(cudaStreamSynchronize(ctx->stream));   // the pinned offsets are rewritten by the next chunk
        total += chunk_total;
    }
    if (counts) std::memcpy(counts, img_count.data(), (size_t)n_images * sizeof(int32_t));
    if (append_frames) {
        // frame table bookkeeping (as mvs_frames_append, for all new frames at once)
        CK(grow_keep(ctx, ctx->d_foff, sizeof(int32_t), (size_t)frames0, (size_t)frames0 + n_images));
        CK(grow_keep(ctx, ctx->d_fcnt, sizeof(int32_t), (size_t)frames0, (size_t)frames0 + n_images));
        std::vector<int32_t> offs(n_images);
        size_t at = used0;
        for (int i = 0; i < n_images; ++i) { offs[i] = (int32_t)at; at += (size_t)img_count[i]; }
        CK(cudaMemcpyAsync(ctx->d_foff.as<int32_t>() + frames0, offs.data(), (size_t)n_images * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_fcnt.as<int32_t>() + frames0, img_count.data(), (size_t)n_images * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < n_images; ++i) { ctx->h_off.push_back(offs[i]); ctx->h_cnt.push_back(img_count[i]); }
        if (first_frame) *first_frame = frames0;
    }
    if ((keypoints || descriptors) && (int64_t)total > capacity)
        return fail(ctx, MVS_E_CAPACITY, "orb_extract: keypoint capacity too small (counts are filled)");
    if (keypoints && total)
        CK(cudaMemcpyAsync(keypoints, ctx->o_kp.p, total * sizeof(mvs_keypoint), cudaMemcpyDeviceToHost, ctx->stream));
    if (descriptors && total)
        CK(cudaMemcpyAsync(descriptors, append_frames ? ctx->d_desc.as<uint8_t>() + used0 * 32 : ctx->o_desc.as<uint8_t>(), total * 32,
                           cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return MVS_OK;
}

int mvs_orb_extract(mvs_ctx *ctx, const uint8_t *const *images, int n_images, int width, int height, int stride_bytes,
                    const mvs_orb_params *params, int append_frames, int32_t *first_frame,
                    int32_t *counts, mvs_keypoint *keypoints, uint8_t *descriptors, int64_t capacity)
{
    return orb_extract_impl(ctx, images, nullptr, n_images, width, height, stride_bytes, params, append_frames, first_frame,
                            counts, keypoints, descriptors, capacity);
}

int mvs_orb_extract_device(mvs_ctx *ctx, const void *d_images, int n_images, int width, int height, int stride_bytes,
                           const mvs_orb_params *params, int append_frames, int32_t *first_frame,
                           int32_t *counts, mvs_keypoint *keypoints, uint8_t *descriptors, int64_t capacity)
{
    return orb_extract_impl(ctx, nullptr, static_cast<const uint8_t *>(d_images), n_images, width, height, stride_bytes, params,
                            append_frames, first_frame, counts, keypoints, descriptors, capacity);
}

// ------------------------------------------------------------------------------------------ pnp_solve
void mvs_pnp_sample_table(uint64_t seed, uint64_t problem_id, uint32_t n_points, int H, uint32_t *out)
{
    pnp_sample_table_host(seed, problem_id, n_points, H, out);
}

static int pnp_impl(mvs_ctx *ctx, const double *world, const double *image, const int32_t *counts, int n_problems,
                    const double K[9], const mvs_pnp_params *params, const uint32_t *samples, mvs_pnp_result *results,
                    uint8_t *inlier_mask, int32_t *all_counts)
{
    if (!ctx) return MVS_E_BAD_ARG;
    if (!counts || n_problems < 1 || !K || !results) return fail(ctx, MVS_E_BAD_ARG, "pnp_solve: null argument or no problem");
    std::vector<int32_t> off(n_problems + 1, 0);
    int max_n = 0;
    for (int i = 0; i < n_problems; ++i) {
        if (counts[i] < 0) return fail(ctx, MVS_E_BAD_ARG, "pnp_solve: negative point count");
        if ((int64_t)off[i] + counts[i] > 0x7FFFFFFF) return fail(ctx, MVS_E_UNSUPPORTED, "pnp_solve: more than 2^31 points");
        off[i + 1] = off[i] + counts[i];
        max_n = std::max(max_n, counts[i]);
    }
    const size_t total = (size_t)off[n_problems];
    if (total && (!world || !image)) return fail(ctx, MVS_E_BAD_ARG, "pnp_solve: null points");
    if (!(K[0] != 0.0) || !(K[4] != 0.0)) return fail(ctx, MVS_E_BAD_ARG, "pnp_solve: zero focal length");
    PnpArgs a{};
    a.H = params && params->n_hypotheses > 0 ? params->n_hypotheses : 100;        // pnp-solve.cpp:50
    a.refine_iters = params && params->refine_iterations >= 0 ? params->refine_iterations : 10;
    const double err = params && params->reprojection_error > 0 ? params->reprojection_error : 0.05;   // :51
    a.thr2 = err * err;
    a.seed = params ? params->seed : 0; a.problem_id_base = params ? params->problem_id_base : 0;
    a.min_inliers = params && params->min_inliers > 0 ? params->min_inliers : 4;
    a.fx = K[0]; a.fy = K[4]; a.cx = K[2]; a.cy = K[5];
    a.tiles = pnp_tiles(max_n);
    if (n_problems > 65535) return fail(ctx, MVS_E_UNSUPPORTED, "pnp_solve: more than 65535 problems per call");
    CK(cudaSetDevice(ctx->device));
    CK(ctx->p_world.ensure(std::max<size_t>(total, 1) * 24));
    CK(ctx->p_image.ensure(std::max<size_t>(total, 1) * 16));
    CK(ctx->p_off.ensure(off.size() * sizeof(int32_t)));
    CK(ctx->p_poses.ensure((size_t)n_problems * a.H * 12 * sizeof(double)));
    CK(ctx->p_valid.ensure((size_t)n_problems * a.H));
    CK(ctx->p_pc.ensure((size_t)n_problems * a.tiles * a.H * sizeof(int32_t)));
    CK(ctx->p_maskws.ensure(std::max<size_t>(total, 1)));
    CK(ctx->p_results.ensure((size_t)n_problems * sizeof(mvs_pnp_result)));
    if (inlier_mask) CK(ctx->p_mask.ensure(std::max<size_t>(total, 1)));
    if (all_counts) CK(ctx->p_counts.ensure((size_t)n_problems * a.H * sizeof(int32_t)));
    if (samples) {
        CK(ctx->p_table.ensure((size_t)a.H * 4 * sizeof(uint32_t)));
        CK(cudaMemcpyAsync(ctx->p_table.p, samples, (size_t)a.H * 4 * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (total) {
        CK(cudaMemcpyAsync(ctx->p_world.p, world, total * 24, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->p_image.p, image, total * 16, cudaMemcpyHostToDevice, ctx->stream));
    }
    CK(cudaMemcpyAsync(ctx->p_off.p, off.data(), off.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    a.world = ctx->p_world.as<double>(); a.image = ctx->p_image.as<double>(); a.offsets = ctx->p_off.as<int32_t>();
    a.table = samples ? ctx->p_table.as<uint32_t>() : nullptr;
    a.poses = ctx->p_poses.as<double>(); a.valid = ctx->p_valid.as<uint8_t>(); a.part_count = ctx->p_pc.as<int32_t>();
    a.mask_ws = ctx->p_maskws.as<uint8_t>(); a.mask = inlier_mask ? ctx->p_mask.as<uint8_t>() : nullptr;
    a.all_counts = all_counts ? ctx->p_counts.as<int32_t>() : nullptr;
    a.results = ctx->p_results.as<mvs_pnp_result>();
    {
        StageTimer t(ctx, MVS_STAGE_PNP, 3);
        launch_pnp(a, n_problems, max_n, ctx->stream);
        CK(cudaGetLastError());
    }
    CK(cudaMemcpyAsync(results, ctx->p_results.p, (size_t)n_problems * sizeof(mvs_pnp_result), cudaMemcpyDeviceToHost, ctx->stream));
    if (inlier_mask && total) CK(cudaMemcpyAsync(inlier_mask, ctx->p_mask.p, total, cudaMemcpyDeviceToHost, ctx->stream));
    if (all_counts) CK(cudaMemcpyAsync(all_counts, ctx->p_counts.p, (size_t)n_problems * a.H * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));   // off is stack/vector memory; results are host-visible on return
    return MVS_OK;
}

int mvs_pnp_solve(mvs_ctx *ctx, const double *world, const double *image, int n, const double K[9],
                  const mvs_pnp_params *params, const uint32_t *samples, mvs_pnp_result *result,
                  uint8_t *inlier_mask, int32_t *all_counts)
{
    const int32_t cnt = n;
    int st = pnp_impl(ctx, world, image, &cnt, 1, K, params, samples, result, inlier_mask, all_counts);
    return st != MVS_OK ? st : result->status;
}

int mvs_pnp_solve_batch(mvs_ctx *ctx, const double *world, const double *image, const int32_t *counts, int n_problems,
                        const double K[9], const mvs_pnp_params *params, const uint32_t *samples,
                        mvs_pnp_result *results, uint8_t *inlier_mask)
{
    return pnp_impl(ctx, world, image, counts, n_problems, K, params, samples, results, inlier_mask, nullptr);
}

// ------------------------------------------------------------------------------------------ bundle adjustment
int mvs_ba_solve_batch(mvs_ctx *ctx, int n_problems, const double K[9],
                       const int32_t *n_frames, const int32_t *n_points, const int32_t *n_obs,
                       const double *pose_R, const double *pose_t, const double *pose_prior_cov,
                       const double *points, const double *point_prior_cov, const mvs_ba_observation *obs,
                       const mvs_ba_params *params,
                       double *pose_R_out, double *pose_t_out, double *pose_cov_out,
                       double *points_out, double *point_cov_out, mvs_ba_result *results)
{
    if (!ctx) return MVS_E_BAD_ARG;
    if (n_problems < 1 || !K || !n_frames || !n_points || !n_obs || !pose_R || !pose_t || !pose_prior_cov || !results)
        return fail(ctx, MVS_E_BAD_ARG, "ba_solve: null argument or no problem");
    std::vector<int32_t> foff(n_problems + 1, 0), poff(n_problems + 1, 0), ooff(n_problems + 1, 0);
    for (int i = 0; i < n_problems; ++i) {
        if (n_frames[i] < 1 || n_points[i] < 0 || n_obs[i] < 0) return fail(ctx, MVS_E_BAD_ARG, "ba_solve: bad problem size");
        if (n_frames[i] > BA_MAX_FRAMES) return fail(ctx, MVS_E_UNSUPPORTED, "ba_solve: more than 16 frames in a problem");
        if ((int64_t)poff[i] + n_points[i] > 0x7FFFFFFF || (int64_t)ooff[i] + n_obs[i] > 0x7FFFFFFF)
            return fail(ctx, MVS_E_UNSUPPORTED, "ba_solve: more than 2^31 points or observations");
        foff[i + 1] = foff[i] + n_frames[i]; poff[i + 1] = poff[i] + n_points[i]; ooff[i + 1] = ooff[i] + n_obs[i];
    }
    const size_t NF = (size_t)foff[n_problems], NP = (size_t)poff[n_problems], NO = (size_t)ooff[n_problems];
    if ((NP && (!points || !point_prior_cov)) || (NO && !obs)) return fail(ctx, MVS_E_BAD_ARG, "ba_solve: null points or observations");
    // observations grouped by point (counting sort per problem; order within a point is the caller's)
    std::vector<mvs_ba_observation> sorted(NO);
    std::vector<int32_t> pobs(NP + 1, 0);
    for (int i = 0; i < n_problems; ++i) {
        const int P = n_points[i], F = n_frames[i];
        for (int o = ooff[i]; o < ooff[i + 1]; ++o) {
            if (obs[o].point < 0 || obs[o].point >= P || obs[o].frame < 0 || obs[o].frame >= F)
                return fail(ctx, MVS_E_BAD_ARG, "ba_solve: observation refers to a frame or point outside its problem");
            ++pobs[(size_t)poff[i] + obs[o].point + 1];
        }
    }
    for (size_t p = 0; p < NP; ++p) pobs[p + 1] += pobs[p];
    {
        std::vector<int32_t> at(pobs.begin(), pobs.end() - 1);
        for (int i = 0; i < n_problems; ++i)
            for (int o = ooff[i]; o < ooff[i + 1]; ++o) sorted[(size_t)at[(size_t)poff[i] + obs[o].point]++] = obs[o];
    }
    CK(cudaSetDevice(ctx->device));
    auto up = [&](DevBuf &b, const void *src, size_t bytes) -> cudaError_t {
        cudaError_t e = b.ensure(std::max<size_t>(bytes, 8));
        if (e != cudaSuccess || !bytes) return e;
        return cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream);
    };
    CK(up(ctx->b_foff, foff.data(), foff.size() * 4)); CK(up(ctx->b_poff, poff.data(), poff.size() * 4));
    CK(up(ctx->b_R, pose_R, NF * 72)); CK(up(ctx->b_t, pose_t, NF * 24)); CK(up(ctx->b_pc, pose_prior_cov, NF * 288));
    CK(up(ctx->b_X, points, NP * 24)); CK(up(ctx->b_xc, point_prior_cov, NP * 72));
    CK(up(ctx->b_obs, sorted.data(), NO * sizeof(mvs_ba_observation))); CK(up(ctx->b_ooff, pobs.data(), pobs.size() * 4));
    CK(ctx->b_ws.ensure(std::max<size_t>(NP, 1) * 48 * 8));
    // problems with more than two frames: their own per-point workspace (12 + 36 F doubles) and one 27-double slot per observation
    int min_frames = n_frames[0], max_frames = n_frames[0];
    std::vector<long long> wsm(n_problems + 1, 0);
    for (int i = 0; i < n_problems; ++i) {
        min_frames = std::min(min_frames, (int)n_frames[i]); max_frames = std::max(max_frames, (int)n_frames[i]);
        wsm[i + 1] = wsm[i] + (n_frames[i] > 2 ? (long long)n_points[i] * (12 + 36 * n_frames[i]) : 0);
    }
    if (max_frames > 2) {
        CK(up(ctx->b_wsm_off, wsm.data(), wsm.size() * sizeof(long long)));
        CK(ctx->b_wsm.ensure(std::max<size_t>((size_t)wsm[n_problems], 1) * 8));
        CK(ctx->b_wsobs.ensure(std::max<size_t>(NO, 1) * 27 * 8));
    }
    CK(ctx->b_Ro.ensure(NF * 72)); CK(ctx->b_to.ensure(NF * 24)); CK(ctx->b_pco.ensure(NF * 288));
    CK(ctx->b_Xo.ensure(std::max<size_t>(NP, 1) * 24)); CK(ctx->b_xco.ensure(std::max<size_t>(NP, 1) * 72));
    CK(ctx->b_res.ensure((size_t)n_problems * sizeof(mvs_ba_result)));
    BaArgs a{};
    a.fx = K[0]; a.fy = K[4]; a.sk = K[1]; a.u0 = K[2]; a.v0 = K[5];
    a.frame_off = ctx->b_foff.as<int32_t>(); a.point_off = ctx->b_poff.as<int32_t>();
    a.pose_R = ctx->b_R.as<double>(); a.pose_t = ctx->b_t.as<double>(); a.pose_prior_cov = ctx->b_pc.as<double>();
    a.points = ctx->b_X.as<double>(); a.point_prior_cov = ctx->b_xc.as<double>();
    a.obs = ctx->b_obs.as<mvs_ba_observation>(); a.point_obs_off = ctx->b_ooff.as<int32_t>();
    a.ws = ctx->b_ws.as<double>();
    if (max_frames > 2) { a.ws_multi = ctx->b_wsm.as<double>(); a.ws_multi_off = ctx->b_wsm_off.as<long long>(); a.ws_obs = ctx->b_wsobs.as<double>(); }
    a.pose_R_out = ctx->b_Ro.as<double>(); a.pose_t_out = ctx->b_to.as<double>(); a.pose_cov_out = ctx->b_pco.as<double>();
    a.points_out = ctx->b_Xo.as<double>(); a.point_cov_out = ctx->b_xco.as<double>();
    a.results = ctx->b_res.as<mvs_ba_result>();
    a.max_iter = params && params->max_iterations > 0 ? params->max_iterations : 100;
    a.lambda0 = params && params->lambda_initial > 0 ? params->lambda_initial : 1e-5;
    a.rel_tol = params && params->relative_tolerance > 0 ? params->relative_tolerance : 1e-5;     // GTSAM relativeErrorTol
    a.abs_tol = !params || params->absolute_tolerance == 0 ? 1e-5 : params->absolute_tolerance;    // GTSAM absoluteErrorTol; < 0: off
    {
        StageTimer t(ctx, MVS_STAGE_BA);
        CK(launch_ba(a, n_problems, min_frames, max_frames, ctx->stream));
    }
    CK(cudaMemcpyAsync(results, ctx->b_res.p, (size_t)n_problems * sizeof(mvs_ba_result), cudaMemcpyDeviceToHost, ctx->stream));
    if (pose_R_out) CK(cudaMemcpyAsync(pose_R_out, ctx->b_Ro.p, NF * 72, cudaMemcpyDeviceToHost, ctx->stream));
    if (pose_t_out) CK(cudaMemcpyAsync(pose_t_out, ctx->b_to.p, NF * 24, cudaMemcpyDeviceToHost, ctx->stream));
    if (pose_cov_out) CK(cudaMemcpyAsync(pose_cov_out, ctx->b_pco.p, NF * 288, cudaMemcpyDeviceToHost, ctx->stream));
    if (points_out && NP) CK(cudaMemcpyAsync(points_out, ctx->b_Xo.p, NP * 24, cudaMemcpyDeviceToHost, ctx->stream));
    if (point_cov_out && NP) CK(cudaMemcpyAsync(point_cov_out, ctx->b_xco.p, NP * 72, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return MVS_OK;
}

}  // extern "C"
