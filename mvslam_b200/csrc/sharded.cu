// sharded.cu — the multi-GPU driver of the path behind the C ABI: image pairs are independent
// (ImagePair touches only its two frames and the global K, reference source/front-end/image-pair.cpp:30-71,143), so the
// pair list is cut into contiguous slices, one per rank (one process per GPU), every rank keeps the frame table resident,
// and the ONLY communication is at the end: one gather of the fixed-size records and one gather of the variable-length
// point clouds (points + indexes + matches) with exclusive-scan offsets.  No collective inside the hot path.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy already mapped into the process, e.g. PyTorch's, or the
// system one), so libmvslam_b200.so has no link-time dependency on it and single-GPU users never load it.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels.h"

using namespace mvs;

extern "C" int mvs_pair_batch_device_only(mvs_ctx *ctx, const int32_t *pairs, int n_pairs, const double K[9],
                                          const mvs_match_params *mparams, const mvs_ransac_params *rparams);
int mvs_ctx_device(const mvs_ctx *ctx);
cudaStream_t mvs_ctx_stream(const mvs_ctx *ctx);
void mvs_ctx_set_error(mvs_ctx *ctx, const std::string &msg);
// device-side outputs of the last pair_batch chunk of this ctx (api.cu)
void mvs_ctx_last_outputs(mvs_ctx *ctx, const mvs_pair_result **res, const mvs_match **matches, const double **points,
                          const uint64_t **indexes, int *stride);

namespace {

struct Nccl {
    void *lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    bool ok = false;
};

Nccl &nccl()
{
    static Nccl n;
    if (n.lib) return n;
    n.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!n.lib) return n;
#define MVS_SYM(name) n.name = (decltype(n.name))dlsym(n.lib, "nccl" #name)
    MVS_SYM(GetUniqueId); MVS_SYM(CommInitRank); MVS_SYM(CommDestroy); MVS_SYM(GroupStart); MVS_SYM(GroupEnd);
    MVS_SYM(Send); MVS_SYM(Recv); MVS_SYM(GetErrorString);
#undef MVS_SYM
    n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.GroupStart && n.GroupEnd && n.Send && n.Recv && n.GetErrorString;
    return n;
}

// order-preserving compaction of one chunk's per-pair rows (stride slots each) into a contiguous array
template <typename T, int W>
__global__ void compact_rows_kernel(const T *src, int stride, const int32_t *counts, const int64_t *offsets, T *dst)
{
    const int pair = blockIdx.x;
    const int n = min(counts[pair], stride);
    const T *s = src + (size_t)pair * stride * W;
    T *d = dst + (size_t)offsets[pair] * W;
    for (int i = threadIdx.x; i < n * W; i += blockDim.x) d[i] = s[i];
}

struct Buf {
    void *p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        void *q = nullptr;
        const size_t want = bytes + bytes / 2 + 4096;
        cudaError_t e = cudaMalloc(&q, want);
        if (e != cudaSuccess) return e;
        if (p) cudaFree(p);
        p = q; cap = want;
        return cudaSuccess;
    }
    cudaError_t ensure_keep(size_t bytes, size_t used, cudaStream_t s)
    {
        if (bytes <= cap) return cudaSuccess;
        void *q = nullptr;
        const size_t want = bytes * 2 + 4096;
        cudaError_t e = cudaMalloc(&q, want);
        if (e != cudaSuccess) return e;
        if (p && used) { cudaMemcpyAsync(q, p, used, cudaMemcpyDeviceToDevice, s); cudaStreamSynchronize(s); }
        if (p) cudaFree(p);
        p = q; cap = want;
        return cudaSuccess;
    }
    ~Buf() { if (p) cudaFree(p); }
};

}  // namespace

struct mvs_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    mvs_ctx *ctx = nullptr;
    Buf rec, pts, idx, mat, cnt, off, all_rec, all_pts, all_idx, all_mat;
};

#define CKC(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) { mvs_ctx_set_error(ctx, std::string(#call) + ": " + cudaGetErrorString(e__)); return MVS_E_CUDA; } \
    } while (0)
#define CKN(call)                                                                                   \
    do {                                                                                            \
        ncclResult_t r__ = (call);                                                                  \
        if (r__ != ncclSuccess) { mvs_ctx_set_error(ctx, std::string(#call) + ": " + nccl().GetErrorString(r__)); return MVS_E_CUDA; } \
    } while (0)

extern "C" {

int mvs_comm_unique_id(uint8_t id[128])
{
    if (!id || !nccl().ok) return MVS_E_UNSUPPORTED;
    ncclUniqueId u;
    static_assert(sizeof(u) == 128, "ncclUniqueId is 128 bytes");
    if (nccl().GetUniqueId(&u) != ncclSuccess) return MVS_E_CUDA;
    std::memcpy(id, &u, 128);
    return MVS_OK;
}

int mvs_comm_create(mvs_comm **out, mvs_ctx *ctx, const uint8_t id[128], int rank, int world)
{
    if (!out || !ctx || !id || world < 1 || rank < 0 || rank >= world) return MVS_E_BAD_ARG;
    *out = nullptr;
    if (!nccl().ok) { mvs_ctx_set_error(ctx, "libnccl.so.2 could not be loaded"); return MVS_E_UNSUPPORTED; }
    CKC(cudaSetDevice(mvs_ctx_device(ctx)));
    mvs_comm *c = new mvs_comm();
    c->rank = rank; c->world = world; c->ctx = ctx;
    ncclUniqueId u;
    std::memcpy(&u, id, 128);
    ncclResult_t r = nccl().CommInitRank(&c->comm, world, u, rank);
    if (r != ncclSuccess) { mvs_ctx_set_error(ctx, std::string("ncclCommInitRank: ") + nccl().GetErrorString(r)); delete c; return MVS_E_CUDA; }
    *out = c;
    return MVS_OK;
}

void mvs_comm_destroy(mvs_comm *c)
{
    if (!c) return;
    if (c->comm) nccl().CommDestroy(c->comm);
    delete c;
}

void mvs_shard_bounds(int64_t n, int world, int rank, int64_t *lo, int64_t *hi)
{
    const int64_t base = n / world, rem = n % world;
    *lo = rank * base + std::min<int64_t>(rank, rem);
    *hi = *lo + base + (rank < rem ? 1 : 0);
}

int mvs_pair_batch_sharded(mvs_ctx *ctx, mvs_comm *c, const int32_t *pairs, int64_t n_pairs_total, const double K[9],
                           const mvs_match_params *mparams, const mvs_ransac_params *rparams, int root,
                           mvs_pair_result *results, int64_t *point_offsets, double *points, uint64_t *indexes,
                           int64_t point_capacity, int64_t *match_offsets, mvs_match *matches, int64_t match_capacity)
{
    if (!ctx || !c || !pairs || !K || n_pairs_total < 1 || root < 0 || root >= c->world) return MVS_E_BAD_ARG;
    const bool is_root = c->rank == root;
    if (is_root && !results) { mvs_ctx_set_error(ctx, "results must be given on the root rank"); return MVS_E_BAD_ARG; }
    const bool want_pts = point_offsets != nullptr, want_mat = match_offsets != nullptr;   // must agree on every rank
    CKC(cudaSetDevice(mvs_ctx_device(ctx)));
    cudaStream_t s = mvs_ctx_stream(ctx);
    int64_t lo, hi;
    mvs_shard_bounds(n_pairs_total, c->world, c->rank, &lo, &hi);
    const int64_t n_local = hi - lo;
    // ---- this rank's slice, chunk by chunk; records, and compacted details, accumulate on the device
    CKC(c->rec.ensure((size_t)std::max<int64_t>(n_local, 1) * sizeof(mvs_pair_result)));
    std::vector<mvs_pair_result> h_rec((want_pts || want_mat) ? (size_t)n_local : 0);
    std::vector<int32_t> h_cnt;
    std::vector<int64_t> h_off;
    int64_t pts_used = 0, mat_used = 0;
    constexpr int64_t kChunk = 4096;
    for (int64_t c0 = 0; c0 < n_local; c0 += kChunk) {
        const int n = (int)std::min(kChunk, n_local - c0);
        mvs_ransac_params rp = rparams ? *rparams : mvs_ransac_params{1, MVS_SCORE_ALGEBRAIC, 0.0, 0, 0, 0, 0};
        rp.pair_id_base += (uint64_t)(lo + c0);       // sampling is keyed by the GLOBAL pair index: results do not depend on the sharding
        // records stay on the device unless the sizes of the variable-length parts are needed here
        int st = (want_pts || want_mat)
                     ? mvs_pair_batch_enqueue(ctx, pairs + 2 * (lo + c0), n, K, mparams, &rp, h_rec.data() + c0, nullptr, nullptr, nullptr, nullptr, 0)
                     : mvs_pair_batch_device_only(ctx, pairs + 2 * (lo + c0), n, K, mparams, &rp);
        if (st != MVS_OK) return st;
        const mvs_pair_result *d_res; const mvs_match *d_mat; const double *d_pts; const uint64_t *d_idx; int stride;
        mvs_ctx_last_outputs(ctx, &d_res, &d_mat, &d_pts, &d_idx, &stride);
        CKC(cudaMemcpyAsync((mvs_pair_result *)c->rec.p + c0, d_res, (size_t)n * sizeof(mvs_pair_result), cudaMemcpyDeviceToDevice, s));
        if (want_pts || want_mat) {
            CKC(cudaStreamSynchronize(s));                // the chunk's records are on the host: sizes of its variable-length parts
            CKC(c->cnt.ensure((size_t)n * sizeof(int32_t))); CKC(c->off.ensure((size_t)n * sizeof(int64_t)));
            for (int pass = 0; pass < 2; ++pass) {
                if (!(pass == 0 ? want_pts : want_mat)) continue;
                h_cnt.resize(n); h_off.resize(n);
                int64_t tot = 0;
                for (int i = 0; i < n; ++i) {
                    const mvs_pair_result &r = h_rec[c0 + i];
                    h_cnt[i] = pass == 0 ? (r.status == MVS_OK ? r.n_points : 0) : std::min(r.n_matches, stride);
                    h_off[i] = tot; tot += h_cnt[i];
                }
                CKC(cudaMemcpyAsync(c->cnt.p, h_cnt.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
                CKC(cudaMemcpyAsync(c->off.p, h_off.data(), (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice, s));
                if (pass == 0) {
                    CKC(c->pts.ensure_keep((size_t)(pts_used + tot + 1) * 24, (size_t)pts_used * 24, s));
                    CKC(c->idx.ensure_keep((size_t)(pts_used + tot + 1) * 8, (size_t)pts_used * 8, s));
                    if (tot) {
                        compact_rows_kernel<double, 3><<<n, 128, 0, s>>>(d_pts, stride, (const int32_t *)c->cnt.p, (const int64_t *)c->off.p, (double *)c->pts.p + pts_used * 3);
                        compact_rows_kernel<uint64_t, 1><<<n, 128, 0, s>>>(d_idx, stride, (const int32_t *)c->cnt.p, (const int64_t *)c->off.p, (uint64_t *)c->idx.p + pts_used);
                    }
                    pts_used += tot;
                } else {
                    CKC(c->mat.ensure_keep((size_t)(mat_used + tot + 1) * sizeof(mvs_match), (size_t)mat_used * sizeof(mvs_match), s));
                    if (tot)
                        compact_rows_kernel<mvs_match, 1><<<n, 128, 0, s>>>(d_mat, stride, (const int32_t *)c->cnt.p, (const int64_t *)c->off.p, (mvs_match *)c->mat.p + mat_used);
                    mat_used += tot;
                }
                CKC(cudaStreamSynchronize(s));            // h_cnt / h_off are reused by the next pass
            }
        }
    }
    CKC(cudaStreamSynchronize(s));
    CKC(cudaGetLastError());
    // ---- gather 1: fixed-size records -> root
    std::vector<int64_t> slo(c->world), shi(c->world);
    for (int r = 0; r < c->world; ++r) mvs_shard_bounds(n_pairs_total, c->world, r, &slo[r], &shi[r]);
    if (is_root) CKC(c->all_rec.ensure((size_t)n_pairs_total * sizeof(mvs_pair_result)));
    CKN(nccl().GroupStart());
    if (is_root)
        for (int r = 0; r < c->world; ++r)
            if (shi[r] > slo[r]) CKN(nccl().Recv((mvs_pair_result *)c->all_rec.p + slo[r], (size_t)(shi[r] - slo[r]) * sizeof(mvs_pair_result), ncclChar, r, c->comm, s));
    if (n_local > 0) CKN(nccl().Send(c->rec.p, (size_t)n_local * sizeof(mvs_pair_result), ncclChar, root, c->comm, s));
    CKN(nccl().GroupEnd());
    if (is_root) CKC(cudaMemcpyAsync(results, c->all_rec.p, (size_t)n_pairs_total * sizeof(mvs_pair_result), cudaMemcpyDeviceToHost, s));
    CKC(cudaStreamSynchronize(s));
    if (!want_pts && !want_mat) return MVS_OK;
    // ---- gather 2: variable-length parts, placed by the exclusive scan of the per-pair counts (the records carry them)
    std::vector<int64_t> p_rank(c->world + 1, 0), m_rank(c->world + 1, 0);
    if (is_root) {
        int64_t pt = 0, mt = 0;
        for (int r = 0; r < c->world; ++r) {
            p_rank[r] = pt; m_rank[r] = mt;
            for (int64_t i = slo[r]; i < shi[r]; ++i) {
                if (want_pts) point_offsets[i] = pt;
                if (want_mat) match_offsets[i] = mt;
                pt += results[i].status == MVS_OK ? results[i].n_points : 0;
                mt += results[i].n_matches;        // stride >= n_matches: the slots hold every match (stride = largest pair frame)
            }
        }
        p_rank[c->world] = pt; m_rank[c->world] = mt;
        if (want_pts) point_offsets[n_pairs_total] = pt;
        if (want_mat) match_offsets[n_pairs_total] = mt;
        if ((want_pts && pt > point_capacity) || (want_mat && mt > match_capacity)) {
            mvs_ctx_set_error(ctx, "point / match capacity too small (offsets[n_pairs_total] holds the size needed)");
            // the other ranks are about to send: receive into scratch so that nobody hangs, then report
        }
        if (want_pts) { CKC(c->all_pts.ensure((size_t)(pt + 1) * 24)); CKC(c->all_idx.ensure((size_t)(pt + 1) * 8)); }
        if (want_mat) CKC(c->all_mat.ensure((size_t)(mt + 1) * sizeof(mvs_match)));
    }
    CKN(nccl().GroupStart());
    if (is_root)
        for (int r = 0; r < c->world; ++r) {
            const int64_t np = p_rank[r + 1] - p_rank[r], nm = m_rank[r + 1] - m_rank[r];
            if (want_pts && np > 0) {
                CKN(nccl().Recv((double *)c->all_pts.p + p_rank[r] * 3, (size_t)np * 24, ncclChar, r, c->comm, s));
                CKN(nccl().Recv((uint64_t *)c->all_idx.p + p_rank[r], (size_t)np * 8, ncclChar, r, c->comm, s));
            }
            if (want_mat && nm > 0) CKN(nccl().Recv((mvs_match *)c->all_mat.p + m_rank[r], (size_t)nm * sizeof(mvs_match), ncclChar, r, c->comm, s));
        }
    if (want_pts && pts_used > 0) {
        CKN(nccl().Send(c->pts.p, (size_t)pts_used * 24, ncclChar, root, c->comm, s));
        CKN(nccl().Send(c->idx.p, (size_t)pts_used * 8, ncclChar, root, c->comm, s));
    }
    if (want_mat && mat_used > 0) CKN(nccl().Send(c->mat.p, (size_t)mat_used * sizeof(mvs_match), ncclChar, root, c->comm, s));
    CKN(nccl().GroupEnd());
    int status = MVS_OK;
    if (is_root) {
        const int64_t pt = p_rank[c->world], mt = m_rank[c->world];
        if ((want_pts && pt > point_capacity) || (want_mat && mt > match_capacity)) status = MVS_E_CAPACITY;
        else {
            if (want_pts && pt > 0) {
                if (points) CKC(cudaMemcpyAsync(points, c->all_pts.p, (size_t)pt * 24, cudaMemcpyDeviceToHost, s));
                if (indexes) CKC(cudaMemcpyAsync(indexes, c->all_idx.p, (size_t)pt * 8, cudaMemcpyDeviceToHost, s));
            }
            if (want_mat && mt > 0 && matches) CKC(cudaMemcpyAsync(matches, c->all_mat.p, (size_t)mt * sizeof(mvs_match), cudaMemcpyDeviceToHost, s));
        }
    }
    CKC(cudaStreamSynchronize(s));
    return status;
}

}  // extern "C"
