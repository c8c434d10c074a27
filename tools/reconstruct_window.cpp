// reconstruct_window — all-pairs matching + two-view geometry over a window of frames, the pairs sharded across the GPUs of
// one box (BASELINE config 5; the per-pair unit is utility/reconstruct-scene.cpp:36-53 of the reference).  Host code is C++
// over the C ABI: one process per GPU, rank / world size from the command line or from RANK / WORLD_SIZE (torchrun, mpirun
// -x), the ncclUniqueId passed through a file.  Rank 0 prints one line per solved pair count and the cloud size.
//
//   reconstruct_window <dir with 1.mvsf .. N.mvsf + camera.config> <n_frames> <id file> [rank world] [max_dist] [H]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

#include "mvslam/camera.hpp"
#include "mvslam/feature-io.hpp"

int main(int argc, char **argv)
{
    if (argc < 4) { std::fprintf(stderr, "usage: %s <dir> <n_frames> <id file> [rank world] [max_dist] [H]\n", argv[0]); return 2; }
    const std::string dir = argv[1], idf = argv[3];
    const int nf = std::atoi(argv[2]);
    int rank = std::getenv("RANK") ? std::atoi(std::getenv("RANK")) : 0, world = std::getenv("WORLD_SIZE") ? std::atoi(std::getenv("WORLD_SIZE")) : 1;
    if (argc > 5) { rank = std::atoi(argv[4]); world = std::atoi(argv[5]); }
    const double max_dist = argc > 6 ? std::atof(argv[6]) : -1.0;
    const int H = argc > 7 ? std::atoi(argv[7]) : 1;
    try {
        using namespace mvSLAM;
        int n_dev = 1;
        mvs_ctx *ctx = nullptr;
        (void)n_dev;
        if (mvs_create(&ctx, std::getenv("LOCAL_RANK") ? std::atoi(std::getenv("LOCAL_RANK")) : rank) != MVS_OK) { std::fprintf(stderr, "mvs_create failed (no CPU fallback)\n"); return 1; }
        // the communicator: rank 0 makes the id, the others read it from the file
        uint8_t id[128];
        if (rank == 0) {
            b200::check(ctx, mvs_comm_unique_id(id), "mvs_comm_unique_id");
            std::ofstream(idf + ".tmp", std::ios::binary).write((const char *)id, 128);
            std::rename((idf + ".tmp").c_str(), idf.c_str());
        } else {
            for (int t = 0; t < 1200; ++t) {
                std::ifstream in(idf, std::ios::binary);
                if (in.read((char *)id, 128)) break;
                std::this_thread::sleep_for(std::chrono::milliseconds(50));
            }
        }
        mvs_comm *comm = nullptr;
        int st = mvs_comm_create(&comm, ctx, id, rank, world);
        if (st != MVS_OK) { std::fprintf(stderr, "mvs_comm_create: %s\n", mvs_last_error(ctx)); return 1; }
        // every rank loads the window (replicated frame table) and the full pair list
        std::vector<VisualFeature> vf;
        for (int i = 1; i <= nf; ++i) vf.push_back(load_visual_feature(dir + "/" + std::to_string(i) + ".mvsf"));
        const PinholeCamera camera(dir + "/camera.config");
        std::vector<const uint8_t *> dp(nf);
        std::vector<std::vector<float>> kp(nf);
        std::vector<const float *> kpp(nf);
        std::vector<int32_t> cnt(nf);
        int cap = 1;
        for (int f = 0; f < nf; ++f) {
            dp[f] = b200::desc_data(vf[f].get_descriptors()); cnt[f] = (int32_t)vf[f].size(); cap = std::max(cap, cnt[f]);
            for (const auto &k : vf[f].get_keypoints()) { kp[f].push_back(k.pt.x); kp[f].push_back(k.pt.y); }
            kpp[f] = kp[f].data();
        }
        b200::check(ctx, mvs_frames_upload(ctx, nf, dp.data(), kpp.data(), cnt.data(), 32), "frames_upload");
        std::vector<int32_t> pairs;
        for (int a = 0; a < nf; ++a) for (int b = a + 1; b < nf; ++b) { pairs.push_back(a); pairs.push_back(b); }
        const int64_t np = (int64_t)pairs.size() / 2;
        std::vector<mvs_pair_result> res(rank == 0 ? np : 0);
        std::vector<int64_t> po(np + 1), mo(np + 1);
        std::vector<double> pts(rank == 0 ? (size_t)np * cap * 3 : 0);
        std::vector<uint64_t> idx(rank == 0 ? (size_t)np * cap : 0);
        const mvs_match_params mp{0.7, max_dist, 0, 0};
        const mvs_ransac_params rp{H, MVS_SCORE_ALGEBRAIC, 0.0, 0, 0, MVS_SOLVER_REFERENCE, 0};
        const auto t0 = std::chrono::steady_clock::now();
        st = mvs_pair_batch_sharded(ctx, comm, pairs.data(), np, b200::rm3(camera.get_intrinsics()).data(), &mp, &rp, 0,
                                    rank == 0 ? res.data() : nullptr, po.data(), rank == 0 ? pts.data() : nullptr,
                                    rank == 0 ? idx.data() : nullptr, rank == 0 ? np * cap : 0, nullptr, nullptr, 0);
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (st != MVS_OK) { std::fprintf(stderr, "rank %d: mvs_pair_batch_sharded: %s: %s\n", rank, mvs_status_string(st), mvs_last_error(ctx)); return 1; }
        if (rank == 0) {
            int ok = 0;
            for (const auto &r : res) ok += r.status == MVS_OK;
            std::printf("ranks = %d, pairs = %lld, solved = %d, cloud points = %lld, %.2f ms\n", world, (long long)np, ok, (long long)po[np], ms);
            for (int64_t i = 0; i < np && i < 4; ++i)
                std::printf("pair (%d,%d): status %d matches %d points %d t = %.6f %.6f %.6f first point %.6f %.6f %.6f\n", pairs[2 * i] + 1, pairs[2 * i + 1] + 1,
                            res[i].status, res[i].n_matches, res[i].n_points, res[i].t2in1[0], res[i].t2in1[1], res[i].t2in1[2],
                            po[i + 1] > po[i] ? pts[3 * po[i]] : 0.0, po[i + 1] > po[i] ? pts[3 * po[i] + 1] : 0.0, po[i + 1] > po[i] ? pts[3 * po[i] + 2] : 0.0);
        }
        mvs_comm_destroy(comm);
        mvs_destroy(ctx);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
