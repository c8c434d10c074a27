// latency_probe — what a C++ caller (the reference's VisualOdometer::add_frame, visual-odometer.cpp:140-148) pays for one
// mvs_pair_batch call on resident frames: wall time of the synchronous call, of the enqueue alone (host issue time), and
// of enqueue + synchronize, for 1 and 10 pairs, REFERENCE H=1 and FAST H=1024.  Prints one JSON object.
//   latency_probe <feature_dir>   (tools/export_features.py npz tests/golden/tsukuba_orb2000.npz <dir>)
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <fstream>
#include <vector>

#include <mvslam/feature-io.hpp>
#include <mvslam_b200.h>

static double now_us()
{
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static double median(std::vector<double> v)
{
    std::sort(v.begin(), v.end());
    return v[v.size() / 2];
}

int main(int argc, char **argv)
{
    if (argc < 2) { std::printf("Usage: %s <feature_dir>\n", argv[0]); return 1; }
    const std::string dir(argv[1]);
    try {
        const mvSLAM::CameraIntrinsics Kc = mvSLAM::load_camera_intrinsics(dir + "/camera.config");
        double K[9];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) K[r * 3 + c] = Kc(r, c);
        std::ifstream list(dir + "/features.txt");
        std::vector<mvSLAM::VisualFeature> feats;
        std::string name;
        while (list >> name) feats.push_back(mvSLAM::load_visual_feature(dir + "/" + name));
        if (feats.size() < 5) { std::fprintf(stderr, "need 5 frames\n"); return 2; }
        mvs_ctx *ctx = nullptr;
        if (mvs_create(&ctx, -1) != MVS_OK) { std::fprintf(stderr, "mvs_create failed\n"); return 2; }
        std::vector<std::vector<float>> kp(feats.size());
        std::vector<const uint8_t *> dp;
        std::vector<const float *> kpp;
        std::vector<int32_t> counts;
        int cap = 0;
        for (size_t f = 0; f < feats.size(); ++f) {
            const auto &kps = feats[f].get_keypoints();
            for (const auto &k : kps) { kp[f].push_back(k.pt.x); kp[f].push_back(k.pt.y); }
            dp.push_back(mvSLAM::b200::desc_data(feats[f].get_descriptors()));
            kpp.push_back(kp[f].data());
            counts.push_back((int32_t)kps.size());
            cap = std::max(cap, (int)kps.size());
        }
        if (mvs_frames_upload(ctx, (int)feats.size(), dp.data(), kpp.data(), counts.data(), 32) != MVS_OK) return 3;
        std::printf("{");
        bool first = true;
        for (int n_pairs : {1, 10})
            for (int cfg = 0; cfg < 2; ++cfg) {
                std::vector<int32_t> pairs;
                for (int i = 0; i < n_pairs; ++i) { pairs.push_back(n_pairs == 1 ? 0 : i % 4); pairs.push_back(n_pairs == 1 ? 1 : 4); }
                mvs_match_params mp{};
                mp.ratio = 0.7; mp.max_dist = 10.0;
                mvs_ransac_params rp{};
                rp.n_hypotheses = cfg == 0 ? 1 : 1024; rp.solver = cfg == 0 ? MVS_SOLVER_REFERENCE : MVS_SOLVER_FAST;
                std::vector<mvs_pair_result> res(n_pairs);
                std::vector<mvs_match> matches((size_t)n_pairs * cap);
                std::vector<uint8_t> mask((size_t)n_pairs * cap);
                std::vector<double> points((size_t)n_pairs * cap * 3);
                std::vector<uint64_t> indexes((size_t)n_pairs * cap);
                std::vector<double> t_sync, t_enq, t_total, t_rec;
                for (int it = 0; it < 240; ++it) {
                    double t0 = now_us();
                    int st = mvs_pair_batch(ctx, pairs.data(), n_pairs, K, &mp, &rp, res.data(), matches.data(), mask.data(), points.data(),
                                            indexes.data(), cap);
                    double t1 = now_us();
                    if (st != MVS_OK) { std::fprintf(stderr, "pair_batch: %s\n", mvs_last_error(ctx)); return 4; }
                    st = mvs_pair_batch_enqueue(ctx, pairs.data(), n_pairs, K, &mp, &rp, res.data(), matches.data(), mask.data(),
                                                points.data(), indexes.data(), cap);
                    double t2 = now_us();
                    mvs_synchronize(ctx);
                    double t3 = now_us();
                    st = mvs_pair_batch(ctx, pairs.data(), n_pairs, K, &mp, &rp, res.data(), nullptr, nullptr, nullptr, nullptr, 0);
                    double t4 = now_us();
                    if (it >= 40) { t_sync.push_back(t1 - t0); t_enq.push_back(t2 - t1); t_total.push_back(t3 - t1); t_rec.push_back(t4 - t3); }
                }
                std::printf("%s\"%d_pair%s_%s\": {\"call_us\": %.1f, \"enqueue_us\": %.1f, \"enqueue_plus_sync_us\": %.1f, \"records_only_call_us\": %.1f, "
                            "\"n_inliers\": %d, \"n_points\": %d}",
                            first ? "" : ", ", n_pairs, n_pairs > 1 ? "s" : "", cfg == 0 ? "reference_h1" : "fast_h1024", median(t_sync),
                            median(t_enq), median(t_total), median(t_rec), res[0].n_inliers, res[0].n_points);
                first = false;
            }
        // the bench's end-to-end step from a C++ caller: every step uploads the five frames again and makes one synchronous
        // 1024-pair call (the four consecutive pairs cycled) with all detail outputs, REFERENCE solver, H = 1, every buffer pinned
        // (mvs_host_alloc): host wall clock per step
        {
            const int n_pairs = 1024, ecap = 256;
            struct Pinned {
                void *p;
                explicit Pinned(size_t b) : p(mvs_host_alloc(b)) {}
                ~Pinned() { mvs_host_free(p); }
            };
            std::vector<Pinned *> keep;
            std::vector<const uint8_t *> pdp;
            std::vector<const float *> pkp;
            for (size_t f = 0; f < feats.size(); ++f) {
                Pinned *d = new Pinned((size_t)counts[f] * 32), *k = new Pinned((size_t)counts[f] * 8);
                if (!d->p || !k->p) { std::fprintf(stderr, "mvs_host_alloc failed\n"); return 5; }
                std::copy(dp[f], dp[f] + (size_t)counts[f] * 32, static_cast<uint8_t *>(d->p));
                std::copy(kp[f].begin(), kp[f].end(), static_cast<float *>(k->p));
                pdp.push_back(static_cast<const uint8_t *>(d->p)); pkp.push_back(static_cast<const float *>(k->p));
                keep.push_back(d); keep.push_back(k);
            }
            Pinned res((size_t)n_pairs * sizeof(mvs_pair_result)), mat((size_t)n_pairs * ecap * sizeof(mvs_match)), msk((size_t)n_pairs * ecap),
                pts((size_t)n_pairs * ecap * 24), idx((size_t)n_pairs * ecap * 8);
            if (!res.p || !mat.p || !msk.p || !pts.p || !idx.p) { std::fprintf(stderr, "mvs_host_alloc failed\n"); return 5; }
            std::vector<int32_t> pairs;
            for (int i = 0; i < n_pairs; ++i) { pairs.push_back(i % 4); pairs.push_back(i % 4 + 1); }
            mvs_match_params mp{};
            mp.ratio = 0.7; mp.max_dist = 10.0;
            mvs_ransac_params rp{};
            rp.n_hypotheses = 1; rp.solver = MVS_SOLVER_REFERENCE;
            std::vector<double> t_step;
            int solved = 0;
            for (int it = 0; it < 60; ++it) {
                const double t0 = now_us();
                if (mvs_frames_upload(ctx, (int)feats.size(), pdp.data(), pkp.data(), counts.data(), 32) != MVS_OK) return 3;
                const int st = mvs_pair_batch(ctx, pairs.data(), n_pairs, K, &mp, &rp, static_cast<mvs_pair_result *>(res.p),
                                              static_cast<mvs_match *>(mat.p), static_cast<uint8_t *>(msk.p), static_cast<double *>(pts.p),
                                              static_cast<uint64_t *>(idx.p), ecap);
                const double t1 = now_us();
                if (st != MVS_OK) { std::fprintf(stderr, "pair_batch: %s\n", mvs_last_error(ctx)); return 4; }
                if (it >= 10) t_step.push_back(t1 - t0);
            }
            for (int i = 0; i < n_pairs; ++i) solved += static_cast<mvs_pair_result *>(res.p)[i].status == MVS_OK;
            std::sort(t_step.begin(), t_step.end());
            std::printf(", \"e2e_1024_pairs_reference_h1\": {\"step_us_median\": %.1f, \"step_us_min\": %.1f, \"step_us_max\": %.1f, "
                        "\"pairs_per_s\": %.0f, \"solved\": %d}",
                        t_step[t_step.size() / 2], t_step.front(), t_step.back(), n_pairs / (t_step[t_step.size() / 2] * 1e-6), solved);
            for (Pinned *q : keep) delete q;
        }
        std::printf("}\n");
        mvs_destroy(ctx);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 2;
    }
    return 0;
}
