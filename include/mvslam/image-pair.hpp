// image-pair.hpp — ImagePair (reference source/front-end/image-pair.hpp:8-84, image-pair.cpp:30-237): construction
// (match + reconstruct) with its batched form, refine() and update().
#pragma once
#include <algorithm>
#include <memory>

#include "ba.hpp"
#include "sfm.hpp"
#include "visual-feature.hpp"

namespace mvSLAM {

struct Frame {   // the fields of front-end/data-type.hpp:19-37 the path reads
    size_t id = static_cast<size_t>(-1);
    VisualFeature visual_feature;
};
using FramePtr = std::shared_ptr<Frame>;

class ImagePair {
public:
    struct MatchedPoint {
        Point3 position;
        size_t vf_idx_in_base, vf_idx_in_pair;
    };
    struct Params {
        ScalarType max_match_inlier_distance = 10;      // image-pair.cpp:22-23
        bool refine_structure_in_constructor = false;   // image-pair.cpp:24-25
    };
    static Params get_default_params() { return Params(); }

    /** image-pair.hpp:38-40: match + reconstruct for one pair (a batch of one). */
    ImagePair(const FramePtr &base_frame_, const FramePtr &pair_frame_, const CameraIntrinsics &K, const Params &params)
        : base_frame(base_frame_), pair_frame(pair_frame_)
    {
        std::vector<ImagePair> one = solve_batch({base_frame_, pair_frame_}, {{0, 1}}, K, params);
        *this = std::move(one[0]);
        if (valid && params.refine_structure_in_constructor) refine();   // image-pair.cpp:62-70
    }

    /** image-pair.cpp:176-237: two-view bundle adjustment of the pair pose and the matched points (sfm_refine). */
    bool refine()
    {
        if (!valid) throw b200::Error(MVS_E_BAD_ARG, "ImagePair::refine: invalid pair (image-pair.cpp:178)");
        const auto e1 = base_frame->visual_feature.get_point_estimates(), e2 = pair_frame->visual_feature.get_point_estimates();
        std::vector<Point2Estimate> b, p;
        std::vector<Point3> pts;
        for (const auto &mp : matched_points) { b.push_back(e1[mp.vf_idx_in_base]); p.push_back(e2[mp.vf_idx_in_pair]); pts.push_back(mp.position); }
        TransformationEstimate T;
        std::vector<Point3Estimate> pe;
        valid = sfm_refine(b, p, m_K, T_pair_to_base, pts, T, pe, error);
        if (valid) {
            T_pair_to_base = T.mean();
            T_pair_to_base_covar = T.covar();
            matched_points_covar.clear();
            for (size_t i = 0; i < matched_points.size(); ++i) { matched_points[i].position = pe[i].mean(); matched_points_covar.push_back(pe[i].covar()); }
            refined = true;
        }
        return valid;
    }

    /** image-pair.cpp:77-113: does @p new_frame make a better pair with base_frame?  Replaces *this if so. */
    bool update(const FramePtr &new_frame)
    {
        if (new_frame->id == base_frame->id || new_frame->id == pair_frame->id) return false;
        Params light = m_params;
        light.refine_structure_in_constructor = false;
        ImagePair candidate(base_frame, new_frame, m_K, light);
        if (!candidate.valid) return false;
        if (candidate.match_inlier_count < match_inlier_count || candidate.match_inlier_ssd < match_inlier_ssd) return false;
        candidate.refine();
        if (candidate.error < error) { std::swap(*this, candidate); return true; }
        return false;
    }

    /** Many (base, pair) constructions in one device pass: frames[i] uploaded once, pairs index into it. */
    static std::vector<ImagePair> solve_batch(const std::vector<FramePtr> &frames,
                                              const std::vector<std::pair<int, int>> &pairs, const CameraIntrinsics &K,
                                              const Params &params)
    {
        mvs_ctx *ctx = b200::Context::thread_default().get();
        const int nf = (int)frames.size(), np = (int)pairs.size();
        std::vector<const uint8_t *> dp(nf);
        std::vector<std::vector<float>> kp(nf);
        std::vector<const float *> kpp(nf);
        std::vector<int32_t> cnt(nf);
        int cap = 1;
        for (int f = 0; f < nf; ++f) {
            const auto &vf = frames[f]->visual_feature;
            dp[f] = vf.get_descriptors().data();
            cnt[f] = (int32_t)vf.size();
            kp[f].resize(vf.size() * 2);
            for (size_t i = 0; i < vf.size(); ++i) { kp[f][2 * i] = vf.get_keypoints()[i].pt.x; kp[f][2 * i + 1] = vf.get_keypoints()[i].pt.y; }
            kpp[f] = kp[f].data();
            cap = std::max(cap, cnt[f]);
        }
        b200::check(ctx, mvs_frames_upload(ctx, nf, dp.data(), kpp.data(), cnt.data(), 32), "ImagePair: frames_upload");
        std::vector<int32_t> pr((size_t)np * 2);
        for (int i = 0; i < np; ++i) { pr[2 * i] = pairs[i].first; pr[2 * i + 1] = pairs[i].second; }
        std::vector<mvs_pair_result> res(np);
        std::vector<mvs_match> matches((size_t)np * cap);
        std::vector<double> pts((size_t)np * cap * 3);
        std::vector<uint64_t> idx((size_t)np * cap);
        const mvs_match_params mp{0.7, params.max_match_inlier_distance, 0, 1};
        b200::check(ctx, mvs_pair_batch(ctx, pr.data(), np, K.m, &mp, &b200::ransac_defaults(), res.data(), matches.data(),
                                        nullptr, pts.data(), idx.data(), cap), "ImagePair: pair_batch");
        std::vector<ImagePair> out;
        out.reserve(np);
        for (int i = 0; i < np; ++i) {
            ImagePair ip;
            ip.m_K = K; ip.m_params = params;
            ip.base_frame = frames[pairs[i].first];
            ip.pair_frame = frames[pairs[i].second];
            ip.status = res[i].status;
            ip.valid = (res[i].status == MVS_OK);
            if (ip.valid) {   // image-pair.cpp:158-167 (with points[] indexed by position, not by original index)
                ip.match_inlier_count = (uint32_t)res[i].n_points;
                ip.match_inlier_ssd = (uint32_t)res[i].match_inlier_ssd;
                Matrix3Type R;
                for (int k = 0; k < 9; ++k) R.m[k] = res[i].R2in1[k];
                ip.T_pair_to_base = SE3(SO3(R), Vector3Type(res[i].t2in1[0], res[i].t2in1[1], res[i].t2in1[2]));
                ip.matched_points.reserve(res[i].n_points);
                for (int j = 0; j < res[i].n_points; ++j) {
                    const double *p = &pts[((size_t)i * cap + j) * 3];
                    const mvs_match &m = matches[(size_t)i * cap + idx[(size_t)i * cap + j]];
                    ip.matched_points.push_back({Point3(p[0], p[1], p[2]), (size_t)m.train, (size_t)m.query});
                }
            }
            out.push_back(std::move(ip));
        }
        return out;
    }

    FramePtr base_frame, pair_frame;
    bool valid = false;
    int status = MVS_E_BAD_ARG;
    uint32_t match_inlier_count = 0;
    uint32_t match_inlier_ssd = 0;
    Transformation T_pair_to_base;
    std::vector<MatchedPoint> matched_points;
    // available after refine() (image-pair.hpp:66-69)
    bool refined = false;
    ScalarType error = infinity;
    TransformationUncertainty T_pair_to_base_covar;
    std::vector<Point3Uncertainty> matched_points_covar;

private:
    ImagePair() = default;
    CameraIntrinsics m_K;
    Params m_params;
};

}  // namespace mvSLAM
