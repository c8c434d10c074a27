// image-pair.hpp — ImagePair (reference source/front-end/image-pair.hpp:8-84, image-pair.cpp:30-237): construction
// (match + reconstruct) with its batched form, refine() and update().
#pragma once
#include <algorithm>
#include <memory>

#include "ba.hpp"
#include "camera.hpp"
#include "sfm.hpp"
#include "visual-feature.hpp"

namespace mvSLAM {

/** front-end/data-type.hpp:11-41: what the front end keeps per image (the debugging image is optional here). */
struct FrontEndTypes {
    using FrameId = Id::Type;
    struct Frame {
        FrameId id = static_cast<FrameId>(-1);          // Id::INVALID
        uint64_t capture_time = 0;                      // timestamp_us_t
        VisualFeature visual_feature;
        ImageGrayscale image;
        Frame() = default;
        Frame(FrameId id_, uint64_t capture_time_, const VisualFeature &vf, const ImageGrayscale &image_ = ImageGrayscale())
            : id(id_), capture_time(capture_time_), visual_feature(vf), image(image_) {}
    };
    using FramePtr = std::shared_ptr<const Frame>;      // "by default, no write access"
};
using Frame = FrontEndTypes::Frame;
using FramePtr = FrontEndTypes::FramePtr;

class ImagePair {
public:
    struct MatchedPoint {
        Point3 position;
        size_t vf_idx_in_base, vf_idx_in_pair;
        MatchedPoint(const Point3 &p, size_t viib, size_t viip) : position(p), vf_idx_in_base(viib), vf_idx_in_pair(viip) {}
    };
    struct Params {
        ScalarType max_match_inlier_distance = 10;      // image-pair.cpp:22-23
        bool refine_structure_in_constructor = false;   // image-pair.cpp:24-25
    };
    static Params get_default_params() { return Params(); }

    /** image-pair.hpp:38-40, the reference's own signature: K comes from CameraManager::get_camera()
     *  (image-pair.cpp:143), exactly what VisualOdometer::add_frame constructs (visual-odometer.cpp:140-148). */
    ImagePair(const FrontEndTypes::FramePtr &base_frame_, const FrontEndTypes::FramePtr &pair_frame_, const Params &params)
        : ImagePair(base_frame_, pair_frame_, CameraManager::get_camera().get_intrinsics(), params) {}

    /** The same with an explicit camera matrix (no global state). */
    ImagePair(const FramePtr &base_frame_, const FramePtr &pair_frame_, const CameraIntrinsics &K, const Params &params)
        : base_frame(base_frame_), pair_frame(pair_frame_)
    {
        if (!base_frame_ || !pair_frame_ || base_frame_->id == pair_frame_->id)
            throw b200::Error(MVS_E_BAD_ARG, "ImagePair: base and pair frame must differ (image-pair.cpp:49)");
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
        // a context of its own: the frame table a caller keeps resident on thread_default() is not replaced behind its back
        mvs_ctx *ctx = b200::Context::thread_scratch().get();
        const int nf = (int)frames.size(), np = (int)pairs.size();
        std::vector<const uint8_t *> dp(nf);
        std::vector<std::vector<float>> kp(nf);
        std::vector<const float *> kpp(nf);
        std::vector<int32_t> cnt(nf);
        int cap = 1;
        for (int f = 0; f < nf; ++f) {
            const auto &vf = frames[f]->visual_feature;
            dp[f] = b200::desc_data(vf.get_descriptors());
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
        b200::check(ctx, mvs_pair_batch(ctx, pr.data(), np, b200::rm3(K).data(), &mp, &b200::ransac_defaults(), res.data(), matches.data(),
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
                // the reference starts the sum at uint32_t(-1) ("deliberate overflow", image-pair.cpp:38) and adds sqr(distance)
                ip.match_inlier_ssd = (uint32_t)(0xFFFFFFFFu + (uint32_t)res[i].match_inlier_ssd);
                ip.T_pair_to_base = SE3(SO3(b200::mat3_from(res[i].R2in1)), Vector3Type(res[i].t2in1[0], res[i].t2in1[1], res[i].t2in1[2]));
                ip.matched_points.reserve(res[i].n_points);
                for (int j = 0; j < res[i].n_points; ++j) {
                    const double *p = &pts[((size_t)i * cap + j) * 3];
                    const mvs_match &m = matches[(size_t)i * cap + idx[(size_t)i * cap + j]];
                    ip.matched_points.emplace_back(Point3(p[0], p[1], p[2]), (size_t)m.train, (size_t)m.query);
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
    uint32_t match_inlier_ssd = 0xFFFFFFFFu;    // image-pair.cpp:38: (uint32_t)-1 until reconstruct() succeeds
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
