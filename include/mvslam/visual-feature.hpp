// visual-feature.hpp — VisualFeature with the reference's matching entry points
// (reference source/vision/visual-feature.hpp:14-93), forwarding to the C ABI.
#pragma once
#include <utility>

#include "types.hpp"

namespace mvSLAM {

class VisualFeature {
public:
    VisualFeature() = default;
    /** keypoints + 32-byte descriptors as produced by cv::ORB (e.g. extracted elsewhere) */
    VisualFeature(VisualFeatureConfig::DetectorResultType kps, VisualFeatureConfig::ExtractorResultType desc,
                  int image_width, int image_height)
        : m_keypoints(std::move(kps)), m_descriptors(std::move(desc)), m_image_width(image_width), m_image_height(image_height)
    {
        if (!b200::desc_ok(m_descriptors, m_keypoints.size())) throw b200::Error(MVS_E_BAD_ARG, "descriptor rows must be 32 contiguous bytes");
    }

    static constexpr int MAX_FEATURE_COUNT = 500;   // visual-feature.cpp:9

    /** visual-feature.hpp:14 / visual-feature.cpp:40-49: cv::ORB::create(MAX_FEATURE_COUNT) detect + compute, on the
     *  device.  Same keypoint set, responses, angles and descriptor bytes as cv::ORB; order is (level, y, x). */
    static VisualFeature extract(const ImageGrayscale &image, int n_features = MAX_FEATURE_COUNT)
    {
        return std::move(extract_batch({image}, n_features)[0]);
    }

    /** Many same-size images in one device pass (what FrameManager::add_frame does per image,
     *  source/front-end/frame-manager.cpp:107-125, for a whole window at once). */
    static std::vector<VisualFeature> extract_batch(const std::vector<ImageGrayscale> &images, int n_features = MAX_FEATURE_COUNT)
    {
        if (images.empty()) return {};
        mvs_ctx *ctx = b200::Context::thread_default().get();
        const int n = (int)images.size(), w = images[0].cols, h = images[0].rows;
        const size_t step = images[0].step;
        std::vector<const uint8_t *> ptrs(n);
        for (int i = 0; i < n; ++i) {
            if (images[i].cols != w || images[i].rows != h || images[i].step != step || !images[i].data)
                throw b200::Error(MVS_E_BAD_ARG, "extract_batch: images must share one size and stride");
            ptrs[i] = images[i].data;
        }
        const mvs_orb_params op{n_features, {0, 0, 0}};
        std::vector<int32_t> counts(n);
        int64_t cap = (int64_t)n * (n_features + 64);
        std::vector<mvs_keypoint> kp((size_t)cap);
        std::vector<uint8_t> desc((size_t)cap * 32);
        int st = mvs_orb_extract(ctx, ptrs.data(), n, w, h, (int)step, &op, 0, nullptr, counts.data(), kp.data(), desc.data(), cap);
        if (st == MVS_E_CAPACITY) {   // many keypoints tied at a cut-off response: counts[] holds the exact sizes
            cap = 0;
            for (int c : counts) cap += c;
            if (cap > (int64_t)kp.size()) {
                kp.resize((size_t)cap); desc.resize((size_t)cap * 32);
                st = mvs_orb_extract(ctx, ptrs.data(), n, w, h, (int)step, &op, 0, nullptr, counts.data(), kp.data(), desc.data(), cap);
            }
        }
        b200::check(ctx, st, "VisualFeature::extract");
        std::vector<VisualFeature> out(n);
        size_t at = 0;
        for (int i = 0; i < n; ++i) {
            VisualFeature &vf = out[i];
            vf.m_image_width = w; vf.m_image_height = h;
            vf.m_keypoints.resize(counts[i]);
            for (int k = 0; k < counts[i]; ++k) {
                const mvs_keypoint &s = kp[at + k];
                KeyPoint &d = vf.m_keypoints[k];
                d.pt.x = s.x; d.pt.y = s.y; d.size = s.size; d.angle = s.angle; d.response = s.response; d.octave = s.octave;
            }
            vf.m_descriptors = b200::desc_make(desc.data() + at * 32, (size_t)counts[i]);
            at += (size_t)counts[i];
        }
        return out;
    }

    /** Matches from 2 to 1 (query = vf2, train = vf1) sorted by ascending distance
     *  (visual-feature.hpp:23-26, visual-feature.cpp:51-80); empty when nothing survives. */
    static VisualFeatureConfig::MatchResultType match_visual_features(const VisualFeature &vf1, const VisualFeature &vf2,
                                                                       ScalarType max_dist = -1)
    {
        if (!vf1.valid() || !vf2.valid()) throw b200::Error(MVS_E_BAD_ARG, "invalid VisualFeature");  // :56 assert
        mvs_ctx *ctx = b200::Context::thread_default().get();
        std::vector<mvs_match> out(vf2.size());
        int n = 0;
        const mvs_match_params mp{0.7, max_dist, 0, 1};   // bounded: identical matches, less work when max_dist >= 0
        int st = mvs_match_hamming(ctx, b200::desc_data(vf2.m_descriptors), (int)vf2.size(), b200::desc_data(vf1.m_descriptors), (int)vf1.size(),
                                   32, &mp, out.data(), (int)out.size(), &n);
        b200::check(ctx, st, "match_visual_features");
        VisualFeatureConfig::MatchResultType r(n);
        for (int i = 0; i < n; ++i) { r[i].queryIdx = out[i].query; r[i].trainIdx = out[i].train; r[i].distance = out[i].distance; }
        return r;
    }

    /** visual-feature.hpp:45-48 / visual-feature.cpp:93-119 (keypoints gathered; the reference's descriptor copy
     *  is buggy (:115) and unused downstream, so descriptors are gathered correctly here). */
    static std::pair<VisualFeature, VisualFeature> match_and_filter_visual_features(const VisualFeature &vf1,
                                                                                     const VisualFeature &vf2,
                                                                                     ScalarType max_dist = -1)
    {
        auto matches = match_visual_features(vf1, vf2, max_dist);
        VisualFeature f1, f2;
        f1.m_image_width = f2.m_image_width = vf1.m_image_width;
        f1.m_image_height = f2.m_image_height = vf1.m_image_height;
        std::vector<uint8_t> d1(matches.size() * 32), d2(matches.size() * 32);
        const uint8_t *s1 = b200::desc_data(vf1.m_descriptors), *s2 = b200::desc_data(vf2.m_descriptors);
        size_t r = 0;
        for (const auto &m : matches) {
            f1.m_keypoints.push_back(vf1.m_keypoints[m.trainIdx]);
            f2.m_keypoints.push_back(vf2.m_keypoints[m.queryIdx]);
            for (int k = 0; k < 32; ++k) { d1[r * 32 + k] = s1[32 * (size_t)m.trainIdx + k]; d2[r * 32 + k] = s2[32 * (size_t)m.queryIdx + k]; }
            ++r;
        }
        f1.m_descriptors = b200::desc_make(d1.data(), matches.size());
        f2.m_descriptors = b200::desc_make(d2.data(), matches.size());
        return std::make_pair(f1, f2);
    }

    size_t size() const { return m_keypoints.size(); }
    const VisualFeatureConfig::DetectorResultType &get_keypoints() const { return m_keypoints; }
    const VisualFeatureConfig::ExtractorResultType &get_descriptors() const { return m_descriptors; }
    std::vector<ImagePoint> get_image_points() const   // visual-feature.cpp:179-190
    {
        std::vector<ImagePoint> r;
        r.reserve(m_keypoints.size());
        for (const auto &kp : m_keypoints) r.emplace_back(kp.pt.x, kp.pt.y);
        return r;
    }
    /** visual-feature.cpp:192-207: ORB keypoint uncertainty, standard deviation 2^octave * 0.5 pixels */
    std::vector<Point2Estimate> get_point_estimates() const
    {
        std::vector<Point2Estimate> r;
        r.reserve(m_keypoints.size());
        for (const auto &kp : m_keypoints) {
            const ScalarType sd = static_cast<ScalarType>(1 << kp.octave) * 0.5;
            Point2 mu; mu[0] = kp.pt.x; mu[1] = kp.pt.y;
            Matrix2Type C = Matrix2Type::Zero(); C(0, 0) = C(1, 1) = sd * sd;
            r.emplace_back(mu, C);
        }
        return r;
    }
    bool valid() const { return size() > 0 && m_image_width > 0 && m_image_height > 0; }

private:
    VisualFeatureConfig::DetectorResultType m_keypoints;
    VisualFeatureConfig::ExtractorResultType m_descriptors;
    int m_image_width = -1, m_image_height = -1;
};

}  // namespace mvSLAM
