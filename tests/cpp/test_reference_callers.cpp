// The two reference callers of the hot path, with their calls into ImagePair / VisualFeature / sfm_solve written EXACTLY
// as the reference writes them:
//   * VisualOdometer::add_frame (reference source/front-end/visual-odometer.cpp:129-148): camera loaded once through
//     CameraManager, then per new frame `m_image_pair_queue.emplace_back(prev_frame, new_frame, image_pair_params)`;
//   * utility/reconstruct-scene.cpp:36-53: extract x2, match_and_filter_visual_features, PinholeCamera(file), sfm_solve.
// Compiled twice by __graft_entry__.build(): against the POD stand-ins and, with MVSLAM_B200_WITH_EIGEN_OPENCV, against
// column-major Eigen-shaped matrices and cv-shaped containers (tests/cpp/mock_eigen_opencv/): the numbers must be the same.
// Input: a feature file pair exported by tools/export_features.py (argv[1], argv[2]) and a camera.config (argv[3]).
#include <cstdio>
#include <deque>
#include <fstream>
#include <memory>
#include <string>

#include "mvslam/image-pair.hpp"
#include "mvslam/feature-io.hpp"

using namespace mvSLAM;

namespace {
// the slice of VisualOdometer that owns the queues (visual-odometer.hpp:129-140), parameters as in visual-odometer.cpp:68-72
struct VisualOdometerSlice {
    struct Params { ScalarType max_match_inlier_distance = 10; size_t frame_queue_size = 10; } m_params;
    std::deque<FrontEndTypes::FramePtr> m_frame_queue;
    std::deque<ImagePair> m_image_pair_queue;

    void add_frame(const FrontEndTypes::FramePtr &new_frame)
    {
        // ---- verbatim from visual-odometer.cpp:137-149
        // update the frame queue
        m_frame_queue.push_back(new_frame);
        // create a new image pair from the last two frames
        if (m_frame_queue.size() > 1)
        {
            const auto &prev_frame = *(m_frame_queue.crbegin() + 1);
            // reconstruction only
            auto image_pair_params = ImagePair::get_default_params();
            image_pair_params.max_match_inlier_distance = m_params.max_match_inlier_distance;
            image_pair_params.refine_structure_in_constructor = false;
            m_image_pair_queue.emplace_back(prev_frame, new_frame, image_pair_params);
        }
        // ---- end of verbatim part
    }
};
}  // namespace

int main(int argc, char **argv)
{
    if (argc != 4) { std::fprintf(stderr, "usage: %s <features 1> <features 2> <camera.config>\n", argv[0]); return 2; }
    try {
        // --- visual-odometer path (utility/visual-odometer.cpp:67-68, then add_frame per image)
        CameraManager::load_from_file(argv[3]);
        VisualOdometerSlice vo;
        for (int i = 1; i <= 2; ++i)
            vo.add_frame(std::make_shared<const FrontEndTypes::Frame>((FrontEndTypes::FrameId)i, (uint64_t)i, load_visual_feature(argv[i])));
        const ImagePair &ip = vo.m_image_pair_queue.back();
        std::printf("vo: valid %d inliers %u ssd %u points %zu\n", (int)ip.valid, ip.match_inlier_count, ip.match_inlier_ssd, ip.matched_points.size());
        const Matrix3Type R = ip.T_pair_to_base.rotation().get_matrix();
        const Vector3Type t = ip.T_pair_to_base.translation();
        std::printf("vo: T_pair_to_base %.17g %.17g %.17g | %.17g %.17g %.17g | %.17g %.17g %.17g | %.17g %.17g %.17g\n", R(0, 0), R(0, 1), R(0, 2),
                    R(1, 0), R(1, 1), R(1, 2), R(2, 0), R(2, 1), R(2, 2), t[0], t[1], t[2]);
        if (ip.valid) std::printf("vo: first point %.17g %.17g %.17g idx %zu %zu\n", ip.matched_points[0].position[0], ip.matched_points[0].position[1],
                                  ip.matched_points[0].position[2], ip.matched_points[0].vf_idx_in_base, ip.matched_points[0].vf_idx_in_pair);

        // --- reconstruct-scene path, calls verbatim from utility/reconstruct-scene.cpp:40-53 (features come from files here)
        auto image1_vf = load_visual_feature(argv[1]);
        auto image2_vf = load_visual_feature(argv[2]);
        ScalarType max_dist = 30;
        std::string camera_intrinsics_fn(argv[3]);
        auto matched_vf_pair = mvSLAM::VisualFeature::match_and_filter_visual_features(
            image1_vf, image2_vf, max_dist);
        mvSLAM::PinholeCamera camera(camera_intrinsics_fn);

        // output
        mvSLAM::Transformation pose2in1_scaled;
        std::vector<mvSLAM::Point3> pointsin1_scaled;
        std::vector<size_t> point_indexes;
        if (!sfm_solve(matched_vf_pair.first.get_image_points(),
                       matched_vf_pair.second.get_image_points(),
                       camera.get_intrinsics(),
                       pose2in1_scaled,
                       pointsin1_scaled,
                       point_indexes))
        {
            std::printf("Reconstruction failed.\n");
            return 1;
        }
        const Vector3Type t2 = pose2in1_scaled.translation();
        std::printf("rs: matches %zu points %zu t %.17g %.17g %.17g first %.17g %.17g %.17g\n", matched_vf_pair.first.size(), pointsin1_scaled.size(),
                    t2[0], t2[1], t2[2], pointsin1_scaled[0][0], pointsin1_scaled[0][1], pointsin1_scaled[0][2]);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
