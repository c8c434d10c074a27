// reconstruct_scene — the flow of the reference's utility/reconstruct-scene.cpp:22-81 on top of the adapters:
//   features(1), features(2) -> match_and_filter_visual_features -> sfm_solve -> print pose + points.
// Image loading and ORB extraction are outside the hot path: the inputs are pre-extracted feature files
// (include/mvslam/feature-io.hpp; tools/export_features.py writes them with cv2.ORB).
#include <cstdio>
#include <iostream>

#include <mvslam/feature-io.hpp>
#include <mvslam/sfm.hpp>

static void print_help(const char *cmdline)
{
    std::printf("Usage: %s <features_1> <features_2> <intrinsics> <max_dist>\n", cmdline);
    std::printf("\tReconstruct scene using the features of two images.\n");
}

int main(int argc, char **argv)
{
    if (argc != 5) { print_help(argv[0]); return 1; }
    try {
        auto vf1 = mvSLAM::load_visual_feature(argv[1]);
        auto vf2 = mvSLAM::load_visual_feature(argv[2]);
        const mvSLAM::CameraIntrinsics K = mvSLAM::load_camera_intrinsics(argv[3]);
        const mvSLAM::ScalarType max_dist = std::stoi(std::string(argv[4]));
        auto matched = mvSLAM::VisualFeature::match_and_filter_visual_features(vf1, vf2, max_dist);
        mvSLAM::Transformation pose2in1_scaled;
        std::vector<mvSLAM::Point3> pointsin1_scaled;
        std::vector<size_t> point_indexes;
        if (!mvSLAM::sfm_solve(matched.first.get_image_points(), matched.second.get_image_points(), K, pose2in1_scaled,
                               pointsin1_scaled, point_indexes)) {
            std::printf("Reconstruction failed.\n");
            return 3;
        }
        std::printf("matches = %zu\n", matched.first.size());
        const auto &R = pose2in1_scaled.rotation().get_matrix();
        const auto &t = pose2in1_scaled.translation();
        std::printf("scaled transformation =\n");
        for (int i = 0; i < 3; ++i) std::printf("%.9f %.9f %.9f | %.9f\n", R(i, 0), R(i, 1), R(i, 2), t[i]);
        std::printf("pointsin1_scaled = %zu\n", pointsin1_scaled.size());
        for (const auto &p : pointsin1_scaled) std::printf("%.6f, %.6f, %.6f\n", p[0], p[1], p[2]);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 2;
    }
    return 0;
}
