// reconstruct_scene — the flow of the reference's utility/reconstruct-scene.cpp:22-81 on top of the adapters:
//   image(1), image(2) -> VisualFeature::extract -> match_and_filter_visual_features -> sfm_solve -> print pose + points.
// Inputs ending in .pgm are 8-bit binary PGM images, extracted on the device like the reference does with OpenCV
// (reconstruct-scene.cpp:36-39); anything else is a pre-extracted feature file (include/mvslam/feature-io.hpp).
#include <cstdio>
#include <iostream>

#include <mvslam/feature-io.hpp>
#include <mvslam/sfm.hpp>

static void print_help(const char *cmdline)
{
    std::printf("Usage: %s <image_1.pgm|features_1> <image_2.pgm|features_2> <intrinsics> <max_dist> [nfeatures=500]\n", cmdline);
    std::printf("\tReconstruct scene using two images (or their pre-extracted features).\n");
}

int main(int argc, char **argv)
{
    if (argc != 5 && argc != 6) { print_help(argv[0]); return 1; }
    try {
        const int nfeatures = argc == 6 ? std::stoi(argv[5]) : mvSLAM::VisualFeature::MAX_FEATURE_COUNT;
        auto load = [&](const std::string &fn) {
            if (fn.size() > 4 && fn.substr(fn.size() - 4) == ".pgm") {
                std::vector<uint8_t> pixels;
                return mvSLAM::VisualFeature::extract(mvSLAM::load_pgm(fn, pixels), nfeatures);
            }
            return mvSLAM::load_visual_feature(fn);
        };
        auto vf1 = load(argv[1]);
        auto vf2 = load(argv[2]);
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
