// visual_odometer_pairs — the hot-path part of the reference's utility/visual-odometer.cpp replay loop:
// for every new frame, the ImagePair constructions that VisualOdometer::add_frame issues
// (source/front-end/visual-odometer.cpp:140-148) and, while initialising, the re-pairings of the queued base frames
// with the new frame (visual-odometer.cpp:289-296 -> image-pair.cpp:77-113), submitted as ONE batch per frame.
// Input directory: camera.config plus either image.txt (one binary PGM file name per line, like the reference's image.txt:
// features are extracted on the device by VisualFeature::extract, frame-manager.cpp:114) or features.txt (one
// pre-extracted feature file per line, see include/mvslam/feature-io.hpp).  Optional 3rd argument: nfeatures.
#include <chrono>
#include <cstdio>
#include <fstream>

#include <mvslam/feature-io.hpp>
#include <mvslam/image-pair.hpp>

int main(int argc, char **argv)
{
    if (argc < 2) { std::printf("Usage: %s <input_directory> [frame_queue_size=10] [nfeatures=500]\n", argv[0]); return 1; }
    const std::string dir(argv[1]);
    const size_t queue_size = argc > 2 ? (size_t)std::stoi(argv[2]) : 10;   // visual-odometer.cpp:71-72
    const int nfeatures = argc > 3 ? std::stoi(argv[3]) : mvSLAM::VisualFeature::MAX_FEATURE_COUNT;
    try {
        const mvSLAM::CameraIntrinsics K = mvSLAM::load_camera_intrinsics(dir + "/camera.config");
        std::ifstream list(dir + "/image.txt");
        const bool from_images = list.good();
        if (!from_images) { list.close(); list.clear(); list.open(dir + "/features.txt"); }
        std::vector<mvSLAM::FramePtr> frames;
        std::vector<uint8_t> pixels;
        std::string name;
        const auto params = mvSLAM::ImagePair::get_default_params();   // max_match_inlier_distance = 10
        while (list >> name) {
            auto f = std::make_shared<mvSLAM::Frame>();
            f->id = frames.size();
            if (from_images) {
                const auto t0 = std::chrono::steady_clock::now();
                f->visual_feature = mvSLAM::VisualFeature::extract(mvSLAM::load_pgm(dir + "/" + name, pixels), nfeatures);
                const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
                std::printf("frame %zu: %zu keypoints extracted in %.0f us (incl. reading the file)\n", frames.size(), f->visual_feature.size(), us);
            } else {
                f->visual_feature = mvSLAM::load_visual_feature(dir + "/" + name);
            }
            frames.push_back(f);
            if (frames.size() < 2) continue;
            // (base_k, new) for the queued frames, newest base first == consecutive-frame pair first
            const int nf = (int)frames.size();
            std::vector<std::pair<int, int>> pairs;
            for (int b = nf - 2; b >= 0 && pairs.size() < queue_size; --b) pairs.emplace_back(b, nf - 1);
            const auto t0 = std::chrono::steady_clock::now();
            auto ips = mvSLAM::ImagePair::solve_batch(frames, pairs, K, params);
            const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
            const auto &ip = ips[0];
            std::printf("frame %d: %zu pair(s) in %.0f us; pair (%d,%d) valid=%d inliers=%u t=(%.6f %.6f %.6f)\n", nf - 1,
                        pairs.size(), us, pairs[0].first, pairs[0].second, (int)ip.valid, ip.match_inlier_count,
                        ip.T_pair_to_base.translation()[0], ip.T_pair_to_base.translation()[1], ip.T_pair_to_base.translation()[2]);
        }
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 2;
    }
    return 0;
}
