// feature-io.hpp — file helpers of the C++ tools, so that they run without OpenCV: a tiny binary container for
// pre-extracted features, binary PGM (P5) images for VisualFeature::extract, and the reference's camera file.
//   feature file = "MVSF" | int32 n | int32 width | int32 height | n x (float x, float y) | n x 32 descriptor bytes
#pragma once
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>

#include "visual-feature.hpp"

namespace mvSLAM {

inline VisualFeature load_visual_feature(const std::string &filename)
{
    std::ifstream in(filename, std::ios::binary);
    char magic[4];
    int32_t n = 0, w = 0, h = 0;
    in.read(magic, 4);
    in.read(reinterpret_cast<char *>(&n), 4); in.read(reinterpret_cast<char *>(&w), 4); in.read(reinterpret_cast<char *>(&h), 4);
    if (!in || std::string(magic, 4) != "MVSF" || n < 0) throw b200::Error(MVS_E_BAD_ARG, "bad feature file " + filename);
    VisualFeatureConfig::DetectorResultType kps(n);
    std::vector<float> xy((size_t)n * 2);
    in.read(reinterpret_cast<char *>(xy.data()), (std::streamsize)(xy.size() * sizeof(float)));
    for (int i = 0; i < n; ++i) { kps[i].pt.x = xy[2 * i]; kps[i].pt.y = xy[2 * i + 1]; }
    std::vector<uint8_t> desc((size_t)n * 32);
    in.read(reinterpret_cast<char *>(desc.data()), (std::streamsize)desc.size());
    if (!in) throw b200::Error(MVS_E_BAD_ARG, "truncated feature file " + filename);
    return VisualFeature(std::move(kps), b200::desc_make(desc.data(), (size_t)n), w, h);
}

/** 8-bit binary PGM (P5, maxval 255) -> pixels; the returned ImageGrayscale points into `pixels`. */
inline ImageGrayscale load_pgm(const std::string &filename, std::vector<uint8_t> &pixels)
{
    std::ifstream in(filename, std::ios::binary);
    std::string magic;
    int w = 0, h = 0, maxval = 0;
    auto next_token = [&](std::string &tok) {
        tok.clear();
        char c;
        while (in.get(c)) {
            if (c == '#') { while (in.get(c) && c != '\n') {} continue; }
            if (!std::isspace((unsigned char)c)) { tok.push_back(c); break; }
        }
        while (in.get(c) && !std::isspace((unsigned char)c)) tok.push_back(c);
    };
    std::string tok;
    next_token(magic);
    next_token(tok); w = std::atoi(tok.c_str());
    next_token(tok); h = std::atoi(tok.c_str());
    next_token(tok); maxval = std::atoi(tok.c_str());
    if (!in || magic != "P5" || w < 1 || h < 1 || maxval != 255) throw b200::Error(MVS_E_BAD_ARG, "bad PGM file " + filename);
    pixels.resize((size_t)w * h);
    in.read(reinterpret_cast<char *>(pixels.data()), (std::streamsize)pixels.size());
    if (!in) throw b200::Error(MVS_E_BAD_ARG, "truncated PGM file " + filename);
#ifdef MVSLAM_B200_WITH_EIGEN_OPENCV
    return ImageGrayscale(h, w, CV_8U, pixels.data());
#else
    return ImageGrayscale(h, w, pixels.data());
#endif
}

/** PinholeCamera::load_from_file (reference source/vision/camera.cpp:105-123): first line "fx fy shear px py". */
inline CameraIntrinsics load_camera_intrinsics(const std::string &filename)
{
    std::ifstream in(filename);
    CameraIntrinsics K = Matrix3Type::Identity();
    in >> K(0, 0) >> K(1, 1) >> K(0, 1) >> K(0, 2) >> K(1, 2);
    if (!in) throw b200::Error(MVS_E_BAD_ARG, "bad camera file " + filename);
    return K;
}

}  // namespace mvSLAM
