// Minimal stand-in for <opencv2/core.hpp> — TEST INFRASTRUCTURE (OpenCV's C++ headers are not installed in the build image).
// cv::Point_, cv::KeyPoint, cv::DMatch and a reference-counted cv::Mat with the members the adapters read.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#define CV_8U 0
namespace cv {
typedef unsigned char uchar;
template <typename T> struct Point_ { T x = 0, y = 0; Point_() = default; Point_(T x_, T y_) : x(x_), y(y_) {} };
typedef Point_<float> Point2f;
struct KeyPoint { Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1; };
struct DMatch { int queryIdx = -1, trainIdx = -1, imgIdx = -1; float distance = 0; };
class Mat {
public:
    int rows = 0, cols = 0;
    std::size_t step = 0;
    uchar *data = nullptr;
    Mat() = default;
    Mat(int r, int c, int /*type CV_8U*/) : rows(r), cols(c), step((std::size_t)c), own(new uchar[(std::size_t)r * c + 1], std::default_delete<uchar[]>()) { data = own.get(); }
    Mat(int r, int c, int /*type*/, void *ext, std::size_t step_ = 0) : rows(r), cols(c), step(step_ ? step_ : (std::size_t)c), data((uchar *)ext) {}
    bool empty() const { return rows == 0 || cols == 0 || !data; }
    bool isContinuous() const { return step == (std::size_t)cols; }
private:
    std::shared_ptr<uchar> own;
};
}  // namespace cv
