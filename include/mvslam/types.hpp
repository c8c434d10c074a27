// types.hpp — minimal stand-ins for the Eigen / OpenCV types that appear in the reference signatures of
// the hot path (reference source/math/matrix.hpp:9-20, source/base/image.hpp:37-51,
// source/base/data-type.hpp:21-27, source/math/lie-group.hpp).  Row-major doubles, value semantics.
// An mvSLAM build that has Eigen/OpenCV converts at the call site; this image
// has neither, so the adapters are written — and tested — against these PODs.
#pragma once
#include <array>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../mvslam_b200.h"

namespace mvSLAM {

using ScalarType = double;                                                     // source/system-config.hpp:6
constexpr ScalarType epsilon = std::numeric_limits<ScalarType>::epsilon();     // :8
constexpr ScalarType tolerance = epsilon * 1000;                               // :10
constexpr ScalarType infinity = std::numeric_limits<ScalarType>::max() / 10;   // :14

struct Vector3Type {
    ScalarType v[3] = {0, 0, 0};
    Vector3Type() = default;
    Vector3Type(ScalarType x, ScalarType y, ScalarType z) : v{x, y, z} {}
    ScalarType &operator[](size_t i) { return v[i]; }
    const ScalarType &operator[](size_t i) const { return v[i]; }
};

struct Matrix3Type {
    ScalarType m[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // row-major (Eigen's default storage is column-major!)
    ScalarType &operator()(size_t r, size_t c) { return m[r * 3 + c]; }
    const ScalarType &operator()(size_t r, size_t c) const { return m[r * 3 + c]; }
    static Matrix3Type Identity() { Matrix3Type I; I.m[0] = I.m[4] = I.m[8] = 1; return I; }
};

using Point3 = Vector3Type;
using CameraIntrinsics = Matrix3Type;
using IdealCameraImagePoint = Vector3Type;

struct ImagePoint {  // cv::Point_<ScalarType>
    ScalarType x = 0, y = 0;
    ImagePoint() = default;
    ImagePoint(ScalarType x_, ScalarType y_) : x(x_), y(y_) {}
};

struct KeyPoint {    // cv::KeyPoint
    struct { float x, y; } pt{0, 0};
    float size = 0, angle = -1, response = 0;
    int octave = 0, class_id = -1;
};

struct ImageGrayscale {   // the fields of a CV_8UC1 cv::Mat that VisualFeature::extract reads (source/base/image.hpp:21-23)
    int rows = 0, cols = 0;
    size_t step = 0;      // bytes per row
    const uint8_t *data = nullptr;
    ImageGrayscale() = default;
    ImageGrayscale(int rows_, int cols_, const uint8_t *data_, size_t step_ = 0)
        : rows(rows_), cols(cols_), step(step_ ? step_ : (size_t)cols_), data(data_) {}
};

struct DMatch {      // cv::DMatch
    int queryIdx = -1, trainIdx = -1, imgIdx = 0;
    float distance = 0;
};

struct VisualFeatureConfig {
    using DetectorResultType = std::vector<KeyPoint>;
    using ExtractorResultType = std::vector<uint8_t>;   // CV_8U rows of 32 bytes, contiguous
    using MatchResultType = std::vector<DMatch>;
};

// SO3/SE3 as far as the path's outputs need them (source/math/lie-group.hpp:24-234): the rotation is
// stored as delivered by the library (already rectified the way the reference's SO3 ctor does).
class SO3 {
public:
    SO3() : _R(Matrix3Type::Identity()) {}
    explicit SO3(const Matrix3Type &m) : _R(m) {}
    const Matrix3Type &get_matrix() const { return _R; }
private:
    Matrix3Type _R;
};

class SE3 {
public:
    SE3() = default;
    SE3(const SO3 &r, const Vector3Type &t) : _R(r), _t(t) {}
    const SO3 &rotation() const { return _R; }
    const Vector3Type &translation() const { return _t; }
private:
    SO3 _R;
    Vector3Type _t;
};
using Transformation = SE3;

// StateEstimate<Mean, Covar> (source/math/state-estimate.hpp:6-60) and the aliases of source/base/data-type.hpp:20-29
template <typename MeanType, typename CovarType>
class StateEstimate {
public:
    StateEstimate() = default;
    StateEstimate(const MeanType &mean, const CovarType &covar) : _mean(mean), _covar(covar) {}
    const MeanType &mean() const { return _mean; }
    const CovarType &covar() const { return _covar; }
private:
    MeanType _mean{};
    CovarType _covar{};
};
struct Point2 { ScalarType v[2] = {0, 0}; ScalarType &operator[](size_t i) { return v[i]; } const ScalarType &operator[](size_t i) const { return v[i]; } };
struct Matrix2Type { ScalarType m[4] = {0, 0, 0, 0}; ScalarType &operator()(size_t r, size_t c) { return m[r * 2 + c]; } const ScalarType &operator()(size_t r, size_t c) const { return m[r * 2 + c]; } };
struct Matrix6Type { ScalarType m[36] = {}; ScalarType &operator()(size_t r, size_t c) { return m[r * 6 + c]; } const ScalarType &operator()(size_t r, size_t c) const { return m[r * 6 + c]; } };
using TransformationUncertainty = Matrix6Type;
using TransformationEstimate = StateEstimate<Transformation, TransformationUncertainty>;
using Point3Uncertainty = Matrix3Type;
using Point3Estimate = StateEstimate<Point3, Point3Uncertainty>;
using Point2Uncertainty = Matrix2Type;
using Point2Estimate = StateEstimate<Point2, Point2Uncertainty>;
namespace Id { using Type = uint64_t; }
using PointIdToPoint2Estimate = std::unordered_map<Id::Type, Point2Estimate>;

namespace b200 {

struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string &what) : std::runtime_error(what), status(s) {}
};

// One context per thread (the C ABI's ctx is not thread-safe); replaces the reference's hidden globals.
class Context {
public:
    explicit Context(int device = 0)
    {
        int st = mvs_create(&_ctx, device);
        if (st != MVS_OK) throw Error(st, std::string("mvs_create: ") + mvs_status_string(st) + " (no CPU fallback)");
    }
    ~Context() { mvs_destroy(_ctx); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    mvs_ctx *get() const { return _ctx; }
    static Context &thread_default()
    {
        thread_local Context c(0);
        return c;
    }
private:
    mvs_ctx *_ctx = nullptr;
};

inline void check(mvs_ctx *ctx, int st, const char *where)
{
    if (st == MVS_E_CUDA || st == MVS_E_BAD_ARG || st == MVS_E_CAPACITY || st == MVS_E_UNSUPPORTED)
        throw Error(st, std::string(where) + ": " + mvs_status_string(st) + ": " + mvs_last_error(ctx));
}

}  // namespace b200
}  // namespace mvSLAM
