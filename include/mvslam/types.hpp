// types.hpp — minimal stand-ins for the Eigen / OpenCV types that appear in the reference signatures of
// the hot path (reference source/math/matrix.hpp:9-20, source/base/image.hpp:37-51,
// source/base/data-type.hpp:21-27, source/math/lie-group.hpp).  Row-major doubles, value semantics.
// With MVSLAM_B200_WITH_EIGEN_OPENCV the reference's own Eigen / OpenCV types are aliased instead (see below and
// INTEGRATION.md); this image has neither library, so by default the adapters are built against these PODs.
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

// MVSLAM_B200_WITH_EIGEN_OPENCV: alias the reference's own Eigen / OpenCV types (source/math/matrix.hpp:9-20,
// source/base/image.hpp:37-51) instead of the stand-ins below, so that these headers take and return exactly what mvSLAM's
// callers hold.  Eigen matrices are column-major: every adapter reads and writes matrices through operator()(row, col)
// and converts to the C ABI's row-major arrays with b200::rm() / b200::mat_from(), never through data().  (Neither library
// is installed in the build image: tests/cpp/mock_eigen_opencv/ provides just enough of both, column-major storage included,
// to compile and run the adapters in this mode: tests/cpp/test_adapters_eigen.cpp.)
#ifdef MVSLAM_B200_WITH_EIGEN_OPENCV
#include <Eigen/Core>
#include <opencv2/core.hpp>
#endif

namespace mvSLAM {

using ScalarType = double;                                                     // source/system-config.hpp:6
constexpr ScalarType epsilon = std::numeric_limits<ScalarType>::epsilon();     // :8
constexpr ScalarType tolerance = epsilon * 1000;                               // :10
constexpr ScalarType infinity = std::numeric_limits<ScalarType>::max() / 10;   // :14

#ifdef MVSLAM_B200_WITH_EIGEN_OPENCV
using Matrix2Type = Eigen::Matrix<ScalarType, 2, 2>;
using Vector2Type = Eigen::Matrix<ScalarType, 2, 1>;
using Matrix3Type = Eigen::Matrix<ScalarType, 3, 3>;
using Vector3Type = Eigen::Matrix<ScalarType, 3, 1>;
using Matrix6Type = Eigen::Matrix<ScalarType, 6, 6>;
using ImagePoint = cv::Point_<ScalarType>;
using KeyPoint = cv::KeyPoint;
using DMatch = cv::DMatch;
using ImageGrayscale = cv::Mat;                      // CV_8UC1; rows, cols, step and data are read
using Point2 = Vector2Type;
struct VisualFeatureConfig {
    using DetectorResultType = std::vector<cv::KeyPoint>;
    using ExtractorResultType = cv::Mat;             // CV_8U, one 32-byte row per keypoint
    using MatchResultType = std::vector<cv::DMatch>;
};
#else
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
    static Matrix3Type Zero() { return Matrix3Type(); }
};

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
struct Point2 { ScalarType v[2] = {0, 0}; ScalarType &operator[](size_t i) { return v[i]; } const ScalarType &operator[](size_t i) const { return v[i]; } };
struct Matrix2Type { ScalarType m[4] = {0, 0, 0, 0}; ScalarType &operator()(size_t r, size_t c) { return m[r * 2 + c]; } const ScalarType &operator()(size_t r, size_t c) const { return m[r * 2 + c]; } static Matrix2Type Zero() { return Matrix2Type(); } };
struct Matrix6Type { ScalarType m[36] = {}; ScalarType &operator()(size_t r, size_t c) { return m[r * 6 + c]; } const ScalarType &operator()(size_t r, size_t c) const { return m[r * 6 + c]; } static Matrix6Type Zero() { return Matrix6Type(); } };
#endif

using Point3 = Vector3Type;
using CameraIntrinsics = Matrix3Type;
using IdealCameraImagePoint = Vector3Type;

namespace b200 {
// Matrices cross the C ABI as row-major arrays; these two are the only places a matrix's storage order matters.
template <int R, int C>
struct RowMajor {
    double a[R * C];
    const double *data() const { return a; }
    double operator[](int i) const { return a[i]; }
};
template <int R, int C, class M>
inline RowMajor<R, C> rm(const M &m)
{
    RowMajor<R, C> o;
    for (int i = 0; i < R; ++i)
        for (int j = 0; j < C; ++j) o.a[i * C + j] = m(i, j);
    return o;
}
inline RowMajor<3, 3> rm3(const Matrix3Type &m) { return rm<3, 3>(m); }
template <class M, int R, int C>
inline M mat_from(const double *p)
{
    M m;
    for (int i = 0; i < R; ++i)
        for (int j = 0; j < C; ++j) m(i, j) = p[i * C + j];
    return m;
}
inline Matrix3Type mat3_from(const double *p) { return mat_from<Matrix3Type, 3, 3>(p); }
inline std::array<double, 3> v3(const Vector3Type &v) { return {{v[0], v[1], v[2]}}; }

// descriptor container (cv::Mat CV_8U [n][32] in the reference, a flat byte vector in the stand-in)
#ifdef MVSLAM_B200_WITH_EIGEN_OPENCV
inline const uint8_t *desc_data(const cv::Mat &d) { return d.data; }
inline size_t desc_rows(const cv::Mat &d) { return d.empty() ? 0 : (size_t)d.rows; }
inline bool desc_ok(const cv::Mat &d, size_t n) { return (n == 0 && d.empty()) || ((size_t)d.rows == n && d.cols == 32 && d.isContinuous()); }
inline cv::Mat desc_make(const uint8_t *p, size_t rows)
{
    cv::Mat m((int)rows, 32, CV_8U);
    for (size_t i = 0; i < rows * 32; ++i) m.data[i] = p[i];
    return m;
}
#else
inline const uint8_t *desc_data(const std::vector<uint8_t> &d) { return d.data(); }
inline size_t desc_rows(const std::vector<uint8_t> &d) { return d.size() / 32; }
inline bool desc_ok(const std::vector<uint8_t> &d, size_t n) { return d.size() == n * 32; }
inline std::vector<uint8_t> desc_make(const uint8_t *p, size_t rows) { return std::vector<uint8_t>(p, p + rows * 32); }
#endif
}  // namespace b200

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
    /** device < 0: the calling thread's current CUDA device (cudaGetDevice), so that a rank of a multi-GPU job that has
     *  selected its GPU gets its context there */
    explicit Context(int device = -1)
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
        thread_local Context c(-1);
        return c;
    }
    /** A second per-thread context for the adapters that upload a scratch frame table of their own (ImagePair): the frame
     *  table a caller keeps resident on thread_default() (mvs_frames_upload / mvs_orb_extract(append_frames)) is never replaced
     *  behind its back. */
    static Context &thread_scratch()
    {
        thread_local Context c(-1);
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
