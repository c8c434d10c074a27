// sfm.hpp — sfm_solve / sfm_triangulate (reference source/vision/sfm.hpp:30-53),
// FundamentalMatrixEstimatorRANSAC (source/vision/estimator-RANSAC.hpp:10-50) and find_fundamental_matrix
// (source/vision/fundamental-matrix.hpp:16-19) with the reference's signatures, forwarding to the C ABI.
#pragma once
#include "types.hpp"

namespace mvSLAM {

namespace b200 {
/** Knobs the reference hard-codes (sfm-solve.cpp:18-23,67): process-wide defaults for the adapters. */
inline mvs_ransac_params &ransac_defaults()
{
    static mvs_ransac_params p{1, MVS_SCORE_ALGEBRAIC, 0.0, 0, 0, MVS_SOLVER_REFERENCE, 0};   // H = 1: the reference's single sample,
                                                                                              // literal A^T A / cv::SVDecomp arithmetic
    return p;
}
inline std::vector<double> flatten(const std::vector<ImagePoint> &p)
{
    std::vector<double> r(p.size() * 2);
    for (size_t i = 0; i < p.size(); ++i) { r[2 * i] = p[i].x; r[2 * i + 1] = p[i].y; }
    return r;
}
}  // namespace b200

/** sfm.hpp:30-35.  Outputs are written only on success (sfm-solve.cpp:364-366). */
inline bool sfm_solve(const std::vector<ImagePoint> &p1, const std::vector<ImagePoint> &p2, const CameraIntrinsics &K,
                      Transformation &pose2in1_scaled, std::vector<Point3> &pointsin1_scaled,
                      std::vector<size_t> &point_indexes)
{
    if (p1.size() != p2.size()) throw b200::Error(MVS_E_BAD_ARG, "sfm_solve: p1.size() != p2.size()");  // :292 assert
    mvs_ctx *ctx = b200::Context::thread_default().get();
    const int n = (int)p1.size();
    auto a = b200::flatten(p1), b = b200::flatten(p2);
    std::vector<double> pts((size_t)n * 3 + 3);
    std::vector<uint64_t> idx((size_t)n + 1);
    mvs_pair_result r;
    int st = mvs_sfm_solve(ctx, a.data(), b.data(), n, b200::rm3(K).data(), &b200::ransac_defaults(), nullptr, &r, nullptr, pts.data(),
                           idx.data(), n);
    b200::check(ctx, st, "sfm_solve");
    if (st != MVS_OK) return false;
    pose2in1_scaled = SE3(SO3(b200::mat3_from(r.R2in1)), Vector3Type(r.t2in1[0], r.t2in1[1], r.t2in1[2]));
    pointsin1_scaled.resize(r.n_points);
    point_indexes.resize(r.n_points);
    for (int i = 0; i < r.n_points; ++i) {
        pointsin1_scaled[i] = Point3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
        point_indexes[i] = (size_t)idx[i];
    }
    return true;
}

/** sfm.hpp:47-53 */
inline void sfm_triangulate(const std::vector<ImagePoint> &p1, const std::vector<ImagePoint> &p2, const CameraIntrinsics &K,
                            const Transformation &pose1, const Transformation &pose2, std::vector<Point3> &points,
                            std::vector<size_t> &point_indexes)
{
    if (p1.size() != p2.size()) throw b200::Error(MVS_E_BAD_ARG, "sfm_triangulate: p1.size() != p2.size()");
    mvs_ctx *ctx = b200::Context::thread_default().get();
    const int n = (int)p1.size();
    auto a = b200::flatten(p1), b = b200::flatten(p2);
    std::vector<double> pts((size_t)n * 3 + 3);
    std::vector<uint64_t> idx((size_t)n + 1);
    int m = 0;
    int st = mvs_sfm_triangulate(ctx, a.data(), b.data(), n, b200::rm3(K).data(), b200::rm3(pose1.rotation().get_matrix()).data(),
                                 b200::v3(pose1.translation()).data(), b200::rm3(pose2.rotation().get_matrix()).data(),
                                 b200::v3(pose2.translation()).data(), b200::ransac_defaults().solver, pts.data(), idx.data(), n, &m);
    b200::check(ctx, st, "sfm_triangulate");
    points.resize(m);
    point_indexes.resize(m);
    for (int i = 0; i < m; ++i) { points[i] = Point3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]); point_indexes[i] = (size_t)idx[i]; }
}

/** fundamental-matrix.hpp:16-19: F21 from exactly 8 correspondences (x, y, 1). */
inline bool find_fundamental_matrix(const std::vector<Vector3Type> &p1_sample, const std::vector<Vector3Type> &p2_sample,
                                    Matrix3Type &F21)
{
    if (p1_sample.size() != 8 || p2_sample.size() != 8) throw b200::Error(MVS_E_BAD_ARG, "find_fundamental_matrix needs 8 points");
    mvs_ctx *ctx = b200::Context::thread_default().get();
    double a[24], b[24];
    for (int i = 0; i < 8; ++i) for (int k = 0; k < 3; ++k) { a[3 * i + k] = p1_sample[i][k]; b[3 * i + k] = p2_sample[i][k]; }
    double F[9];
    int st = mvs_find_fundamental_matrix(ctx, a, b, 1, b200::ransac_defaults().solver, F);
    b200::check(ctx, st, "find_fundamental_matrix");
    if (st == MVS_OK) F21 = b200::mat3_from(F);
    return st == MVS_OK;
}

/** estimator-RANSAC.hpp:10-50.  max_iteration = number of rows of the seeded sample table (row 0 is the
 *  reference's un-shuffled sample {0..7}, estimator-RANSAC.cpp:41-48). */
class FundamentalMatrixEstimatorRANSAC {
public:
    FundamentalMatrixEstimatorRANSAC(ScalarType max_error_sq_, size_t max_iteration_, uint64_t seed = 0)
        : max_error_sq(max_error_sq_), max_iteration(max_iteration_), m_seed(seed)
    {
        if (!(max_error_sq > epsilon) || max_iteration == 0) throw b200::Error(MVS_E_BAD_ARG, "bad RANSAC parameters");  // .cpp:12-13
    }
    bool compute(const std::vector<Vector3Type> &p1, const std::vector<Vector3Type> &p2, Matrix3Type &F21,
                 std::vector<uint8_t> &inlier_mask)
    {
        if (p1.size() != p2.size()) throw b200::Error(MVS_E_BAD_ARG, "compute: p1.size() != p2.size()");
        mvs_ctx *ctx = b200::Context::thread_default().get();
        const int n = (int)p1.size();
        std::vector<double> a((size_t)n * 3), b((size_t)n * 3);
        for (int i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) { a[3 * i + k] = p1[i][k]; b[3 * i + k] = p2[i][k]; }
        std::vector<uint8_t> mask((size_t)n + 1);
        const mvs_ransac_params rp{(int32_t)max_iteration, MVS_SCORE_ALGEBRAIC, max_error_sq, m_seed, 0,
                                   b200::ransac_defaults().solver, 0};
        double F[9];
        int cnt = 0, bh = -1;
        double res = 0;
        int st = mvs_ransac_fundamental(ctx, a.data(), b.data(), n, nullptr, &rp, F, mask.data(), &cnt, &res, &bh, nullptr);
        b200::check(ctx, st, "FundamentalMatrixEstimatorRANSAC::compute");
        if (st == MVS_E_TOO_FEW_POINTS) return false;   // estimator-RANSAC.cpp:25-29
        F21 = b200::mat3_from(F);
        mask.resize(n);
        inlier_mask.swap(mask);
        return cnt > 0;                                  // :89
    }
private:
    const ScalarType max_error_sq;
    const size_t max_iteration;
    const uint64_t m_seed;
};

}  // namespace mvSLAM
