// pnp.hpp — pnp_solve with the reference's signature (reference source/vision/pnp.hpp:22-26,
// source/vision/pnp-solve.cpp:16-104), forwarding to the C ABI.  pnp_refine (GTSAM) is outside the path.
#pragma once
#include <vector>

#include "types.hpp"

namespace mvSLAM {

constexpr size_t PNP_MIN_POINT_COUNT = 7;   // pnp-solve.cpp:13-14

namespace b200 {
inline mvs_pnp_params &pnp_defaults()
{
    // iterationsCount 100, reprojectionError 0.05 (pnp-solve.cpp:50-51); 10 Gauss-Newton steps on the inliers
    static thread_local mvs_pnp_params p{100, 10, 0.05, 0, 0, 0, 0};
    return p;
}
}  // namespace b200

/** Camera pose (camera to world) from 3D-2D correspondences; false when no pose has enough support.
 *  inlier_point_indexes: indexes of the correspondences consistent with the winning minimal-sample pose.
 *  Outputs are written only on success. */
inline bool pnp_solve(const std::vector<Point3> &world_points, const std::vector<ImagePoint> &image_points,
                      const CameraIntrinsics &K, Transformation &pose, std::vector<size_t> &inlier_point_indexes)
{
    if (world_points.size() < PNP_MIN_POINT_COUNT || world_points.size() != image_points.size())
        throw b200::Error(MVS_E_BAD_ARG, "pnp_solve: need >= 7 correspondences of equal count");   // :21-22 asserts
    mvs_ctx *ctx = b200::Context::thread_default().get();
    const int n = (int)world_points.size();
    std::vector<double> w((size_t)n * 3), im((size_t)n * 2);
    for (int i = 0; i < n; ++i) {
        for (int k = 0; k < 3; ++k) w[3 * i + k] = world_points[i][k];
        im[2 * i] = image_points[i].x; im[2 * i + 1] = image_points[i].y;
    }
    std::vector<uint8_t> mask(n);
    mvs_pnp_result r;
    const int st = mvs_pnp_solve(ctx, w.data(), im.data(), n, b200::rm3(K).data(), &b200::pnp_defaults(), nullptr, &r, mask.data(), nullptr);
    b200::check(ctx, st, "pnp_solve");
    if (st != MVS_OK) return false;
    pose = SE3(SO3(b200::mat3_from(r.R_c2w)), Vector3Type(r.t_c2w[0], r.t_c2w[1], r.t_c2w[2]));
    inlier_point_indexes.clear();
    for (int i = 0; i < n; ++i) if (mask[i]) inlier_point_indexes.push_back((size_t)i);
    return true;
}

}  // namespace mvSLAM
