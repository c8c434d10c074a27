// ba.hpp — ba_frame_pose_and_point, sfm_refine and pnp_refine with the reference's signatures
// (reference source/vision/ba.hpp:25-36, source/vision/sfm.hpp:69-76, source/vision/pnp.hpp:41-46), forwarding to
// mvs_ba_solve_batch.  The reference builds a GTSAM factor graph (ba.cpp:26-156); here the same cost function is minimised
// on the device (see mvslam_b200.h).  Up to 16 frames per problem (all three reference callers use one or two).
#pragma once
#include <cmath>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "types.hpp"

namespace mvSLAM {

namespace b200 {
inline mvs_ba_params &ba_defaults()
{
    // gtsam::LevenbergMarquardtParams defaults, which ba.cpp:124 uses: lambdaInitial 1e-5, relativeErrorTol 1e-5, absoluteErrorTol 1e-5
    static thread_local mvs_ba_params p{100, 0, 1e-5, 1e-5, 1e-5};
    return p;
}
}  // namespace b200

/** ba.hpp:25-36.  Frames/points are taken in ascending id order; outputs are filled only on success (throws on
 *  violated preconditions, like the reference's asserts). */
inline void ba_frame_pose_and_point(const CameraIntrinsics &ci, const std::unordered_set<Id::Type> &frame_id,
                                    const std::unordered_set<Id::Type> &point_id,
                                    const std::unordered_map<Id::Type, Transformation> &frame_pose_guess,
                                    const std::unordered_map<Id::Type, TransformationUncertainty> &frame_pose_prior,
                                    const std::unordered_map<Id::Type, Point3> &point_guess,
                                    const std::unordered_map<Id::Type, Point3Uncertainty> &point_prior,
                                    const std::unordered_map<Id::Type, PointIdToPoint2Estimate> &frame_observation,
                                    std::unordered_map<Id::Type, TransformationEstimate> &frame_pose_estimate,
                                    std::unordered_map<Id::Type, Point3Estimate> &point_estimate, ScalarType &final_error)
{
    if (frame_id.empty() || point_id.empty() || frame_pose_guess.size() != frame_id.size() || point_guess.size() != point_id.size() ||
        frame_pose_prior.size() + point_prior.size() < 2 || frame_observation.empty())
        throw b200::Error(MVS_E_BAD_ARG, "ba_frame_pose_and_point: precondition violated (ba.cpp:39-44)");
    std::vector<Id::Type> fids(frame_id.begin(), frame_id.end()), pids(point_id.begin(), point_id.end());
    std::sort(fids.begin(), fids.end()); std::sort(pids.begin(), pids.end());
    std::unordered_map<Id::Type, int32_t> fidx, pidx;
    for (size_t i = 0; i < fids.size(); ++i) fidx[fids[i]] = (int32_t)i;
    for (size_t i = 0; i < pids.size(); ++i) pidx[pids[i]] = (int32_t)i;
    const int32_t F = (int32_t)fids.size(), P = (int32_t)pids.size();
    const double nan = std::nan("");
    std::vector<double> R((size_t)F * 9), t((size_t)F * 3), pc((size_t)F * 36, nan), X((size_t)P * 3), xc((size_t)P * 9, nan);
    for (int32_t f = 0; f < F; ++f) {
        const Transformation &T = frame_pose_guess.at(fids[f]);
        { const auto Rr = b200::rm3(T.rotation().get_matrix()); for (int k = 0; k < 9; ++k) R[f * 9 + k] = Rr[k]; }
        for (int k = 0; k < 3; ++k) t[f * 3 + k] = T.translation()[k];
        auto it = frame_pose_prior.find(fids[f]);
        if (it != frame_pose_prior.end()) { const auto Cr = b200::rm<6, 6>(it->second); for (int k = 0; k < 36; ++k) pc[(size_t)f * 36 + k] = Cr[k]; }   // passed to GTSAM as is
    }
    for (int32_t j = 0; j < P; ++j) {
        const Point3 &p = point_guess.at(pids[j]);
        for (int k = 0; k < 3; ++k) X[(size_t)j * 3 + k] = p[k];
        auto it = point_prior.find(pids[j]);
        if (it != point_prior.end()) { const auto Cr = b200::rm3(it->second); for (int k = 0; k < 9; ++k) xc[(size_t)j * 9 + k] = Cr[k]; }
    }
    std::vector<mvs_ba_observation> obs;
    for (const auto &fo : frame_observation)
        for (const auto &po : fo.second) {
            mvs_ba_observation o;
            o.frame = fidx.at(fo.first); o.point = pidx.at(po.first);
            o.uv[0] = po.second.mean()[0]; o.uv[1] = po.second.mean()[1];
            o.cov[0] = po.second.covar()(0, 0); o.cov[1] = po.second.covar()(0, 1); o.cov[2] = po.second.covar()(1, 1);
            obs.push_back(o);
        }
    const int32_t O = (int32_t)obs.size();
    std::vector<double> Ro(R.size()), to(t.size()), pco(pc.size()), Xo(X.size()), xco(xc.size());
    mvs_ba_result res;
    mvs_ctx *ctx = b200::Context::thread_default().get();
    int st = mvs_ba_solve_batch(ctx, 1, b200::rm3(ci).data(), &F, &P, &O, R.data(), t.data(), pc.data(), X.data(), xc.data(), obs.data(),
                                &b200::ba_defaults(), Ro.data(), to.data(), pco.data(), Xo.data(), xco.data(), &res);
    b200::check(ctx, st, "ba_frame_pose_and_point");
    if (st != MVS_OK || res.status != MVS_OK) throw b200::Error(res.status, "ba_frame_pose_and_point: " + std::string(mvs_status_string(res.status)));
    frame_pose_estimate.clear();
    for (int32_t f = 0; f < F; ++f) {
        const Matrix3Type Rm = b200::mat3_from(&Ro[(size_t)f * 9]);
        const Matrix6Type C = b200::mat_from<Matrix6Type, 6, 6>(&pco[(size_t)f * 36]);
        frame_pose_estimate[fids[f]] = TransformationEstimate(SE3(SO3(Rm), Vector3Type(to[f * 3], to[f * 3 + 1], to[f * 3 + 2])), C);
    }
    point_estimate.clear();
    for (int32_t j = 0; j < P; ++j) {
        const Matrix3Type C = b200::mat3_from(&xco[(size_t)j * 9]);
        point_estimate[pids[j]] = Point3Estimate(Point3(Xo[(size_t)j * 3], Xo[(size_t)j * 3 + 1], Xo[(size_t)j * 3 + 2]), C);
    }
    final_error = res.final_error;
}

namespace detail {
inline Matrix6Type diag6(ScalarType a, ScalarType b) { Matrix6Type C = Matrix6Type::Zero(); for (int i = 0; i < 3; ++i) { C(i, i) = a * a; C(i + 3, i + 3) = b * b; } return C; }
inline Matrix3Type diag3(ScalarType a) { Matrix3Type C = Matrix3Type::Zero(); for (int i = 0; i < 3; ++i) C(i, i) = a * a; return C; }
}  // namespace detail

/** sfm.hpp:69-76 / sfm-refine.cpp:20-139: camera 1 anchored at the origin (1e-5), camera 2 and every point regularised (1e-2). */
inline bool sfm_refine(const std::vector<Point2Estimate> &p1_estimate, const std::vector<Point2Estimate> &p2_estimate,
                       const CameraIntrinsics &ci, const Transformation &pose2in1_guess, const std::vector<Point3> pointsin1_guess,
                       TransformationEstimate &pose2in1_estimate, std::vector<Point3Estimate> &pointsin1_estimate, ScalarType &error)
{
    if (p1_estimate.size() != p2_estimate.size() || p1_estimate.size() != pointsin1_guess.size() || p1_estimate.empty())
        throw b200::Error(MVS_E_BAD_ARG, "sfm_refine: size mismatch (sfm-refine.cpp:29-30)");
    const size_t n = p1_estimate.size();
    std::unordered_set<Id::Type> fid{0, 1}, pid;
    std::unordered_map<Id::Type, Transformation> fg{{0, SE3()}, {1, pose2in1_guess}};
    std::unordered_map<Id::Type, TransformationUncertainty> fp{{0, detail::diag6(1e-5, 1e-5)}, {1, detail::diag6(1e-2, 1e-2)}};
    std::unordered_map<Id::Type, Point3> pg;
    std::unordered_map<Id::Type, Point3Uncertainty> pp;
    std::unordered_map<Id::Type, PointIdToPoint2Estimate> ob;
    for (size_t i = 0; i < n; ++i) { pid.insert(i); pg[i] = pointsin1_guess[i]; pp[i] = detail::diag3(1e-2); ob[0][i] = p1_estimate[i]; ob[1][i] = p2_estimate[i]; }
    std::unordered_map<Id::Type, TransformationEstimate> fe;
    std::unordered_map<Id::Type, Point3Estimate> pe;
    ba_frame_pose_and_point(ci, fid, pid, fg, fp, pg, pp, ob, fe, pe, error);
    pose2in1_estimate = fe[1];
    pointsin1_estimate.clear();
    for (size_t i = 0; i < n; ++i) pointsin1_estimate.push_back(pe[i]);
    return true;
}

/** pnp.hpp:41-46 / pnp-refine.cpp:16-110: one regularised camera (1e-2), points with their own covariances as priors;
 *  the point estimates are not returned (pnp-refine.cpp:104-105). */
inline bool pnp_refine(const std::vector<Point3Estimate> &world_point_estimates, const std::vector<Point2Estimate> &image_point_estimates,
                       const CameraIntrinsics &ci, const Transformation &pose_guess, TransformationEstimate &pose_estimate, ScalarType &error)
{
    if (world_point_estimates.size() != image_point_estimates.size() || world_point_estimates.empty())
        throw b200::Error(MVS_E_BAD_ARG, "pnp_refine: size mismatch (pnp-refine.cpp:23)");
    const size_t n = world_point_estimates.size();
    std::unordered_set<Id::Type> fid{0}, pid;
    std::unordered_map<Id::Type, Transformation> fg{{0, pose_guess}};
    std::unordered_map<Id::Type, TransformationUncertainty> fp{{0, detail::diag6(1e-2, 1e-2)}};
    std::unordered_map<Id::Type, Point3> pg;
    std::unordered_map<Id::Type, Point3Uncertainty> pp;
    std::unordered_map<Id::Type, PointIdToPoint2Estimate> ob;
    for (size_t i = 0; i < n; ++i) { pid.insert(i); pg[i] = world_point_estimates[i].mean(); pp[i] = world_point_estimates[i].covar(); ob[0][i] = image_point_estimates[i]; }
    std::unordered_map<Id::Type, TransformationEstimate> fe;
    std::unordered_map<Id::Type, Point3Estimate> pe;
    ba_frame_pose_and_point(ci, fid, pid, fg, fp, pg, pp, ob, fe, pe, error);
    pose_estimate = fe[0];
    return true;
}

}  // namespace mvSLAM
