"""Python mirror of the reference's entry points for this path (same names / argument meaning):

  VisualFeature.extract(image)                              source/vision/visual-feature.hpp:14
  VisualFeature.match_visual_features(vf1, vf2, max_dist)   source/vision/visual-feature.hpp:23-26
  sfm_solve(p1, p2, K)                                      source/vision/sfm.hpp:30-35
  sfm_triangulate(p1, p2, K, pose1, pose2)                  source/vision/sfm.hpp:47-53
  pnp_solve(world_points, image_points, K)                  source/vision/pnp.hpp:22-26
  image_pairs(frames, pairs, K, params)                     ImagePair ctor, source/front-end/image-pair.hpp:38-40

All of them forward to the C ABI (libmvslam_b200.so); none computes anything on the CPU.
(The C++ equivalents for an mvSLAM build are the headers under include/mvslam/.)
"""
import numpy as np

from . import capi

_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = capi.Context(-1)      # the process's current CUDA device (torch.cuda.set_device / cudaSetDevice)
    return _default_ctx


class VisualFeature:
    """Keypoints + descriptors of one frame (source/vision/visual-feature.hpp:91-93)."""

    def __init__(self, keypoints_xy, descriptors, image_width=-1, image_height=-1):
        self.keypoints = np.ascontiguousarray(keypoints_xy, np.float32).reshape(-1, 2)
        self.descriptors = np.ascontiguousarray(descriptors, np.uint8)
        self.image_width, self.image_height = image_width, image_height

    @staticmethod
    def extract(image, n_features=500, ctx=None):
        """cv::ORB detect + compute on the device (visual-feature.cpp:40-49; MAX_FEATURE_COUNT = 500, :9)."""
        ctx = ctx or default_context()
        _, kp, desc, _ = ctx.orb_extract([image], n_features)
        vf = VisualFeature(np.stack([kp["x"], kp["y"]], 1), desc, image.shape[1], image.shape[0])
        vf.cv_keypoints = kp
        return vf

    def size(self):
        return self.keypoints.shape[0]

    def get_image_points(self):
        return self.keypoints.astype(np.float64)      # visual-feature.cpp:179-190

    @staticmethod
    def match_visual_features(vf1, vf2, max_dist=-1.0, ctx=None):
        return match_visual_features(vf1, vf2, max_dist, ctx)


def match_visual_features(vf1, vf2, max_dist=-1.0, ctx=None):
    """Matches from 2 to 1 (query = vf2, train = vf1), ascending distance; empty if none."""
    ctx = ctx or default_context()
    return ctx.match_hamming(vf2.descriptors, vf1.descriptors, 0.7, float(max_dist), False, bounded=True)


def sfm_solve(p1, p2, K, ctx=None, **kw):
    """Returns (ok, (R2in1, t2in1), points, indexes) like the reference's bool + out-params."""
    ctx = ctx or default_context()
    r = ctx.sfm_solve(p1, p2, K, **kw)
    ok = r["status"] == capi.OK
    return ok, (r["R2in1"], r["t2in1"]), r["points"], r["indexes"]


def sfm_triangulate(p1, p2, K, pose1, pose2, ctx=None):
    ctx = ctx or default_context()
    return ctx.sfm_triangulate(p1, p2, K, pose1[0], pose1[1], pose2[0], pose2[1])


def pnp_solve(world_points, image_points, K, ctx=None, **kw):
    """Returns (ok, (R, t) camera-to-world pose, inlier_point_indexes) like the reference's bool + out-params
    (pnp-solve.cpp:16-104: cv::solvePnPRansac with P3P, 100 iterations, reprojection error 0.05)."""
    ctx = ctx or default_context()
    r = ctx.pnp_solve(world_points, image_points, K, **kw)
    return r["status"] == capi.OK, (r["R_c2w"], r["t_c2w"]), np.nonzero(r["mask"])[0].astype(np.uint64)


def image_pairs(frames, pairs, K, max_match_inlier_distance=10.0, ctx=None, **kw):
    """Batch of ImagePair constructions: frames = [VisualFeature], pairs = [(base, pair)]."""
    ctx = ctx or default_context()
    ctx.frames_upload([f.descriptors for f in frames], [f.keypoints for f in frames])
    kw.setdefault("bounded", True)
    return ctx.pair_batch(pairs, K, max_dist=max_match_inlier_distance, **kw)
