"""Oracle for bundle adjustment — numpy restatement of ba_frame_pose_and_point and its callers sfm_refine / pnp_refine.

TEST INFRASTRUCTURE ONLY (never imported by the product package).

The reference builds a gtsam::NonlinearFactorGraph and runs gtsam::LevenbergMarquardtOptimizer + gtsam::Marginals
(source/vision/ba.cpp:26-156); GTSAM is an un-vendored third-party dependency that is absent from this image, so
**parity with GTSAM itself is unpinned**.  What the reference's own code fixes is the *cost function*, restated here:

  PriorFactor<Pose3>(x_f, guess_f, Gaussian::Covariance(C_f))          ba.cpp:57-72    1/2 |Local(guess_f, T_f)|^2_{C_f}
  PriorFactor<Point3>(p_j, guess_j, Gaussian::Covariance(C_j))         ba.cpp:75-93    1/2 |X_j - guess_j|^2_{C_j}
  GenericProjectionFactor<Pose3, Point3, Cal3_S2>(z, Covariance(C_z))  ba.cpp:96-117   1/2 |K pi(T_f^-1 X_j) - z|^2_{C_z}

with T_f the camera-to-world pose, Cal3_S2 = (fx, fy, skew, u0, v0), Pose3 local coordinates [rotation, translation]
= [Log(R_g^T R), R_g^T (t - t_g)] (GTSAM's default chart; its Cayley variant differs at third order in a deviation the
priors keep below 1e-2).  final_error = the value of that cost at the result (optimizer.error(), ba.cpp:154), the
estimates' covariances = the blocks of the inverse Gauss-Newton Hessian at the result (gtsam::Marginals, ba.cpp:123).

Pinning (tests/test_ba_oracle.py): the analytic Jacobians against finite differences, the minimiser against
scipy.optimize.least_squares on the same residual vector, and the reference's own known-answer tests
(test/test-sfm.cpp sfm_refine_L_shape, test/test-pnp.cpp pnp_refine_L_shape; tolerance 0.025).
"""
import numpy as np


def hat(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], float)


def so3_exp(w):
    th = np.linalg.norm(w)
    W = hat(w)
    if th < 1e-10:
        return np.eye(3) + W + 0.5 * W @ W
    return np.eye(3) + np.sin(th) / th * W + (1 - np.cos(th)) / th ** 2 * W @ W


def so3_log(R):
    c = np.clip((np.trace(R) - 1) / 2, -1, 1)
    th = np.arccos(c)
    v = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    if th < 1e-10:
        return 0.5 * v
    return th / (2 * np.sin(th)) * v


def jr_inv(phi):
    """inverse right Jacobian of SO(3): Log(R Exp(w)) ~ Log(R) + Jr^-1(Log R) w"""
    th = np.linalg.norm(phi)
    P = hat(phi)
    if th < 1e-6:
        return np.eye(3) + 0.5 * P + P @ P / 12.0
    return np.eye(3) + 0.5 * P + (1 / th ** 2 - (1 + np.cos(th)) / (2 * th * np.sin(th))) * P @ P


class Problem:
    """K: 3x3; poses: list of (R, t) camera-to-world; pose_prior: dict f -> 6x6 covariance (order rotation, translation;
    the prior mean is the guess); points: [P,3]; point_prior: dict j -> 3x3 covariance; obs: list of
    (frame, point, uv[2], cov 2x2)."""

    def __init__(self, K, poses, pose_prior, points, point_prior, obs):
        self.K = np.asarray(K, float)
        self.poses0 = [(np.asarray(R, float).copy(), np.asarray(t, float).copy()) for R, t in poses]
        self.pose_prior = {f: np.linalg.inv(np.asarray(C, float)) for f, C in pose_prior.items()}
        self.points0 = np.asarray(points, float).reshape(-1, 3).copy()
        self.point_prior = {j: np.linalg.inv(np.asarray(C, float)) for j, C in point_prior.items()}
        self.obs = [(int(f), int(j), np.asarray(z, float), np.linalg.inv(np.asarray(C, float))) for f, j, z, C in obs]
        self.F, self.P = len(self.poses0), len(self.points0)

    # ---- residual blocks with Jacobians w.r.t. the local perturbation (R <- R Exp(w), t <- t + R v; X <- X + d)
    def project(self, pose, X):
        R, t = pose
        p = R.T @ (X - t)
        fx, fy, s, u0, v0 = self.K[0, 0], self.K[1, 1], self.K[0, 1], self.K[0, 2], self.K[1, 2]
        x, y, z = p
        uv = np.array([fx * x / z + s * y / z + u0, fy * y / z + v0])
        dpi = np.array([[fx / z, s / z, -(fx * x + s * y) / z ** 2], [0, fy / z, -fy * y / z ** 2]])
        Jcam = dpi @ np.hstack([hat(p), -np.eye(3)])     # d p / d w = [p]x, d p / d v = -I
        JX = dpi @ R.T
        return uv, Jcam, JX

    def cost(self, poses, points):
        c = 0.0
        for f, info in self.pose_prior.items():
            Rg, tg = self.poses0[f]
            e = np.concatenate([so3_log(Rg.T @ poses[f][0]), Rg.T @ (poses[f][1] - tg)])
            c += 0.5 * e @ info @ e
        for j, info in self.point_prior.items():
            e = points[j] - self.points0[j]
            c += 0.5 * e @ info @ e
        for f, j, z, info in self.obs:
            e = self.project(poses[f], points[j])[0] - z
            c += 0.5 * e @ info @ e
        return c

    def normal_equations(self, poses, points):
        n = 6 * self.F + 3 * self.P
        H = np.zeros((n, n)); g = np.zeros(n)
        for f, info in self.pose_prior.items():
            Rg, tg = self.poses0[f]
            R, t = poses[f]
            phi = so3_log(Rg.T @ R)
            e = np.concatenate([phi, Rg.T @ (t - tg)])
            J = np.zeros((6, 6)); J[:3, :3] = jr_inv(phi); J[3:, 3:] = Rg.T @ R
            s = slice(6 * f, 6 * f + 6)
            H[s, s] += J.T @ info @ J; g[s] += J.T @ info @ e
        for j, info in self.point_prior.items():
            s = slice(6 * self.F + 3 * j, 6 * self.F + 3 * j + 3)
            H[s, s] += info; g[s] += info @ (points[j] - self.points0[j])
        for f, j, z, info in self.obs:
            uv, Jc, JX = self.project(poses[f], points[j])
            e = uv - z
            sc = slice(6 * f, 6 * f + 6); sp = slice(6 * self.F + 3 * j, 6 * self.F + 3 * j + 3)
            H[sc, sc] += Jc.T @ info @ Jc; H[sp, sp] += JX.T @ info @ JX
            H[sc, sp] += Jc.T @ info @ JX; H[sp, sc] += JX.T @ info @ Jc
            g[sc] += Jc.T @ info @ e; g[sp] += JX.T @ info @ e
        return H, g

    @staticmethod
    def retract(poses, points, d):
        F = len(poses)
        newp = []
        for f, (R, t) in enumerate(poses):
            w, v = d[6 * f:6 * f + 3], d[6 * f + 3:6 * f + 6]
            newp.append((R @ so3_exp(w), t + R @ v))
        return newp, points + d[6 * F:].reshape(-1, 3)

    def solve(self, max_iter=100, lam=1e-5):
        """Levenberg-Marquardt (lambda * I damping like GTSAM's default), iterated to a tight tolerance."""
        poses, points = self.poses0, self.points0.copy()
        c = self.cost(poses, points)
        it = 0
        for it in range(max_iter):
            H, g = self.normal_equations(poses, points)
            improved = False
            while lam < 1e12:
                try:
                    d = -np.linalg.solve(H + lam * np.eye(len(g)), g)
                except np.linalg.LinAlgError:
                    lam *= 10; continue
                np_, nx = self.retract(poses, points, d)
                cn = self.cost(np_, nx)
                if cn <= c:
                    improved = True
                    break
                lam *= 10
            if not improved:
                break
            rel = (c - cn) / max(c, 1e-300)
            poses, points, c = np_, nx, cn
            lam = max(lam / 10, 1e-12)
            if rel < 1e-13 or np.abs(d).max() < 1e-14:
                break
        H, _ = self.normal_equations(poses, points)
        cov = np.linalg.inv(H)
        pose_cov = [cov[6 * f:6 * f + 6, 6 * f:6 * f + 6] for f in range(self.F)]
        o = 6 * self.F
        point_cov = [cov[o + 3 * j:o + 3 * j + 3, o + 3 * j:o + 3 * j + 3] for j in range(self.P)]
        return dict(poses=poses, points=points, error=c, pose_cov=pose_cov, point_cov=point_cov, iterations=it + 1)

    # ---- the same cost as a plain whitened residual vector over a minimal parametrisation, for the scipy pin
    def residual_vector(self, x):
        poses, points = self.retract(self.poses0, self.points0, x)
        r = []
        for f, info in self.pose_prior.items():
            Rg, tg = self.poses0[f]
            e = np.concatenate([so3_log(Rg.T @ poses[f][0]), Rg.T @ (poses[f][1] - tg)])
            r.append(np.linalg.cholesky(info).T @ e)
        for j, info in self.point_prior.items():
            r.append(np.linalg.cholesky(info).T @ (points[j] - self.points0[j]))
        for f, j, z, info in self.obs:
            r.append(np.linalg.cholesky(info).T @ (self.project(poses[f], points[j])[0] - z))
        return np.concatenate(r)


# ---------------------------------------------------------------- the reference's two wrappers
SFM_ANCHOR_STDDEV = 1e-5       # sfm-refine.cpp:10-13 (position and orientation)
SFM_REGULATOR_STDDEV = 1e-2    # :14-17
PNP_REGULATOR_STDDEV = 1e-2    # pnp-refine.cpp:11-14


def sfm_refine_problem(p1, p1_cov, p2, p2_cov, K, pose2in1_guess, pointsin1_guess):
    """sfm_refine (source/vision/sfm-refine.cpp:20-139): camera 1 anchored at the origin, camera 2 and every point
    regularised around their guesses."""
    n = len(p1)
    prior = {0: np.eye(6) * SFM_ANCHOR_STDDEV ** 2, 1: np.eye(6) * SFM_REGULATOR_STDDEV ** 2}
    pprior = {j: np.eye(3) * SFM_REGULATOR_STDDEV ** 2 for j in range(n)}
    obs = [(0, j, p1[j], p1_cov[j]) for j in range(n)] + [(1, j, p2[j], p2_cov[j]) for j in range(n)]
    return Problem(K, [(np.eye(3), np.zeros(3)), pose2in1_guess], prior, pointsin1_guess, pprior, obs)


def pnp_refine_problem(world, world_cov, image, image_cov, K, pose_guess):
    """pnp_refine (source/vision/pnp-refine.cpp:16-110): one regularised camera, points with their own priors."""
    n = len(world)
    return Problem(K, [pose_guess], {0: np.eye(6) * PNP_REGULATOR_STDDEV ** 2}, world,
                   {j: world_cov[j] for j in range(n)}, [(0, j, image[j], image_cov[j]) for j in range(n)])
