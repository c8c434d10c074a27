"""Seeded bundle-adjustment test problems shared by the oracle-pin (CPU) and parity (GPU) tests: the reference's
L-shape rig (test/unit-test-helper.cpp:57-72; test/test-sfm.cpp:157-290, test/test-pnp.cpp:65-160) and random scenes."""
import numpy as np

from pnp_scenes import rodrigues


def l_shape_rig():
    """get_rig_points(L_SHAPE, SO3(yaw 1.5, pitch 0.7, roll 0), (0.6, 0, 3), 0.5)"""
    pts = np.array([[1, 0, 0], [0, 0, 0], [0, 2, 0], [1, 0, 3], [0, 0, 3], [0, 2, 3], [0.5, 0, 1.5], [0, 1, 1.5]], float)
    y, p = 1.5, 0.7                                   # SO3(yaw, pitch, roll) = Rz(yaw) Ry(pitch) Rx(roll)
    Rz = np.array([[np.cos(y), -np.sin(y), 0], [np.sin(y), np.cos(y), 0], [0, 0, 1]])
    Ry = np.array([[np.cos(p), 0, np.sin(p)], [0, 1, 0], [-np.sin(p), 0, np.cos(p)]])
    return 0.5 * pts @ (Rz @ Ry).T + [0.6, 0.0, 3.0]


def two_view(seed, n=None, noise=5e-3, K=None, X=None):
    """sfm_refine_L_shape-style problem: camera 1 at the origin, camera 2 at x = +1; noisy observations and guesses."""
    r = np.random.default_rng(seed)
    K = np.eye(3) if K is None else K
    X = l_shape_rig() if X is None else X
    if n is not None:
        X = np.stack([r.uniform(-1.5, 1.5, n), r.uniform(-1.5, 1.5, n), r.uniform(2.5, 5.0, n)], 1)
    R2, t2 = rodrigues(r.normal(size=3) * 0.02), np.array([1.0, 0, 0])          # camera 2 in camera 1 (camera to world)

    def proj(R, t, P):
        pc = (P - t) @ R                                                           # R^T (P - t)
        return np.stack([K[0, 0] * pc[:, 0] / pc[:, 2] + K[0, 1] * pc[:, 1] / pc[:, 2] + K[0, 2],
                         K[1, 1] * pc[:, 1] / pc[:, 2] + K[1, 2]], 1)
    sig = noise * max(K[0, 0], 1.0) if K[0, 0] > 10 else noise
    p1 = proj(np.eye(3), np.zeros(3), X) + r.normal(size=(len(X), 2)) * sig
    p2 = proj(R2, t2, X) + r.normal(size=(len(X), 2)) * sig
    cov = [np.eye(2) * sig ** 2] * len(X)
    guess = (R2 @ rodrigues(r.normal(size=3) * 1e-2), t2 + r.normal(size=3) * 5e-3)
    Xg = X + r.normal(size=X.shape) * 5e-3
    return dict(K=K, p1=p1, p2=p2, cov=cov, pose_guess=guess, points_guess=Xg, truth=(R2, t2, X))


def one_view(seed, noise=5e-3):
    """pnp_refine_L_shape-style problem: camera at x = +1 looking at the L-shape rig, noisy points with priors."""
    r = np.random.default_rng(seed)
    X = l_shape_rig()
    R, t = np.eye(3), np.array([1.0, 0, 0])
    pc = (X - t) @ R
    uv = pc[:, :2] / pc[:, 2:] + r.normal(size=(len(X), 2)) * noise
    Xn = X + r.normal(size=X.shape) * 5e-3
    guess = (rodrigues(r.normal(size=3) * 5e-3), t + r.normal(size=3) * 5e-3)
    return dict(K=np.eye(3), world=Xn, world_cov=[np.eye(3) * 25e-6] * len(X), image=uv, image_cov=[np.eye(2) * noise ** 2] * len(X),
                pose_guess=guess, truth=(R, t, X))


def multi_view(seed, n_frames=4, n=60, noise=3e-3, K=None, drop=0.3, unprior_points=0.25):
    """A window of n_frames cameras on a slow arc looking at a random cloud (ba_frame_pose_and_point with more than two
    frames, source/vision/ba.cpp:26-156): every point is seen by at least two frames, `drop` of the other observations are
    missing, frame 0 is anchored, the others and most points carry a loose prior.  Returns what oracle.ba_np.Problem takes."""
    r = np.random.default_rng(seed)
    K = np.eye(3) if K is None else K
    X = np.stack([r.uniform(-1.5, 1.5, n), r.uniform(-1.2, 1.2, n), r.uniform(3.0, 6.0, n)], 1)
    poses = []
    for f in range(n_frames):
        poses.append((rodrigues(np.array([0.01, -0.04, 0.005]) * f + r.normal(size=3) * 0.005), np.array([0.35 * f, 0.02 * f, 0.03 * f])))
    sig = noise * K[0, 0] if K[0, 0] > 10 else noise
    obs = []
    for j in range(n):
        seen = [f for f in range(n_frames) if r.uniform() > drop]
        while len(seen) < 2:
            f = int(r.integers(n_frames))
            if f not in seen:
                seen.append(f)
        for f in sorted(seen):
            R, t = poses[f]
            pc = R.T @ (X[j] - t)
            uv = np.array([K[0, 0] * pc[0] / pc[2] + K[0, 1] * pc[1] / pc[2] + K[0, 2], K[1, 1] * pc[1] / pc[2] + K[1, 2]])
            obs.append((f, j, uv + r.normal(size=2) * sig, np.eye(2) * sig ** 2))
    order = r.permutation(len(obs))                                    # the caller's observation order is arbitrary
    obs = [obs[i] for i in order]
    guess = [poses[0]] + [(R @ rodrigues(r.normal(size=3) * 5e-3), t + r.normal(size=3) * 5e-3) for R, t in poses[1:]]
    Xg = X + r.normal(size=X.shape) * 5e-3
    pose_prior = {0: np.eye(6) * 1e-10}
    pose_prior.update({f: np.eye(6) * 1e-4 for f in range(1, n_frames)})
    point_prior = {j: np.eye(3) * 1e-4 for j in range(n) if r.uniform() > unprior_points}
    return dict(K=K, poses=guess, pose_prior=pose_prior, points=Xg, point_prior=point_prior, obs=obs, truth=(poses, X))
