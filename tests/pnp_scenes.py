"""Seeded PnP test scenes shared by the oracle-pin (CPU) and parity (GPU) tests."""
import numpy as np

K_PNP = np.array([[700.0, 0, 640.0], [0, 700.0, 360.0], [0, 0, 1.0]])


def rodrigues(v):
    th = np.linalg.norm(v)
    if th < 1e-12:
        return np.eye(3)
    k = v / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx


def scene(n, outl=0.0, noise=0.0, seed=0, K=K_PNP):
    """n world points seen by a camera with pose (R, t) world->camera; the first outl*n image points are replaced by
    uniform clutter, then everything is permuted.  Returns world[n,3], image[n,2], R, t, good[n] (bool)."""
    r = np.random.default_rng(seed)
    R = rodrigues(r.normal(size=3) * 0.3)
    t = r.normal(size=3) * 0.5 + np.array([0, 0, 6.0])
    X = r.uniform(-2, 2, (n, 3))
    Xc = X @ R.T + t
    uv = (Xc[:, :2] / Xc[:, 2:]) * [K[0, 0], K[1, 1]] + [K[0, 2], K[1, 2]] + r.normal(size=(n, 2)) * noise
    no = int(outl * n)
    uv[:no] = r.uniform(0, 1, (no, 2)) * [1280, 720]
    p = r.permutation(n)
    return X[p], uv[p], R, t, p >= no


def cube_rig():
    """get_rig_points(CUBE, identity, (0.6, 0, 3), 1) of the reference's test/unit-test-helper.cpp, seen from the camera at
    x = +1 with K = I (test/test-pnp.cpp:14-63)."""
    pts = np.array([[x, y, z] for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)], float) + [0.6, 0.0, 3.0]
    cam = pts - [1.0, 0, 0]
    return pts, cam[:, :2] / cam[:, 2:], np.eye(3)
