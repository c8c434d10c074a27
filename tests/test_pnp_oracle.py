"""CPU: pins oracle/pnp_oracle.c — the restatement of the reference's pnp_solve (source/vision/pnp-solve.cpp:16-104 =
cv::solvePnPRansac with SOLVEPNP_P3P, 100 iterations, reprojection error 0.05) — against the live third-party routine
(cv2.solveP3P, cv2.solvePnPRansac) and the reference's own known-answer test (test/test-pnp.cpp:14-63)."""
import numpy as np
import pytest

from oracle import cbind as orc
from pnp_scenes import K_PNP, cube_rig, rodrigues, scene

cv2 = pytest.importorskip("cv2")


def test_quartic_roots():
    rng = np.random.default_rng(0)
    for _ in range(1500):
        kind = rng.integers(0, 3)
        if kind == 0:
            r = list(rng.uniform(-3, 3, 4))
        elif kind == 1:
            z = complex(rng.uniform(-2, 2), rng.uniform(0.001, 1)); r = [rng.uniform(-3, 3), rng.uniform(-3, 3), z, z.conjugate()]
        else:
            z = complex(rng.uniform(-2, 2), rng.uniform(0.001, 1)); w = complex(rng.uniform(-2, 2), rng.uniform(0.001, 1))
            r = [z, z.conjugate(), w, w.conjugate()]
        c = np.poly(r).real[::-1] * rng.uniform(0.1, 10)
        real = np.sort([complex(x).real for x in r if complex(x).imag == 0])
        got = np.sort(orc.solve_quartic(c))
        assert len(got) == len(real)
        if len(real):
            assert np.abs(got - real).max() < 1e-6


def test_p3p_contains_cv2_solutions_and_the_true_pose():
    rng = np.random.default_rng(1)
    K = K_PNP
    trials = missed_true = missed_cv = cv_total = 0
    for _ in range(400):
        R = rodrigues(rng.normal(size=3) * 0.3); t = rng.normal(size=3) * 0.5 + [0, 0, 6.0]
        X = rng.uniform(-2, 2, (3, 3)); Xc = X @ R.T + t
        uv = (Xc[:, :2] / Xc[:, 2:]) * 700 + [640, 360]
        f = np.concatenate([(uv - [640, 360]) / 700, np.ones((3, 1))], 1); f /= np.linalg.norm(f, axis=1, keepdims=True)
        Rs, ts = orc.p3p(f, X)
        trials += 1
        missed_true += min([np.abs(a - R).max() + np.abs(b - t).max() for a, b in zip(Rs, ts)] or [9]) > 1e-6
        _, rv, tv = cv2.solveP3P(X.reshape(3, 1, 3), uv.reshape(3, 1, 2), K, None, flags=cv2.SOLVEPNP_P3P)
        for r_, t_ in zip(rv, tv):
            Rc = cv2.Rodrigues(r_)[0]; tc = t_.ravel()
            pc = X @ Rc.T + tc
            if np.abs((pc[:, :2] / pc[:, 2:]) * 700 + [640, 360] - uv).max() > 1e-6:
                continue                      # cv2 sometimes returns a pose that does not reproject its own input
            cv_total += 1
            missed_cv += min([np.abs(a - Rc).max() + np.abs(b - tc).max() for a, b in zip(Rs, ts)] or [9]) > 1e-5
    assert missed_true <= 2 and missed_cv <= 2 and cv_total > trials      # degenerate triangles only


def test_reference_known_answer_cube():
    """test/test-pnp.cpp:14-63: cube rig, camera at x = +1, K = I, all 8 points inliers, se3 within 1e-3."""
    X, uv, K = cube_rig()
    o = orc.pnp_solve(X, uv, K, H=100, seed=0)
    assert o["status"] == 0 and o["n_inliers"] == 8 and o["mask"].all()
    assert np.abs(o["R"] - np.eye(3)).max() < 1e-9 and np.abs(o["t"] - [1, 0, 0]).max() < 1e-9


@pytest.mark.parametrize("n,outl,noise", [(50, 0.0, 0.0), (200, 0.3, 0.0), (500, 0.5, 0.0), (200, 0.3, 0.01)])
def test_matches_cv2_solvepnpransac(n, outl, noise):
    X, uv, R, t, good = scene(n, outl, noise, seed=n + int(100 * outl))
    o = orc.pnp_solve(X, uv, K_PNP, H=100, seed=1, reproj_error=0.05)
    ok, rv, tv, inl = cv2.solvePnPRansac(X.reshape(-1, 1, 3), uv.reshape(-1, 1, 2), K_PNP, None, iterationsCount=100,
                                         reprojectionError=0.05, confidence=0.95, flags=cv2.SOLVEPNP_P3P)
    assert ok and o["status"] == 0
    Rc = cv2.Rodrigues(rv)[0]
    Rm, tm = o["R"].T, -o["R"].T @ o["t"]                # oracle pose is camera-to-world like the reference's output
    tol = 1e-3                                            # the reference test's own tolerance (test-pnp.cpp:16)
    assert np.abs(Rm - Rc).max() < tol and np.abs(tm - tv.ravel()).max() < tol
    assert np.abs(Rm - R).max() < 1e-4 and np.abs(tm - t).max() < 1e-4
    if noise == 0.0:                                      # clean data + gross outliers: the inlier sets coincide
        assert set(inl.ravel()) == set(np.nonzero(o["mask"])[0]) == set(np.nonzero(good)[0])
        assert np.abs(Rm - R).max() < 1e-9


def test_failure_statuses():
    X, uv, *_ = scene(3, seed=3)
    assert orc.pnp_solve(X, uv, K_PNP)["status"] == 2           # < 4 points
    r = np.random.default_rng(0)
    o = orc.pnp_solve(r.uniform(-1, 1, (30, 3)) + [0, 0, 5], r.uniform(0, 700, (30, 2)), K_PNP, H=50)
    assert o["status"] in (0, 3)                                # pure clutter: at best the 4 sample points agree
    assert o["n_inliers"] <= 5
