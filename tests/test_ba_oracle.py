"""CPU: pins oracle/ba_np.py (the restatement of ba_frame_pose_and_point's cost function, source/vision/ba.cpp:26-156,
and of its callers sfm_refine / pnp_refine) — Jacobians against finite differences, the minimiser against
scipy.optimize.least_squares, and the reference's own known-answer tests (tolerance 0.025)."""
import numpy as np
import pytest

from ba_scenes import multi_view, one_view, two_view
from oracle import ba_np as B

scipy_opt = pytest.importorskip("scipy.optimize")


def make_two_view(seed, **kw):
    s = two_view(seed, **kw)
    return s, B.sfm_refine_problem(s["p1"], s["cov"], s["p2"], s["cov"], s["K"], s["pose_guess"], s["points_guess"])


def test_jacobians_match_finite_differences():
    s, prob = make_two_view(3)
    x0 = np.random.default_rng(0).normal(size=6 * prob.F + 3 * prob.P) * 1e-3
    poses, points = prob.retract(prob.poses0, prob.points0, x0)
    H, g = prob.normal_equations(poses, points)
    # gradient of the cost w.r.t. a local perturbation at (poses, points)
    num = np.zeros_like(g)
    for k in range(len(g)):
        d = np.zeros(len(g)); d[k] = 1e-6
        cp = prob.cost(*prob.retract(poses, points, d)); cm = prob.cost(*prob.retract(poses, points, -d))
        num[k] = (cp - cm) / 2e-6
    assert np.abs(num - g).max() < 1e-5 * max(1.0, np.abs(g).max())


@pytest.mark.parametrize("seed", [1, 2])
def test_minimiser_matches_scipy(seed):
    s, prob = make_two_view(seed)
    mine = prob.solve()
    ref = scipy_opt.least_squares(prob.residual_vector, np.zeros(6 * prob.F + 3 * prob.P), method="lm", xtol=1e-15, ftol=1e-15, gtol=1e-15)
    poses_ref, points_ref = prob.retract(prob.poses0, prob.points0, ref.x)
    assert abs(mine["error"] - ref.cost) < 1e-9 * max(ref.cost, 1.0)
    assert np.abs(mine["points"] - points_ref).max() < 1e-7
    for (Ra, ta), (Rb, tb) in zip(mine["poses"], poses_ref):
        assert np.abs(Ra - Rb).max() < 1e-7 and np.abs(ta - tb).max() < 1e-7


def test_reference_known_answer_sfm_refine_L_shape():
    """test/test-sfm.cpp:157-290: pose and points recovered within 0.025."""
    for seed in range(5):
        s, prob = make_two_view(10 + seed)
        r = prob.solve()
        R2, t2, X = s["truth"]
        assert np.abs(r["poses"][1][1] - t2).max() < 0.025 and np.abs(B.so3_log(R2.T @ r["poses"][1][0])).max() < 0.025
        assert np.abs(r["points"] - X).max() < 0.025
        assert np.abs(r["poses"][0][1]).max() < 1e-6                       # anchored camera 1
        assert all(np.all(np.linalg.eigvalsh(c) > 0) for c in r["pose_cov"] + r["point_cov"])


def test_reference_known_answer_pnp_refine_L_shape():
    """test/test-pnp.cpp:65-160."""
    for seed in range(5):
        s = one_view(20 + seed)
        prob = B.pnp_refine_problem(s["world"], s["world_cov"], s["image"], s["image_cov"], s["K"], s["pose_guess"])
        r = prob.solve()
        R, t, _ = s["truth"]
        assert np.abs(r["poses"][0][1] - t).max() < 0.025 and np.abs(B.so3_log(R.T @ r["poses"][0][0])).max() < 0.025


@pytest.mark.parametrize("n_frames", [3, 5])
def test_minimiser_matches_scipy_with_more_than_two_frames(n_frames):
    """ba_frame_pose_and_point takes any number of frames (ba.cpp:26-156): the same pin for a window of cameras."""
    s = multi_view(70 + n_frames, n_frames=n_frames, n=25)
    prob = B.Problem(s["K"], s["poses"], s["pose_prior"], s["points"], s["point_prior"], s["obs"])
    mine = prob.solve()
    ref = scipy_opt.least_squares(prob.residual_vector, np.zeros(6 * prob.F + 3 * prob.P), method="lm", xtol=1e-15, ftol=1e-15, gtol=1e-15)
    poses_ref, points_ref = prob.retract(prob.poses0, prob.points0, ref.x)
    assert abs(mine["error"] - ref.cost) < 1e-9 * max(ref.cost, 1.0)
    assert np.abs(mine["points"] - points_ref).max() < 1e-6
    for (Ra, ta), (Rb, tb) in zip(mine["poses"], poses_ref):
        assert np.abs(Ra - Rb).max() < 1e-7 and np.abs(ta - tb).max() < 1e-7
