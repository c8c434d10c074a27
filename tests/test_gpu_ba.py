"""GPU parity tests of bundle adjustment: mvs_ba_solve_batch (CUDA, through the C ABI) against the numpy oracle
(oracle/ba_np.py: the cost function of source/vision/ba.cpp:26-156, minimiser pinned against scipy) on the reference's
sfm_refine / pnp_refine problem shapes, and against the reference's known-answer tolerances.

Bar (floating point): final cost within 1e-9 relative, poses and points within 1e-8, marginal covariances within 1e-6
relative of the oracle's (both iterate Levenberg-Marquardt to a 1e-13 relative cost decrease)."""
import numpy as np
import pytest

import mvslam_b200 as mvs
from ba_scenes import multi_view, one_view, two_view
from oracle import ba_np as B

pytestmark = pytest.mark.gpu
NAN6 = np.full((6, 6), np.nan)


@pytest.fixture(scope="module")
def ctx():
    c = mvs.Context(0)
    yield c
    c.close()


def to_abi(prob, pose_prior_cov, point_prior_cov):
    """oracle Problem -> the C ABI's arrays (covariances, NaN = no prior)."""
    F, P = prob.F, prob.P
    pc = np.stack([pose_prior_cov.get(f, NAN6) for f in range(F)])
    xc = np.stack([point_prior_cov.get(j, np.full((3, 3), np.nan)) for j in range(P)]) if P else np.zeros((0, 3, 3))
    obs = np.zeros(len(prob.obs), mvs.BA_OBS_DTYPE)
    for i, (f, j, z, info) in enumerate(prob.obs):
        C = np.linalg.inv(info)
        obs[i] = (f, j, z, (C[0, 0], C[0, 1], C[1, 1]))
    return dict(pose_R=np.stack([R for R, _ in prob.poses0]), pose_t=np.stack([t for _, t in prob.poses0]), pose_prior_cov=pc,
                points=prob.points0, point_prior_cov=xc, obs=obs)


def sfm_case(seed, **kw):
    s = two_view(seed, **kw)
    n = len(s["p1"])
    prob = B.sfm_refine_problem(s["p1"], s["cov"], s["p2"], s["cov"], s["K"], s["pose_guess"], s["points_guess"])
    pp = {0: np.eye(6) * B.SFM_ANCHOR_STDDEV ** 2, 1: np.eye(6) * B.SFM_REGULATOR_STDDEV ** 2}
    xp = {j: np.eye(3) * B.SFM_REGULATOR_STDDEV ** 2 for j in range(n)}
    return s, prob, to_abi(prob, pp, xp)


def pnp_case(seed):
    s = one_view(seed)
    prob = B.pnp_refine_problem(s["world"], s["world_cov"], s["image"], s["image_cov"], s["K"], s["pose_guess"])
    return s, prob, to_abi(prob, {0: np.eye(6) * B.PNP_REGULATOR_STDDEV ** 2}, {j: s["world_cov"][j] for j in range(len(s["world"]))})


def multi_case(seed, **kw):
    s = multi_view(seed, **kw)
    prob = B.Problem(s["K"], s["poses"], s["pose_prior"], s["points"], s["point_prior"], s["obs"])
    return s, prob, to_abi(prob, s["pose_prior"], s["point_prior"])


def check(g, o):
    assert g["status"] == mvs.OK
    assert abs(g["final_error"] - o["error"]) <= 1e-9 * max(o["error"], 1e-12), (g["final_error"], o["error"])
    for f, (R, t) in enumerate(o["poses"]):
        assert np.abs(g["pose_R"][f] - R).max() < 1e-8 and np.abs(g["pose_t"][f] - t).max() < 1e-8
        assert np.abs(g["pose_cov"][f] - o["pose_cov"][f]).max() <= 1e-6 * np.abs(o["pose_cov"][f]).max()
    assert np.abs(g["points"] - o["points"]).max() < 1e-8
    for j, C in enumerate(o["point_cov"]):
        assert np.abs(g["point_cov"][j] - C).max() <= 1e-6 * np.abs(C).max()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_sfm_refine_L_shape_matches_oracle(ctx, seed):
    s, prob, abi = sfm_case(seed)
    g = ctx.ba_solve_batch(s["K"], [abi])[0]
    check(g, prob.solve())
    R2, t2, X = s["truth"]                                     # the reference's own bar: 0.025 (test-sfm.cpp:159)
    assert np.abs(g["pose_t"][1] - t2).max() < 0.025 and np.abs(g["points"] - X).max() < 0.025
    assert g["final_error"] <= g["initial_error"]


def test_pnp_refine_L_shape_matches_oracle(ctx):
    for seed in (20, 21):
        s, prob, abi = pnp_case(seed)
        g = ctx.ba_solve_batch(s["K"], [abi])[0]
        check(g, prob.solve())
        assert np.abs(g["pose_t"][0] - s["truth"][1]).max() < 0.025      # test-pnp.cpp:67


def test_pixel_intrinsics_with_skew_and_many_points(ctx):
    K = np.array([[700.0, 1.5, 640.0], [0, 690.0, 360.0], [0, 0, 1.0]])
    s, prob, abi = sfm_case(5, n=300, noise=0.5 / 700, K=K)
    check(ctx.ba_solve_batch(K, [abi])[0], prob.solve())


def test_batch_of_mixed_problems(ctx):
    cases = [sfm_case(30), pnp_case(31), sfm_case(32, n=40), pnp_case(33), sfm_case(34)]
    K = np.eye(3)
    res = ctx.ba_solve_batch(K, [c[2] for c in cases])
    for (s, prob, abi), g in zip(cases, res):
        check(g, prob.solve())
    again = ctx.ba_solve_batch(K, [c[2] for c in cases])       # deterministic: bit-identical on a second run
    for a, b in zip(res, again):
        assert a["final_error"] == b["final_error"] and np.array_equal(a["points"], b["points"])


def test_argument_checks(ctx):
    s, prob, abi = sfm_case(40)
    bad = dict(abi, obs=abi["obs"].copy()); bad["obs"]["point"][0] = 99
    with pytest.raises(mvs.MvsError):
        ctx.ba_solve_batch(np.eye(3), [bad])
    many = dict(abi, pose_R=np.stack([np.eye(3)] * 17), pose_t=np.zeros((17, 3)), pose_prior_cov=np.stack([np.eye(6)] * 17))
    with pytest.raises(mvs.MvsError) as e:
        ctx.ba_solve_batch(np.eye(3), [many])
    assert e.value.status == mvs.E_UNSUPPORTED


@pytest.mark.parametrize("n_frames,n", [(3, 40), (4, 60), (8, 120), (16, 80)])
def test_more_than_two_frames_matches_oracle(ctx, n_frames, n):
    """ba_frame_pose_and_point accepts any number of frames (ba.cpp:26-156); 3..16 run in ba_solve_multi_kernel."""
    s, prob, abi = multi_case(100 + n_frames, n_frames=n_frames, n=n)
    g = ctx.ba_solve_batch(s["K"], [abi])[0]
    o = prob.solve()
    check(g, o)
    assert g["final_error"] < g["initial_error"]
    poses, X = s["truth"]
    assert max(np.abs(g["pose_t"][f] - poses[f][1]).max() for f in range(n_frames)) < 0.05


def test_mixed_frame_counts_in_one_batch(ctx):
    """problems of 1, 2 and more frames in one call: two kernels, each takes its own problems; deterministic."""
    K = np.array([[700.0, 0.8, 320.0], [0, 705.0, 240.0], [0, 0, 1.0]])
    cases = [multi_case(200, n_frames=5, n=50, K=K, noise=0.4 / 700), sfm_case(201, n=30, noise=0.5 / 700, K=K),
             multi_case(202, n_frames=3, n=25, K=K, noise=0.4 / 700, unprior_points=1.0)]
    res = ctx.ba_solve_batch(K, [c[2] for c in cases])
    for (s, prob, abi), g in zip(cases, res):
        check(g, prob.solve())
    again = ctx.ba_solve_batch(K, [c[2] for c in cases])
    for a, b in zip(res, again):
        assert a["final_error"] == b["final_error"] and np.array_equal(a["points"], b["points"]) and np.array_equal(a["pose_cov"], b["pose_cov"])
