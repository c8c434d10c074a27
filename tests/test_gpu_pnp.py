"""GPU parity tests of pnp_solve: mvs_pnp_solve (CUDA, through the C ABI) against the CPU oracle (oracle/pnp_oracle.c)
on the same seeded inputs and sample tables, and against the reference's known-answer test (test/test-pnp.cpp:14-63).

Bars: per-hypothesis consensus sizes, the winning hypothesis, its minimal-sample pose and the inlier set are bit-exact
(shared arithmetic contract: + - * / sqrt in IEEE double, same order); the refined pose agrees to 1e-9 (summation order)."""
import numpy as np
import pytest

import mvslam_b200 as mvs
from oracle import cbind as orc
from pnp_scenes import K_PNP, cube_rig, scene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = mvs.Context(0)
    yield c
    c.close()


def check_against_oracle(g, o):
    assert g["status"] == o["status"]
    assert np.array_equal(g["all_counts"], o["all_counts"])
    assert g["n_inliers"] == o["n_inliers"] and g["best_hypothesis"] == o["best_h"]
    if g["status"] == mvs.OK:
        assert np.array_equal(g["mask"], o["mask"])
        assert np.array_equal(g["R_w2c_p3p"], o["R_p3p"]) and np.array_equal(g["t_w2c_p3p"], o["t_p3p"])
        assert np.abs(g["R_c2w"] - o["R"]).max() < 1e-9 and np.abs(g["t_c2w"] - o["t"]).max() < 1e-9


def test_host_sample_table_matches_oracle():
    for seed, pid, n, H in [(0, 0, 4, 8), (1, 5, 91, 100), (2**63, 2**40, 5000, 300)]:
        assert np.array_equal(mvs.pnp_sample_table(seed, pid, n, H), orc.pnp_sample_table(seed, pid, n, H))


def test_reference_known_answer_cube(ctx):
    X, uv, K = cube_rig()
    g = ctx.pnp_solve(X, uv, K, H=100, want_all=True)
    assert g["status"] == mvs.OK and g["n_inliers"] == 8 and g["mask"].all()
    assert np.abs(g["R_c2w"] - np.eye(3)).max() < 1e-9 and np.abs(g["t_c2w"] - [1, 0, 0]).max() < 1e-9
    check_against_oracle(g, orc.pnp_solve(X, uv, K, H=100))


@pytest.mark.parametrize("n,outl,noise,H", [(7, 0.0, 0.0, 100), (50, 0.0, 0.0, 100), (200, 0.3, 0.0, 100), (300, 0.3, 0.02, 257),
                                            (1000, 0.5, 0.01, 1024), (5000, 0.6, 0.02, 512)])
def test_matches_oracle(ctx, n, outl, noise, H):
    X, uv, R, t, good = scene(n, outl, noise, seed=7 * n + H)
    g = ctx.pnp_solve(X, uv, K_PNP, H=H, seed=3, problem_id=11, want_all=True)
    o = orc.pnp_solve(X, uv, K_PNP, H=H, seed=3, problem_id=11)
    check_against_oracle(g, o)
    assert g["status"] == mvs.OK
    Rw2c, tw2c = g["R_c2w"].T, -g["R_c2w"].T @ g["t_c2w"]
    assert np.abs(Rw2c - R).max() < 2e-4 and np.abs(tw2c - t).max() < 2e-3


def test_explicit_sample_table_and_no_refinement(ctx):
    X, uv, *_ = scene(120, 0.25, 0.0, seed=5)
    tab = orc.pnp_sample_table(9, 0, 120, 64)
    g = ctx.pnp_solve(X, uv, K_PNP, samples=tab, refine_iters=0, want_all=True)
    o = orc.pnp_solve(X, uv, K_PNP, samples=tab, refine_iters=0)
    check_against_oracle(g, o)
    assert np.array_equal(g["R_c2w"], g["R_w2c_p3p"].T)          # no refinement: the pose is the P3P pose inverted


def test_failures(ctx):
    X, uv, *_ = scene(3, seed=3)
    assert ctx.pnp_solve(X, uv, K_PNP)["status"] == mvs.E_TOO_FEW_POINTS
    r = np.random.default_rng(0)
    Xc = r.uniform(-1, 1, (30, 3)) + [0, 0, 5]; uc = r.uniform(0, 700, (30, 2))
    g = ctx.pnp_solve(Xc, uc, K_PNP, H=50, want_all=True)
    o = orc.pnp_solve(Xc, uc, K_PNP, H=50)
    check_against_oracle(g, o)
    with pytest.raises(mvs.MvsError):
        ctx.pnp_solve(X, uv, np.zeros((3, 3)))


def test_batch_equals_single_calls(ctx):
    probs = [scene(n, 0.3, 0.005, seed=40 + i) for i, n in enumerate([60, 7, 3, 500, 0, 129, 256, 257])]
    res, masks = ctx.pnp_solve_batch([p[0] for p in probs], [p[1] for p in probs], K_PNP, H=128, seed=2, problem_id_base=100)
    for i, p in enumerate(probs):
        if len(p[0]) == 0:
            assert res["status"][i] == mvs.E_TOO_FEW_POINTS
            continue
        g = ctx.pnp_solve(p[0], p[1], K_PNP, H=128, seed=2, problem_id=100 + i)
        assert res["status"][i] == g["status"] and res["n_inliers"][i] == g["n_inliers"]
        assert res["best_hypothesis"][i] == g["best_hypothesis"]
        if g["status"] == mvs.OK:
            assert np.array_equal(masks[i], g["mask"]) and np.array_equal(res["R_c2w"][i], g["R_c2w"])


def test_python_mirror_of_the_reference_signature(ctx):
    X, uv, R, t, good = scene(150, 0.2, 0.0, seed=77)
    ok, (Rc2w, tc2w), inliers = mvs.pnp_solve(X, uv, K_PNP, ctx=ctx)
    assert ok and set(inliers.tolist()) == set(np.nonzero(good)[0].tolist())
    assert np.abs(Rc2w - R.T).max() < 1e-9 and np.abs(tc2w + R.T @ t).max() < 1e-9
