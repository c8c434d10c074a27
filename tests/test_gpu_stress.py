"""GPU parity, wider nets: many seeded sizes and degenerate configurations for the stages of SURVEY §8(f) against their
CPU oracles — same bars as the focused tests (bit-exact extraction; bit-exact pnp consensus; BA within 1e-8)."""
import numpy as np
import pytest

import mvslam_b200 as mvs
from mvslam_b200 import synth
from oracle import ba_np as B
from oracle import cbind as orc
from oracle import orb_np as O
from pnp_scenes import K_PNP, rodrigues, scene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = mvs.Context(0)
    yield c
    c.close()


def test_extraction_random_sizes_match_oracle(ctx):
    rng = np.random.default_rng(2024)
    for k in range(10):
        w, h = int(rng.integers(64, 700)), int(rng.integers(64, 500))
        nf = int(rng.choice([50, 300, 500, 1200, 3000]))
        img = synth.synthetic_image(3000 + k, w, h, n_shapes=int(rng.integers(20, 300)))
        if k % 3 == 0:                                     # hard contrast: saturated regions and exact ties
            img = np.where(img > 128, 255, 0).astype(np.uint8)
        counts, kp, desc, _ = ctx.orb_extract([img], nf)
        ref = O.orb_extract(img, nf)
        what = f"{w}x{h} nfeatures {nf}"
        assert counts[0] == len(ref["pt"]), what
        assert np.array_equal(kp["x"], ref["pt"][:, 0]) and np.array_equal(kp["y"], ref["pt"][:, 1]), what
        assert np.array_equal(kp["octave"], ref["octave"]) and np.array_equal(kp["response"], ref["response"]), what
        assert np.array_equal(kp["angle"], ref["angle"]) and np.array_equal(desc, ref["desc"]), what


def test_extraction_mixed_batch_sizes_share_a_context(ctx):
    """Geometry is cached per (width, height, nfeatures): alternate shapes and check nothing leaks between them."""
    a = synth.synthetic_image(41, 320, 240); b = synth.synthetic_image(42, 200, 320)
    ra, rb = O.orb_extract(a, 400), O.orb_extract(b, 900)
    for _ in range(2):
        _, _, da, _ = ctx.orb_extract([a, a], 400)
        _, _, db, _ = ctx.orb_extract([b], 900)
        assert np.array_equal(da[:len(ra["desc"])], ra["desc"]) and np.array_equal(db, rb["desc"])


def check_pnp(g, o):
    assert g["status"] == o["status"] and np.array_equal(g["all_counts"], o["all_counts"])
    assert g["best_hypothesis"] == o["best_h"]
    if g["status"] == mvs.OK:
        assert np.array_equal(g["mask"], o["mask"]) and np.array_equal(g["R_w2c_p3p"], o["R_p3p"])
        assert np.abs(g["R_c2w"] - o["R"]).max() < 1e-9 and np.abs(g["t_c2w"] - o["t"]).max() < 1e-9


def test_pnp_random_and_degenerate_scenes_match_oracle(ctx):
    rng = np.random.default_rng(7)
    for k in range(25):
        n = int(rng.integers(4, 400))
        X, uv, *_ = scene(n, float(rng.choice([0.0, 0.2, 0.6])), float(rng.choice([0.0, 0.01, 0.5])), seed=900 + k)
        if k % 5 == 1:
            X[:, 2] = 1.0                                  # planar scene (images unchanged: inconsistent but well defined)
        if k % 5 == 2:
            X[:4] = X[0] + np.outer(np.arange(4), [1.0, 0.5, 0.25])     # the default sample {0,1,2,3} is collinear
        if k % 5 == 3:
            X[1] = X[0]; uv[1] = uv[0]                      # duplicated correspondence inside the default sample
        H = int(rng.choice([1, 33, 100, 300]))
        g = ctx.pnp_solve(X, uv, K_PNP, H=H, seed=k, problem_id=k, want_all=True)
        check_pnp(g, orc.pnp_solve(X, uv, K_PNP, H=H, seed=k, problem_id=k))


def test_ba_partial_priors_and_single_view_points(ctx):
    """Points without a prior, points seen by one camera only, a skewed K: still the oracle's minimum and covariances."""
    rng = np.random.default_rng(11)
    K = np.array([[650.0, 2.0, 300.0], [0, 640.0, 250.0], [0, 0, 1.0]])
    n = 60
    X = np.stack([rng.uniform(-1.5, 1.5, n), rng.uniform(-1.5, 1.5, n), rng.uniform(3, 6, n)], 1)
    R2, t2 = rodrigues(rng.normal(size=3) * 0.03), np.array([0.8, 0.1, -0.05])
    proj = lambda R, t, P: (lambda pc: np.stack([K[0, 0] * pc[:, 0] / pc[:, 2] + K[0, 1] * pc[:, 1] / pc[:, 2] + K[0, 2],  # noqa: E731
                                                  K[1, 1] * pc[:, 1] / pc[:, 2] + K[1, 2]], 1))((P - t) @ R)
    sig = 0.4
    p1 = proj(np.eye(3), np.zeros(3), X) + rng.normal(size=(n, 2)) * sig
    p2 = proj(R2, t2, X) + rng.normal(size=(n, 2)) * sig
    cov = np.array([[sig * sig, 0.02], [0.02, 1.5 * sig * sig]])
    obs = [(0, j, p1[j], cov) for j in range(n)] + [(1, j, p2[j], cov) for j in range(n) if j % 4]      # every 4th: one view
    pose_prior = {0: np.eye(6) * 1e-10, 1: np.diag([1e-3, 1e-3, 1e-3, 1e-2, 1e-2, 1e-2])}
    point_prior = {j: np.diag([1e-2, 2e-2, 4e-2]) for j in range(n) if j % 3 == 0 or j % 4 == 0}      # others: none
    guess2 = (R2 @ rodrigues(rng.normal(size=3) * 5e-3), t2 + rng.normal(size=3) * 1e-2)
    prob = B.Problem(K, [(np.eye(3), np.zeros(3)), guess2], pose_prior, X + rng.normal(size=X.shape) * 1e-2, point_prior, obs)
    nan6, nan3 = np.full((6, 6), np.nan), np.full((3, 3), np.nan)
    o_arr = np.zeros(len(obs), mvs.BA_OBS_DTYPE)
    for i, (f, j, z, C) in enumerate(obs):
        o_arr[i] = (f, j, z, (C[0, 0], C[0, 1], C[1, 1]))
    abi = dict(pose_R=np.stack([np.eye(3), guess2[0]]), pose_t=np.stack([np.zeros(3), guess2[1]]),
               pose_prior_cov=np.stack([pose_prior.get(f, nan6) for f in range(2)]), points=prob.points0,
               point_prior_cov=np.stack([point_prior.get(j, nan3) for j in range(n)]), obs=o_arr)
    g = ctx.ba_solve_batch(K, [abi])[0]
    o = prob.solve()
    assert g["status"] == mvs.OK and abs(g["final_error"] - o["error"]) <= 1e-8 * o["error"]
    assert np.abs(g["points"] - o["points"]).max() < 1e-7 and np.abs(g["pose_t"][1] - o["poses"][1][1]).max() < 1e-8
    for j in range(n):
        assert np.abs(g["point_cov"][j] - o["point_cov"][j]).max() <= 1e-5 * np.abs(o["point_cov"][j]).max()


def test_ba_many_random_problems_in_one_batch(ctx):
    """40 seeded problems of mixed shape (one- and two-frame, 6..150 points) in a single call: cost, poses, points and
    every marginal covariance against the oracle."""
    from test_gpu_ba import pnp_case, sfm_case
    cases = []
    for k in range(40):
        if k % 4 == 3:
            cases.append(pnp_case(500 + k))
        else:
            n = [None, 6, 25, 150][k % 4] if k % 4 else None
            cases.append(sfm_case(500 + k, n=n) if n else sfm_case(500 + k))
    res = ctx.ba_solve_batch(np.eye(3), [c[2] for c in cases])
    for k, ((s, prob, abi), g) in enumerate(zip(cases, res)):
        o = prob.solve()
        assert g["status"] == mvs.OK, k
        assert abs(g["final_error"] - o["error"]) <= 1e-8 * max(o["error"], 1e-12), k
        for f, (R, t) in enumerate(o["poses"]):
            assert np.abs(g["pose_R"][f] - R).max() < 1e-7 and np.abs(g["pose_t"][f] - t).max() < 1e-7, k
            assert np.abs(g["pose_cov"][f] - o["pose_cov"][f]).max() <= 1e-5 * np.abs(o["pose_cov"][f]).max(), k
        assert np.abs(g["points"] - o["points"]).max() < 1e-7, k
        for j, C in enumerate(o["point_cov"]):
            assert np.abs(g["point_cov"][j] - C).max() <= 1e-5 * np.abs(C).max(), (k, j)
