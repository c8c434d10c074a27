"""GPU parity tests: the CUDA path, called through the C ABI (libmvslam_b200.so via ctypes),
against the CPU oracle (oracle/mvs_oracle.c) on the same seeded inputs, against the committed golden
fixtures, and — at the BASELINE sizes — through size-independent properties.

Bars: Hamming matches bit-exact (indices, distances, order); identical inlier sets / counts for
identical sample tables; E within 1e-5 (Frobenius-normalised, sign-free), pose within 1e-6,
points within 1e-4 relative (the north-star tolerances), and much tighter where stated.
"""
import os

import numpy as np
import pytest

import mvslam_b200 as mvs
from mvslam_b200 import shard, synth
from oracle import cbind as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = mvs.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def ctx_popc():
    """A context on the integer-pipe matcher (knn2_hamming_kernel): the kernel choice is read at mvs_create."""
    old = os.environ.get("MVS_MATCHER")
    os.environ["MVS_MATCHER"] = "popc"
    try:
        c = mvs.Context(0)
    finally:
        if old is None:
            del os.environ["MVS_MATCHER"]
        else:
            os.environ["MVS_MATCHER"] = old
    yield c
    c.close()


def same_up_to_scale(Fa, Fb, atol):
    Fa = Fa / np.linalg.norm(Fa); Fb = Fb / np.linalg.norm(Fb)
    return min(np.abs(Fa - Fb).max(), np.abs(Fa + Fb).max()) <= atol


def as_mvs(m):
    return m.view(mvs.MATCH_DTYPE)


def general_scene(n, seed, noise=0.0, outl=0.0):
    r = np.random.default_rng(seed)
    K = synth.K_S8K
    X = np.stack([r.uniform(-4, 4, n), r.uniform(-4, 4, n), r.uniform(4, 12, n)], 1)
    rv = r.normal(size=3); rv *= 0.1 / np.linalg.norm(rv)
    R = synth._rodrigues(rv); t = r.normal(size=3); t *= 0.5 / np.linalg.norm(t)
    x1, _ = synth._project(K, np.eye(3), np.zeros(3), X)
    x2, _ = synth._project(K, R, t, X)
    x1 = x1 + r.normal(size=x1.shape) * noise; x2 = x2 + r.normal(size=x2.shape) * noise
    no = int(outl * n)
    x2[:no] = r.uniform(0, 1, (no, 2)) * [1280, 720]
    p = r.permutation(n)
    return K, x1[p], x2[p]


# ------------------------------------------------------------------------------------------ matcher
@pytest.mark.parametrize("nq,nt,rand_bytes", [(1, 2, 32), (7, 3, 32), (257, 300, 32), (500, 411, 3), (1759, 1748, 32),
                                              (3000, 2500, 2), (256, 256, 32), (513, 8192, 32)])
def test_knn2_hamming_bit_exact(ctx, nq, nt, rand_bytes):
    rng = np.random.default_rng(nq * 7919 + nt)
    q = np.zeros((nq, 32), np.uint8); t = np.zeros((nt, 32), np.uint8)
    q[:, :rand_bytes] = rng.integers(0, 256, (nq, rand_bytes)); t[:, :rand_bytes] = rng.integers(0, 256, (nt, rand_bytes))
    t[nt // 2] = t[0]
    ig, dg = ctx.knn2_hamming(q, t)
    io, do = orc.knn2_hamming(q, t)
    assert np.array_equal(ig, io) and np.array_equal(dg, do)


def test_tensor_matcher_extreme_distances_and_duplicates(ctx):
    """knn2_hamming_tc_kernel: S = 256 - 2 d must be exact at both ends (d = 0, d = 256), with duplicated train rows
    (lowest index wins) and a train set that ends inside a 128-row tile."""
    rng = np.random.default_rng(5)
    t = rng.integers(0, 256, (200, 32), dtype=np.uint8)
    t[0] = 0; t[1] = 255; t[7] = t[150]; t[199] = t[3]
    q = np.concatenate([t[:5], ~t[:5], np.zeros((1, 32), np.uint8), np.full((1, 32), 255, np.uint8)])
    ig, dg = ctx.knn2_hamming(q, t)
    io, do = orc.knn2_hamming(q, t)
    assert np.array_equal(ig, io) and np.array_equal(dg, do)
    assert dg[0, 0] == 0 and dg[10, 0] == 0 and dg[11, 0] == 0
    only = ctx.knn2_hamming(np.zeros((3, 32), np.uint8), np.full((2, 32), 255, np.uint8))
    assert np.array_equal(only[1], np.full((3, 2), 256)) and np.array_equal(only[0], np.tile([0, 1], (3, 1)))


@pytest.mark.parametrize("nq,nt,rand_bytes", [(129, 128, 32), (128, 129, 32), (1000, 127, 1), (2000, 2047, 32), (300, 4097, 2)])
@pytest.mark.parametrize("cross", [False, True])
def test_tensor_matcher_equals_popc_matcher(ctx, ctx_popc, nq, nt, rand_bytes, cross):
    rng = np.random.default_rng(nq + 3 * nt)
    q = np.zeros((nq, 32), np.uint8); t = np.zeros((nt, 32), np.uint8)
    q[:, :rand_bytes] = rng.integers(0, 256, (nq, rand_bytes)); t[:, :rand_bytes] = rng.integers(0, 256, (nt, rand_bytes))
    a, b = ctx.knn2_hamming(q, t), ctx_popc.knn2_hamming(q, t)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    ma, mb = ctx.match_hamming(q, t, 0.8, -1.0, cross), ctx_popc.match_hamming(q, t, 0.8, -1.0, cross)
    assert np.array_equal(ma, mb) and np.array_equal(ma, as_mvs(orc.match_hamming(q, t, 0.8, -1.0, cross)))


@pytest.mark.parametrize("cross", [False, True])
def test_tensor_matcher_second_neighbour_is_a_stream_mate(ctx, cross):
    """The tensor-core epilogue keeps one maximum per column stream (train index mod 4 inside a 64-row block) and K2 looks the
    best's stream-mates up again (refine_second_warp): plant the two nearest neighbours of every query in ONE stream, with
    equal and with different distances, at block edges and in a ragged last block."""
    rng = np.random.default_rng(42)
    nt, nq = 1000, 600
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    for i in range(nq):
        blk = int(rng.integers(0, (nt + 63) // 64)) * 64
        r = int(rng.integers(0, 4))
        slots = [x for x in range(blk + r, min(blk + 64, nt), 4)]
        if len(slots) < 2:
            continue
        a, b = rng.choice(len(slots), 2, replace=False)
        q[i] = t[slots[a]]                                   # distance 0 to slots[a] (t rows may be shared between queries)
        if i % 3 == 0:
            t[slots[b]] = t[slots[a]]                        # ... and 0 to its stream-mate: tie, lowest index first
        elif i % 3 == 1:
            t[slots[b]] = t[slots[a]]; t[slots[b], 0] ^= 1   # ... and 1 to its stream-mate
    ig, dg = ctx.knn2_hamming(q, t)
    io, do = orc.knn2_hamming(q, t)
    assert np.array_equal(ig, io) and np.array_equal(dg, do)
    for md in (-1.0, 5.0):
        mg = ctx.match_hamming(q, t, max_dist=md, cross_check=cross)
        mo = orc.match_hamming(q, t, max_dist=md, cross_check=cross)
        assert np.array_equal(mg, as_mvs(mo))


def test_more_than_32768_train_descriptors_stay_on_the_tensor_cores(ctx, ctx_popc):
    """The tensor-core epilogue key is `hamming << 22 | trainIdx` (22 index bits, the library-wide limit): train sets beyond
    32768 rows -- round 1's limit -- run on knn2_hamming_tc_kernel too, with the same bytes as the integer-pipe kernel."""
    rng = np.random.default_rng(9)
    q = rng.integers(0, 256, (200, 32), dtype=np.uint8); t = rng.integers(0, 256, (33000, 32), dtype=np.uint8)
    t[32900] = q[5]
    ig, dg = ctx.knn2_hamming(q, t)
    io, do = orc.knn2_hamming(q, t)
    assert np.array_equal(ig, io) and np.array_equal(dg, do) and ig[5, 0] == 32900
    ip, dp = ctx_popc.knn2_hamming(q, t)
    assert np.array_equal(ig, ip) and np.array_equal(dg, dp)
    ctx.profile_enable(True); ctx.profile_read(reset=True)
    ctx.knn2_hamming(q, t)
    assert ctx.profile_read()["knn"][1] == 2        # expand + tensor-core kernel (the popc path is one launch)
    ctx.profile_enable(False)


def test_pair_batch_records_identical_on_both_matchers(ctx, ctx_popc, tsukuba):
    descs = [tsukuba[f"desc{i}"] for i in range(1, 6)]; kps = [tsukuba[f"kp{i}"] for i in range(1, 6)]
    pairs = [(a, b) for a in range(5) for b in range(5) if a != b]
    out = []
    for c in (ctx, ctx_popc):
        c.frames_upload(descs, kps)
        out.append(c.pair_batch(pairs, tsukuba["K"], max_dist=30.0, H=64, seed=3, cross_check=True))
        out.append(c.pair_batch(pairs, tsukuba["K"], max_dist=10.0, H=64, seed=3))
    for k in (0, 1):
        ra, rb = out[k][0], out[k + 2][0]
        for f in ra.dtype.names:
            assert np.array_equal(ra[f], rb[f], equal_nan=True), (k, f, ra[f], rb[f])
        for i in range(len(pairs)):
            m, n = ra["n_matches"][i], ra["n_points"][i]
            assert np.array_equal(out[k][1]["matches"][i][:m], out[k + 2][1]["matches"][i][:m])
            assert np.array_equal(out[k][1]["mask"][i][:m], out[k + 2][1]["mask"][i][:m])
            assert np.array_equal(out[k][1]["points"][i][:n], out[k + 2][1]["points"][i][:n])
            assert np.array_equal(out[k][1]["indexes"][i][:n], out[k + 2][1]["indexes"][i][:n])


@pytest.mark.parametrize("max_dist,cross", [(-1.0, False), (10.0, False), (30.0, False), (-1.0, True), (64.0, True)])
def test_match_hamming_vs_oracle(ctx, max_dist, cross):
    d1, _, d2, _, _ = synth.synthetic_pair(11, n=3000)
    g = ctx.match_hamming(d2, d1, 0.7, max_dist, cross)
    o = orc.match_hamming(d2, d1, 0.7, max_dist, cross)
    assert len(g) == len(o) > 0
    assert np.array_equal(g, as_mvs(o))


def test_match_hamming_golden_tsukuba(ctx, tsukuba, tsukuba_golden):
    for a in range(1, 5):
        for md in (10, 30, -1):
            m = ctx.match_hamming(tsukuba[f"desc{a + 1}"], tsukuba[f"desc{a}"], 0.7, float(md))
            tag = f"p{a}{a + 1}_md{md}_"
            assert np.array_equal(m["query"], tsukuba_golden[tag + "q"])
            assert np.array_equal(m["train"], tsukuba_golden[tag + "t"])
            assert np.array_equal(m["distance"], tsukuba_golden[tag + "d"])


def test_match_edge_cases(ctx):
    rng = np.random.default_rng(0)
    q = rng.integers(0, 256, (5, 32), dtype=np.uint8)
    with pytest.raises(mvs.MvsError) as e:          # knnMatch(k=2) with a single train row: reference UB -> BAD_ARG
        ctx.match_hamming(q, q[:1])
    assert e.value.status == mvs.E_BAD_ARG
    with pytest.raises(mvs.MvsError):
        ctx.match_hamming(np.zeros((4, 16), np.uint8), np.zeros((4, 16), np.uint8))   # only 256-bit descriptors
    m = ctx.match_hamming(q, np.repeat(q[:1], 4, axis=0))    # all train identical: ratio test rejects everything
    assert len(m) == 0
    m = ctx.match_hamming(q, q)                               # self match: d1 = 0 < 0.7 * d2
    assert np.array_equal(m["query"], m["train"]) and len(m) == 5 and np.all(m["distance"] == 0)


def test_match_full_size_properties(ctx):
    """BASELINE config 3 size (8192 x 8192 x 256 bit): bit-exact vs oracle on a query subset, plus
    idempotence and agreement of first neighbours with the planted permutation."""
    d1, _, d2, _, tr = synth.synthetic_pair(0, n=8192)
    ig, dg = ctx.knn2_hamming(d2, d1)
    sub = np.random.default_rng(1).choice(8192, 256, replace=False)
    io, do = orc.knn2_hamming(d2[sub], d1)
    assert np.array_equal(ig[sub], io) and np.array_equal(dg[sub], do)
    ig2, dg2 = ctx.knn2_hamming(d2, d1)
    assert np.array_equal(ig, ig2) and np.array_equal(dg, dg2)
    planted = tr["inlier"][tr["perm"]]
    assert (ig[planted, 0] == tr["perm"][planted]).mean() > 0.999
    assert np.all(dg[:, 0] <= dg[:, 1])


# ------------------------------------------------------------------------------------------ 8-point / RANSAC
def test_find_fundamental_matrix_vs_oracle(ctx, solver):
    K, x1, x2 = general_scene(400, 3, noise=0.3)
    p1 = orc.normalize_points(K, x1); p2 = orc.normalize_points(K, x2)
    tab = orc.sample_table(9, 1, 400, 200)
    Fg = ctx.find_fundamental_matrix(p1[tab], p2[tab])
    exact = 0
    for h in range(200):
        Fo = orc.find_fundamental_matrix(p1[tab[h]], p2[tab[h]])
        assert same_up_to_scale(Fg[h], Fo, 1e-9)
        exact += np.array_equal(Fg[h], Fo)
    print(f"bit-identical F ({solver}): {exact}/200")
    assert exact == 200      # same IEEE operations in the same order on both sides: every hypothesis, every bit


@pytest.mark.parametrize("mode", [mvs.SCORE_ALGEBRAIC, mvs.SCORE_SAMPSON])
@pytest.mark.parametrize("n,H,noise,outl,thr", [(300, 64, 0.0, 0.3, 1e-7), (1000, 513, 0.3, 0.3, 1e-3),
                                                (1500, 300, 0.5, 0.5, 1e-6), (8, 1, 0.0, 0.0, 1e-7)])
def test_ransac_identical_inlier_sets(ctx, solver, mode, n, H, noise, outl, thr):
    K, x1, x2 = general_scene(n, n + H, noise, outl)
    p1 = orc.normalize_points(K, x1); p2 = orc.normalize_points(K, x2)
    tab = orc.sample_table(4, 2, n, H)
    g = ctx.ransac_fundamental(p1, p2, samples=tab, max_error_sq=thr, mode=mode, want_all=True)
    o = orc.ransac_fundamental(p1, p2, tab, thr, mode, want_all=True)
    assert g["status"] == o["status"]
    assert np.array_equal(g["all_counts"], o["all_counts"])      # every hypothesis, not only the winner
    if noise > 0:                                                # noise-free: residual ties are pure round-off
        assert g["best_h"] == o["best_h"]
        assert np.array_equal(g["mask"], o["mask"])
        assert same_up_to_scale(g["F"], o["F"], 1e-9)
        assert np.isclose(g["residual"], o["residual"], rtol=1e-9)
    assert g["count"] == o["count"]


def test_ransac_general_z_points(ctx, solver):
    """FundamentalMatrixEstimatorRANSAC::compute takes arbitrary homogeneous 3-vectors."""
    K, x1, x2 = general_scene(200, 77, 0.2, 0.2)
    p1 = orc.normalize_points(K, x1); p2 = orc.normalize_points(K, x2)
    s = np.random.default_rng(0).uniform(0.5, 2.0, (200, 1))
    p1 = p1 * s; p2 = p2 * s[::-1]
    tab = orc.sample_table(1, 1, 200, 50)
    g = ctx.ransac_fundamental(p1, p2, samples=tab, max_error_sq=1e-3, want_all=True)
    o = orc.ransac_fundamental(p1, p2, tab, 1e-3, 0, want_all=True)
    assert np.array_equal(g["all_counts"], o["all_counts"]) and g["best_h"] == o["best_h"]
    assert np.array_equal(g["mask"], o["mask"])


def test_ransac_seeded_table_on_device_equals_host_table(ctx, solver):
    K, x1, x2 = general_scene(500, 5, 0.3, 0.3)
    p1 = orc.normalize_points(K, x1); p2 = orc.normalize_points(K, x2)
    a = ctx.ransac_fundamental(p1, p2, samples=None, H=128, seed=99, max_error_sq=1e-3, want_all=True)
    b = ctx.ransac_fundamental(p1, p2, samples=mvs.sample_table(99, 0, 500, 128), max_error_sq=1e-3, want_all=True)
    assert np.array_equal(a["all_counts"], b["all_counts"]) and np.array_equal(a["F"], b["F"])


# ------------------------------------------------------------------------------------------ sfm_solve / triangulate
def check_solution(g, o, pose_tol=1e-9, strict=True, exact=False):
    assert g["status"] == o["status"]
    if o["status"] != orc.OK:
        return
    if exact:   # REFERENCE solver: the device executes the oracle's IEEE operations one for one
        for k in ("F", "E", "mask", "R1to2", "t1to2", "R2in1", "t2in1", "points", "indexes"):
            assert np.array_equal(g[k], o[k]), k
        assert g["n_inliers"] == o["n_inliers"] and g["best_hypothesis"] == o["best_hypothesis"]
        assert g["candidate"] == o["candidate"] and np.isclose(g["residual"], o["residual"], rtol=1e-12)
        return
    assert g["n_inliers"] == o["n_inliers"] and g["n_points"] == o["n_points"]
    if strict:
        assert g["best_hypothesis"] == o["best_hypothesis"] and g["candidate"] == o["candidate"]
        assert np.array_equal(g["mask"], o["mask"]) and np.array_equal(g["indexes"], o["indexes"])
        assert same_up_to_scale(g["E"], o["E"], 1e-9)
        assert np.allclose(g["points"], o["points"], rtol=1e-7, atol=1e-9)
    assert np.allclose(g["R2in1"], o["R2in1"], atol=pose_tol) and np.allclose(g["t2in1"], o["t2in1"], atol=pose_tol)


@pytest.mark.parametrize("mode", [mvs.SCORE_ALGEBRAIC, mvs.SCORE_SAMPSON])
def test_sfm_solve_vs_oracle_and_golden(ctx, solver, synthetic_golden, mode):
    s = synthetic_golden
    tagm = "alg_" if mode == mvs.SCORE_ALGEBRAIC else "smp_"
    for c in range(4):
        xy1, xy2, K, tab = s[f"s{c}_xy1"], s[f"s{c}_xy2"], s[f"s{c}_K"], s[f"s{c}_tab"]
        g = ctx.sfm_solve(xy1, xy2, K, samples=tab, mode=mode)
        o = orc.sfm_solve(xy1, xy2, K, samples=tab, mode=mode)
        check_solution(g, o, strict=(c >= 2), exact=(solver == "reference"))
        t = f"s{c}_{tagm}"
        assert (g["status"] == mvs.OK) == bool(s[t + "ok"])
        if g["status"] == mvs.OK and solver == "reference":
            for k in ("F", "E", "mask", "R2in1", "t2in1", "points", "indexes"):
                assert np.array_equal(g[k], s[t + k]), (t, k)
        if g["status"] == mvs.OK:      # Oracle-A (cv2.SVDecomp) golden: north-star tolerances
            assert g["n_inliers"] == int(s[t + "n_inliers"])
            assert np.allclose(g["R2in1"], s[t + "R2in1"], atol=1e-6)
            if c >= 2:
                assert same_up_to_scale(g["E"], s[t + "E"], 1e-5)
                assert np.allclose(g["points"], s[t + "points"], rtol=1e-4, atol=1e-6)


def test_sfm_solve_reference_single_sample_lshape(ctx, solver, synthetic_golden):
    """H=1 == the reference's behaviour (sample {0..7}); L-shape rig of test/test-sfm.cpp."""
    s = synthetic_golden
    g = ctx.sfm_solve(s["lshape_xy1"], s["lshape_xy2"], np.eye(3))
    assert g["status"] == mvs.OK and g["n_points"] == 8
    assert np.allclose(g["t2in1"], [1, 0, 0], atol=1e-9) and np.allclose(g["R2in1"], np.eye(3), atol=1e-9)
    assert np.allclose(g["points"], s["lshape_P"], atol=1e-9)
    check_solution(g, orc.sfm_solve(s["lshape_xy1"], s["lshape_xy2"], np.eye(3)), strict=False, exact=(solver == "reference"))


def test_sfm_solve_tsukuba_golden(ctx, solver, tsukuba, tsukuba_golden):
    K = tsukuba["K"]; gl = tsukuba_golden
    for a in range(1, 5):
        for md, H in ((10, 1), (30, 1), (-1, 1)):
            tag = f"p{a}{a + 1}_md{md}_"
            xy1 = tsukuba[f"kp{a}"][gl[tag + "t"]]; xy2 = tsukuba[f"kp{a + 1}"][gl[tag + "q"]]
            g = ctx.sfm_solve(xy1, xy2, K)
            o = orc.sfm_solve(xy1, xy2, K)
            check_solution(g, o, pose_tol=1e-7, exact=(solver == "reference"))
            if solver == "reference":
                # the reference's own configuration (H = 1, sample {0..7}) against the numpy + real cv2.SVDecomp
                # goldens: identical inlier sets AND identical bits in F, E, pose and points, pair 4-5 included
                t1 = tag + "h1_"
                assert g["n_inliers"] == int(gl[t1 + "n_inliers"])
                for k in ("F", "E", "mask", "R2in1", "t2in1", "points", "indexes"):
                    assert np.array_equal(g[k], gl[t1 + k]), (t1, k)
            # reference expectation (test/test-image-pair.cpp:40-45): pose ~ (I,(1,0,0)) to 1e-3
            assert np.allclose(g["t2in1"], [1, 0, 0], atol=1e-3) and np.allclose(g["R2in1"], np.eye(3), atol=1e-3)
        tag = f"p{a}{a + 1}_md30_"
        xy1 = tsukuba[f"kp{a}"][gl[tag + "t"]]; xy2 = tsukuba[f"kp{a + 1}"][gl[tag + "q"]]
        g = ctx.sfm_solve(xy1, xy2, K, samples=gl[tag + "tab256"])
        if solver == "reference":
            assert g["n_inliers"] == int(gl[tag + "h256_n_inliers"]) and g["best_hypothesis"] == int(gl[tag + "h256_best_h"])
            for k in ("F", "E", "mask", "R2in1", "t2in1", "points"):
                assert np.array_equal(g[k], gl[tag + "h256_" + k]), (tag, k)
        else:
            assert abs(g["n_inliers"] - int(gl[tag + "h256_n_inliers"])) <= 2
            assert np.allclose(g["t2in1"], gl[tag + "h256_t2in1"], atol=1e-5)


def test_sfm_triangulate_cube_known_answer(ctx, solver, synthetic_golden):
    """test/test-sfm.cpp:92-155"""
    s = synthetic_golden
    pts, idx = ctx.sfm_triangulate(s["cube_xy1"], s["cube_xy2"], np.eye(3), np.eye(3), np.zeros(3), np.eye(3),
                                   np.array([1.0, 0, 0]))
    assert np.array_equal(idx, np.arange(8)) and np.allclose(pts, s["cube_P"], atol=1e-3)
    po, io = orc.sfm_triangulate(s["cube_xy1"], s["cube_xy2"], np.eye(3), np.eye(3), np.zeros(3), np.eye(3),
                                 np.array([1.0, 0, 0]))
    assert np.allclose(pts, po, atol=1e-12)
    if solver == "reference":
        assert np.array_equal(pts, po) and np.array_equal(pts, s["cube_tri_pts"])


def test_sfm_triangulate_general_and_behind_camera(ctx, solver):
    K, x1, x2 = general_scene(700, 21, 0.2, 0.3)
    R2 = synth._rodrigues(np.array([0.02, -0.05, 0.01])); t2 = np.array([0.4, 0.1, -0.2])
    pg, ig = ctx.sfm_triangulate(x1, x2, K, np.eye(3), np.zeros(3), R2, t2)
    po, io = orc.sfm_triangulate(x1, x2, K, np.eye(3), np.zeros(3), R2, t2)
    assert 0 < len(io) < 700 and np.array_equal(ig, io)
    assert np.allclose(pg, po, rtol=1e-9, atol=1e-10)
    if solver == "reference":
        assert np.array_equal(pg, po)


def test_failure_codes(ctx, solver):
    xy = np.random.default_rng(0).uniform(0, 100, (5, 2))
    assert ctx.sfm_solve(xy, xy, np.eye(3))["status"] == mvs.E_TOO_FEW_POINTS
    r = np.random.default_rng(1)
    xa = r.uniform(0, 1000, (60, 2)); xb = r.uniform(0, 1000, (60, 2))
    g = ctx.sfm_solve(xa, xb, synth.K_S8K, H=16, seed=3)
    o = orc.sfm_solve(xa, xb, synth.K_S8K, H=16, seed=3)
    assert g["status"] == o["status"] and g["status"] in (mvs.E_TOO_FEW_INLIERS, mvs.E_NO_MODEL)
    assert g["points"].shape[0] == 0


# ------------------------------------------------------------------------------------------ batched pairs
def test_pair_batch_tsukuba_all_pairs(ctx, solver, tsukuba):
    """ImagePair ctor + reconstruct over every ordered pair of the 5 bundled frames, VO default max_dist=10."""
    descs = [tsukuba[f"desc{i}"] for i in range(1, 6)]; kps = [tsukuba[f"kp{i}"] for i in range(1, 6)]
    K = tsukuba["K"]
    pairs = [(a, b) for a in range(5) for b in range(5) if a != b]
    ctx.frames_upload(descs, kps)
    for md, H in ((10.0, 1), (30.0, 64)):
        res, det = ctx.pair_batch(pairs, K, max_dist=md, H=H, seed=7)
        for i, (a, b) in enumerate(pairs):
            o = orc.image_pair(descs[a], kps[a], descs[b], kps[b], K, max_dist=md, H=H, seed=7, pair_id=i)
            r = res[i]; m = r["n_matches"]
            assert m == o["n_matches"] and np.array_equal(det["matches"][i][:m], as_mvs(o["matches"]))
            assert r["status"] == o["status"]
            if o["status"] != orc.OK:
                continue
            assert r["n_inliers"] == o["n_inliers"] and r["best_hypothesis"] == o["best_hypothesis"]
            assert np.array_equal(det["mask"][i][:m], o["mask"])
            n = r["n_points"]
            assert n == o["n_points"] and np.array_equal(det["indexes"][i][:n], o["indexes"])
            assert np.allclose(det["points"][i][:n], o["points"], rtol=1e-6, atol=1e-8)
            assert np.allclose(r["R2in1"], o["R2in1"], atol=1e-7) and np.allclose(r["t2in1"], o["t2in1"], atol=1e-7)
            if solver == "reference":
                assert np.array_equal(det["points"][i][:n], o["points"])
                for k in ("F", "E", "R1to2", "t1to2", "R2in1", "t2in1"):
                    assert np.array_equal(r[k], o[k]), (a, b, k)
            ssd = int((o["matches"]["distance"][o["indexes"].astype(int)].astype(np.int64) ** 2).sum())
            assert int(r["match_inlier_ssd"]) == ssd


@pytest.mark.parametrize("mode,thr", [(mvs.SCORE_SAMPSON, 0.0), (mvs.SCORE_ALGEBRAIC, 1e-3)])
def test_pair_batch_synthetic_s8k_shape(ctx, solver, mode, thr):
    """Config-3-shaped pairs (general motion, 0.5 px noise, outliers), ragged keypoint counts in one batch,
    sharding-invariant sampling through pair_id_base."""
    sizes = [2048, 1500, 2048, 777]
    frames_d, frames_k, pairs = [], [], []
    for p, n in enumerate(sizes):
        d1, k1, d2, k2, _ = synth.synthetic_pair(100 + p, n=n)
        frames_d += [d1, d2]; frames_k += [k1, k2]; pairs.append((2 * p, 2 * p + 1))
    ctx.frames_upload(frames_d, frames_k)
    res, det = ctx.pair_batch(pairs, synth.K_S8K, H=256, seed=5, mode=mode, max_error_sq=thr)
    for i, (a, b) in enumerate(pairs):
        o = orc.image_pair(frames_d[a], frames_k[a], frames_d[b], frames_k[b], synth.K_S8K, H=256, seed=5, pair_id=i,
                           mode=mode, max_error_sq=thr)
        r = res[i]; m = r["n_matches"]
        assert r["status"] == o["status"] == mvs.OK
        assert m == o["n_matches"] and np.array_equal(det["matches"][i][:m], as_mvs(o["matches"]))
        assert r["n_inliers"] == o["n_inliers"] and r["best_hypothesis"] == o["best_hypothesis"]
        assert np.array_equal(det["mask"][i][:m], o["mask"])
        assert r["n_points"] == o["n_points"]
        assert same_up_to_scale(r["E"], o["E"], 1e-9)
        assert np.allclose(r["R2in1"], o["R2in1"], atol=1e-9) and np.allclose(r["t2in1"], o["t2in1"], atol=1e-9)
        assert np.allclose(det["points"][i][:r["n_points"]], o["points"], rtol=1e-7, atol=1e-9)
        if solver == "reference":
            assert np.array_equal(det["points"][i][:r["n_points"]], o["points"]) and np.array_equal(r["E"], o["E"])
            assert np.array_equal(r["R2in1"], o["R2in1"]) and np.array_equal(r["t2in1"], o["t2in1"])
    # the second half of the batch alone, with pair_id_base=2, must reproduce entries 2,3
    res2, _ = ctx.pair_batch(pairs[2:], synth.K_S8K, H=256, seed=5, mode=mode, max_error_sq=thr, pair_id_base=2)
    for k in ("n_inliers", "best_hypothesis", "n_points", "R2in1", "t2in1", "F"):
        assert np.array_equal(res2[k], res[2:][k])


def test_pair_batch_full_size_8k_h4096(ctx, solver):
    """BASELINE config 3 at full size (8192 kpts, H=4096): oracle check of the whole pipeline on one pair."""
    d1, k1, d2, k2, tr = synth.synthetic_pair(1, n=8192)
    ctx.frames_upload([d1, d2], [k1, k2])
    res, det = ctx.pair_batch([(0, 1)], synth.K_S8K, H=4096, seed=1, mode=mvs.SCORE_SAMPSON)
    r = res[0]
    o = orc.image_pair(d1, k1, d2, k2, synth.K_S8K, H=4096, seed=1, mode=orc.SCORE_SAMPSON)
    assert r["status"] == o["status"] == 0 and r["n_matches"] == o["n_matches"]
    assert r["n_inliers"] == o["n_inliers"] and r["best_hypothesis"] == o["best_hypothesis"]
    assert np.array_equal(det["mask"][0][:r["n_matches"]], o["mask"]) and r["n_points"] == o["n_points"]
    assert np.allclose(r["R2in1"], o["R2in1"], atol=1e-9) and np.allclose(r["t2in1"], o["t2in1"], atol=1e-9)
    # sanity against the planted motion (pose2in1 = (R,t)^-1, |t| = 1)
    tt = -tr["R"].T @ tr["t"]; tt /= np.linalg.norm(tt)
    assert np.allclose(r["R2in1"], tr["R"].T, atol=2e-2) and np.allclose(r["t2in1"], tt, atol=0.15)


def test_pair_batch_bad_pairs_do_not_abort_batch(ctx):
    d1, k1, d2, k2, _ = synth.synthetic_pair(5, n=512, noise_px=1e-4)
    rng = np.random.default_rng(0)
    junk = rng.integers(0, 256, (300, 32), dtype=np.uint8); junk_k = rng.uniform(0, 700, (300, 2)).astype(np.float32)
    ctx.frames_upload([d1, d2, junk], [k1, k2, junk_k])
    res, _ = ctx.pair_batch([(0, 1), (0, 2), (1, 0)], synth.K_S8K, H=32, seed=1, details=False)
    assert res[0]["status"] == mvs.OK and res[2]["status"] == mvs.OK
    assert res[1]["status"] in (mvs.E_TOO_FEW_POINTS, mvs.E_TOO_FEW_INLIERS, mvs.E_NO_MODEL)
    with pytest.raises(mvs.MvsError):
        ctx.pair_batch([(0, 0)], synth.K_S8K)
    with pytest.raises(mvs.MvsError):
        ctx.pair_batch([(0, 9)], synth.K_S8K)


def test_profile_and_launch_counters(ctx, tsukuba):
    descs = [tsukuba[f"desc{i}"] for i in range(1, 3)]; kps = [tsukuba[f"kp{i}"] for i in range(1, 3)]
    ctx.frames_upload(descs, kps)
    ctx.profile_enable(True); ctx.profile_read(reset=True)
    n0 = ctx.kernel_launches()
    ctx.pair_batch([(0, 1)] * 8, tsukuba["K"], max_dist=10.0, H=64, details=False)
    prof = ctx.profile_read()
    # first batch after an upload: the 7 stage kernels + the one-off expansion of the new frames for the tensor-core matcher
    assert ctx.kernel_launches() - n0 == 8 and prof["knn"][1] == 2
    assert all(prof[s][1] == 1 and prof[s][0] > 0 for s in mvs.STAGES[1:7])
    n0 = ctx.kernel_launches()
    ctx.profile_read(reset=True)
    ctx.pair_batch([(0, 1)] * 8, tsukuba["K"], max_dist=10.0, H=64, details=False)
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    assert ctx.kernel_launches() - n0 == 7
    assert all(prof[s][1] == 1 and prof[s][0] > 0 for s in mvs.STAGES[:7])


def test_cpp_adapters_reference_signatures():
    """The C++ headers under include/mvslam/ (VisualFeature::match_visual_features, sfm_solve, sfm_triangulate,
    FundamentalMatrixEstimatorRANSAC, ImagePair incl. refine, VisualFeature::extract, pnp_solve, sfm_refine, pnp_refine) driven like the reference's own tests (tests/cpp/test_adapters.cpp)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cpp", "test_adapters")
    assert os.path.exists(exe), "run __graft_entry__.build() first"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("PASSED") == 6 and "FAILED" not in r.stdout


def test_two_gpu_sharded_equals_single_gpu(tmp_path, tsukuba):
    """world_size-2 NCCL run of shard.solve_pairs_sharded vs the same batch on one GPU (skipped on 1-GPU boxes)."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "w.py"
    script.write_text('''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
import mvslam_b200 as mvs
from mvslam_b200 import shard
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
f = np.load(os.path.join(sys.argv[1], "tests", "golden", "tsukuba_orb2000.npz"))
descs = [f[f"desc{i}"] for i in range(1, 6)]; kps = [f[f"kp{i}"] for i in range(1, 6)]
pairs = np.array([(a, b) for a in range(5) for b in range(5) if a != b], np.int32)
ctx = mvs.Context(local)
full = shard.solve_pairs_sharded(ctx, descs, kps, pairs, f["K"], dist=dist, device=torch.device("cuda", local),
                                 max_dist=30.0, H=128, seed=11)
if dist.get_rank() == 0:
    np.save(sys.argv[2], full)
dist.barrier(); dist.destroy_process_group()
''')
    out = tmp_path / "multi.npy"
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                    "127.0.0.1", "--master-port", "29711", str(script), root, str(out)], check=True, timeout=600)
    multi = np.load(out)
    descs = [tsukuba[f"desc{i}"] for i in range(1, 6)]; kps = [tsukuba[f"kp{i}"] for i in range(1, 6)]
    pairs = np.array([(a, b) for a in range(5) for b in range(5) if a != b], np.int32)
    with mvs.Context(0) as c:
        single = shard.solve_pairs_sharded(c, descs, kps, pairs, tsukuba["K"], max_dist=30.0, H=128, seed=11)
    assert multi.tobytes() == single.tobytes()


def test_two_gpu_c_abi_sharded_with_clouds_equals_single_gpu(tmp_path, tsukuba):
    """mvs_pair_batch_sharded (C ABI, NCCL bound at run time; no torch.distributed anywhere): two processes, one GPU each,
    exchange the ncclUniqueId through a file; the gathered records AND the variable-length clouds (points, indexes, matches
    placed by exclusive-scan offsets) must equal the single-GPU bytes.  Skipped on 1-GPU boxes."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "w.py"
    script.write_text('''
import os, sys, time
sys.path.insert(0, sys.argv[1])
import numpy as np
import mvslam_b200 as mvs
rank, world, idf, out = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
f = np.load(os.path.join(sys.argv[1], "tests", "golden", "tsukuba_orb2000.npz"))
descs = [f[f"desc{i}"] for i in range(1, 6)]; kps = [f[f"kp{i}"] for i in range(1, 6)]
pairs = np.array([(a, b) for a in range(5) for b in range(5) if a != b], np.int32)
ctx = mvs.Context(rank)
if rank == 0:
    uid = mvs.Comm.unique_id()
    open(idf + ".tmp", "wb").write(uid); os.rename(idf + ".tmp", idf)
else:
    while not os.path.exists(idf): time.sleep(0.05)
    uid = open(idf, "rb").read()
comm = mvs.Comm(ctx, uid, rank, world)
ctx.frames_upload(descs, kps)
res, det = comm.pair_batch_sharded(pairs, f["K"], max_dist=30.0, H=128, seed=11, solver="fast")
if rank == 0:
    np.savez(out, res=res, **det)
comm.close(); ctx.close()
''')
    out = str(tmp_path / "multi.npz"); idf = str(tmp_path / "nccl.id")
    procs = [subprocess.Popen([sys.executable, str(script), root, str(r), "2", idf, out]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=600) == 0
    multi = np.load(out)
    descs = [tsukuba[f"desc{i}"] for i in range(1, 6)]; kps = [tsukuba[f"kp{i}"] for i in range(1, 6)]
    pairs = np.array([(a, b) for a in range(5) for b in range(5) if a != b], np.int32)
    with mvs.Context(0) as c:
        c.frames_upload(descs, kps)
        single, det = c.pair_batch(pairs, tsukuba["K"], max_dist=30.0, H=128, seed=11, solver="fast")
    assert multi["res"].tobytes() == single.tobytes()
    po, mo = multi["point_offsets"], multi["match_offsets"]
    for i in range(len(pairs)):
        n = int(single[i]["n_points"]) if single[i]["status"] == 0 else 0
        m = int(single[i]["n_matches"])
        assert po[i + 1] - po[i] == n and mo[i + 1] - mo[i] == m
        assert np.array_equal(multi["points"][po[i]:po[i + 1]], det["points"][i][:n])
        assert np.array_equal(multi["indexes"][po[i]:po[i + 1]], det["indexes"][i][:n])
        assert np.array_equal(multi["matches"][mo[i]:mo[i + 1]], det["matches"][i][:m])


def test_pair_batch_large_batch_is_chunked_consistently(ctx, tsukuba):
    """More pairs than one launch carries (grid.z / workspace chunks of 8192): same records as small batches,
    sampling tied to the global pair index."""
    descs = [tsukuba[f"desc{i}"][:300] for i in range(1, 6)]; kps = [tsukuba[f"kp{i}"][:300] for i in range(1, 6)]
    base = [(a, b) for a in range(5) for b in range(5) if a != b]
    pairs = np.array([base[i % len(base)] for i in range(9000)], np.int32)
    ctx.frames_upload(descs, kps)
    big, _ = ctx.pair_batch(pairs, tsukuba["K"], max_dist=40.0, H=8, seed=2, details=False)
    lo, hi = 8150, 8250                      # straddles the chunk boundary
    small, _ = ctx.pair_batch(pairs[lo:hi], tsukuba["K"], max_dist=40.0, H=8, seed=2, details=False, pair_id_base=lo)
    assert big[lo:hi].tobytes() == small.tobytes()
    assert (big["status"] == mvs.OK).sum() > 0


@pytest.mark.parametrize("max_dist,cross", [(0.0, False), (10.0, False), (30.0, False), (64.0, True), (10.0, True), (200.0, False)])
def test_bounded_search_returns_identical_matches(ctx, ctx_popc, tsukuba, max_dist, cross):
    """mvs_match_params.bounded: early-abandoned train descriptors can never change the filtered result (the early
    abandon lives in knn2_hamming_kernel; the tensor-core matcher evaluates every pair and ignores the flag)."""
    sets = [(tsukuba["desc2"], tsukuba["desc1"])]
    d1, _, d2, _, _ = synth.synthetic_pair(21, n=2500)
    sets.append((d2, d1))
    rng = np.random.default_rng(3)                      # tie-heavy, small distances: many candidates inside the bound
    a = np.zeros((700, 32), np.uint8); b = np.zeros((650, 32), np.uint8)
    a[:, :2] = rng.integers(0, 256, (700, 2)); b[:, :2] = rng.integers(0, 256, (650, 2))
    sets.append((a, b))
    for q, t in sets:
        want = as_mvs(orc.match_hamming(q, t, 0.7, max_dist, cross))
        for c in (ctx, ctx_popc):
            full = c.match_hamming(q, t, 0.7, max_dist, cross, bounded=False)
            fast = c.match_hamming(q, t, 0.7, max_dist, cross, bounded=True)
            assert np.array_equal(full, fast) and np.array_equal(full, want)


def test_bounded_pair_batch_identical_records(ctx_popc, tsukuba):
    ctx = ctx_popc
    descs = [tsukuba[f"desc{i}"] for i in range(1, 6)]; kps = [tsukuba[f"kp{i}"] for i in range(1, 6)]
    pairs = [(a, b) for a in range(5) for b in range(5) if a != b]
    ctx.frames_upload(descs, kps)
    for md in (10.0, 30.0):
        r0, d0 = ctx.pair_batch(pairs, tsukuba["K"], max_dist=md, H=32, seed=1)
        r1, d1 = ctx.pair_batch(pairs, tsukuba["K"], max_dist=md, H=32, seed=1, bounded=True)
        for f in r0.dtype.names:
            assert np.array_equal(r0[f], r1[f], equal_nan=True), f
        for i in range(len(pairs)):        # entries beyond the counts are unspecified
            m, n = r0["n_matches"][i], r0["n_points"][i]
            assert np.array_equal(d0["matches"][i][:m], d1["matches"][i][:m]) and np.array_equal(d0["mask"][i][:m], d1["mask"][i][:m])
            assert np.array_equal(d0["points"][i][:n], d1["points"][i][:n]) and np.array_equal(d0["indexes"][i][:n], d1["indexes"][i][:n])


def test_large_synchronous_batch_trims_detail_copies(ctx, tsukuba):
    """A synchronous mvs_pair_batch with more than 1 MB of details copies only as many entries per pair as the fullest
    pair holds (api.cu pair_batch_chunk); everything inside the counts must equal the small-batch result."""
    descs = [tsukuba[f"desc{i}"] for i in range(1, 6)]; kps = [tsukuba[f"kp{i}"] for i in range(1, 6)]
    ctx.frames_upload(descs, kps)
    base = [(a, b) for a in range(5) for b in range(5) if a != b]
    pairs = [base[i % len(base)] for i in range(60)]              # 60 pairs x 1759 slots x 45 B = 4.7 MB of details
    big, dbig = ctx.pair_batch(pairs, tsukuba["K"], max_dist=30.0, H=16, seed=4)
    for i in range(0, 60, 7):
        one, done = ctx.pair_batch([pairs[i]], tsukuba["K"], max_dist=30.0, H=16, seed=4, pair_id_base=i)
        assert one.tobytes() == big[i:i + 1].tobytes()
        m, n = int(one["n_matches"][0]), int(one["n_points"][0])
        assert m > 20
        assert np.array_equal(dbig["matches"][i][:m], done["matches"][0][:m]) and np.array_equal(dbig["mask"][i][:m], done["mask"][0][:m])
        assert np.array_equal(dbig["points"][i][:n], done["points"][0][:n]) and np.array_equal(dbig["indexes"][i][:n], done["indexes"][0][:n])
    most = int(big["n_matches"].max())
    assert not dbig["matches"][:, most:]["distance"].any()         # slots beyond the fullest pair were not copied (host zeros)


def test_frames_upload_from_pinned_memory_equals_pageable(ctx, tsukuba):
    """Frames in pinned host memory are gathered by one kernel (api.cu gather_frames_kernel) instead of two copies per frame;
    an empty frame, a frame at an address the kernel's 16-byte loads cannot take (falls back to the copies) and more frames
    than one launch carries give the same table as the pageable upload."""
    torch = pytest.importorskip("torch")
    base_d = [tsukuba[f"desc{i}"] for i in range(1, 6)]; base_k = [tsukuba[f"kp{i}"] for i in range(1, 6)]
    descs = base_d + [base_d[0][:0]] + [base_d[i % 5][: 40 + i] for i in range(70)]
    kps = base_k + [base_k[0][:0]] + [base_k[i % 5][: 40 + i] for i in range(70)]
    pairs = [(0, 1), (1, 2), (3, 4), (6, 0), (75, 1), (2, 70), (5, 1)]
    kw = dict(max_dist=30.0, H=8, seed=3, solver="fast")
    ctx.frames_upload(descs, kps)
    ref, dref = ctx.pair_batch(pairs, tsukuba["K"], **kw)
    pd = [torch.from_numpy(np.ascontiguousarray(d)).pin_memory() for d in descs]
    pk = [torch.from_numpy(np.ascontiguousarray(k)).pin_memory() for k in kps]
    ctx.frames_upload([t.numpy() for t in pd], [t.numpy() for t in pk])
    res, det = ctx.pair_batch(pairs, tsukuba["K"], **kw)
    assert res.tobytes() == ref.tobytes()
    for i in range(len(pairs)):
        m = int(ref["n_matches"][i])
        assert np.array_equal(det["matches"][i][:m], dref["matches"][i][:m])
    # a descriptor block that starts 4 bytes into a pinned buffer: not 16-byte aligned, so this upload takes the copy path
    raw = torch.zeros(descs[0].size + 4, dtype=torch.uint8).pin_memory()
    odd = raw.numpy()[4:].reshape(descs[0].shape); odd[:] = descs[0]
    ctx.frames_upload([odd] + [t.numpy() for t in pd[1:]], [t.numpy() for t in pk])
    res2, _ = ctx.pair_batch(pairs, tsukuba["K"], **kw)
    assert res2.tobytes() == ref.tobytes()


@pytest.mark.parametrize("cross", [False, True])
def test_small_batches_split_the_train_dimension(ctx, cross):
    """Launches with fewer (pair, query tile) items than SMs also cut the train frames into splits (match_hamming_tc.cu
    tc_train_splits): ragged frames -- the trailing splits of a short base frame own no tile and must export "no neighbour" --
    and every small batch size give the oracle's matches bit for bit."""
    rng = np.random.default_rng(11)
    sizes = [1764, 130, 900, 2, 1300, 257]
    descs = [rng.integers(0, 256, (n, 32), dtype=np.uint8) for n in sizes]
    for d in descs[1:]:                         # planted near-duplicates, so that matches survive the ratio test
        k = min(len(d), 100)
        d[:k] = descs[0][:k] ^ (rng.integers(0, 256, (k, 32), dtype=np.uint8) & rng.integers(0, 256, (k, 32), dtype=np.uint8)
                                & rng.integers(0, 256, (k, 32), dtype=np.uint8) & rng.integers(0, 256, (k, 32), dtype=np.uint8))
    kps = [rng.uniform(0, 700, (n, 2)).astype(np.float32) for n in sizes]
    ctx.frames_upload(descs, kps)
    all_pairs = [(a, b) for a in range(6) for b in range(6) if a != b]
    for n_pairs in (1, 2, 5, 9, 30):
        pairs = all_pairs[:n_pairs] if n_pairs != 2 else [(1, 0), (0, 3)]
        res, det = ctx.pair_batch(pairs, synth.K_S8K, max_dist=-1.0, cross_check=cross, H=4, seed=1)
        for i, (a, b) in enumerate(pairs):
            o = orc.match_hamming(descs[b], descs[a], max_dist=-1.0, cross_check=cross) if len(descs[a]) >= 2 else np.zeros(0, mvs.MATCH_DTYPE)
            m = int(res[i]["n_matches"])
            assert m == len(o), (n_pairs, a, b, m, len(o))
            assert np.array_equal(det["matches"][i][:m], as_mvs(o)), (n_pairs, a, b)


@pytest.mark.parametrize("enqueue", [False, True])
@pytest.mark.parametrize("cap", [2048, 40])
def test_pinned_detail_outputs_are_written_by_the_device(ctx, tsukuba, enqueue, cap):
    """Detail buffers in pinned host memory are filled by the library's export kernel (exactly the entries every pair owns,
    api.cu export_details_kernel) instead of strided copies: same bytes as the pageable-buffer call inside the counts, nothing
    written beyond them, truncation at `capacity` as documented; a bad pair leaves its rows alone."""
    torch = pytest.importorskip("torch")
    descs = [tsukuba[f"desc{i}"] for i in range(1, 6)] + [tsukuba["desc1"][:1]]; kps = [tsukuba[f"kp{i}"] for i in range(1, 6)] + [tsukuba["kp1"][:1]]
    ctx.frames_upload(descs, kps)
    pairs = [(0, 1), (1, 2), (5, 2), (3, 4), (4, 0), (2, 3)]       # (5, 2): a one-keypoint base frame cannot be matched
    n = len(pairs)
    kw = dict(max_dist=30.0, H=32, seed=9, solver="fast")
    ref, dref = ctx.pair_batch(pairs, tsukuba["K"], **kw)
    assert ref["status"][2] != 0 and (ref["status"][[0, 1, 3, 4, 5]] == 0).all()
    assert cap == 2048 or ref["n_matches"].max() > cap          # the small capacity truncates
    item = mvs.RESULT_DTYPE.itemsize
    res_t = torch.zeros(n * item, dtype=torch.uint8).pin_memory()
    mat_t = torch.full((n * cap * 12,), 0xEE, dtype=torch.uint8).pin_memory(); msk_t = torch.full((n * cap,), 0xEE, dtype=torch.uint8).pin_memory()
    pts_t = torch.full((n * cap * 3,), -7.0, dtype=torch.float64).pin_memory(); idx_t = torch.full((n * cap,), -7, dtype=torch.int64).pin_memory()
    ctx.pair_batch(pairs, tsukuba["K"], enqueue_only=enqueue, out=dict(
        results=res_t.data_ptr(), matches=mat_t.data_ptr(), mask=msk_t.data_ptr(), points=pts_t.data_ptr(), indexes=idx_t.data_ptr(),
        capacity=cap), **kw)
    ctx.synchronize()
    res = np.frombuffer(res_t.numpy(), dtype=mvs.RESULT_DTYPE)
    assert res.tobytes() == ref.tobytes()
    mat = np.frombuffer(mat_t.numpy(), dtype=mvs.MATCH_DTYPE).reshape(n, cap); msk = msk_t.numpy().reshape(n, cap)
    pts = pts_t.numpy().reshape(n, cap, 3); idx = idx_t.numpy().reshape(n, cap)
    for i in range(n):
        m, k = min(int(ref["n_matches"][i]), cap), min(int(ref["n_points"][i]), cap)
        assert np.array_equal(mat[i][:m], dref["matches"][i][:m]) and np.array_equal(msk[i][:m], dref["mask"][i][:m])
        assert np.array_equal(pts[i][:k], dref["points"][i][:k]) and np.array_equal(idx[i][:k].astype(np.uint64), dref["indexes"][i][:k])
        assert (mat_t.numpy().reshape(n, cap * 12)[i][m * 12:] == 0xEE).all() and (msk[i][m:] == 0xEE).all()      # untouched
        assert (pts[i][k:] == -7.0).all() and (idx[i][k:] == -7).all()


def test_empty_and_degenerate_inputs(ctx):
    """Empty / minimal inputs: status codes instead of the reference's asserts or undefined behaviour."""
    q = np.zeros((0, 32), np.uint8); t = np.random.default_rng(0).integers(0, 256, (10, 32), dtype=np.uint8)
    with pytest.raises(mvs.MvsError) as e:
        ctx.match_hamming(q, t)                                   # visual-feature.cpp:56 assert(valid())
    assert e.value.status == mvs.E_BAD_ARG
    pts, idx = ctx.sfm_triangulate(np.zeros((0, 2)), np.zeros((0, 2)), np.eye(3), np.eye(3), np.zeros(3), np.eye(3), np.ones(3))
    assert len(pts) == 0 and len(idx) == 0
    assert ctx.sfm_solve(np.zeros((0, 2)), np.zeros((0, 2)), np.eye(3))["status"] == mvs.E_TOO_FEW_POINTS
    r = ctx.ransac_fundamental(np.ones((7, 3)), np.ones((7, 3)), H=4, max_error_sq=1e-3)
    assert r["status"] == mvs.E_TOO_FEW_POINTS                     # estimator-RANSAC.cpp:25-29
    # all correspondences identical: rank-deficient 8-point system, must not crash or hang
    same = np.tile(np.array([[10.0, 20.0]]), (20, 1))
    g = ctx.sfm_solve(same, same, synth.K_S8K, H=4, seed=1)
    assert g["status"] in (mvs.OK, mvs.E_NO_MODEL, mvs.E_TOO_FEW_INLIERS, mvs.E_NO_CHEIRALITY)
    # a frame without keypoints (a dark image in a window) only fails its own pairs: no matches, MVS_E_TOO_FEW_POINTS,
    # an all-zero mask row; the other pairs of the batch are solved as usual
    d1, k1, d2, k2, _ = synth.synthetic_pair(5, n=512, noise_px=1e-4)
    e8 = np.zeros((0, 32), np.uint8); e2 = np.zeros((0, 2), np.float32)
    ctx.frames_upload([d1, e8, d2, d1[:1]], [k1, e2, k2, k1[:1]])
    res, det = ctx.pair_batch([(0, 1), (0, 2), (1, 0), (3, 2), (2, 3)], synth.K_S8K, H=8, seed=1)
    assert [int(r["status"]) for r in res] == [mvs.E_TOO_FEW_POINTS, mvs.OK, mvs.E_TOO_FEW_POINTS, mvs.E_TOO_FEW_POINTS,
                                               mvs.E_TOO_FEW_POINTS]
    assert [int(r["n_matches"]) for r in res[:4]] == [0, int(res[1]["n_matches"]), 0, 0] and res[1]["n_matches"] > 100
    assert int(res[4]["n_matches"]) <= 1      # a single query keypoint may well find its match: still too few points
    assert all(not det["mask"][i][:int(res[i]["n_matches"])].any() for i in (0, 2, 3, 4))   # within the counts: no stale data
    assert res["n_points"][[0, 2, 3, 4]].sum() == 0 and res["n_inliers"][[0, 2, 3, 4]].sum() == 0
    alone, _ = ctx.pair_batch([(0, 2)], synth.K_S8K, H=8, seed=1, pair_id_base=1)
    assert alone.tobytes() == res[1:2].tobytes()
    # bad RANSAC parameters
    with pytest.raises(mvs.MvsError):
        ctx.sfm_solve(np.zeros((10, 2)), np.zeros((10, 2)), np.eye(3), H=0)


def test_cpp_tools_reconstruct_scene_and_vo_pairs(tmp_path, tsukuba, tsukuba_golden):
    """tools/reconstruct_scene and tools/visual_odometer_pairs (C++ over the adapters) on the bundled Tsukuba features:
    the pose the reference's tests expect (test-image-pair.cpp:40-45, test-visual-odometer.cpp:98-102: (I,(1,0,0)) to 1e-3)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / "tsu")
    subprocess.run([sys.executable, os.path.join(root, "tools", "export_features.py"), "npz",
                    os.path.join(root, "tests", "golden", "tsukuba_orb2000.npz"), out], check=True)
    r = subprocess.run([os.path.join(root, "tools", "reconstruct_scene"), out + "/1.mvsf", out + "/2.mvsf", out + "/camera.config", "30"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.splitlines()
    assert lines[0] == f"matches = {len(tsukuba_golden['p12_md30_q'])}"
    rows = [l for l in lines if "|" in l]
    Rt = np.array([[float(x) for x in l.replace("|", " ").split()] for l in rows])
    assert np.allclose(Rt[:, :3], np.eye(3), atol=1e-3) and np.allclose(Rt[:, 3], [1, 0, 0], atol=1e-3)
    npts = int([l for l in lines if l.startswith("pointsin1_scaled")][0].split("=")[1])
    o = orc.image_pair(tsukuba["desc1"], tsukuba["kp1"], tsukuba["desc2"], tsukuba["kp2"], tsukuba["K"], max_dist=30.0)
    assert npts == o["n_points"]
    r = subprocess.run([os.path.join(root, "tools", "visual_odometer_pairs"), out], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    fl = [l for l in r.stdout.splitlines() if l.startswith("frame")]
    assert len(fl) == 4
    for i, l in enumerate(fl):
        assert f"{i + 1} pair(s)" in l and "valid=1" in l
        t = [float(x) for x in l.split("t=(")[1].rstrip(")").split()]
        assert np.allclose(t, [1, 0, 0], atol=1e-3)


def test_reference_callers_verbatim_pod_and_eigen_shaped(tmp_path, tsukuba, tsukuba_golden):
    """VisualOdometer::add_frame's pair construction (3-argument ImagePair, K from CameraManager) and reconstruct-scene's call
    sequence, verbatim (tests/cpp/test_reference_callers.cpp), built against the POD stand-ins and against column-major
    Eigen-shaped / cv-shaped types: identical output, and the reference-solver numbers of the committed goldens."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / "tsu")
    subprocess.run([sys.executable, os.path.join(root, "tools", "export_features.py"), "npz",
                    os.path.join(root, "tests", "golden", "tsukuba_orb2000.npz"), out], check=True)
    outs = []
    for exe in ("test_reference_callers", "test_reference_callers_eigen"):
        r = subprocess.run([os.path.join(root, "tests", "cpp", exe), out + "/1.mvsf", out + "/2.mvsf", out + "/camera.config"],
                           capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stdout + r.stderr
        outs.append(r.stdout)
    assert outs[0] == outs[1]                                  # storage order of the matrix type is invisible
    lines = outs[0].splitlines()
    gl = tsukuba_golden
    n10 = len(gl["p12_md10_h1_indexes"])
    ssd = (int((gl["p12_md10_d"][gl["p12_md10_h1_indexes"].astype(int)].astype(np.int64) ** 2).sum()) + 0xFFFFFFFF) & 0xFFFFFFFF
    assert lines[0] == f"vo: valid 1 inliers {n10} ssd {ssd} points {n10}"
    T = [float(x) for x in lines[1].split("T_pair_to_base")[1].replace("|", " ").split()]
    assert np.array_equal(np.array(T[:9]).reshape(3, 3), gl["p12_md10_h1_R2in1"]) and np.array_equal(T[9:], gl["p12_md10_h1_t2in1"])
    first = [float(x) for x in lines[2].split("first point")[1].split("idx")[0].split()]
    assert np.array_equal(first, gl["p12_md10_h1_points"][0])
    rs = lines[3].split()
    assert int(rs[2]) == len(gl["p12_md30_q"]) and int(rs[4]) == len(gl["p12_md30_h1_indexes"])
    assert np.array_equal([float(x) for x in rs[6:9]], gl["p12_md30_h1_t2in1"])


def test_cpp_reconstruct_window_tool_sharded(tmp_path, tsukuba):
    """tools/reconstruct_window: the all-pairs job of BASELINE config 5 as a C++ program over the C ABI, one process per GPU
    (one here, two when the box has them): same records whatever the number of ranks."""
    import subprocess
    import sys
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / "tsu")
    subprocess.run([sys.executable, os.path.join(root, "tools", "export_features.py"), "npz",
                    os.path.join(root, "tests", "golden", "tsukuba_orb2000.npz"), out], check=True)
    exe = os.path.join(root, "tools", "reconstruct_window")
    outs = []
    for world in ([1, 2] if torch.cuda.device_count() >= 2 else [1]):
        idf = str(tmp_path / f"id{world}")
        procs = [subprocess.Popen([exe, out, "5", idf, str(r), str(world), "30", "1"], stdout=subprocess.PIPE, text=True) for r in range(world)]
        texts = [p.communicate(timeout=300)[0] for p in procs]
        assert all(p.returncode == 0 for p in procs), texts
        outs.append([l for l in texts[0].splitlines() if not l.startswith("NCCL version")])   # NCCL_DEBUG=VERSION prints to stdout
    head = outs[0][0]
    assert "pairs = 10" in head and "solved = 10" in head
    o = orc.image_pair(tsukuba["desc1"], tsukuba["kp1"], tsukuba["desc2"], tsukuba["kp2"], tsukuba["K"], max_dist=30.0)
    assert f"matches {o['n_matches']} points {o['n_points']}" in outs[0][1]
    for other in outs[1:]:      # identical per-pair lines (the head line carries the rank count and the wall time)
        assert other[1:] == outs[0][1:] and other[0].split(",")[1:4] == head.split(",")[1:4]


def test_cpp_vo_tool_from_images(tmp_path):
    """tools/visual_odometer_pairs fed with the Tsukuba frames themselves (PGM): VisualFeature::extract on the device,
    then the per-frame pair batches; the 2000-feature run recovers the (I, (1, 0, 0)) motion of test-visual-odometer.cpp:98-102."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / "tsu_img")
    subprocess.run([sys.executable, os.path.join(root, "tools", "export_features.py"), "pgm",
                    os.path.join(root, "tests", "golden", "tsukuba_gray.npz"), out], check=True)
    r = subprocess.run([os.path.join(root, "tools", "visual_odometer_pairs"), out, "10", "2000"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    ext = [l for l in r.stdout.splitlines() if "keypoints extracted" in l]
    assert len(ext) == 5 and "1748 keypoints" in ext[0] and "1759 keypoints" in ext[1]      # cv2.ORB_create(2000) counts
    fl = [l for l in r.stdout.splitlines() if "pair(s)" in l]
    assert len(fl) == 4
    for l in fl:
        assert "valid=1" in l
        t = [float(x) for x in l.split("t=(")[1].rstrip(")").split()]
        assert np.allclose(t, [1, 0, 0], atol=1e-3)
    # utility/reconstruct-scene.cpp on two images: extract (2000 features), match_and_filter, sfm_solve
    r = subprocess.run([os.path.join(root, "tools", "reconstruct_scene"), out + "/1.pgm", out + "/2.pgm", out + "/camera.config", "30", "2000"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    rows = [l for l in r.stdout.splitlines() if "|" in l]
    Rt = np.array([[float(x) for x in l.replace("|", " ").split()] for l in rows])
    assert np.allclose(Rt[:, :3], np.eye(3), atol=1e-3) and np.allclose(Rt[:, 3], [1, 0, 0], atol=1e-3)


def test_frames_append_equals_bulk_upload(ctx, solver, tsukuba):
    """Streaming use (FrameManager::add_frame per image): appended frames behave exactly like a bulk upload."""
    descs = [tsukuba[f"desc{i}"] for i in range(1, 6)]; kps = [tsukuba[f"kp{i}"] for i in range(1, 6)]
    pairs = [(0, 1), (1, 2), (0, 4), (3, 2)]
    ctx.frames_upload(descs, kps)
    ref, dref = ctx.pair_batch(pairs, tsukuba["K"], max_dist=30.0, H=16, seed=4)
    ctx.frames_clear()
    for i in range(5):
        assert ctx.frames_append(descs[i], kps[i]) == i
        if i == 1:
            one, _ = ctx.pair_batch([(0, 1)], tsukuba["K"], max_dist=30.0, H=16, seed=4)
            assert one.tobytes() == ref[:1].tobytes()
    got, dgot = ctx.pair_batch(pairs, tsukuba["K"], max_dist=30.0, H=16, seed=4)
    assert got.tobytes() == ref.tobytes()
    assert all(np.array_equal(dgot["matches"][i][:m], dref["matches"][i][:m]) for i, m in enumerate(ref["n_matches"]))


def test_python_mirror_of_the_reference_entry_points(ctx, tsukuba):
    """mvslam_b200/vision.py carries the reference's names (VisualFeature.match_visual_features, sfm_solve, sfm_triangulate,
    image_pairs) over the same C ABI: each returns what the Context method it forwards to returns."""
    vf1 = mvs.VisualFeature(tsukuba["kp1"], tsukuba["desc1"], 384, 288); vf2 = mvs.VisualFeature(tsukuba["kp2"], tsukuba["desc2"], 384, 288)
    K = tsukuba["K"]
    m = mvs.VisualFeature.match_visual_features(vf1, vf2, 30.0, ctx=ctx)
    assert np.array_equal(m, ctx.match_hamming(tsukuba["desc2"], tsukuba["desc1"], 0.7, 30.0, False)) and len(m) > 100
    assert np.array_equal(m, as_mvs(orc.match_hamming(tsukuba["desc2"], tsukuba["desc1"], max_dist=30.0)))
    p1 = vf1.get_image_points()[m["train"]]; p2 = vf2.get_image_points()[m["query"]]
    ok, (R, t), pts, idx = mvs.sfm_solve(p1, p2, K, ctx=ctx)
    r = ctx.sfm_solve(p1, p2, K)
    assert ok and np.array_equal(R, r["R2in1"]) and np.array_equal(t, r["t2in1"]) and np.array_equal(pts, r["points"])
    assert np.allclose(t / np.linalg.norm(t), [1, 0, 0], atol=5e-2)          # test-image-pair.cpp:42-44: pure x translation
    tri_pts, tri_idx = mvs.sfm_triangulate(p1, p2, K, (np.eye(3), np.zeros(3)), (R, t), ctx=ctx)
    assert tri_pts.shape[0] == tri_idx.shape[0] >= 0.8 * len(pts)
    res, det = mvs.image_pairs([vf1, vf2], [(0, 1)], K, max_match_inlier_distance=30.0, ctx=ctx)
    assert res[0]["status"] == mvs.OK and res[0]["n_matches"] == len(m) and np.array_equal(det["matches"][0][:len(m)], m)
