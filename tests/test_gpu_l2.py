"""GPU parity of the float-descriptor (NORM_L2) matcher — tensor-core contraction + exact FP32 re-rank —
against the CPU oracle (double-precision brute force rounded to float, pinned to cv2.batchDistance in
tests/test_oracle_pinning.py).  Tolerance: indices equal except where the two candidate
distances differ by less than 1e-6 relative; distances within 1e-5 relative."""
import numpy as np
import pytest

import mvslam_b200 as mvs
from mvslam_b200 import synth
from oracle import cbind as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = mvs.Context(0)
    yield c
    c.close()


def exact_dist(q, t, qi, ti):
    return np.sqrt(((q[qi].astype(np.float64) - t[ti].astype(np.float64)) ** 2).sum(-1))


def check_knn(q, t, ig, dg, io, do):
    assert ig.shape == io.shape
    assert np.allclose(dg, do, rtol=1e-5, atol=1e-7)
    bad = np.argwhere(ig != io)
    for r, c in bad:      # an index may differ only between (near-)equidistant candidates
        a = exact_dist(q, t, r, ig[r, c]); b = exact_dist(q, t, r, io[r, c])
        assert abs(a - b) <= 1e-6 * max(a, b) + 1e-9, (r, c, a, b)
    return len(bad)


@pytest.mark.parametrize("nq,nt,dim", [(1000, 1500, 64), (300, 200, 33), (257, 129, 128), (5, 2, 64), (129, 4100, 32),
                                       (2000, 2000, 64)])
def test_knn2_l2_vs_oracle(ctx, nq, nt, dim):
    rng = np.random.default_rng(nq + nt + dim)
    q = rng.normal(size=(nq, dim)).astype(np.float32); t = rng.normal(size=(nt, dim)).astype(np.float32)
    m = min(nq, nt) // 2
    t[:m] = q[:m] + 0.05 * rng.normal(size=(m, dim)).astype(np.float32)
    ig, dg = ctx.knn2_l2(q, t)
    st = ctx.l2_stats()
    io, do = orc.knn2_l2(q, t)
    check_knn(q, t, ig, dg, io, do)
    if nt >= 1000:      # the tensor-core path must carry the result: the exact fallback is the rare exception
        assert st["fallback_fwd"] <= 0.02 * nq, st


def test_knn2_l2_near_duplicates_force_exact_fallback(ctx):
    """Clusters of almost identical train rows: TF32 cannot separate them, the proof fails, the exact kernel decides."""
    rng = np.random.default_rng(7)
    base = rng.normal(size=(40, 64)).astype(np.float32)
    t = (base[:, None, :] + 1e-4 * rng.normal(size=(40, 30, 64)).astype(np.float32)).reshape(-1, 64)
    q = base + 1e-4 * rng.normal(size=base.shape).astype(np.float32)
    ig, dg = ctx.knn2_l2(q, t)
    assert ctx.l2_stats()["fallback_fwd"] > 0
    io, do = orc.knn2_l2(q, t)
    check_knn(q, t, ig, dg, io, do)
    t[7] = t[3]                                  # exact duplicates: lowest index first
    ig, dg = ctx.knn2_l2(t[3:4], t)
    assert list(ig[0]) == [3, 7] and np.all(dg[0] == 0)


@pytest.mark.parametrize("cross", [False, True])
def test_match_l2_vs_oracle(ctx, cross):
    q, t = synth.synthetic_l2(1500, 1800, 64, seed=5)
    g = ctx.match_l2(q, t, 0.7, -1.0, cross)
    o = orc.match_l2(q, t, 0.7, -1.0, cross)
    assert len(g) == len(o) > 100
    assert np.array_equal(g["query"], o["query"]) and np.array_equal(g["train"], o["train"])
    assert np.allclose(g["distance"], o["distance"], rtol=1e-5)


def test_knn2_l2_full_size_32k(ctx):
    """BASELINE config 4 (32768 x 32768 x 64): oracle check on a query subset + planted-match recovery + idempotence."""
    q, t = synth.synthetic_l2(32768, 32768, 64)
    ig, dg = ctx.knn2_l2(q, t)
    st = ctx.l2_stats()
    print("l2 32k stats:", st, "TFLOP/s (tf32 GEMM kernel):", 2 * 32768 * 32768 * 64 / (st["gemm_us"] * 1e-6) / 1e12)
    assert st["fallback_fwd"] <= 0.01 * 32768, st
    sub = np.random.default_rng(0).choice(32768, 64, replace=False)
    io, do = orc.knn2_l2(q[sub], t)
    check_knn(q[sub], t, ig[sub], dg[sub], io, do)
    assert np.all(dg[:, 0] <= dg[:, 1]) and ig.min() >= 0 and ig.max() < 32768
    # the first half of the queries has a planted noisy copy in T (distance ~ 0.05*sqrt(64)/|.| ~ 0.37): found as 1st neighbour
    assert (dg[:16384, 0] < 0.6).mean() > 0.999
    ig2, dg2 = ctx.knn2_l2(q, t)
    st = ctx.l2_stats()
    print("l2 32k warm stats:", st, "TFLOP/s (tf32 GEMM kernel):", 2 * 32768 * 32768 * 64 / (st["gemm_us"] * 1e-6) / 1e12)
    assert np.array_equal(ig, ig2) and np.array_equal(dg, dg2)


@pytest.mark.parametrize("nq,nt,dim", [(2000, 3000, 64), (700, 900, 128), (500, 600, 48)])
def test_knn2_l2_device_pointers_equal_host_call(ctx, nq, nt, dim):
    """Descriptor sets already resident in HBM are read in place when their rows are TMA-compatible (dim % 32 == 0), padded
    through the workspace otherwise; either way the result equals the host-buffer call bit for bit, and the inputs are intact."""
    torch = pytest.importorskip("torch")
    q, t = synth.synthetic_l2(nq, nt, dim, seed=77)
    ih, dh = ctx.knn2_l2(q, t)
    dq, dt = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    di = torch.empty((nq, 2), dtype=torch.int32, device="cuda"); dd = torch.empty((nq, 2), dtype=torch.float32, device="cuda")
    ctx.knn2_l2_ptr(dq.data_ptr(), nq, dt.data_ptr(), nt, dim, di.data_ptr(), dd.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(di.cpu().numpy(), ih) and np.array_equal(dd.cpu().numpy(), dh)
    assert np.array_equal(dq.cpu().numpy(), q) and np.array_equal(dt.cpu().numpy(), t)
