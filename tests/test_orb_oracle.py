"""CPU: pins oracle/orb_np.py (the numpy restatement of cv::ORB as VisualFeature::extract uses it, reference
source/vision/visual-feature.cpp:9-17,40-49) against the real third-party routine: live cv2 of this image, stage by
stage and end to end, and the committed goldens made by tools/make_golden_orb.py.  Everything is compared bit-for-bit."""
import os

import numpy as np
import pytest

from mvslam_b200 import synth
from oracle import orb_np as O

cv2 = pytest.importorskip("cv2")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIELDS = ("pt", "level_xy", "octave", "size", "angle", "response", "desc")


def cv2_canonical(img, nf, two_calls=True):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), "..", "tools"))
    from make_golden_orb import cv2_orb_canonical
    return cv2_orb_canonical(img, nf, two_calls)


def assert_same(a, b, what=""):
    assert len(a["pt"]) == len(b["pt"]), f"{what}: {len(a['pt'])} vs {len(b['pt'])} keypoints"
    for k in FIELDS:
        assert np.array_equal(a[k], b[k]), f"{what}: field {k} differs"


def test_pattern_table():
    p = O.load_pattern()
    assert p.shape == (512, 2) and p.min() >= -13 and p.max() <= 13
    assert p[:4].ravel().tolist() == [8, -3, 9, 5, 4, 2, 7, -12]


def test_quota_and_geometry():
    assert O.features_per_level(500) == [109, 90, 75, 63, 52, 44, 36, 31]
    assert sum(O.features_per_level(2000)) == 2000
    assert O.level_sizes(384, 288)[1:3] == [(320, 240), (267, 200)]
    assert O.umax_table()[:16] == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]


@pytest.mark.parametrize("w,h", [(384, 288), (640, 480), (333, 257)])
def test_resize_matches_cv2_linear_exact(w, h):
    rng = np.random.default_rng(w)
    prev = rng.integers(0, 256, (h, w), dtype=np.uint8)
    for l, sz in enumerate(O.level_sizes(w, h)[1:], 1):
        ref = cv2.resize(prev, sz, interpolation=cv2.INTER_LINEAR_EXACT)
        assert np.array_equal(ref, O.resize_linear_exact(prev, *sz)), l
        prev = ref


def test_fast_matches_cv2():
    for seed in range(3):
        img = synth.synthetic_image(seed, 320, 240)
        kp = cv2.FastFeatureDetector_create(O.FAST_THRESHOLD, True).detect(img)
        ref = sorted((int(k.pt[1]), int(k.pt[0]), int(k.response)) for k in kp)
        sc = O.fast_score_map(img)
        ys, xs = np.nonzero(O.fast_nms(sc))
        assert ref == sorted((int(y), int(x), int(sc[y, x])) for y, x in zip(ys, xs))
        assert len(ref) > 100


def test_fast_atan2_matches_cv2():
    rng = np.random.default_rng(5)
    for y, x in rng.integers(-200000, 200000, (2000, 2)):
        assert np.float32(cv2.fastAtan2(float(y), float(x))) == O.fast_atan2(y, x), (y, x)


def test_blur_matches_cv2_float_separable_filter():
    k = O.gaussian_kernel_f32()
    assert np.array_equal(k, cv2.getGaussianKernel(7, 2, cv2.CV_32F).ravel())
    for seed in range(2):
        img = synth.synthetic_image(10 + seed, 320, 240)
        ref = cv2.sepFilter2D(img, cv2.CV_8U, k.reshape(-1, 1), k.reshape(-1, 1), borderType=cv2.BORDER_REFLECT_101)
        assert np.array_equal(ref, O.gaussian_blur_7x7(img))


def test_oracle_matches_golden_tsukuba():
    gray = np.load(os.path.join(GOLDEN, "tsukuba_gray.npz"))["gray"]
    gold = np.load(os.path.join(GOLDEN, "orb_golden.npz"))
    for nf, frames in ((500, (1, 3, 5)), (2000, (1,))):
        for f in frames:
            r = O.orb_extract(gray[f - 1], nf)
            assert_same({k: gold[f"n{nf}_f{f}_{k}"] for k in FIELDS}, r, f"tsukuba frame {f}, nfeatures {nf}")


@pytest.mark.parametrize("seed,w,h,nf", [(1, 640, 480, 500), (2, 333, 257, 2000), (3, 200, 150, 1000)])
def test_oracle_matches_live_cv2(seed, w, h, nf):
    img = synth.synthetic_image(seed, w, h)
    ref = cv2_canonical(img, nf)
    assert len(ref["pt"]) > 50
    assert_same(ref, O.orb_extract(img, nf), f"synthetic {w}x{h}")


def test_detect_then_compute_equals_detect_and_compute():
    """The reference calls detect() and compute() separately (visual-feature.cpp:44-45)."""
    img = synth.synthetic_image(4, 400, 300)
    assert_same(cv2_canonical(img, 500, True), cv2_canonical(img, 500, False))


def test_degenerate_images():
    flat = np.full((120, 160), 77, np.uint8)
    assert len(O.orb_extract(flat, 500)["pt"]) == 0 and len(cv2_canonical(flat, 500)["pt"]) == 0
    small = synth.synthetic_image(6, 70, 66)      # only level 0 has an interior beyond the 31-pixel border
    assert_same(cv2_canonical(small, 500), O.orb_extract(small, 500), "70x66")
