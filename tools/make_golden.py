#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/.

Run in the build container (needs /root/reference/data/tsukuba and cv2); the GPU box never
runs this — it only reads the committed .npz files.

  tsukuba_orb2000.npz   ORB (nfeatures=2000, cv2 defaults otherwise) keypoints+descriptors of the
                        reference's five bundled New Tsukuba frames (data/tsukuba/{1..5}.jpg) and
                        the camera matrix from data/tsukuba/camera.config.  ORB extraction is
                        outside the hot path and stays on the host.
  tsukuba_golden.npz    Oracle-A (cv2.batchDistance + cv2.SVDecomp restatement of the reference's
                        own branch, oracle/oracle_np.py) outputs for the 4 consecutive pairs:
                        matches at max_dist 10/30/-1, and sfm_solve with the reference's single
                        sample {0..7} (H=1) and with a 256-row seeded table.
  synthetic_golden.npz  Oracle-A outputs on small general-motion scenes (noise-free and noisy,
                        with outliers) and the L-shape / cube rigs of the reference's tests.
  cv_svd_golden.npz     cv2.SVDecomp(MODIFY_A | FULL_UV) outputs (w, u, vt) for 3x3, 4x4 and 9x9 inputs of the
                        kinds the path produces (general, A^T A of 8-point samples, rank-deficient, exactly
                        singular): the known-answer vectors of the bit-for-bit restatement of OpenCV's
                        small-matrix SVD (oracle orc_cv_svd, device cv_svd_full / MVS_SOLVER_REFERENCE).
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_np as A  # noqa: E402
from oracle import cbind as B      # noqa: E402  (only for the shared seeded sample table)

REF = "/root/reference/data/tsukuba"
OUT = os.path.join(ROOT, "tests", "golden")


def tsukuba_features():
    out = {}
    orb = cv2.ORB_create(2000)
    for i in range(1, 6):
        im = cv2.imread(os.path.join(REF, f"{i}.jpg"), cv2.IMREAD_GRAYSCALE)
        kp, d = orb.detectAndCompute(im, None)
        out[f"desc{i}"] = d
        out[f"kp{i}"] = np.array([k.pt for k in kp], np.float32)
        out["image_hw"] = np.array(im.shape, np.int32)
    fx, fy, sh, px, py = [float(v) for v in open(os.path.join(REF, "camera.config")).readline().split()]
    out["K"] = np.array([[fx, sh, px], [0, fy, py], [0, 0, 1]], np.float64)
    return out


def solve_pack(prefix, out, xy1, xy2, K, samples, mode="algebraic"):
    r = A.sfm_solve(xy1, xy2, K, samples, mode)
    out[prefix + "ok"] = np.array(r is not None)
    if r is None:
        return
    for k in ("F", "E", "mask", "R1to2", "t1to2", "R2in1", "t2in1", "points", "indexes"):
        out[prefix + k] = r[k]
    out[prefix + "n_inliers"] = np.array(r["n_inliers"])
    out[prefix + "residual"] = np.array(r["residual"])
    out[prefix + "best_h"] = np.array(r["best_h"])


def main():
    os.makedirs(OUT, exist_ok=True)
    feats = tsukuba_features()
    np.savez_compressed(os.path.join(OUT, "tsukuba_orb2000.npz"), **feats)
    K = feats["K"]

    g = {}
    for a in range(1, 5):
        b = a + 1
        for md in (10, 30, -1):
            q, t, d = A.match_visual_features(feats[f"desc{a}"], feats[f"desc{b}"], float(md))
            tag = f"p{a}{b}_md{md}_"
            g[tag + "q"], g[tag + "t"], g[tag + "d"] = q, t, d
            xy1 = feats[f"kp{a}"][t].astype(np.float64)
            xy2 = feats[f"kp{b}"][q].astype(np.float64)
            solve_pack(tag + "h1_", g, xy1, xy2, K, None)
            if md == 30:
                tab = B.sample_table(0, a - 1, len(q), 256)
                g[tag + "tab256"] = tab
                solve_pack(tag + "h256_", g, xy1, xy2, K, tab)
                solve_pack(tag + "h256s_", g, xy1, xy2, K, tab, "sampson")
    np.savez_compressed(os.path.join(OUT, "tsukuba_golden.npz"), **g)

    s = {}
    Ks = np.array([[700, 0, 640], [0, 700, 360], [0, 0, 1.0]])
    for case, (n, noise, outl) in enumerate([(64, 0.0, 0.0), (300, 0.0, 0.3), (300, 1e-4, 0.3), (400, 0.5, 0.3)]):
        r = np.random.default_rng(100 + case)
        X = np.stack([r.uniform(-4, 4, n), r.uniform(-4, 4, n), r.uniform(4, 12, n)], 1)
        rv = r.normal(size=3); rv = rv / np.linalg.norm(rv) * 0.1 * r.uniform() ** (1 / 3)
        R = A.rodrigues(rv); t = r.normal(size=3); t = t / np.linalg.norm(t) * 0.5
        x1 = A.project_points(Ks, np.eye(3), np.zeros(3), X) + r.normal(size=(n, 2)) * noise
        x2 = A.project_points(Ks, R, t, X) + r.normal(size=(n, 2)) * noise
        no = int(outl * n)
        x2[:no] = r.uniform(0, 1, (no, 2)) * [1280, 720]
        perm = r.permutation(n)
        x1, x2, X = x1[perm], x2[perm], X[perm]
        tab = B.sample_table(7, case, n, 128)
        tag = f"s{case}_"
        s[tag + "xy1"], s[tag + "xy2"], s[tag + "K"], s[tag + "tab"] = x1, x2, Ks, tab
        s[tag + "R"], s[tag + "t"], s[tag + "X"] = R, t, X
        solve_pack(tag + "alg_", s, x1, x2, Ks, tab)
        solve_pack(tag + "smp_", s, x1, x2, Ks, tab, "sampson")
    # rigs of test/test-sfm.cpp: cam2 at x=+1 (pose2in1 = (I,(1,0,0))), K = I
    for name, Rr, sc in (("lshape", A.so3_from_rpy(1.5, 0.7, 0.0), 0.5), ("cube", np.eye(3), 1.0)):
        P = A.get_rig_points(name, Rr, np.array([0.6, 0.0, 3.0]), sc)
        x1 = A.project_points(np.eye(3), np.eye(3), np.zeros(3), P)
        x2 = A.project_points(np.eye(3), np.eye(3), np.array([-1.0, 0, 0]), P)
        s[name + "_P"], s[name + "_xy1"], s[name + "_xy2"] = P, x1, x2
        solve_pack(name + "_", s, x1, x2, np.eye(3), None)
        pts, idx = A.sfm_triangulate(x1, x2, np.eye(3), np.eye(3), np.zeros(3), np.eye(3), np.array([1.0, 0, 0]))
        s[name + "_tri_pts"], s[name + "_tri_idx"] = pts, idx
    np.savez_compressed(os.path.join(OUT, "synthetic_golden.npz"), **s)
    v = {}
    r = np.random.default_rng(2024)
    for n in (3, 4, 9):
        mats = []
        for t in range(96):
            M = r.normal(size=(n, n))
            kind = t % 6
            if kind == 1:
                M[:, -1] = M[:, 0] * 2                       # rank n-1, sigma_n ~ round-off
            elif kind == 2:
                M = M @ np.diag([1.0] * (n - 1) + [0.0])     # exactly singular: cv::RNG regenerates a left vector
            elif kind == 3:
                M = np.zeros((n, n)); M[t % n, (t // n) % n] = 1.0 + t
            elif kind == 4 and n == 9:                       # A^T A of a noisy 8-point sample
                x1 = r.normal(size=(8, 2)); x2 = x1 + r.normal(size=(8, 2)) * 0.05
                Am = np.stack([x2[:, 0] * x1[:, 0], x2[:, 0] * x1[:, 1], x2[:, 0], x2[:, 1] * x1[:, 0],
                               x2[:, 1] * x1[:, 1], x2[:, 1], x1[:, 0], x1[:, 1], np.ones(8)], 1)
                M = np.zeros((9, 9))
                for k in range(8):
                    M = M + np.outer(Am[k], Am[k])
            elif kind == 5:
                M = M * 10.0 ** r.integers(-8, 8)
            mats.append(M)
        mats = np.array(mats)
        ws, us, vts = [], [], []
        for M in mats:
            w, u, vt = cv2.SVDecomp(M.copy(), flags=cv2.SVD_MODIFY_A | cv2.SVD_FULL_UV)
            ws.append(w.ravel()); us.append(u); vts.append(vt)
        v[f"A{n}"], v[f"w{n}"], v[f"u{n}"], v[f"vt{n}"] = mats, np.array(ws), np.array(us), np.array(vts)
    np.savez_compressed(os.path.join(OUT, "cv_svd_golden.npz"), **v)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
