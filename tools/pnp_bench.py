#!/usr/bin/env python
"""pnp_solve throughput on one GPU: mvs_pnp_solve_batch over seeded synthetic problems (30 % gross outliers, 0.01 px
noise), device time of the three kernels and wall time of the call with host buffers; cv2.solvePnPRansac (the
third-party routine the reference calls, pnp-solve.cpp:53-66) timed on the host beside it.  Prints one JSON object."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mvslam_b200 as mvs  # noqa: E402
from pnp_scenes import K_PNP, scene  # noqa: E402


def run(ctx, n_prob=1024, n_pts=500, H=100, steps=10, cpu=True):
    probs = [scene(n_pts, 0.3, 0.01, seed=1000 + i) for i in range(min(n_prob, 64))]
    worlds = [probs[i % len(probs)][0] for i in range(n_prob)]
    images = [probs[i % len(probs)][1] for i in range(n_prob)]
    for _ in range(3):
        res, _ = ctx.pnp_solve_batch(worlds, images, K_PNP, H=H, seed=1)
    counts = np.array([len(w) for w in worlds], np.int32)
    wcat, icat = np.ascontiguousarray(np.concatenate(worlds)), np.ascontiguousarray(np.concatenate(images))
    ctx.profile_enable(True); ctx.profile_read(True)
    t0 = time.perf_counter()
    for _ in range(steps):
        res, _ = ctx.pnp_solve_batch(wcat, icat, K_PNP, H=H, seed=1, counts=counts)
    wall = (time.perf_counter() - t0) / steps
    prof = ctx.profile_read(True); ctx.profile_enable(False)
    dev_ms = prof["pnp"][0] / steps
    out = dict(problems=n_prob, points_per_problem=n_pts, hypotheses=H, solved=int((res["status"] == 0).sum()),
               device_ms_per_batch=dev_ms, problems_per_s_device=n_prob / (dev_ms * 1e-3),
               hyp_pt_evals_per_s=n_prob * H * n_pts / (dev_ms * 1e-3),
               e2e_problems_per_s=n_prob / wall, e2e_note="one mvs_pnp_solve_batch call on concatenated host arrays: H2D, kernels, D2H of results and masks",
               mean_inliers=float(res["n_inliers"].mean()))
    t0 = time.perf_counter()
    for _ in range(50):
        ctx.pnp_solve(worlds[0], images[0], K_PNP, H=H, seed=1)
    out["single_problem_latency_us"] = (time.perf_counter() - t0) / 50 * 1e6
    if cpu:
        try:
            import cv2
            cv2.setNumThreads(os.cpu_count())
            t0 = time.perf_counter(); k = 0
            while time.perf_counter() - t0 < 3.0:
                X, uv = worlds[k % n_prob], images[k % n_prob]
                cv2.solvePnPRansac(X.reshape(-1, 1, 3), uv.reshape(-1, 1, 2), K_PNP, None, iterationsCount=H,
                                   reprojectionError=0.05, confidence=0.95, flags=cv2.SOLVEPNP_P3P)
                k += 1
            out["cv2_solvepnpransac_problems_per_s"] = k / (time.perf_counter() - t0)
            out["cv2_note"] = "one host thread per call (cv2 does not parallelise solvePnPRansac); early exit by confidence 0.95"
        except ImportError:
            out["cv2_solvepnpransac_problems_per_s"] = None
    return out


if __name__ == "__main__":
    n_prob = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    n_pts = int(sys.argv[2]) if len(sys.argv) > 2 else 500
    H = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    with mvs.Context(0) as ctx:
        print(json.dumps(run(ctx, n_prob, n_pts, H)))
