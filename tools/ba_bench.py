#!/usr/bin/env python
"""Bundle-adjustment throughput on one GPU: mvs_ba_solve_batch over seeded two-view (sfm_refine-shaped) problems;
device time of the kernel, wall time of the call, and the numpy oracle on the host beside it.  Prints one JSON object."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mvslam_b200 as mvs  # noqa: E402
from ba_scenes import two_view  # noqa: E402

K = np.array([[700.0, 0, 640.0], [0, 700.0, 360.0], [0, 0, 1.0]])
NAN6 = np.full((6, 6), np.nan)


def make(seed, n):
    """sfm_refine-shaped problem in the C ABI's layout: camera 1 anchored (1e-5), camera 2 and the points regularised (1e-2)."""
    s = two_view(seed, n=n, noise=0.5 / 700, K=K)
    obs = np.zeros(2 * n, mvs.BA_OBS_DTYPE)
    for f, pts in enumerate((s["p1"], s["p2"])):
        for j in range(n):
            C = s["cov"][j]
            obs[f * n + j] = (f, j, pts[j], (C[0, 0], C[0, 1], C[1, 1]))
    abi = dict(pose_R=np.stack([np.eye(3), s["pose_guess"][0]]), pose_t=np.stack([np.zeros(3), s["pose_guess"][1]]),
               pose_prior_cov=np.stack([np.eye(6) * 1e-10, np.eye(6) * 1e-4]), points=s["points_guess"],
               point_prior_cov=np.stack([np.eye(3) * 1e-4] * n), obs=obs)
    return s, abi


def run(ctx, n_prob=1024, n_pts=200, steps=5, cpu=True):
    base = [make(500 + i, n_pts) for i in range(min(n_prob, 16))]
    problems = [base[i % len(base)][1] for i in range(n_prob)]
    for _ in range(2):
        res = ctx.ba_solve_batch(K, problems)
    ctx.profile_enable(True); ctx.profile_read(True)
    pk = ctx.ba_pack(problems)
    t0 = time.perf_counter()
    for _ in range(steps):
        ctx.ba_solve_packed(K, pk)
    wall = (time.perf_counter() - t0) / steps
    prof = ctx.profile_read(True); ctx.profile_enable(False)
    dev_ms = prof["ba"][0] / steps
    its = np.array([r["iterations"] for r in res])
    out = dict(problems=n_prob, frames_per_problem=2, points_per_problem=n_pts, observations_per_problem=2 * n_pts,
               solved=int(sum(r["status"] == 0 for r in res)), mean_lm_iterations=float(its.mean()),
               device_ms_per_batch=dev_ms, problems_per_s_device=n_prob / (dev_ms * 1e-3),
               e2e_problems_per_s=n_prob / wall, e2e_note="one mvs_ba_solve_batch call on packed host arrays: grouping observations by point, H2D, kernel, D2H of poses, points and covariances")
    t0 = time.perf_counter()
    for _ in range(30):
        ctx.ba_solve_batch(K, problems[:1])
    out["single_problem_latency_us"] = (time.perf_counter() - t0) / 30 * 1e6
    if cpu:      # the CPU checker, timed as the baseline (the only use of oracle/ here)
        from oracle import ba_np as B
        probs = [B.sfm_refine_problem(s["p1"], s["cov"], s["p2"], s["cov"], K, s["pose_guess"], s["points_guess"]) for s, _ in base[:4]]
        t0 = time.perf_counter(); k = 0
        while time.perf_counter() - t0 < 3.0:
            probs[k % len(probs)].solve(); k += 1
        out["cpu_oracle_problems_per_s"] = k / (time.perf_counter() - t0)
        out["cpu_note"] = "oracle/ba_np.py (dense numpy Levenberg-Marquardt, one thread); the reference's GTSAM is not in this image"
    return out


if __name__ == "__main__":
    n_prob = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    n_pts = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    with mvs.Context(0) as ctx:
        print(json.dumps(run(ctx, n_prob, n_pts)))
