#!/usr/bin/env python
"""Single-pair mvs_pair_batch latency: wall time per call with / without detail outputs and the per-stage device times
(run on the GPU box)."""
import sys, os, time, json
import numpy as np
sys.path.insert(0, os.getcwd())
import mvslam_b200 as mvs
f = np.load("tests/golden/tsukuba_orb2000.npz")
descs = [f[f"desc{i}"] for i in range(1, 6)]; kps = [f[f"kp{i}"] for i in range(1, 6)]
ctx = mvs.Context(0)
ctx.frames_upload(descs, kps)
for H, slv in ((1, "reference"), (1024, "fast")):
    kw = dict(max_dist=10.0, H=H, seed=0, solver=slv)
    for _ in range(20): ctx.pair_batch([(0, 1)], f["K"], **kw)
    ctx.profile_enable(True); ctx.profile_read(True)
    t0 = time.perf_counter()
    for _ in range(200): ctx.pair_batch([(0, 1)], f["K"], **kw)
    wall = (time.perf_counter() - t0) / 200 * 1e6
    p = ctx.profile_read(True); ctx.profile_enable(False)
    print(H, slv, "wall (profiling on) us", round(wall, 1), {k: round(v[0] / 200 * 1e3, 1) for k, v in p.items() if v[0] > 0}, "sum", round(sum(v[0] for v in p.values()) / 200 * 1e3, 1))
    t0 = time.perf_counter()
    for _ in range(200): ctx.pair_batch([(0, 1)], f["K"], **kw)
    print("  wall us", round((time.perf_counter() - t0) / 200 * 1e6, 1))
    # without details
    t0 = time.perf_counter()
    for _ in range(200): ctx.pair_batch([(0, 1)], f["K"], details=False, **kw)
    print("  no details wall us", round((time.perf_counter() - t0) / 200 * 1e6, 1))
    import torch
    res_t = torch.empty(mvs.RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    t0 = time.perf_counter()
    for _ in range(200): ctx.pair_batch([(0, 1)], f["K"], out=dict(results=res_t.data_ptr()), **kw)
    print("  pinned results only wall us", round((time.perf_counter() - t0) / 200 * 1e6, 1))
