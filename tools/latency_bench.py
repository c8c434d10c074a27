#!/usr/bin/env python
"""BASELINE configs[1] in streaming form: per-call latency of the hot path through the C ABI for the batches a
visual odometer issues — 1 pair (add_frame) and 10 pairs (initialize re-pairings), Tsukuba ~2k ORB keypoints."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mvslam_b200 as mvs

f = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "tsukuba_orb2000.npz"))
descs = [f[f"desc{i}"] for i in range(1, 6)]; kps = [f[f"kp{i}"] for i in range(1, 6)]
ctx = mvs.Context(0)
ctx.frames_upload(descs, kps)
out = {}
for name, pairs, H in (("1_pair_H1", [(0, 1)], 1), ("1_pair_H1024", [(0, 1)], 1024),
                       ("10_pairs_H1024", [(i % 4, 4) for i in range(10)], 1024), ("10_pairs_H1024_bounded", [(i % 4, 4) for i in range(10)], 1024)):
    kw = dict(max_dist=10.0, H=H, seed=0, bounded=name.endswith("bounded"))
    for _ in range(20):
        ctx.pair_batch(pairs, f["K"], **kw)
    ts = []
    for _ in range(200):
        t0 = time.perf_counter(); res, det = ctx.pair_batch(pairs, f["K"], **kw); ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e6
    out[name] = dict(median_us=float(np.median(ts)), p90_us=float(np.percentile(ts, 90)), ok=int((res["status"] == 0).sum()))
# upload of one new frame (what add_frame adds per step)
ts = []
for _ in range(50):
    t0 = time.perf_counter(); ctx.frames_upload(descs, kps); ts.append(time.perf_counter() - t0)
out["frames_upload_5_frames_us"] = float(np.median(np.array(ts) * 1e6))
print(json.dumps(out))
