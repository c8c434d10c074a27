#!/usr/bin/env python
"""BASELINE config 4: 32k x 32k x 64 float L2 kNN(2).  Prints warm device timings of mvs_knn2_l2."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mvslam_b200 as mvs
from mvslam_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
q, t = synth.synthetic_l2(n, n, 64)
ctx = mvs.Context(0)
out = []
for r in range(reps):
    t0 = time.perf_counter(); ctx.knn2_l2(q, t); wall = time.perf_counter() - t0
    st = ctx.l2_stats(); st["wall_ms"] = wall * 1e3
    out.append(st)
best = min(out[1:] or out, key=lambda s: s["gemm_us"])
flops = 2.0 * n * n * 64
print(json.dumps(dict(n=n, dim=64, gemm_ms=best["gemm_us"] / 1e3, total_device_ms=best["total_us"] / 1e3,
                      wall_ms=best["wall_ms"], gemm_tflops=flops / (best["gemm_us"] * 1e-6) / 1e12,
                      fallbacks=best["fallback_fwd"], runs=out)))
