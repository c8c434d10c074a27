#!/usr/bin/env python
"""Stage times of the default workload (1024 Tsukuba VO pairs) from the library's own CUDA-event profile; used to
A/B matcher experiments.  Prints one JSON line."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mvslam_b200 as mvs  # noqa: E402
from mvslam_b200 import capi  # noqa: E402

if os.environ.get("MVS_LIB_OVERRIDE"):     # experiment builds (e.g. -DMVS_TC_PROBE) live outside the package directory
    capi.lib_path = lambda: os.path.abspath(os.environ["MVS_LIB_OVERRIDE"])


def main():
    npairs = int(os.environ.get("PAIRS", "1024")); H = int(os.environ.get("H", "1024")); solver = os.environ.get("SOLVER", "fast")
    f = np.load(os.path.join(ROOT, "tests", "golden", "tsukuba_orb2000.npz"))
    descs = [np.ascontiguousarray(f[f"desc{i}"]) for i in range(1, 6)]
    kps = [np.ascontiguousarray(f[f"kp{i}"]) for i in range(1, 6)]
    pairs = np.array([[(i % 4), (i % 4) + 1] for i in range(npairs)], np.int32)
    with mvs.Context(0) as ctx:
        ctx.frames_upload(descs, kps)
        for _ in range(3):
            ctx.pair_batch(pairs, f["K"], max_dist=10.0, H=H, details=False, solver=solver)
        ctx.profile_enable(True)
        ctx.profile_read()
        reps = int(os.environ.get("REPS", "20"))
        for _ in range(reps):
            res, _ = ctx.pair_batch(pairs, f["K"], max_dist=10.0, H=H, details=False, solver=solver)
        pr = ctx.profile_read()
        out = {k: round(v[0] / reps, 4) for k, v in pr.items() if v[0] > 0}
        out["total_ms"] = round(sum(out.values()), 4)
        out["checksum"] = int(res["n_inliers"].astype(np.int64).sum()); out["n_matches"] = int(res["n_matches"].sum())
        out["tag"] = os.environ.get("TAG", "")
        L = capi.load_library()
        if hasattr(L, "mvs_debug_tc_probe_dump"):   # -DMVS_TC_PROBE build: counters of the last matcher launch
            import ctypes
            buf = np.zeros((160, 10), np.uint64)
            L.mvs_debug_tc_probe_dump(buf.ctypes.data_as(ctypes.c_void_p))
            b = buf[:148].astype(np.float64)
            out["probe"] = dict(issuer_clocks_max=b[:, 0].max(), issuer_clocks_mean=b[:, 0].mean(), cta_wall_us_max=b[:, 7].max() / 1e3,
                                sm_mhz_during_kernel=float(np.median(b[:, 0] / np.maximum(b[:, 7], 1) * 1e3)),
                                issuerA_wait_query=b[:, 1].mean(), issuerA_wait_train=b[:, 2].mean(), issuerA_wait_acc=b[:, 3].mean(),
                                issuerB_wait_query=b[:, 4].mean(), issuerB_wait_train=b[:, 5].mean(), issuerB_wait_acc=b[:, 6].mean(),
                                epilogue_wait_acc=b[:, 8].mean())
        print(json.dumps(out))


if __name__ == "__main__":
    main()
