#!/usr/bin/env python
"""Feature-extraction throughput on one GPU: mvs_orb_extract over a batch of Tsukuba-size frames, device-resident and
end to end from host images, with the per-stage device times; cv2.ORB (the third-party routine the reference calls)
timed on the host beside it.  Prints one JSON object."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mvslam_b200 as mvs  # noqa: E402
from mvslam_b200 import synth  # noqa: E402


def main():
    n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    nf = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    gray = np.load(os.path.join(ROOT, "tests", "golden", "tsukuba_gray.npz"))["gray"]
    imgs = [gray[i % 5] for i in range(n_img)]
    h, w = imgs[0].shape
    out = dict(images=n_img, width=w, height=h, n_features=nf, steps=steps)
    with mvs.Context(0) as ctx:
        s = torch.cuda.Stream()
        ctx.set_stream(s.cuda_stream)
        dev = torch.from_numpy(np.stack(imgs)).cuda()
        for _ in range(3):
            counts, _, _, _ = ctx.orb_extract(None, nf, want=False, device_ptr=dev.data_ptr(), shape=(n_img, h, w, w))
        out["keypoints_per_image"] = float(counts.mean())
        ctx.profile_enable(True); ctx.profile_read(True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(steps):
            ctx.orb_extract(None, nf, want=False, device_ptr=dev.data_ptr(), shape=(n_img, h, w, w))
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        prof = ctx.profile_read(True); ctx.profile_enable(False)
        out["device_resident_frames_per_s"] = n_img * steps / dt
        out["device_resident_ms_per_batch"] = dt / steps * 1e3
        out["stage_ms_per_batch"] = {k: round(v[0] / steps, 4) for k, v in prof.items() if k.startswith("orb")}
        for _ in range(2):
            ctx.orb_extract(imgs, nf)
        t0 = time.perf_counter()
        for _ in range(steps):
            ctx.orb_extract(imgs, nf)
        dt = time.perf_counter() - t0
        out["e2e_frames_per_s"] = n_img * steps / dt
        # single-frame latency (the VO case)
        for _ in range(5):
            ctx.orb_extract(imgs[:1], nf)
        lat = []
        for _ in range(50):
            t0 = time.perf_counter(); ctx.orb_extract(imgs[:1], nf); lat.append(time.perf_counter() - t0)
        out["single_frame_latency_us_median"] = float(np.median(lat) * 1e6)
    try:
        import cv2
        cv2.setNumThreads(os.cpu_count())
        orb = cv2.ORB_create(nf)
        t0 = time.perf_counter(); n = 0
        while time.perf_counter() - t0 < 5.0:
            kp = orb.detect(imgs[n % n_img], None); orb.compute(imgs[n % n_img], kp); n += 1
        out["cv2_orb_frames_per_s"] = n / (time.perf_counter() - t0)
        out["cv2_threads"] = cv2.getNumThreads()
    except ImportError:
        out["cv2_orb_frames_per_s"] = None
    print(json.dumps(out))


if __name__ == "__main__":
    main()
