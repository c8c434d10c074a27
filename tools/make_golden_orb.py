#!/usr/bin/env python
"""Generate the committed feature-extraction fixtures under tests/golden/ from the real third-party routine the
reference calls (cv::ORB through cv2, source/vision/visual-feature.cpp:9-17,40-49).

Run in the build container (needs /root/reference/data/tsukuba and cv2 4.13.0); the GPU box only reads the .npz.

  tsukuba_gray.npz   the reference's five bundled New Tsukuba frames decoded to 8-bit grayscale (the input of
                     VisualFeature::extract; utility/visual-odometer.cpp reads them with cv::IMREAD_GRAYSCALE)
  orb_golden.npz     cv2.ORB_create(nf).detect + .compute (the reference's two calls) for nf = 500 (reference's
                     MAX_FEATURE_COUNT, all 5 frames) and nf = 2000 (BASELINE.json's "~2k ORB keypoints", frames 1-2),
                     re-ordered into the canonical order (level, y, x in level coordinates)
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import orb_np as O  # noqa: E402

REF = "/root/reference/data/tsukuba"
OUT = os.path.join(ROOT, "tests", "golden")


def cv2_orb_canonical(img, nf, two_calls=True):
    orb = cv2.ORB_create(nf)
    if two_calls:
        kps = orb.detect(img, None)
        kps, desc = orb.compute(img, kps)
    else:
        kps, desc = orb.detectAndCompute(img, None)
    if len(kps) == 0:
        z = np.zeros(0, np.float32)
        return dict(pt=np.zeros((0, 2), np.float32), level_xy=np.zeros((0, 2), np.int32), octave=np.zeros(0, np.int32),
                    size=z, angle=z, response=z, desc=np.zeros((0, 32), np.uint8))
    scales = O.level_scales()
    pt = np.array([k.pt for k in kps], np.float32)
    octv = np.array([k.octave for k in kps], np.int32)
    inv = np.array([np.float32(1) / scales[o] for o in octv], np.float32)
    lxy = np.rint(pt * inv[:, None]).astype(np.int32)        # computeOrbDescriptors: cvRound(pt * (1 / scale))
    order = np.lexsort((lxy[:, 0], lxy[:, 1], octv))
    return dict(pt=pt[order], level_xy=lxy[order], octave=octv[order],
                size=np.array([k.size for k in kps], np.float32)[order],
                angle=np.array([k.angle for k in kps], np.float32)[order],
                response=np.array([k.response for k in kps], np.float32)[order], desc=desc[order])


def main():
    gray = np.stack([cv2.imread(os.path.join(REF, f"{i}.jpg"), cv2.IMREAD_GRAYSCALE) for i in range(1, 6)])
    np.savez_compressed(os.path.join(OUT, "tsukuba_gray.npz"), gray=gray)
    out = {}
    for nf, frames in ((500, range(5)), (2000, range(2))):
        for f in frames:
            r = cv2_orb_canonical(gray[f], nf)
            for k, v in r.items():
                out[f"n{nf}_f{f + 1}_{k}"] = v
            print(nf, f + 1, len(r["pt"]))
    out["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(OUT, "orb_golden.npz"), **out)


if __name__ == "__main__":
    main()
