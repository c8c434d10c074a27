#!/usr/bin/env python
"""Write feature files for the C++ tools (include/mvslam/feature-io.hpp).
  export_features.py npz <tsukuba_orb2000.npz> <out_dir>      five frames + camera.config + features.txt
  export_features.py pgm <tsukuba_gray.npz> <out_dir>         five binary PGM frames + camera.config + image.txt
  export_features.py image <image> <out_file> [nfeatures]     cv2.ORB on the host"""
import os, struct, sys
import numpy as np


def write(path, kp, desc, w, h):
    with open(path, "wb") as f:
        f.write(b"MVSF" + struct.pack("<iii", len(kp), int(w), int(h)))
        f.write(np.ascontiguousarray(kp, np.float32).tobytes()); f.write(np.ascontiguousarray(desc, np.uint8).tobytes())


if sys.argv[1] == "npz":
    z = np.load(sys.argv[2]); out = sys.argv[3]; os.makedirs(out, exist_ok=True)
    h, w = [int(v) for v in z["image_hw"]]
    names = []
    for i in range(1, 6):
        write(os.path.join(out, f"{i}.mvsf"), z[f"kp{i}"], z[f"desc{i}"], w, h); names.append(f"{i}.mvsf")
    K = z["K"]
    open(os.path.join(out, "camera.config"), "w").write(f"{K[0,0]:g} {K[1,1]:g} {K[0,1]:g} {K[0,2]:g} {K[1,2]:g}\n0 0 0 1.5708 0 0\n")
    open(os.path.join(out, "features.txt"), "w").write("\n".join(names) + "\n")
elif sys.argv[1] == "pgm":
    g = np.load(sys.argv[2])["gray"]; out = sys.argv[3]; os.makedirs(out, exist_ok=True)
    for i, im in enumerate(g, 1):
        with open(os.path.join(out, f"{i}.pgm"), "wb") as f:
            f.write(b"P5\n%d %d\n255\n" % (im.shape[1], im.shape[0])); f.write(np.ascontiguousarray(im, np.uint8).tobytes())
    open(os.path.join(out, "camera.config"), "w").write("350 350 0 192 144\n0 0 0 1.5708 0 0\n")     # data/tsukuba/camera.config
    open(os.path.join(out, "image.txt"), "w").write("\n".join(f"{i}.pgm" for i in range(1, len(g) + 1)) + "\n")
else:
    import cv2
    im = cv2.imread(sys.argv[2], cv2.IMREAD_GRAYSCALE)
    kp, d = cv2.ORB_create(int(sys.argv[4]) if len(sys.argv) > 4 else 500).detectAndCompute(im, None)
    write(sys.argv[3], np.array([k.pt for k in kp], np.float32), d, im.shape[1], im.shape[0])
