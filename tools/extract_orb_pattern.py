#!/usr/bin/env python
"""Write mvslam_b200/csrc/orb_pattern.h: the 256 x 4 rBRIEF sampling table of cv::ORB (`bit_pattern_31_`,
OpenCV modules/features2d/src/orb.cpp — learned test locations published with the ORB paper).  The table is data of
the un-vendored third-party dependency the reference calls (source/vision/visual-feature.cpp:12-17); it is read out of
the installed OpenCV binary (located by its well-known first eight entries) so that no OpenCV source is needed."""
import glob
import os

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
head = np.array([8, -3, 9, 5, 4, 2, 7, -12], "<i4").tobytes()
for so in glob.glob(os.path.join(os.path.dirname(cv2.__file__), "*.so")):
    blob = open(so, "rb").read()
    at = blob.find(head)
    if at >= 0:
        break
else:
    raise SystemExit("pattern not found")
tab = np.frombuffer(blob[at:at + 4096], "<i4").reshape(256, 4)
assert tab.min() >= -15 and tab.max() <= 15
with open(os.path.join(ROOT, "mvslam_b200", "csrc", "orb_pattern.h"), "w") as f:
    f.write("// rBRIEF test locations of cv::ORB (bit_pattern_31_), 256 tests x (x0, y0, x1, y1) in a 31x31 patch.\n"
            f"// Data extracted from OpenCV {cv2.__version__} by tools/extract_orb_pattern.py; do not edit.\n"
            "#pragma once\n#include <stdint.h>\nstatic const int8_t MVS_ORB_PATTERN[256 * 4] = {\n")
    for r in tab:
        f.write("    %d, %d, %d, %d,\n" % tuple(r))
    f.write("};\n")
print("wrote", tab.shape)
