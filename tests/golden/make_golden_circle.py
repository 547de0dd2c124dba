"""Generates tests/golden/circle_cv2_4_13.npz: masks drawn by cv2.circle(..., thickness=cv2.FILLED) (OpenCV 4.13.0), the
rasteriser behind the keyframe-overlap masks (Frontend.cpp:1084-1088, ViSlamBackend.cpp:2369). Run in the authoring
container (cv2 importable); the fixture travels, cv2 is not needed by the tests."""
import os

import cv2
import numpy as np

rng = np.random.default_rng(0)
out = {}
cases = []
for (rows, cols) in ((48, 75), (102, 102), (54, 72), (7, 9)):
    for radius in list(range(0, 11)) + [13]:
        img = np.zeros((rows, cols), np.uint8)
        centers = [(0, 0), (cols - 1, rows - 1), (-3, 5), (cols + 2, rows // 2), (cols // 2, -radius), (cols // 2, rows // 2)]
        centers += [(int(rng.integers(-12, cols + 12)), int(rng.integers(-12, rows + 12))) for _ in range(6)]
        for (cx, cy) in centers:
            one = np.zeros((rows, cols), np.uint8)
            cv2.circle(one, (cx, cy), radius, 255, cv2.FILLED)
            cases.append((rows, cols, radius, cx, cy))
            out[f"m{len(cases) - 1}"] = np.packbits(one > 0)
out["cases"] = np.array(cases, np.int32)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "circle_cv2_4_13.npz"), **out)
print(len(cases), "masks")
