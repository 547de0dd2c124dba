"""Generates tests/golden/*.npz with OpenCV 4.13 (cv2) -- the importable bit-exact reference of the detect/describe
variant named by BASELINE.json. Run in the authoring container:  python tests/golden/make_golden.py
The oracle (oracle/brisk_oracle.c) and the CUDA path are both checked against these files; cv2 is NOT needed to run
the tests.
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from okvis2_b200.synth import synth_frame  # noqa: E402

REF_IMG = "/root/reference/okvis_multisensor_processing/test/testImage.jpg"


def kp_array(kps):
    a = np.zeros((len(kps), 7), np.float64)
    for i, k in enumerate(kps):
        a[i] = (k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave, k.class_id)
    return a.astype(np.float32)


def brisk(img, thr, octv):
    b = cv2.BRISK_create(thr, octv, 1.0)
    kps, desc = b.detectAndCompute(img, None)
    return kp_array(kps), desc


def main():
    cv2.setNumThreads(1)
    real = cv2.imread(REF_IMG, 0)
    assert real is not None
    out = {}
    # real image, EuRoC size (kept as raw pixels so that JPEG decoding is out of the loop)
    img = cv2.resize(real, (752, 480), interpolation=cv2.INTER_AREA)
    out["real752_img"] = img
    for thr, octv in [(30, 0), (30, 3), (60, 2)]:
        kp, d = brisk(img, thr, octv)
        out[f"real752_t{thr}_o{octv}_kp"] = kp
        out[f"real752_t{thr}_o{octv}_desc"] = d
    # pyramid layers of the real image through cv2.resize (what cv::BRISK's half/two-third sampling calls)
    l1 = cv2.resize(img, (2 * (752 // 3), 2 * (480 // 3)), interpolation=cv2.INTER_AREA)
    l2 = cv2.resize(img, (376, 240), interpolation=cv2.INTER_AREA)
    l3 = cv2.resize(l1, (l1.shape[1] // 2, l1.shape[0] // 2), interpolation=cv2.INTER_AREA)
    out["real752_layer1"], out["real752_layer2"], out["real752_layer3"] = l1, l2, l3
    # AGAST integer scores (sparse) on the real image
    for name, typ in [("oast916", cv2.AgastFeatureDetector_OAST_9_16), ("agast58", cv2.AgastFeatureDetector_AGAST_5_8)]:
        det = cv2.AgastFeatureDetector_create(20, False, typ)
        ks = det.detect(img)
        out[f"real752_{name}_t20"] = np.array([(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in ks], np.int32)
    # an odd-sized crop exercises the general INTER_AREA path on every level (341 -> 170 is not a factor 2)
    crop = np.ascontiguousarray(cv2.resize(real, (341, 255), interpolation=cv2.INTER_AREA))
    out["real341_img"] = crop
    kp, d = brisk(crop, 25, 2)
    out["real341_t25_o2_kp"], out["real341_t25_o2_desc"] = kp, d
    # synthetic frames are regenerated from their seed by the tests; only the cv2 answers are stored
    for seed, W, H, thr, octv in [(1000, 752, 480, 30, 3), (1001, 752, 480, 30, 0), (2000, 1024, 1024, 30, 3),
                                  (3000, 720, 540, 30, 3)]:
        im = synth_frame(seed, W, H)
        kp, d = brisk(im, thr, octv)
        out[f"synth{seed}_{W}x{H}_t{thr}_o{octv}_kp"] = kp
        out[f"synth{seed}_{W}x{H}_t{thr}_o{octv}_desc"] = d
        out[f"synth{seed}_{W}x{H}_crc"] = np.array([int(im.astype(np.uint64).sum()), int((im.astype(np.uint64) * np.arange(im.size).reshape(im.shape) % 65521).sum())])
    np.savez_compressed(os.path.join(HERE, "brisk_cv2_4_13.npz"), **out)
    print("wrote", len(out), "arrays;", os.path.getsize(os.path.join(HERE, "brisk_cv2_4_13.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
