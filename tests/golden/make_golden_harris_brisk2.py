"""Generates tests/golden/harris_brisk2_oracle.npz: frozen outputs of THIS REPOSITORY'S restatement of the D = 48 mode
(oracle/brisk_oracle.c section 6) -- NOT outputs of smartroboticslab/brisk, whose source is absent from the reference tree (the mode
is parity-unpinned against it). The vectors freeze the definition DESIGN.md section 2b documents, so that a later change of the oracle
(or of its constants) shows up as a test failure instead of silently moving both sides of the CUDA parity tests.
Run in the authoring container:  python tests/golden/make_golden_harris_brisk2.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import oracle  # noqa: E402
from okvis2_b200.synth import synth_frame  # noqa: E402

EUROC0 = [458.654880721, 457.296696463, 367.215803962, 248.37534061, -0.28340811217, 0.0739590738929, 0.000193595028569, 1.76187114545e-05]
CASES = [("plain_752", 61, 752, 480, 38.0, 150, 700, False), ("aware_752", 62, 752, 480, 12.0, 40, 500, True),
         ("plain_341", 63, 341, 255, 8.0, 20, 0, False)]


def main():
    out = {}
    for name, seed, W, H, radius, thr, max_kp, aware in CASES:
        img = synth_frame(seed, W, H)
        o = oracle.HarrisBrisk2(radius, thr, max_kp)
        args = ()
        if aware:
            rays, jac = oracle.camera_awareness_maps(1, EUROC0, W, H)
            d = np.array([0.05, 0.99, -0.1], np.float32); d /= np.linalg.norm(d)
            args = (rays, jac, float(np.float32(EUROC0[0])), d)
            out[name + "_dir"] = d
        kp, desc = o.detect_and_compute(img, *args)
        sc = o.scores(img)
        out[name + "_cfg"] = np.array([seed, W, H, radius, thr, max_kp, int(aware)], np.float64)
        out[name + "_kp"] = kp.view(np.uint8).reshape(len(kp), 28)
        out[name + "_desc"] = desc
        out[name + "_score_sum"] = np.array([int(sc.astype(np.int64).sum()), int(np.abs(sc.astype(np.int64)).sum()), len(o.maxima(sc))], np.int64)
    np.savez_compressed(os.path.join(HERE, "harris_brisk2_oracle.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
