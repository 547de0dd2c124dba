"""CPU suite: pins the oracle (oracle/brisk_oracle.c) against the committed OpenCV-4.13 golden vectors
(tests/golden/make_golden.py) and, when cv2 is importable, against cv2 directly."""
import numpy as np
import pytest

import oracle
from conftest import assert_same_features, kp_struct
from okvis2_b200.synth import synth_frame

CASES = [("real752", 30, 0), ("real752", 30, 3), ("real752", 60, 2), ("real341", 25, 2)]


@pytest.mark.parametrize("name,thr,octv", CASES)
def test_oracle_matches_cv2_golden_real(golden, name, thr, octv):
    img = golden[f"{name}_img"]
    kp, desc = oracle.Brisk(thr, octv).detect_and_compute(img)
    assert_same_features(kp, desc, kp_struct(golden[f"{name}_t{thr}_o{octv}_kp"]), golden[f"{name}_t{thr}_o{octv}_desc"], name)


@pytest.mark.parametrize("seed,W,H,thr,octv", [(1000, 752, 480, 30, 3), (1001, 752, 480, 30, 0), (2000, 1024, 1024, 30, 3),
                                               (3000, 720, 540, 30, 3)])
def test_oracle_matches_cv2_golden_synth(golden, seed, W, H, thr, octv):
    img = synth_frame(seed, W, H)
    crc = golden[f"synth{seed}_{W}x{H}_crc"]
    assert int(img.astype(np.uint64).sum()) == int(crc[0]), "synthetic generator is not reproducible on this machine"
    kp, desc = oracle.Brisk(thr, octv).detect_and_compute(img)
    key = f"synth{seed}_{W}x{H}_t{thr}_o{octv}"
    assert_same_features(kp, desc, kp_struct(golden[key + "_kp"]), golden[key + "_desc"], key)


def test_pyramid_layers_match_cv2_resize(golden):
    img = golden["real752_img"]
    b = oracle.Brisk(30, 2)
    b.detect_raw(img)
    layers = b.layers()
    assert [l[0].shape for l in layers] == [(480, 752), (320, 500), (240, 376), (160, 250)]
    for i in (1, 2, 3):
        assert np.array_equal(layers[i][0], golden[f"real752_layer{i}"]), f"layer {i}"
    assert [l[2] for l in layers] == [1.0, 1.5, 2.0, 3.0]


def test_agast_scores_match_cv2(golden):
    img = golden["real752_img"]
    L = oracle.lib()
    for name, fn, margin in [("oast916", L.okvo_oast916_bstar, 3), ("agast58", L.okvo_agast58_bstar, 1)]:
        ref = golden[f"real752_{name}_t20"]
        got = {}
        H, W = img.shape
        for y in range(margin, H - margin):
            row = img[y]
            for x in range(margin, W - margin):
                pass
        # dense evaluation through the C function on the golden positions + a raster sample of non-corners
        base = img.ctypes.data
        for x, y, s in ref:
            assert fn(base + int(y) * W + int(x), W) == s
        refset = {(int(x), int(y)) for x, y, _ in ref}
        rng = np.random.default_rng(0)
        for _ in range(20000):
            x, y = int(rng.integers(margin, W - margin)), int(rng.integers(margin, H - margin))
            if (x, y) not in refset:
                assert fn(base + y * W + x, W) < 20


def test_cap_keeps_strongest_in_detection_order():
    img = synth_frame(7, 752, 480)
    b = oracle.Brisk(30, 3)
    raw = b.detect_raw(img)
    assert len(raw) > 1000
    kp = raw.copy()
    n = oracle.lib().okvo_cap_strongest(kp.ctypes.data, len(kp), 500)
    kp = kp[:n]
    assert n == 500
    thr = np.sort(raw["response"])[-500]
    assert (kp["response"] >= thr).all()
    # order preserved: (octave, y, x) keys of a subsequence of raw
    pos = [np.nonzero((raw["x"] == k["x"]) & (raw["y"] == k["y"]) & (raw["octave"] == k["octave"]))[0][0] for k in kp]
    assert pos == sorted(pos)


def test_empty_and_flat_images():
    b = oracle.Brisk(30, 3)
    for img in (np.zeros((480, 752), np.uint8), np.full((100, 120), 200, np.uint8)):
        kp, d = b.detect_and_compute(img)
        assert len(kp) == 0 and d.shape == (0, 64)


def test_oracle_vs_cv2_live():
    cv2 = pytest.importorskip("cv2")
    img = synth_frame(4242, 640, 400)
    kps, desc = cv2.BRISK_create(25, 3, 1.0).detectAndCompute(img, None)
    kp, d = oracle.Brisk(25, 3).detect_and_compute(img)
    assert_same_features(kp, d, oracle.cv_keypoints_to_array(kps), desc, "live cv2")
