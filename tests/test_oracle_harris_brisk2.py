"""CPU suite of the D = 48 mode (Harris + uniformity-enforcement detector, 48-byte BRISK2 extractor; reference call
sites okvis_frontend/src/Frontend.cpp:2406-2412, 232-251). PARITY UNPINNED vs smartroboticslab/brisk@1ef8b42a: these tests pin the
oracle's own definition (oracle/brisk_oracle.c section 6) by known answers computed independently in numpy, and check that the
parallel formulation the CUDA kernels implement (tests/emul, same per-element code) reproduces the sequential oracle bit for bit."""
import ctypes as C

import numpy as np
import pytest

import oracle
from okvis2_b200.synth import synth_frame

EUROC0 = dict(model=1, intr=[458.654880721, 457.296696463, 367.215803962, 248.37534061, -0.28340811217, 0.0739590738929,
                             0.000193595028569, 1.76187114545e-05])


def numpy_scores(img):
    """The Harris pipeline restated with array arithmetic (independent of the C loops)."""
    I = img.astype(np.int64)
    H, W = I.shape
    sx = np.zeros_like(I); sy = np.zeros_like(I)
    sx[1:-1, 1:-1] = 3 * (I[:-2, 2:] - I[:-2, :-2]) + 10 * (I[1:-1, 2:] - I[1:-1, :-2]) + 3 * (I[2:, 2:] - I[2:, :-2])
    sy[1:-1, 1:-1] = 3 * (I[2:, :-2] - I[:-2, :-2]) + 10 * (I[2:, 1:-1] - I[:-2, 1:-1]) + 3 * (I[2:, 2:] - I[:-2, 2:])
    gx, gy = sx >> 5, sy >> 5

    def smooth(p):
        out = np.zeros_like(p)
        w = [[1, 2, 1], [2, 4, 2], [1, 2, 1]]
        for j in range(3):
            for i in range(3):
                out[1:-1, 1:-1] += w[j][i] * p[j:H - 2 + j, i:W - 2 + i]
        return out >> 4
    a, b, c = smooth(gx * gx), smooth(gy * gy), smooth(gx * gy)
    s = a * b - c * c - (((a + b) * (a + b)) >> 4)
    out = np.zeros_like(s)
    out[2:-2, 2:-2] = s[2:-2, 2:-2]
    return out.astype(np.int32)


def test_scores_equal_numpy_restatement():
    for seed, W, H in [(3, 160, 120), (4, 201, 97)]:
        img = synth_frame(seed, W, H)
        assert np.array_equal(oracle.HarrisBrisk2.scores(img), numpy_scores(img))
    # a checkerboard corner scores high, a straight edge and a flat patch do not
    img = np.full((32, 32), 40, np.uint8); img[:16, :16] = 200; img[16:, 16:] = 200
    s = oracle.HarrisBrisk2.scores(img)
    assert s[15:17, 15:17].max() == s.max() and s.max() > 10000
    assert s[8, 14:18].max() <= 0 and s[4, 4] == 0


def test_maxima_rule():
    o = oracle.HarrisBrisk2(38.0, 10, 0)
    s = np.zeros((12, 20), np.int32)
    s[5, 5] = 50                      # isolated maximum
    s[7, 10:15] = 30                  # plateau of five equal values: the scan keeps every second one (10, 12, 14)
    s[3, 15] = 9                      # below the threshold
    s[9, 3] = 40; s[8, 4] = 41        # diagonal neighbour strictly greater
    s[1, 8] = 99                      # outside the [2, size - 2) window
    xy = o.maxima(s)
    assert sorted(map(tuple, xy)) == sorted([(5, 5), (10, 7), (12, 7), (14, 7), (4, 8)])


def test_uniformity_known_answers():
    o = oracle.HarrisBrisk2(20.0, 1, 0)
    lut = np.array([[oracle.lib().okvo_harris_lut(20.0, dx, dy) for dx in range(-15, 16)] for dy in range(-15, 16)], np.float32)
    assert lut[15, 15] == 1.0 and lut[15, 25] == 0.0 and abs(lut[15, 20] - 0.5) < 1e-7   # radius / 2 = 10 half-resolution cells
    # two blobs: the weaker one survives only outside the stronger one's suppression field
    def frame(dist, weak):
        img = np.full((120, 200), 30, np.uint8)
        img[40:60, 40:60] = 230
        img[40:60, 40 + dist:60 + dist] = np.maximum(img[40:60, 40 + dist:60 + dist], weak)
        return img
    kp = o.detect(frame(90, 120))
    assert len(kp) >= 8 and np.all(np.diff(kp["response"]) <= 0), "descending score order"
    full = oracle.HarrisBrisk2(20.0, 1, 0).detect(frame(90, 120))
    capped = oracle.HarrisBrisk2(20.0, 1, 3).detect(frame(90, 120))
    assert len(capped) == 3 and capped.tobytes() == full[:3].tobytes(), "the cap keeps the strongest accepted ones"
    # a larger radius suppresses more
    assert len(oracle.HarrisBrisk2(38.0, 1, 0).detect(synth_frame(5, 320, 240))) < len(oracle.HarrisBrisk2(12.0, 1, 0).detect(synth_frame(5, 320, 240)))


def test_brisk2_pattern_and_plain_descriptor():
    o = oracle.HarrisBrisk2()
    ns, nl = C.c_int(), C.c_int()
    oracle.lib().okvo_brisk_num_pairs(o.h, C.byref(ns), C.byref(nl))
    assert (ns.value, nl.value, o.D) == (383, 870, 48)
    assert oracle.lib().okvo_brisk2_basic_scale() == 17
    img = synth_frame(7, 400, 300)
    kp = o.detect(img)
    k2, d = o.compute(img, kp)
    assert 0 < len(k2) <= len(kp) and d.shape == (len(k2), 48)
    assert np.all(d[:, 47] < 128), "bit 383 is never set"
    # at one scale and without the camera maps the descriptor is plain BRISK: the first 383 comparisons are a subset of BRISK-512's
    assert np.all((k2["angle"] >= 0) & (k2["angle"] < 360))
    # rotation invariance of the plain mode: a 180-degree turn of the image turns the angle by 180 and keeps most bits
    img2 = np.ascontiguousarray(img[::-1, ::-1])
    kr = k2.copy(); kr["x"] = img.shape[1] - 1 - k2["x"]; kr["y"] = img.shape[0] - 1 - k2["y"]
    k3, d3 = o.compute(img2, kr)
    assert len(k3) >= len(k2) - 8


def test_brisk2_warp_identity_and_direction():
    L = oracle.lib()
    M = np.zeros(4, np.float32)

    def warp(e, J, d, fu):
        e = np.asarray(e, np.float32); J = np.asarray(J, np.float32); d = np.asarray(d, np.float32)
        ok = L.okvo_brisk2_warp(e.ctypes.data, J.ctypes.data, d.ctypes.data, C.c_float(fu), M.ctypes.data)
        return ok, M.copy()
    fu = 400.0
    J0 = [fu, 0, 0, 0, fu, 0]    # ideal pinhole at the principal point: d(u, v)/d(ray)
    ok, m = warp([0, 0, 1], J0, [0, 1, 0], fu)          # direction = image +y: identity
    assert ok and np.allclose(m, [1, 0, 0, 1], atol=1e-6)
    ok, m = warp([0, 0, 1], J0, [1, 0, 0], fu)          # direction = image +x: the pattern's y axis points right, x axis up
    assert ok and np.allclose(m, [0, 1, -1, 0], atol=1e-6)
    ok, m = warp([0, 0, 1], J0, [0, 0, -1], fu)         # looking along the direction: falls back to the image's +y
    assert ok and np.allclose(m, [1, 0, 0, 1], atol=1e-6)
    ok, _ = warp([0, 0, 0], J0, [0, 1, 0], fu)          # pixel without a ray
    assert not ok


def test_camera_aware_descriptor_near_plain_at_the_centre():
    """With an undistorted pinhole and the direction along image +y the warp is the identity near the principal point: the camera-aware
    samples coincide with the rotation-0 pattern."""
    W, H = 320, 240
    img = synth_frame(9, W, H)
    o = oracle.HarrisBrisk2(20.0, 50, 0)
    rays, jac = oracle.camera_awareness_maps(0, [300.0, 300.0, 160.0, 120.0, 0, 0, 0, 0], W, H)
    kp = o.detect(img)
    k2, d2 = o.compute(img, kp, rays, jac, 300.0, [0, 1, 0])
    assert len(k2) > 20
    near = (np.abs(k2["x"] - 160) < 12) & (np.abs(k2["y"] - 120) < 12)
    assert np.allclose(k2["angle"][near], 0.0, atol=0.5) or np.allclose(np.minimum(k2["angle"][near], 360 - k2["angle"][near]), 0.0, atol=0.5)


@pytest.fixture(scope="module")
def emul():
    from conftest import build_emul
    lib = C.CDLL(build_emul())
    lib.okb_emul_harris_brisk2.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_float,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]

    def run(img, radius, thr, max_kp, rays=None, jac=None, fu=0.0, direction=None, cap=1 << 14):
        img = np.ascontiguousarray(img)
        kp = np.zeros(cap, oracle.KP_DTYPE); d = np.zeros((cap, 48), np.uint8); st = np.zeros(3, np.int32)
        if rays is not None:
            rays = np.ascontiguousarray(rays, np.float32); jac = np.ascontiguousarray(jac, np.float32)
            direction = np.ascontiguousarray(direction, np.float32)
        n = lib.okb_emul_harris_brisk2(img.ctypes.data, img.shape[1], img.shape[0], radius, thr, max_kp,
                                       None if rays is None else rays.ctypes.data, None if rays is None else jac.ctypes.data, fu,
                                       None if rays is None else direction.ctypes.data, kp.ctypes.data, d.ctypes.data, cap, st.ctypes.data)
        assert 0 <= n <= cap
        return kp[:n], d[:n], st
    return run


@pytest.mark.parametrize("seed,W,H,radius,thr,max_kp,aware", [(21, 752, 480, 38.0, 150, 700, False), (22, 752, 480, 12.0, 20, 0, True),
                                                              (23, 341, 255, 20.0, 50, 120, True), (24, 640, 400, 6.0, 5, 0, False)])
def test_parallel_formulation_equals_oracle(emul, seed, W, H, radius, thr, max_kp, aware):
    img = synth_frame(seed, W, H)
    o = oracle.HarrisBrisk2(radius, thr, max_kp)
    args = ()
    if aware:
        intr = list(EUROC0["intr"]); intr[2] *= W / 752; intr[3] *= H / 480
        rays, jac = oracle.camera_awareness_maps(EUROC0["model"], intr, W, H)
        d = np.array([0.05, 0.99, -0.1], np.float32); d /= np.linalg.norm(d)
        args = (rays, jac, float(np.float32(intr[0])), d)
    rk, rd = o.detect_and_compute(img, *args)
    kp, d48, st = emul(img, radius, thr, max_kp, *args)
    assert st[0] > 100 and st[1] > 1, "the case must need more than one round"
    assert len(rk) == len(kp) and rk.tobytes() == kp.tobytes()
    assert np.array_equal(rd, d48)


def test_oracle_equals_its_frozen_vectors():
    """tests/golden/harris_brisk2_oracle.npz (made by tests/golden/make_golden_harris_brisk2.py) freezes the definition of DESIGN.md 2b:
    outputs of this repository's own restatement, NOT of smartroboticslab/brisk (parity unpinned)."""
    import os
    from conftest import ROOT
    g = np.load(os.path.join(ROOT, "tests", "golden", "harris_brisk2_oracle.npz"))
    names = sorted(k[:-4] for k in g.files if k.endswith("_cfg"))
    assert len(names) == 3
    for name in names:
        seed, W, H, radius, thr, max_kp, aware = g[name + "_cfg"]
        img = synth_frame(int(seed), int(W), int(H))
        o = oracle.HarrisBrisk2(float(radius), int(thr), int(max_kp))
        args = ()
        if aware:
            rays, jac = oracle.camera_awareness_maps(EUROC0["model"], EUROC0["intr"], int(W), int(H))
            args = (rays, jac, float(np.float32(EUROC0["intr"][0])), g[name + "_dir"])
        kp, desc = o.detect_and_compute(img, *args)
        assert kp.view(np.uint8).reshape(len(kp), 28).tobytes() == g[name + "_kp"].tobytes(), name
        assert np.array_equal(desc, g[name + "_desc"]), name
        sc = o.scores(img).astype(np.int64)
        assert [int(sc.sum()), int(np.abs(sc).sum()), len(o.maxima(sc.astype(np.int32)))] == list(g[name + "_score_sum"]), name
