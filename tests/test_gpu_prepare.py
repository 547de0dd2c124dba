"""GPU suite: P1 (landmark-candidate preparation, Frontend.cpp:1196-1360) through the C ABI against the oracle
transcription. Bit-exact: kept landmarks, projections, is3d, pool rows, e_W / r_W, observation ids -- and the pool feeds
the device M1 matcher without a round trip."""
import ctypes as C

import numpy as np
import pytest

import oracle
from okvis2_b200 import lib as L
from okvis2_b200.frontend import Frontend
from okvis2_b200.synth import landmark_scene

pytestmark = pytest.mark.gpu

MODELS = {0: "none", 1: "radialtangential", 2: "equidistant"}


def make(s, model, cams=2):
    fe = Frontend(cams, s["W"], s["H"])
    for c in range(cams):
        k = s["intr"][4:] if model == 1 else ([-0.01, 0.003, -0.002, 0.0004] if model == 2 else [0, 0, 0, 0])
        fe.setCameraModel(c, MODELS[model], (s["intr"][0], s["intr"][1]), (s["intr"][2], s["intr"][3]), list(k))
    fe.configureFeatureStore(s["n_slots"], s["D"])
    for slot in range(s["n_slots"]):
        for c in range(s["n_cams"]):
            t = slot * s["n_cams"] + c
            fe.storeFrame(slot, c, s["desc_tab"][t], s["ray_tab"][t])
    return fe


def compare(got, ref, bit_exact_proj=True):
    assert np.array_equal(got["lm"], ref["lm"])
    assert np.array_equal(got["lm_is3d"], ref["lm_is3d"])
    assert np.array_equal(got["desc_begin"], ref["desc_begin"])
    assert np.array_equal(got["cand_lm"], ref["cand_lm"])
    assert np.array_equal(got["kid"], ref["kid"])
    assert np.array_equal(got["cand_desc"], ref["cand_desc"])
    for k in ("p_W", "e_W", "r_W") + (("lm_proj",) if bit_exact_proj else ()):
        assert np.array_equal(got[k].view(np.uint64), ref[k].view(np.uint64)), k
    if not bit_exact_proj:   # equidistant model: device atan is ulp-close to libm (as for D4)
        assert np.allclose(got["lm_proj"], ref["lm_proj"], rtol=0, atol=1e-9)


@pytest.mark.parametrize("model,D,exclusive,thr", [(1, 64, False, 20.0), (0, 48, False, 150.0), (1, 64, True, 20.0), (2, 64, False, 20.0)])
def test_prepare_equals_oracle(model, D, exclusive, thr):
    s = landmark_scene(11 + model, n_lm=3000, D=D)
    intr = s["intr"].copy()
    if model == 2:
        intr[4:] = [-0.01, 0.003, -0.002, 0.0004]
    if model == 0:
        intr[4:] = 0
    fe = make(s, model)
    try:
        got = fe.prepareLandmarksToMatch(0, s["T_WC1"], s["T_CW1"], s["W"], s["H"], s["hp_W"], s["quality"], s["obs_begin"], s["obs"],
                                         s["T_WC_old"], reprThreshold=thr, exclusive=exclusive)
        ref = oracle.prepare_landmarks(s["hp_W"], s["quality"], s["obs_begin"], s["obs"], s["n_cams"], s["T_WC_old"], s["desc_tab"],
                                       s["ray_tab"], D, s["T_WC1"], s["T_CW1"], model, intr, s["W"], s["H"], thr, exclusive)
        assert len(ref["lm"]) > 300
        compare(got, ref, bit_exact_proj=(model != 2))
    finally:
        fe.close()


def test_empty_and_degenerate_inputs():
    s = landmark_scene(5, n_lm=50)
    fe = make(s, 1)
    try:
        r = fe.prepareLandmarksToMatch(0, s["T_WC1"], s["T_CW1"], s["W"], s["H"], np.zeros((0, 4)), np.zeros(0), np.zeros(1, np.int32),
                                       np.zeros((0, 3), np.int32), s["T_WC_old"])
        assert len(r["lm"]) == 0 and len(r["cand_desc"]) == 0
        # landmarks without observations never survive
        r = fe.prepareLandmarksToMatch(0, s["T_WC1"], s["T_CW1"], s["W"], s["H"], s["hp_W"], s["quality"], np.zeros(51, np.int32),
                                       np.zeros((0, 3), np.int32), s["T_WC_old"])
        assert len(r["lm"]) == 0
        # an observation outside the store is an argument error, not a device fault
        bad = s["obs"].copy(); bad[0, 2] = 10 ** 6
        with pytest.raises(L.OkbError):
            fe.prepareLandmarksToMatch(0, s["T_WC1"], s["T_CW1"], s["W"], s["H"], s["hp_W"], s["quality"], s["obs_begin"], bad, s["T_WC_old"])
    finally:
        fe.close()


def test_prepared_pool_feeds_device_m1():
    """store the features of real detections, prepare the pool, match the current frame against it on the device (M1)
    straight from the device-resident result, and compare with oracle P1 -> oracle M1."""
    import torch
    from okvis2_b200.frontend import MultiFrame
    from okvis2_b200.synth import synth_stereo
    W, H = 752, 480
    fe = Frontend(2, W, H)
    try:
        fe.configure(threshold=30, octaves=3, max_keypoints=600)
        intr = np.array([458.0, 457.0, W / 2 - 8.8, H / 2 + 8.4, -0.2834, 0.0740, 0.00019, 1.76e-05])
        for c in range(2):
            fe.setCameraModel(c, "radialtangential", (intr[0], intr[1]), (intr[2], intr[3]), list(intr[4:]))
        n_slots = 4
        fe.configureFeatureStore(n_slots, 64)
        feats = {}
        for slot in range(n_slots):
            left, right = synth_stereo(20 + slot, W, H)
            mf = MultiFrame(2)
            for c, img in enumerate((left, right)):
                mf.setImage(c, img); fe.detectAndDescribe(c, mf)
                fe.computeBackProjections(mf, c)
                fe.storeLastFrame(slot, c)      # device -> device
                feats[(slot, c)] = (mf.frames[c].descriptors.copy(), mf.frames[c].backProjections.copy())
        s = landmark_scene(9, n_lm=1500, n_slots=n_slots, n_kp=min(len(v[0]) for v in feats.values()))
        desc_tab = [feats[(t // 2, t % 2)][0] for t in range(2 * n_slots)]
        ray_tab = [feats[(t // 2, t % 2)][1] for t in range(2 * n_slots)]
        got = fe.prepareLandmarksToMatch(0, s["T_WC1"], s["T_CW1"], W, H, s["hp_W"], s["quality"], s["obs_begin"], s["obs"], s["T_WC_old"])
        ref = oracle.prepare_landmarks(s["hp_W"], s["quality"], s["obs_begin"], s["obs"], 2, s["T_WC_old"], desc_tab, ray_tab, 64,
                                       s["T_WC1"], s["T_CW1"], 1, intr, W, H)
        compare(got, ref)
        # current frame = a new detection of camera 0; M1 on the device from the prepared pool
        left, _ = synth_stereo(20, W, H)
        mf = MultiFrame(2); mf.setImage(0, left); fe.detectAndDescribe(0, mf)
        fr = mf.frames[0]
        lib = L.lib()
        p = [C.c_void_p() for _ in range(4)]; nc = C.c_int32(); nl = C.c_int32()
        L.check(lib.okb_prepared_device(fe.ctx, *[C.byref(x) for x in p], C.byref(nc), C.byref(nl)))
        assert nc.value == len(ref["cand_desc"]) and nl.value == len(ref["lm"])
        cap = C.c_int(0)
        lib.okb_device_features(fe.ctx, 0, None, None, None, C.byref(cap))
        d_dist = torch.zeros(cap.value, dtype=torch.int32, device="cuda"); d_lm = torch.zeros(cap.value, dtype=torch.int32, device="cuda")
        L.check(lib.okb_match_map3d_device(fe.ctx, 0, 64, 1, nc.value, p[0], p[1], nl.value, p[2], p[3], 20.0, 60, d_dist.data_ptr(), d_lm.data_ptr()))
        L.check(lib.okb_sync(fe.ctx))
        # the same through the mirror's one-call form
        mdist, midx, _ = fe.matchToMap(0, s["T_WC1"], s["T_CW1"], W, H, s["hp_W"], s["quality"], s["obs_begin"], s["obs"], s["T_WC_old"])
        n = len(fr.keypoints)
        xy = np.stack([fr.keypoints["x"], fr.keypoints["y"]], 1).astype(np.float64)
        rdist, rlm = oracle.match_map3d(fr.descriptors, xy, None, ref["cand_desc"], ref["cand_lm"], ref["lm_proj"], ref["lm_is3d"], 20.0, 60)
        assert np.array_equal(d_dist.cpu().numpy()[:n].astype(np.uint32), rdist.astype(np.uint32))
        assert np.array_equal(d_lm.cpu().numpy()[:n], rlm)
        assert np.array_equal(mdist[:n], rdist.astype(np.uint32))
        assert np.array_equal(midx[:n], np.where(rlm >= 0, ref["lm"][np.clip(rlm, 0, None)], -1))
    finally:
        fe.close()
