"""GPU suite: the five matchers through the C ABI against the oracle restatement of the reference loops. Bit-exact
(indices, distances, triangulated points, flags)."""
import numpy as np
import pytest

import oracle
from okvis2_b200.frontend import Frontend
from okvis2_b200.synth import map_scene, stereo_scene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fe():
    f = Frontend(0)
    yield f
    f.close()


def eq(a, b, what):
    for i, (x, y) in enumerate(zip(a, b)):
        if isinstance(x, np.ndarray):
            if x.dtype == np.float64:
                assert np.array_equal(x.view(np.uint64), y.view(np.uint64)), f"{what}[{i}]: {np.nonzero((x != y).reshape(len(x), -1).any(1))[0][:5]}"
            else:
                assert np.array_equal(x, y), f"{what}[{i}]: {np.nonzero(x != y)[0][:5]}"
        else:
            assert x == y, what


def test_hamming_matrix_real_descriptors(fe, voc_desc):
    h = fe.hammingMatrix(voc_desc, voc_desc)
    assert np.array_equal(h, oracle.hamming_matrix(voc_desc, voc_desc))
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (100, 64), dtype=np.uint8)
    assert np.array_equal(fe.hammingMatrix(a, a[:37]), oracle.hamming_matrix(a, a[:37]))


@pytest.mark.parametrize("D,n_kp,n_lm,use_imu", [(64, 1000, 3000, True), (48, 700, 2500, True), (64, 333, 1000, False), (48, 5, 1, True)])
def test_m1_match_to_map(fe, D, n_kp, n_lm, use_imu):
    rng = np.random.default_rng(D + n_kp)
    kp_xy = rng.uniform(0, 752, (n_kp, 2)); kp_xy[:, 1] *= 480 / 752
    kd = rng.integers(0, 256, (n_kp, D), dtype=np.uint8)
    use = (rng.random(n_kp) > 0.1).astype(np.uint8)
    m = map_scene(n_lm, kp_xy, kd, n_lm, W=752, H=480, frac_near=0.3)
    got = fe.matchToMapByThread(kd, kp_xy, use, m["cand_desc"], m["cand_lm"], m["lm_proj"], m["lm_is3d"], use_imu)
    ref = oracle.match_map3d(kd, kp_xy, use, m["cand_desc"], m["cand_lm"], m["lm_proj"], m["lm_is3d"], 20.0 if use_imu else 150.0, 60)
    eq(got, ref, "M1")
    if n_lm > 100:
        assert (ref[1] >= 0).sum() > 10 and (ref[1][use == 0] == -1).all()


def test_m1_at_tumvi_size(fe):
    """BASELINE config 2 at full size: 2000 keypoints against the pool of a 50 000-landmark map (~100 000 rows), both
    reprojection thresholds, exact equality with the oracle; plus the size-independent property that dropping every
    pool row outside the gate radius of all keypoints changes nothing."""
    rng = np.random.default_rng(7)
    n_kp = 2000
    kp_xy = rng.uniform(0, 1024, (n_kp, 2))
    kd = rng.integers(0, 256, (n_kp, 64), dtype=np.uint8)
    m = map_scene(3, kp_xy, kd, 50000, W=1024, H=1024, frac_near=0.1)
    assert len(m["cand_lm"]) > 90000
    for use_imu in (True, False):
        thr = 20.0 if use_imu else 150.0
        got = fe.matchToMapByThread(kd, kp_xy, None, m["cand_desc"], m["cand_lm"], m["lm_proj"], m["lm_is3d"], use_imu)
        ref = oracle.match_map3d(kd, kp_xy, None, m["cand_desc"], m["cand_lm"], m["lm_proj"], m["lm_is3d"], thr, 60)
        eq(got, ref, "M1 full size")
        assert (ref[1] >= 0).sum() > 500
    # rows whose landmark projects farther than 20 px from every keypoint cannot matter
    from scipy.spatial import cKDTree
    near = np.array([len(x) > 0 for x in cKDTree(kp_xy).query_ball_point(np.nan_to_num(m["lm_proj"], nan=-1e9, posinf=1e9, neginf=-1e9), 20.0 + 1e-6)])
    keep = near[m["cand_lm"]] | ~np.isfinite(m["lm_proj"][m["cand_lm"]]).all(1)
    got2 = fe.matchToMapByThread(kd, kp_xy, None, m["cand_desc"][keep], m["cand_lm"][keep], m["lm_proj"], m["lm_is3d"], True)
    ref = oracle.match_map3d(kd, kp_xy, None, m["cand_desc"], m["cand_lm"], m["lm_proj"], m["lm_is3d"], 20.0, 60)
    eq(got2, ref, "M1 pruned pool")


def test_m1_ties_first_in_order_wins(fe):
    # identical descriptors everywhere: every candidate has distance 0 -> the lowest landmark passing the gate must win
    kd = np.zeros((40, 64), np.uint8)
    kp_xy = np.stack([np.arange(40) * 10.0, np.zeros(40)], 1)
    cand_lm = np.repeat(np.arange(300, dtype=np.int32), 3)
    cand = np.zeros((900, 64), np.uint8)
    lm_proj = np.stack([(np.arange(300) % 50) * 8.0, np.zeros(300)], 1)
    is3d = (np.arange(300) % 3 != 0).astype(np.uint8)
    got = fe.matchToMapByThread(kd, kp_xy, None, cand, cand_lm, lm_proj, is3d, True)
    ref = oracle.match_map3d(kd, kp_xy, None, cand, cand_lm, lm_proj, is3d, 20.0, 60)
    eq(got, ref, "M1 ties")
    assert (ref[0] == 0).all()


@pytest.mark.parametrize("D,loop", [(64, False), (48, False), (64, True)])
def test_m2_match_to_map_uninitialised(fe, D, loop):
    s = stereo_scene(21, 800, 10, D=D)
    rng = np.random.default_rng(5)
    kd, ke = s["desc0"], s["e0_W"]
    m = map_scene(17, np.zeros((len(kd), 2)), kd, 2500, frac_3d=0.4, flip_p=0.05)
    # give the copied candidates geometry that triangulates with their source keypoint: ray from another centre
    P = ke[m["src"][m["cand_lm"]]] * rng.uniform(0.1, 30.0, (len(m["cand_lm"]), 1)) + np.array([0.05, 0, 0])
    r = rng.normal(0, 0.4, P.shape)
    e = P - r; e /= np.linalg.norm(e, axis=1, keepdims=True)
    cp = m["is_copy"][m["cand_lm"]] & (rng.random(len(P)) < 0.8)
    m["cand_e_W"][cp] = e[cp]; m["cand_r_W"][cp] = r[cp]
    use = (rng.random(len(kd)) > 0.05).astype(np.uint8)
    prev = None
    if loop:
        prev = np.where(rng.random(len(kd)) < 0.5, m["src"].argsort()[np.searchsorted(np.sort(m["src"]), np.arange(len(kd))).clip(0, len(m["src"]) - 1)], -1).astype(np.int32)
    r1 = np.array([0.05, 0.0, 0.0])
    got = fe.matchToMapByThreadUnitialised(kd, ke, use, m["cand_desc"], m["cand_lm"], m["cand_e_W"], m["cand_r_W"], m["lm_is3d"],
                                           r1, 458.0, prev)
    ref = oracle.match_map_uninit(kd, ke, use, prev, m["cand_desc"], m["cand_lm"], m["cand_e_W"], m["cand_r_W"], m["lm_is3d"],
                                  r1, 1.0 / 458.0, 60)
    eq(got, ref, "M2")
    assert (ref[1] >= 0).sum() > 20 and (ref[2][:, 3] > 0).sum() > 5


@pytest.mark.parametrize("D,n0,n1", [(64, 1000, 1000), (48, 700, 650), (64, 2000, 1777), (64, 3, 0)])
def test_m3_m4_stereo(fe, D, n0, n1):
    s = stereo_scene(n0 + D, n0, max(n1, 1), D=D)
    if n1 == 0:
        for k in ("desc1", "e1_W", "sof1", "valid1"):
            s[k] = s[k][:0]
    got = fe.matchStereo(s["desc0"], s["valid0"], s["e0_W"], s["sof0"], s["desc1"], s["valid1"], s["e1_W"], s["sof1"],
                         s["r_WC0"], s["r_WC1"], s["T_CW0"], s["T_CW1"])
    ref = oracle.match_stereo(s["desc0"], s["valid0"], s["e0_W"], s["sof0"], s["desc1"], s["valid1"], s["e1_W"], s["sof1"],
                              s["r_WC0"], s["r_WC1"], s["T_CW0"], s["T_CW1"], 60)
    eq(got, ref, "M4")
    got = fe.matchMotionStereo(s["desc0"], s["valid0"], s["e0_W"], s["sof0"], s["desc1"], s["valid1"], s["e1_W"],
                               s["r_WC0"], s["r_WC1"], s["T_CW0"], s["T_CW1"])
    ref3 = oracle.match_motion_stereo(s["desc0"], s["valid0"], s["e0_W"], s["sof0"], s["desc1"], s["valid1"], s["e1_W"],
                                      s["r_WC0"], s["r_WC1"], s["T_CW0"], s["T_CW1"], 60)
    eq(got, ref3, "M3")
    if n1 > 100:
        assert (ref[0] >= 0).sum() > 50 and ref[3].sum() > 10 and (ref3[0] >= 0).sum() > 50


def test_m5_place_recognition(fe, voc_desc):
    rng = np.random.default_rng(1)
    kp = voc_desc[rng.permutation(len(voc_desc))[:500]].copy()
    bits = np.unpackbits(kp[:200], axis=1)
    noisy = np.packbits(bits ^ (rng.random(bits.shape) < 0.04).astype(np.uint8), axis=1)
    lm_desc = np.concatenate([noisy, voc_desc[500:560]])
    counts = rng.integers(0, 4, 120); counts[-1] = len(lm_desc) - counts[:-1].sum() if counts[:-1].sum() < len(lm_desc) else 0
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    offs = np.minimum(offs, len(lm_desc)).astype(np.int32)
    got = fe.verifyRecognisedPlaceMatch(offs, lm_desc, kp)
    ref = oracle.match_place(offs, lm_desc, kp, 60)
    eq(got, ref, "M5")
    assert (ref[0] >= 0).sum() > 30


def test_matcher_argument_errors(fe):
    from okvis2_b200.lib import OkbError
    with pytest.raises(OkbError):
        fe.hammingMatrix(np.zeros((2, 32), np.uint8), np.zeros((2, 32), np.uint8))  # D must be 48 or 64


def test_m1_degenerate_projections_follow_the_reference(fe):
    """NaN projections are not gated by `reprDist.dot(reprDist) > thr` (false for NaN), infinite ones always are."""
    rng = np.random.default_rng(4)
    kd = rng.integers(0, 256, (50, 64), dtype=np.uint8)
    kp_xy = rng.uniform(0, 700, (50, 2))
    cand = np.concatenate([kd[:10], kd[10:20], kd[20:30]])          # exact copies -> distance 0 where allowed
    cand_lm = np.arange(30, dtype=np.int32)
    proj = np.zeros((30, 2)); proj[:10] = np.nan; proj[10:20] = np.inf; proj[20:30] = kp_xy[20:30] + 3.0
    is3d = np.ones(30, np.uint8)
    got = fe.matchToMapByThread(kd, kp_xy, None, cand, cand_lm, proj, is3d, True)
    ref = oracle.match_map3d(kd, kp_xy, None, cand, cand_lm, proj, is3d, 20.0, 60)
    eq(got, ref, "M1 degenerate")
    assert (ref[1][:10] == np.arange(10)).all() and (ref[1][10:20] == -1).all() and (ref[1][20:30] == np.arange(20, 30)).all()


def test_matchers_from_two_host_threads(fe):
    import threading
    s = [stereo_scene(70 + i, 500, 480) for i in range(2)]
    out = [None, None]

    def work(i):
        x = s[i]
        for _ in range(5):
            out[i] = fe.matchStereo(x["desc0"], x["valid0"], x["e0_W"], x["sof0"], x["desc1"], x["valid1"], x["e1_W"], x["sof1"],
                                    x["r_WC0"], x["r_WC1"], x["T_CW0"], x["T_CW1"])
    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in th]; [t.join() for t in th]
    for i in range(2):
        x = s[i]
        ref = oracle.match_stereo(x["desc0"], x["valid0"], x["e0_W"], x["sof0"], x["desc1"], x["valid1"], x["e1_W"], x["sof1"],
                                  x["r_WC0"], x["r_WC1"], x["T_CW0"], x["T_CW1"], 60)
        eq(out[i], ref, f"thread {i}")


def test_pool_validation_and_descriptor_width_mismatch():
    """Bad pool indices and a pool whose descriptor width differs from the camera's are argument errors, not device faults."""
    import ctypes as C
    from okvis2_b200 import lib as okl
    from okvis2_b200.lib import OkbError
    fe = Frontend(1, 752, 480)
    try:
        rng = np.random.default_rng(1)
        d = rng.integers(0, 256, (10, 64), dtype=np.uint8); xy = rng.uniform(0, 400, (10, 2))
        cd = rng.integers(0, 256, (4, 64), dtype=np.uint8)
        proj = rng.uniform(0, 400, (2, 2)); is3d = np.ones(2, np.uint8)
        for bad in ([0, 1, 2, 1], [0, 0, 1, 0], [-1, 0, 1, 1]):
            with pytest.raises(OkbError) as e:
                fe.matchToMapByThread(d, xy, None, cd, np.array(bad, np.int32), proj, is3d)
            assert e.value.status == okl.OKB_ERR_ARGUMENT
        # D = 48 pool against a camera that describes with 64 bytes
        L_ = okl.lib()
        import torch
        z = torch.zeros(4096, dtype=torch.uint8, device="cuda")
        rc = L_.okb_match_map3d_device(fe.ctx, 0, 48, 1, 4, z.data_ptr(), z.data_ptr(), 2, z.data_ptr(), z.data_ptr(), 20.0, 60, z.data_ptr(), z.data_ptr())
        assert rc == okl.OKB_ERR_ARGUMENT and b"48 bytes" in L_.okb_last_error()
    finally:
        fe.close()
