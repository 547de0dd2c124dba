"""GPU parity of the device-resident M3 sequence (okb_match_motion_stereo_device_ptr: Frontend::matchMotionStereo over the
older keyframes, Frontend.cpp:1775-1958) against the oracle transcription, batched over several current frames."""
import ctypes as C

import numpy as np
import pytest

import oracle
from okvis2_b200 import lib as okl
from okvis2_b200.frontend import Frontend
from okvis2_b200.synth import motion_scene
from test_oracle_motion_sequence import kp_of, oracle_views

pytestmark = pytest.mark.gpu


def run_device(fe, scenes, cap0, cap1, n_older):
    import torch
    L_ = okl.lib()
    B = len(scenes)
    intr = scenes[0]["intr"]
    m = okl.CameraModel(); m.model = 1; m.fu, m.fv, m.cu, m.cv = intr[:4]
    for i in range(4):
        m.k[i] = intr[4 + i]
    kp = np.zeros((B, cap1), okl.KP_DTYPE); desc = np.zeros((B, cap1, 64), np.uint8); cnt = np.zeros(B, np.int32)
    matched = np.zeros((B, cap1), np.uint8)
    Tw1 = np.zeros((B, 12)); Tc1 = np.zeros((B, 12))
    keep = []
    views = (okl.OlderView * (B * n_older))()
    for b, s in enumerate(scenes):
        cur = s["cur"]; n1 = len(cur["desc"])
        kp[b, :n1] = kp_of(cur["xy"], cur["size"]); desc[b, :n1] = cur["desc"]; cnt[b] = n1; matched[b, :n1] = cur["matched"]
        Tw1[b] = s["T_WC1"]; Tc1[b] = s["T_CW1"]
        for v, ov in enumerate(oracle_views(s)[:n_older]):
            t = [torch.from_numpy(np.ascontiguousarray(ov[k])).cuda() for k in ("desc", "rays", "valid", "size")]
            tu = torch.from_numpy(np.ascontiguousarray(ov["use"])).cuda() if ov["use"] is not None else None
            keep += t + [tu]
            e = views[b * n_older + v]
            e.d_desc, e.d_rays, e.d_valid, e.d_size = (x.data_ptr() if len(x) else None for x in t)
            e.d_use = tu.data_ptr() if tu is not None and len(tu) else None
            e.n = len(ov["desc"])
            e.T_WC[:] = list(ov["T_WC"]); e.T_CW[:] = list(ov["T_CW"])
    d_kp = torch.from_numpy(kp.view(np.uint8).reshape(B, cap1 * 28)).cuda(); d_desc = torch.from_numpy(desc).cuda()
    d_cnt = torch.from_numpy(cnt).cuda(); d_m = torch.from_numpy(matched).cuda()
    n = B * n_older * cap0
    k1 = torch.zeros(n, dtype=torch.int32, device="cuda"); dist = torch.zeros(n, dtype=torch.int32, device="cuda")
    hp = torch.zeros(n * 4, dtype=torch.float64, device="cuda"); fl = torch.zeros(n, dtype=torch.uint8, device="cuda")
    okl.check(L_.okb_match_motion_stereo_device_ptr(fe.ctx, B, cap1, d_kp.data_ptr(), d_desc.data_ptr(), d_cnt.data_ptr(), C.addressof(m),
                                                    scenes[0]["W"], scenes[0]["H"], Tw1.ctypes.data, Tc1.ctypes.data, n_older, views, cap0, 60, None,
                                                    d_m.data_ptr(), k1.data_ptr(), dist.data_ptr(), hp.data_ptr(), fl.data_ptr()))
    okl.check(L_.okb_sync(fe.ctx)); torch.cuda.synchronize()
    sh = (B, n_older, cap0)
    return (k1.cpu().numpy().reshape(sh), dist.cpu().numpy().view(np.uint32).reshape(sh), hp.cpu().numpy().reshape(sh + (4,)),
            fl.cpu().numpy().reshape(sh), d_m.cpu().numpy())


@pytest.mark.parametrize("fused,mma", [(1, 2), (0, 2), (1, 1), (0, 1), (1, 0), (0, 0)])
def test_motion_stereo_sequence_equals_oracle(fused, mma):
    """both forms of the per-view step (one launch per view, the default; separate kernels) x both forms of the
    Hamming scan (2 tcgen05 + TMEM, the default; 1 legacy integer MMA; 0 POPC)"""
    fe = Frontend(0)
    okl.lib().okb_m3_set_fused(fused)
    okl.lib().okb_scan_set_mma(mma)
    try:
        scenes = [motion_scene(21, n_views=5, n0=500, n1=700), motion_scene(22, n_views=5, n0=640, n1=520, premated=0.5),
                  motion_scene(23, n_views=5, n0=100, n1=64, premated=0.0)]
        # an empty older view and a frame without eligible keypoints
        scenes[1]["views"][2] = dict(xy=np.zeros((0, 2), np.float32), desc=np.zeros((0, 64), np.uint8), size=np.zeros(0, np.float32),
                                     use=np.zeros(0, np.uint8), T_WC=scenes[1]["views"][2]["T_WC"], T_CW=scenes[1]["views"][2]["T_CW"])
        scenes[2]["views"][0]["use"][:] = 0
        # near-duplicate descriptors in frame 1 (every pair is below the threshold): the hit list of the Hamming scan overflows and
        # the sequential-replay kernel redoes that frame
        rng = np.random.default_rng(3)
        base = rng.integers(0, 256, 64, dtype=np.uint8)

        def near(n):
            d = np.tile(base, (n, 1))
            for i in range(n):
                for b in rng.choice(512, 6, replace=False):
                    d[i, b // 8] ^= 1 << (b % 8)
            return d
        scenes[1]["cur"]["desc"] = near(len(scenes[1]["cur"]["desc"]))
        for v in (0, 1):
            scenes[1]["views"][v]["desc"] = near(len(scenes[1]["views"][v]["desc"]))
        cap0, cap1, n_older = 640, 704, 5
        k1, dist, hp, fl, m1 = run_device(fe, scenes, cap0, cap1, n_older)
        inserted = 0
        for b, s in enumerate(scenes):
            intr = s["intr"]; cur = s["cur"]
            rays1, valid1 = oracle.back_project(1, intr[0], intr[1], intr[2], intr[3], list(intr[4:8]), kp_of(cur["xy"], cur["size"]))
            ref, rm = oracle.match_motion_stereo_sequence(oracle_views(s), cur["desc"], rays1, valid1, cur["xy"], s["T_WC1"], s["T_CW1"], 1,
                                                          intr, s["W"], s["H"], 60, cur["matched"])
            for v, (rk1, rdist, rhp, rfl) in enumerate(ref):
                n = len(rk1)
                assert np.array_equal(k1[b, v, :n], rk1), (b, v)
                assert np.array_equal(dist[b, v, :n], rdist), (b, v)
                assert np.array_equal(hp[b, v, :n].view(np.uint64), rhp.view(np.uint64)), (b, v)
                assert np.array_equal(fl[b, v, :n], rfl), (b, v)
                assert (k1[b, v, n:] == -1).all() and (fl[b, v, n:] == 0).all()
                inserted += int(((rfl & 4) != 0).sum())
            assert np.array_equal(m1[b, :len(rm)], rm)
        assert inserted > 150
    finally:
        okl.lib().okb_m3_set_fused(-1)
        okl.lib().okb_scan_set_mma(2)
        fe.close()


@pytest.mark.parametrize("mma", [2, 1])
def test_motion_stereo_sequence_batch_of_nine(mma):
    """a batch large enough for the batched forms of the scan (tcgen05: one CTA per candidate tile streams the eligible keypoints of
    all views; more than one candidate tile; several query tiles per view)"""
    fe = Frontend(0)
    okl.lib().okb_scan_set_mma(mma)
    try:
        scenes = [motion_scene(40 + i, n_views=4, n0=260 + 37 * (i % 3), n1=300 + 29 * (i % 4), premated=0.3 if i % 2 else 0.0) for i in range(9)]
        cap0, cap1, n_older = 384, 448, 4
        k1, dist, hp, fl, m1 = run_device(fe, scenes, cap0, cap1, n_older)
        inserted = 0
        for b, s in enumerate(scenes):
            intr = s["intr"]; cur = s["cur"]
            rays1, valid1 = oracle.back_project(1, intr[0], intr[1], intr[2], intr[3], list(intr[4:8]), kp_of(cur["xy"], cur["size"]))
            ref, rm = oracle.match_motion_stereo_sequence(oracle_views(s), cur["desc"], rays1, valid1, cur["xy"], s["T_WC1"], s["T_CW1"], 1,
                                                          intr, s["W"], s["H"], 60, cur["matched"])
            for v, (rk1, rdist, rhp, rfl) in enumerate(ref):
                n = len(rk1)
                assert np.array_equal(k1[b, v, :n], rk1), (b, v)
                assert np.array_equal(dist[b, v, :n], rdist), (b, v)
                assert np.array_equal(hp[b, v, :n].view(np.uint64), rhp.view(np.uint64)), (b, v)
                assert np.array_equal(fl[b, v, :n], rfl), (b, v)
                inserted += int(((rfl & 4) != 0).sum())
            assert np.array_equal(m1[b, :len(rm)], rm)
        assert inserted > 300
    finally:
        okl.lib().okb_scan_set_mma(2)
        fe.close()
