"""GPU suite of the D = 48 mode: Harris + uniformity-enforcement detector and 48-byte BRISK2 extractor through the C ABI (k_harris_score,
k_harris_maxima, k_uniformity, k_describe48) against oracle/brisk_oracle.c section 6. Bit-exact for positions, responses and descriptor
rows; the reported angle (double atan2 on the device, rounded to float) within 1e-4 degrees. PARITY UNPINNED vs smartroboticslab/brisk."""
import ctypes as C

import numpy as np
import pytest

import oracle
from okvis2_b200 import lib as _l
from okvis2_b200.frontend import Frontend, MultiFrame
from okvis2_b200.lib import OkbError
from okvis2_b200.synth import map_scene, synth_frame

pytestmark = pytest.mark.gpu

EUROC = [dict(distortion_type="radialtangential", focal_length=(458.654880721, 457.296696463), principal_point=(367.215803962, 248.37534061),
              distortion_coefficients=[-0.28340811217, 0.0739590738929, 0.000193595028569, 1.76187114545e-05]),
         dict(distortion_type="equidistant", focal_length=(380.81, 380.81), principal_point=(510.29, 514.33),
              distortion_coefficients=[0.0103, -0.0046, 0.0024, -0.0008])]


def same48(kp, d, rk, rd, what=""):
    assert len(kp) == len(rk), f"{what}: {len(kp)} keypoints vs {len(rk)}"
    for f in ("x", "y", "size", "response", "octave", "class_id"):
        a, b = kp[f].view(np.uint32), rk[f].view(np.uint32)
        bad = np.nonzero(a != b)[0]
        assert len(bad) == 0, f"{what}: field {f} differs at {bad[:5]}: {kp[f][bad[:5]]} vs {rk[f][bad[:5]]}"
    da = np.abs(kp["angle"] - rk["angle"]); da = np.minimum(da, 360 - da)
    assert da.max(initial=0) <= 1e-4, f"{what}: angle differs by {da.max()}"
    assert d.shape == rd.shape == (len(rk), 48), what
    bad = np.nonzero((d != rd).any(1))[0]
    assert len(bad) == 0, f"{what}: {len(bad)} descriptor rows differ (first {bad[:5]})"


def make(W, H, radius, thr, max_kp, max_batch=1, n_cams=1):
    fe = Frontend(n_cams, W, H, max_batch=max_batch, descriptor_bytes=48)
    fe.configure(threshold=radius, absolute_threshold=thr, octaves=0, max_keypoints=max_kp)
    return fe


def run(fe, img, cam=0, T_WC=None):
    mf = MultiFrame(fe.numCameras)
    mf.setImage(cam, img)
    assert fe.detectAndDescribe(cam, mf, T_WC, None) is True
    fr = mf.frames[cam]
    assert fr.descriptors.flags["C_CONTIGUOUS"] and fr.descriptors.shape[1] == 48 and (fr.landmarkIds == 0).all()
    return fr


@pytest.mark.parametrize("seed,W,H,radius,thr,max_kp", [(31, 752, 480, 38.0, 150, 700), (32, 752, 480, 12.0, 20, 0), (33, 1024, 1024, 20.0, 100, 2000),
                                                        (34, 341, 255, 6.0, 5, 0), (35, 720, 540, 38.0, 150, 100)])
def test_plain_mode_equals_oracle(seed, W, H, radius, thr, max_kp):
    img = synth_frame(seed, W, H)
    fe = make(W, H, radius, thr, max_kp)
    fr = run(fe, img)
    rk, rd = oracle.HarrisBrisk2(radius, thr, max_kp).detect_and_compute(img)
    assert len(rk) > 20
    same48(fr.keypoints, fr.descriptors, rk, rd, f"seed {seed}")
    fe.close()


def test_real_image_equals_oracle(golden):
    img = golden["real752_img"]
    fe = make(752, 480, 38.0, 150, 700)
    fr = run(fe, img)
    rk, rd = oracle.HarrisBrisk2(38.0, 150, 700).detect_and_compute(img)
    same48(fr.keypoints, fr.descriptors, rk, rd, "real752")
    fe.close()


@pytest.mark.parametrize("cam_model,W,H", [(0, 752, 480), (1, 1024, 1024)])
def test_camera_aware_mode_equals_oracle(cam_model, W, H):
    """Frontend::detectAndDescribe with the camera-awareness maps and the gravity direction (Frontend.cpp:232-251)."""
    m = EUROC[cam_model]
    fe = make(W, H, 25.0, 80, 800)
    fe.setCameraModel(0, **m)
    rays, jac = fe.cameraAwarenessMaps(0)       # D5 on the device; the maps stay there for the extractor
    a = 0.3; Cx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    b = -0.2; Cz = np.array([[np.cos(b), -np.sin(b), 0], [np.sin(b), np.cos(b), 0], [0, 0, 1]])
    T_WC = np.eye(4); T_WC[:3, :3] = Cz @ Cx @ np.array([[1, 0, 0], [0, 0, 1], [0, -1, 0.]])   # camera looking roughly horizontally
    for seed in (41, 42):
        img = synth_frame(seed, W, H)
        fr = run(fe, img, 0, T_WC)
        d = fr.extractionDirection
        assert abs(np.linalg.norm(d) - 1) < 1e-5
        intr = [*m["focal_length"], *m["principal_point"], *m["distortion_coefficients"]]
        orays, ojac = oracle.camera_awareness_maps(Frontend.MODELS[m["distortion_type"]], intr, W, H)
        # the oracle is fed the DEVICE maps (D5's own parity, incl. the equidistant model's 1e-13 interior tolerance, is test_gpu_rig's)
        rk, rd = oracle.HarrisBrisk2(25.0, 80, 800).detect_and_compute(img, rays, jac, float(np.float32(m["focal_length"][0])), d)
        assert len(rk) > 50
        same48(fr.keypoints, fr.descriptors, rk, rd, f"aware seed {seed}")
        if cam_model == 0:   # radtan maps are bit-exact, so the oracle's own maps give the same features
            rk2, rd2 = oracle.HarrisBrisk2(25.0, 80, 800).detect_and_compute(img, orays, ojac, float(np.float32(m["focal_length"][0])), d)
            same48(fr.keypoints, fr.descriptors, rk2, rd2, f"aware/oracle maps seed {seed}")
        # the warp differs from the plain mode: the descriptors are not the plain ones
        pk, pd = oracle.HarrisBrisk2(25.0, 80, 800).detect_and_compute(img)
        assert len(pk) != len(rk) or not np.array_equal(pd, rd)
    fe.close()


def test_batch_and_m1_with_48_byte_rows():
    W, H, B = 752, 480, 6
    imgs = np.stack([synth_frame(50 + i, W, H) for i in range(B)])
    imgs[3] = 17                                   # a frame without a single corner
    fe = make(W, H, 30.0, 100, 500, max_batch=B)
    out = fe.detectAndDescribeBatch(0, imgs)
    o = oracle.HarrisBrisk2(30.0, 100, 500)
    for b in range(B):
        rk, rd = o.detect_and_compute(imgs[b])
        same48(out[b][0], out[b][1], rk, rd, f"batch frame {b}")
    assert len(out[3][0]) == 0
    # M1 (host-buffer form, D = 48) on the rows this mode produced
    L = _l.lib()
    kp, d = out[0]
    xy = np.stack([kp["x"], kp["y"]], 1).astype(np.float64)
    m = map_scene(5, xy, d, 1500, W=W, H=H, frac_near=0.3)
    assert m["cand_desc"].shape[1] == 48
    rdist, rlm = oracle.match_map3d(d, xy, None, m["cand_desc"], m["cand_lm"], m["lm_proj"], m["lm_is3d"], 20.0, 60)
    dist, lm = fe.matchToMapByThread(d, xy, None, m["cand_desc"], m["cand_lm"], m["lm_proj"], m["lm_is3d"])
    assert np.array_equal(dist, rdist) and np.array_equal(lm, rlm) and (lm >= 0).sum() > 20
    fe.close()


def test_rejections():
    with pytest.raises(OkbError) as e:
        fe = Frontend(1, 752, 480, descriptor_bytes=48)
        fe.configure(threshold=38.0, absolute_threshold=150, octaves=2, max_keypoints=100)
    assert e.value.status == _l.OKB_ERR_UNSUPPORTED
    with pytest.raises(OkbError) as e:
        fe = Frontend(1, 752, 480, descriptor_bytes=48)
        fe.configure(threshold=0.0, absolute_threshold=150, octaves=0, max_keypoints=100)
    assert e.value.status == _l.OKB_ERR_ARGUMENT
    with pytest.raises(OkbError) as e:
        Frontend(1, 752, 480, descriptor_bytes=32)
    assert e.value.status == _l.OKB_ERR_ARGUMENT


def test_device_resident_matchers_on_48_byte_rows():
    """A D = 48 camera keeps its rows a second time in 64-byte slots with a zero tail, so the device-resident forms built for 64-byte rows
    (the tensor-core Hamming scans of M4 and of the M3 sequence, the feature block of the camera-sharded exchange) run on them unchanged:
    M4 (okb_match_stereo_device), the export block and the M3 sequence (okb_match_motion_stereo_device) against the oracle on the 48-byte rows."""
    import torch
    from okvis2_b200 import sharding as sh
    from okvis2_b200.synth import pose12, rot, synth_stereo
    from test_gpu_camera import EUROC as EU, T_CW, oracle_bp, world_rays
    B, W, H = 3, 752, 480
    fe = make(W, H, 10.0, 30, 800, max_batch=B, n_cams=2)
    L_ = _l.lib()
    for c in range(2):
        fe.setCameraModel(c, **EU[c])
        fe.cameraAwarenessMaps(c)
    T_WC = np.eye(4); T_WC[:3, :3] = np.array([[1, 0, 0], [0, 0, 1], [0, -1, 0.]])
    for c in range(2):
        _l.check(L_.okb_set_extraction_direction(fe.ctx, c, np.ascontiguousarray(T_WC[:3, :3]).ctypes.data))
    imgs = [np.stack([synth_stereo(700 + t, W, H)[c] for t in range(B)]) for c in range(2)]
    d_imgs = [torch.from_numpy(a).cuda() for a in imgs]
    for c in range(2):
        _l.check(L_.okb_detect_describe_batch_device(fe.ctx, c, B, d_imgs[c].data_ptr()))
    cap = C.c_int(0); L_.okb_device_features(fe.ctx, 0, None, None, None, C.byref(cap)); cap = cap.value
    feats = [[None] * B for _ in range(2)]
    for c in range(2):
        for b in range(B):
            kp = np.zeros(cap, _l.KP_DTYPE); d = np.zeros((cap, 48), np.uint8); n = C.c_int(0)
            _l.check(L_.okb_fetch_features(fe.ctx, c, b, kp.ctypes.data, d.ctypes.data, cap, C.byref(n)))
            feats[c][b] = (kp[:n.value].copy(), d[:n.value].copy())
    assert min(len(f[0]) for fc in feats for f in fc) > 300
    # ---- M4 on the tensor-core scan
    C0 = np.eye(3); r0 = np.zeros(3); a = 0.01
    C1 = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]); r1 = np.array([0.11, 0.001, -0.002])
    k1 = torch.zeros((B, cap), dtype=torch.int32, device="cuda"); dist = torch.zeros((B, cap), dtype=torch.int32, device="cuda")
    hp = torch.zeros((B, cap, 4), dtype=torch.float64, device="cuda"); init = torch.zeros((B, cap), dtype=torch.uint8, device="cuda")
    _l.check(L_.okb_match_stereo_device(fe.ctx, 0, 1, B, C0.ctypes.data, r0.ctypes.data, np.ascontiguousarray(C1).ctypes.data, r1.ctypes.data, 60,
                                        k1.data_ptr(), dist.data_ptr(), hp.data_ptr(), init.data_ptr()))
    _l.check(L_.okb_sync(fe.ctx)); torch.cuda.synchronize()
    k1, dist, hp, init = k1.cpu().numpy(), dist.cpu().numpy().view(np.uint32), hp.cpu().numpy(), init.cpu().numpy()
    total = 0
    for b in range(B):
        (kp0, d0), (kp1, d1) = feats[0][b], feats[1][b]
        rays0, v0 = oracle_bp(EU[0], kp0); rays1, v1 = oracle_bp(EU[1], kp1)
        f0 = 0.5 * sum(EU[0]["focal_length"]); f1 = 0.5 * sum(EU[1]["focal_length"])
        ref = oracle.match_stereo(d0, v0, world_rays(C0, rays0), kp0["size"].astype(np.float64) / f0, d1, v1, world_rays(C1, rays1),
                                  kp1["size"].astype(np.float64) / f1, r0, r1, T_CW(C0, r0), T_CW(C1, r1), 60)
        n0 = len(kp0)
        assert np.array_equal(k1[b, :n0], ref[0]) and np.array_equal(dist[b, :n0], ref[1])
        assert np.array_equal(hp[b, :n0].view(np.uint64), ref[2].view(np.uint64)) and np.array_equal(init[b, :n0], ref[3])
        total += int((ref[0] >= 0).sum())
    assert total > 5
    # ---- the export block: 64-byte slots, the first 48 bytes are the row, the tail is zero
    blk = torch.zeros(L_.okb_feature_block_bytes(B, cap), dtype=torch.uint8, device="cuda")
    _l.check(L_.okb_export_features(fe.ctx, 1, B, blk.data_ptr()))
    _l.check(L_.okb_sync(fe.ctx)); torch.cuda.synchronize()
    counts, kps, descs = sh.unpack_block(blk.cpu().numpy(), B, cap, _l.KP_DTYPE)
    for b in range(B):
        kp1, d1 = feats[1][b]
        assert counts[b] == len(kp1) and kps[b].tobytes() == kp1.tobytes()
        assert np.array_equal(descs[b][:, :48], d1) and not descs[b][:, 48:].any()
    # ---- the M3 sequence of camera 0 against three older views made from frame 0's own features (device blocks in 64-byte slots)
    kp0, d0 = feats[0][0]
    rays0, v0 = oracle_bp(EU[0], kp0)
    rng = np.random.default_rng(5)
    Tw1, Tc1 = pose12(np.eye(3), np.zeros(3))
    e = rays0 / np.linalg.norm(rays0, axis=1, keepdims=True)
    P = e * np.exp(rng.uniform(np.log(1.5), np.log(25.0), len(kp0)))[:, None]
    views = []
    for v in range(3):
        Cv = rot((0, 1, 0), 0.012 * (v + 1)); rv = np.array([-0.06 * (v + 1), 0.01 * v, -0.02 * (v + 1)])
        pc = (P - rv) @ Cv
        idx = np.nonzero((v0 != 0) & (pc[:, 2] > 0.3) & (rng.random(len(kp0)) < 0.5))[0]
        bits = np.unpackbits(d0[idx], axis=1)
        d = np.packbits(bits ^ (rng.random(bits.shape) < 0.04).astype(np.uint8), axis=1)
        d[:, 47] &= 0x7f                                            # bit 383 does not exist
        ry = np.stack([pc[idx, 0] / pc[idx, 2], pc[idx, 1] / pc[idx, 2], np.ones(len(idx))], 1)
        d = np.concatenate([d, rng.integers(0, 256, (200, 48), dtype=np.uint8)])
        ry = np.concatenate([ry, np.stack([rng.uniform(-0.7, 0.7, 200), rng.uniform(-0.45, 0.45, 200), np.ones(200)], 1)])
        Tw, Tc = pose12(Cv, rv)
        views.append(dict(desc=np.ascontiguousarray(d), rays=np.ascontiguousarray(ry), valid=(rng.random(len(d)) > 0.02).astype(np.uint8),
                          size=(rng.choice([12.0, 18.0], len(d)) * rng.uniform(0.9, 1.1, len(d))).astype(np.float32),
                          use=(rng.random(len(d)) > 0.2).astype(np.uint8), T_WC=Tw, T_CW=Tc))
    cap0 = (max(len(v["desc"]) for v in views) + 63) // 64 * 64
    tab = (_l.OlderView * (B * 3))(); keep = []
    for b in range(B):
        for vi, v in enumerate(views):
            slots = np.zeros((len(v["desc"]), 64), np.uint8); slots[:, :48] = v["desc"]
            t = dict(desc=torch.from_numpy(slots).cuda(), **{k: torch.from_numpy(np.ascontiguousarray(v[k])).cuda() for k in ("rays", "valid", "size", "use")})
            keep.append(t)
            en = tab[b * 3 + vi]
            en.d_desc, en.d_rays, en.d_valid, en.d_size, en.d_use = (t[k].data_ptr() for k in ("desc", "rays", "valid", "size", "use"))
            en.n = len(v["desc"]); en.T_WC[:] = list(v["T_WC"]); en.T_CW[:] = list(v["T_CW"])
    TwB = np.ascontiguousarray(np.broadcast_to(Tw1, (B, 12))); TcB = np.ascontiguousarray(np.broadcast_to(Tc1, (B, 12)))
    mask = torch.zeros((B, cap), dtype=torch.uint8, device="cuda")
    mk1 = torch.zeros((B, 3, cap0), dtype=torch.int32, device="cuda"); md = torch.zeros((B, 3, cap0), dtype=torch.int32, device="cuda")
    mhp = torch.zeros((B, 3, cap0, 4), dtype=torch.float64, device="cuda"); mfl = torch.zeros((B, 3, cap0), dtype=torch.uint8, device="cuda")
    _l.check(L_.okb_match_motion_stereo_device(fe.ctx, 0, B, TwB.ctypes.data, TcB.ctypes.data, 3, tab, cap0, 60, mask.data_ptr(), mk1.data_ptr(),
                                               md.data_ptr(), mhp.data_ptr(), mfl.data_ptr()))
    _l.check(L_.okb_sync(fe.ctx)); torch.cuda.synchronize()
    mk1, md, mhp, mfl = mk1.cpu().numpy(), md.cpu().numpy().view(np.uint32), mhp.cpu().numpy(), mfl.cpu().numpy()
    m = EU[0]
    intr = np.array(list(m["focal_length"]) + list(m["principal_point"]) + list(m["distortion_coefficients"]))
    inserted = 0
    for b in range(B):
        kpb, db = feats[0][b]
        raysb, vb = oracle_bp(EU[0], kpb)
        ref, _ = oracle.match_motion_stereo_sequence([dict(v) for v in views], db, raysb, vb, np.stack([kpb["x"], kpb["y"]], 1), Tw1, Tc1, 1, intr, W, H, 60,
                                                     np.zeros(len(kpb), np.uint8))
        for v, (rk1, rdist, rhp, rfl) in enumerate(ref):
            n = len(rk1)
            assert np.array_equal(mk1[b, v, :n], rk1) and np.array_equal(md[b, v, :n], rdist), (b, v)
            assert np.array_equal(mhp[b, v, :n].view(np.uint64), rhp.view(np.uint64)) and np.array_equal(mfl[b, v, :n], rfl), (b, v)
            inserted += int(((rfl & 4) != 0).sum())
    assert inserted > 50
    fe.close()


@pytest.mark.parametrize("use_graph", [True, False])
def test_process_multiframe_in_the_48_byte_mode(use_graph):
    """The live path (okb_process_multiframe: one call per stereo frame, CUDA-graph replay from the third call on) with D = 48 cameras:
    camera-aware extraction whose direction changes with every frame (it reaches a replayed graph through a page-locked mirror), M1 on a
    48-byte pool, the M3 sequence against older views in 64-byte slots, M4 -- every stage against the oracle."""
    import torch
    from okvis2_b200.synth import pose12, rot, synth_stereo
    from test_gpu_camera import EUROC as EU, T_CW, oracle_bp, world_rays
    W, H, MAXKP, CAP_M, NV, RAD, THR = 752, 480, 700, 256, 3, 12.0, 40
    fe = make(W, H, RAD, THR, MAXKP, n_cams=2)
    L_ = _l.lib()
    maps = []
    for c in range(2):
        fe.setCameraModel(c, **EU[c])
        maps.append(fe.cameraAwarenessMaps(c))
    _l.check(L_.okb_stream_use_graph(fe.ctx, 1 if use_graph else 0))
    n_calls = 5
    frames = [synth_stereo(500 + t % 2, W, H) for t in range(n_calls)]
    o = oracle.HarrisBrisk2(RAD, THR, MAXKP)
    intr = [np.array(list(EU[c]["focal_length"]) + list(EU[c]["principal_point"]) + list(EU[c]["distortion_coefficients"])) for c in range(2)]
    fu = [float(np.float32(EU[c]["focal_length"][0])) for c in range(2)]

    def C_WC_at(t):   # the camera rolls a little from frame to frame: the extraction direction changes
        return rot((0, 0, 1), 0.05 * t) @ np.array([[1, 0, 0], [0, 0, 1], [0, -1, 0.]])

    def direction(c, t):
        _l.check(L_.okb_set_extraction_direction(fe.ctx, c, np.ascontiguousarray(C_WC_at(t)).ctypes.data))
        d = np.zeros(3, np.float32); _l.check(L_.okb_get_extraction_direction(fe.ctx, c, d.ctypes.data))
        return d
    poses = [pose12(np.eye(3), np.array([0.11 * c, 0.0, 0.0])) for c in range(2)]
    pools, views, dviews = [], [], []
    rng = np.random.default_rng(11)
    for c in range(2):
        kp, d = o.detect_and_compute(frames[0][c], maps[c][0], maps[c][1], fu[c], direction(c, 0))
        rays, valid = oracle_bp(EU[c], kp)
        pools.append(map_scene(7 + c, np.stack([kp["x"], kp["y"]], 1).astype(np.float64), d, 1200, W=W, H=H, frac_near=0.3))
        e = rays / np.linalg.norm(rays, axis=1, keepdims=True)
        P = poses[c][0][9:] + e * np.exp(rng.uniform(np.log(1.5), np.log(25.0), len(kp)))[:, None]
        vs = []
        for v in range(NV):
            Cv = rot((0, 1, 0), 0.012 * (v + 1)); rv = poses[c][0][9:] + np.array([-0.06 * (v + 1), 0.01 * v, -0.02 * (v + 1)])
            pc = (P - rv) @ Cv
            idx = np.nonzero((valid != 0) & (pc[:, 2] > 0.3) & (rng.random(len(kp)) < 0.4))[0]
            bits = np.unpackbits(d[idx], axis=1)
            dd = np.packbits(bits ^ (rng.random(bits.shape) < 0.04).astype(np.uint8), axis=1)
            ry = np.stack([pc[idx, 0] / pc[idx, 2], pc[idx, 1] / pc[idx, 2], np.ones(len(idx))], 1)
            dd = np.concatenate([dd, rng.integers(0, 256, (200, 48), dtype=np.uint8)])
            ry = np.concatenate([ry, np.stack([rng.uniform(-0.7, 0.7, 200), rng.uniform(-0.45, 0.45, 200), np.ones(200)], 1)])
            Tw, Tc = pose12(Cv, rv)
            vs.append(dict(desc=np.ascontiguousarray(dd), rays=np.ascontiguousarray(ry), valid=(rng.random(len(dd)) > 0.02).astype(np.uint8),
                           size=(rng.choice([12.0, 18.0], len(dd)) * rng.uniform(0.9, 1.1, len(dd))).astype(np.float32),
                           use=(rng.random(len(dd)) > 0.2).astype(np.uint8), T_WC=Tw, T_CW=Tc))
        views.append(vs)
        dv = []
        for v in vs:
            slots = np.zeros((len(v["desc"]), 64), np.uint8); slots[:, :48] = v["desc"]
            dv.append(dict(desc=torch.from_numpy(slots).cuda(), **{k: torch.from_numpy(np.ascontiguousarray(v[k])).cuda() for k in ("rays", "valid", "size", "use")}))
        dviews.append(dv)
    cap0 = (max(len(v["desc"]) for vs in views for v in vs) + 63) // 64 * 64
    cap = MAXKP
    C0 = np.eye(3); r0 = np.zeros(3); r1 = np.array([0.11, 0.0, 0.0])
    totals = dict(m1=0, m3=0, m4=0)
    try:
        for t in range(n_calls):
            io = (_l.MultiframeCam * 2)(); bufs = []; tabs = []; dirs = []
            for c in range(2):
                dirs.append(direction(c, t))
                q = io[c]
                tab = (_l.OlderView * NV)()
                for vi in range(NV):
                    v, dvv = views[c][(vi + t) % NV], dviews[c][(vi + t) % NV]
                    e = tab[vi]
                    e.d_desc, e.d_rays, e.d_valid, e.d_size, e.d_use = (dvv[k].data_ptr() for k in ("desc", "rays", "valid", "size", "use"))
                    e.n = len(v["desc"]); e.T_WC[:] = list(v["T_WC"]); e.T_CW[:] = list(v["T_CW"])
                tabs.append(tab)
                img = np.ascontiguousarray(frames[t][c]); m = pools[c]
                proj = np.ascontiguousarray(m["lm_proj"] + 0.7 * t)
                b = dict(img=img, proj=proj, kp=np.zeros(cap, _l.KP_DTYPE), desc=np.zeros((cap, 48), np.uint8), rays=np.zeros((cap, 3)),
                         valid=np.zeros(cap, np.uint8), m1d=np.zeros(cap, np.uint32), m1l=np.zeros(cap, np.int32), m3n=np.zeros(NV, np.int32),
                         k0=np.zeros((NV, CAP_M), np.int32), k1=np.zeros((NV, CAP_M), np.int32), fl=np.zeros((NV, CAP_M), np.uint8),
                         hp=np.zeros((NV, CAP_M, 4)), Tw=np.ascontiguousarray(poses[c][0]), Tc=np.ascontiguousarray(poses[c][1]),
                         cd=np.ascontiguousarray(m["cand_desc"]), cl=np.ascontiguousarray(m["cand_lm"]), c3=np.ascontiguousarray(m["lm_is3d"]))
                bufs.append(b)
                q.image = img.ctypes.data; q.stride_bytes = W
                q.n_cand = len(b["cl"]); q.n_lm = len(b["c3"]); q.pool_changed = 1 if t == 0 else 0
                q.cand_desc, q.cand_lm, q.lm_is3d, q.lm_proj = b["cd"].ctypes.data, b["cl"].ctypes.data, b["c3"].ctypes.data, proj.ctypes.data
                q.T_WC1, q.T_CW1 = b["Tw"].ctypes.data, b["Tc"].ctypes.data
                q.n_older, q.cap0, q.older = NV, cap0, C.addressof(tab)
                q.cap = cap; q.kp, q.desc, q.rays, q.rays_valid = (b[k].ctypes.data for k in ("kp", "desc", "rays", "valid"))
                q.m1_dist, q.m1_lm = b["m1d"].ctypes.data, b["m1l"].ctypes.data
                q.cap_m = CAP_M; q.m3_n, q.m3_k0, q.m3_k1, q.m3_flags, q.m3_hp_W = (b[k].ctypes.data for k in ("m3n", "k0", "k1", "fl", "hp"))
            st = _l.MultiframeStereo(); st.cam0, st.cam1 = 0, 1
            st.C_WC0[:] = [1, 0, 0, 0, 1, 0, 0, 0, 1]; st.C_WC1[:] = [1, 0, 0, 0, 1, 0, 0, 0, 1]; st.r_WC0[:] = [0, 0, 0]; st.r_WC1[:] = [0.11, 0, 0]
            sb = dict(k1=np.zeros(cap, np.int32), dist=np.zeros(cap, np.uint32), hp=np.zeros((cap, 4)), init=np.zeros(cap, np.uint8))
            st.k1, st.dist, st.hp_W, st.initialisable = (sb[k].ctypes.data for k in ("k1", "dist", "hp", "init"))
            _l.check(L_.okb_process_multiframe(fe.ctx, 2, io, 1, C.byref(st), 20.0, 60))
            feats = []
            for c in range(2):
                b = bufs[c]; n = io[c].n
                rk, rd = o.detect_and_compute(frames[t][c], maps[c][0], maps[c][1], fu[c], dirs[c])
                same48(b["kp"][:n], b["desc"][:n], rk, rd, f"call {t} camera {c}")
                rays, valid = oracle_bp(EU[c], rk)
                assert np.array_equal(b["rays"][:n].view(np.uint64), rays.view(np.uint64)) and np.array_equal(b["valid"][:n], valid)
                m = pools[c]
                xy = np.stack([rk["x"], rk["y"]], 1).astype(np.float64)
                rdist, rlm = oracle.match_map3d(rd, xy, None, m["cand_desc"], m["cand_lm"], b["proj"], m["lm_is3d"], 20.0, 60)
                assert np.array_equal(b["m1d"][:n], rdist.astype(np.uint32)) and np.array_equal(b["m1l"][:n], rlm), (t, c)
                totals["m1"] += int((rlm >= 0).sum())
                ov = [dict(views[c][(vi + t) % NV]) for vi in range(NV)]
                ref, _ = oracle.match_motion_stereo_sequence(ov, rd, rays, valid, np.stack([rk["x"], rk["y"]], 1), poses[c][0], poses[c][1], 1, intr[c], W, H, 60,
                                                             (rlm >= 0).astype(np.uint8))
                for v, (k1, dist, hp, fl) in enumerate(ref):
                    k0s = np.nonzero(fl & 1)[0]
                    assert b["m3n"][v] == len(k0s), (t, c, v)
                    assert np.array_equal(b["k0"][v, :len(k0s)], k0s) and np.array_equal(b["k1"][v, :len(k0s)], k1[k0s])
                    assert np.array_equal(b["fl"][v, :len(k0s)], fl[k0s]) and np.array_equal(b["hp"][v, :len(k0s)].view(np.uint64), hp[k0s].view(np.uint64))
                    totals["m3"] += int(((fl & 4) != 0).sum())
                f = 0.5 * sum(EU[c]["focal_length"])
                feats.append((rd, valid, world_rays(C0, rays), rk["size"].astype(np.float64) / f))
            ref = oracle.match_stereo(*feats[0], *feats[1], r0, r1, T_CW(C0, r0), T_CW(C0, r1), 60)
            n0 = io[0].n
            assert np.array_equal(sb["k1"][:n0], ref[0]) and np.array_equal(sb["dist"][:n0], ref[1])
            assert np.array_equal(sb["hp"][:n0].view(np.uint64), ref[2].view(np.uint64)) and np.array_equal(sb["init"][:n0], ref[3])
            totals["m4"] += int((ref[0] >= 0).sum())
        g, dcalls = C.c_longlong(), C.c_longlong()
        _l.check(L_.okb_stream_stats(fe.ctx, C.byref(g), C.byref(dcalls)))
        assert (g.value, dcalls.value) == ((n_calls - 1, 1) if use_graph else (0, n_calls))
    finally:
        fe.close()
    assert totals["m1"] > 50 and totals["m3"] > 30 and totals["m4"] > 5, totals


def test_cuda_equals_the_frozen_vectors():
    """the CUDA path against tests/golden/harris_brisk2_oracle.npz (frozen outputs of this repository's restatement; parity unpinned)"""
    import os
    from conftest import ROOT
    g = np.load(os.path.join(ROOT, "tests", "golden", "harris_brisk2_oracle.npz"))
    euroc0 = EUROC[0]
    for name in sorted(k[:-4] for k in g.files if k.endswith("_cfg")):
        seed, W, H, radius, thr, max_kp, aware = g[name + "_cfg"]
        W, H = int(W), int(H)
        fe = make(W, H, float(radius), int(thr), int(max_kp))
        T_WC = None
        if aware:
            fe.setCameraModel(0, **euroc0)
            fe.cameraAwarenessMaps(0)
            # a rotation whose third row is minus the frozen direction: gravity in the camera frame = R_CW (0, 0, -1) = -(row 2 of C_WC)
            d = g[name + "_dir"].astype(np.float64)
            a = np.cross(d, [1.0, 0, 0]); a /= np.linalg.norm(a); b = np.cross(d, a)
            T_WC = np.stack([a, b, -d])
            if np.linalg.det(T_WC) < 0:
                T_WC[0] = -T_WC[0]
        fr = run(fe, synth_frame(int(seed), W, H), 0, T_WC)
        rk = np.frombuffer(g[name + "_kp"].tobytes(), _l.KP_DTYPE)
        if aware:
            assert np.array_equal(fr.extractionDirection, g[name + "_dir"]), "the test's rotation must reproduce the frozen direction bit for bit"
        same48(fr.keypoints, fr.descriptors, rk, g[name + "_desc"], name)
        fe.close()


def test_candidate_overflow_is_a_loud_error():
    """more maxima than the uniformity kernel ranks (16384 per frame): OKB_ERR_CAPACITY, never a silently truncated result"""
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (1024, 1024), dtype=np.uint8)
    fe = make(1024, 1024, 20.0, 1, 500)
    with pytest.raises(OkbError) as e:
        run(fe, img)
    assert e.value.status == _l.OKB_ERR_CAPACITY
    # the context stays usable: a frame with a handful of corners
    img2 = np.full((1024, 1024), 20, np.uint8)
    for i in range(6):
        for j in range(6):
            img2[100 + 140 * i:160 + 140 * i, 100 + 140 * j:160 + 140 * j] = 220
    fr = run(fe, img2)
    rk, rd = oracle.HarrisBrisk2(20.0, 1, 500).detect_and_compute(img2)
    assert len(rk) >= 36
    same48(fr.keypoints, fr.descriptors, rk, rd, "after the overflow")
    fe.close()
