"""GPU suite: D4 back-projection and the device-resident stereo pipeline (detect -> describe -> back-project -> M4)."""
import ctypes as C

import numpy as np
import pytest

import oracle
from okvis2_b200 import lib as okl
from okvis2_b200.frontend import Frontend, MultiFrame
from okvis2_b200.synth import map_scene, synth_stereo

pytestmark = pytest.mark.gpu

EUROC = [dict(distortion_type="radialtangential", focal_length=(458.654880721, 457.296696463),
              principal_point=(367.215803962, 248.37534061),
              distortion_coefficients=[-0.28340811217, 0.0739590738929, 0.000193595028569, 1.76187114545e-05]),
         dict(distortion_type="radialtangential", focal_length=(457.587426604, 456.13442556),
              principal_point=(379.99944652, 255.238185386),
              distortion_coefficients=[-0.283683654496, 0.0745128430929, -0.000104738949098, -3.55590700274e-05])]
HILTI0 = dict(distortion_type="equidistant", focal_length=(351.31400364193297, 351.4911744656785),
              principal_point=(367.8522793375995, 253.84021449809963),
              distortion_coefficients=[-0.03696737352869157, -0.008917880497032812, 0.008912969593422046, -0.0037685977496087313])


def oracle_bp(cfg, kp):
    return oracle.back_project(Frontend.MODELS[cfg["distortion_type"]], *cfg["focal_length"], *cfg["principal_point"],
                               cfg["distortion_coefficients"], kp)


def test_back_projection_radtan_exact_and_equidistant_close():
    fe = Frontend(1, 752, 480)
    rng = np.random.default_rng(0)
    mf = MultiFrame(1)
    kp = np.zeros(3000, okl.KP_DTYPE)
    kp["x"] = rng.uniform(0, 752, 3000).astype(np.float32); kp["y"] = rng.uniform(0, 480, 3000).astype(np.float32)
    mf.frames[0].keypoints = kp
    fe.setCameraModel(0, **EUROC[0])
    n_ok = fe.computeBackProjections(mf, 0)
    rays, valid = oracle_bp(EUROC[0], kp)
    assert np.array_equal(mf.frames[0].backProjections.view(np.uint64), rays.view(np.uint64))   # bit-exact
    assert np.array_equal(mf.frames[0].backProjectionsValid, valid) and n_ok == valid.sum() > 2900
    fe.setCameraModel(0, "none", (458.0, 457.0), (367.0, 248.0), [])
    fe.computeBackProjections(mf, 0)
    rays, valid = oracle.back_project(0, 458.0, 457.0, 367.0, 248.0, [0, 0, 0, 0], kp)
    assert np.array_equal(mf.frames[0].backProjections.view(np.uint64), rays.view(np.uint64))
    # equidistant needs atan: device and libm agree to the last bits only (tolerance stated: 1e-13 absolute)
    fe.setCameraModel(0, **HILTI0)
    kp["x"] = rng.uniform(0, 720, 3000).astype(np.float32); kp["y"] = rng.uniform(0, 540, 3000).astype(np.float32)
    mf.frames[0].keypoints = kp
    fe.computeBackProjections(mf, 0)
    rays, valid = oracle_bp(HILTI0, kp)
    assert np.array_equal(mf.frames[0].backProjectionsValid, valid)
    assert np.abs(mf.frames[0].backProjections - rays).max() <= 1e-13
    fe.close()


def world_rays(C_WC, rays):
    """(C_WC * e_C).normalized() with the library's association order"""
    x, y, z = rays[:, 0], rays[:, 1], rays[:, 2]
    w = [(C_WC[i, 0] * x + C_WC[i, 1] * y) + C_WC[i, 2] * z for i in range(3)]
    n = np.sqrt((w[0] * w[0] + w[1] * w[1]) + w[2] * w[2])
    return np.ascontiguousarray(np.stack([w[0] / n, w[1] / n, w[2] / n], 1))


def T_CW(C_WC, r):
    T = np.zeros((3, 4))
    for i in range(3):
        T[i, :3] = C_WC[:, i]
        T[i, 3] = -((C_WC[0, i] * r[0] + C_WC[1, i] * r[1]) + C_WC[2, i] * r[2])
    return T.reshape(12)


@pytest.fixture(params=["umma", "mma", "popc"])
def scan_form(request):
    """the three forms of the Hamming scan of the device-resident matchers: tcgen05 + TMEM (default), legacy integer MMA, POPC"""
    okl.lib().okb_scan_set_mma({"umma": 2, "mma": 1, "popc": 0}[request.param])
    yield request.param
    okl.lib().okb_scan_set_mma(2)


def test_device_resident_stereo_pipeline_equals_oracle(scan_form):
    import torch
    B = 3
    fe = Frontend(2, 752, 480, max_batch=B)
    fe.configure(threshold=30, octaves=3, max_keypoints=1000)
    for c in range(2):
        fe.setCameraModel(c, **EUROC[c])
    L_ = okl.lib()
    imgs = [np.stack([synth_stereo(700 + t, 752, 480)[c] for t in range(B)]) for c in range(2)]
    d_imgs = [torch.from_numpy(a).cuda() for a in imgs]
    for c in range(2):
        okl.check(L_.okb_detect_describe_batch_device(fe.ctx, c, B, d_imgs[c].data_ptr()))
    cap = C.c_int(0)
    L_.okb_device_features(fe.ctx, 0, None, None, None, C.byref(cap)); cap = cap.value
    C0 = np.eye(3); r0 = np.zeros(3)
    a = 0.01
    C1 = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]); r1 = np.array([0.11, 0.001, -0.002])
    k1 = torch.zeros((B, cap), dtype=torch.int32, device="cuda"); dist = torch.zeros((B, cap), dtype=torch.int32, device="cuda")
    hp = torch.zeros((B, cap, 4), dtype=torch.float64, device="cuda"); init = torch.zeros((B, cap), dtype=torch.uint8, device="cuda")
    okl.check(L_.okb_match_stereo_device(fe.ctx, 0, 1, B, C0.ctypes.data, r0.ctypes.data, np.ascontiguousarray(C1).ctypes.data,
                                         r1.ctypes.data, 60, k1.data_ptr(), dist.data_ptr(), hp.data_ptr(), init.data_ptr()))
    okl.check(L_.okb_sync(fe.ctx)); torch.cuda.synchronize()
    k1, dist, hp, init = k1.cpu().numpy(), dist.cpu().numpy().view(np.uint32), hp.cpu().numpy(), init.cpu().numpy()
    total = 0
    for b in range(B):
        feats = []
        for c in range(2):
            kp = np.zeros(cap, okl.KP_DTYPE); d = np.zeros((cap, 64), np.uint8); n = C.c_int(0)
            okl.check(L_.okb_fetch_features(fe.ctx, c, b, kp.ctypes.data, d.ctypes.data, cap, C.byref(n)))
            feats.append((kp[:n.value], d[:n.value]))
        (kp0, d0), (kp1, d1) = feats
        rk, rd = oracle.Brisk(30, 3).detect_and_compute(imgs[0][b], 1000)
        assert kp0.tobytes() == rk.tobytes() and np.array_equal(d0, rd)
        rays0, v0 = oracle_bp(EUROC[0], kp0); rays1, v1 = oracle_bp(EUROC[1], kp1)
        f0 = 0.5 * sum(EUROC[0]["focal_length"]); f1 = 0.5 * sum(EUROC[1]["focal_length"])
        ref = oracle.match_stereo(d0, v0, world_rays(C0, rays0), kp0["size"].astype(np.float64) / f0, d1, v1, world_rays(C1, rays1),
                                  kp1["size"].astype(np.float64) / f1, r0, r1, T_CW(C0, r0), T_CW(C1, r1), 60)
        n0 = len(kp0)
        assert np.array_equal(k1[b, :n0], ref[0]) and np.array_equal(dist[b, :n0], ref[1])
        assert np.array_equal(hp[b, :n0].view(np.uint64), ref[2].view(np.uint64)) and np.array_equal(init[b, :n0], ref[3])
        assert (k1[b, n0:] == -1).all()
        total += (ref[0] >= 0).sum()
    assert total > 5
    fe.close()


@pytest.mark.parametrize("pinned", [False, True])
def test_host_buffer_batch_pipeline_equals_oracle(pinned):
    """Replay step through HOST buffers: okb_detect_describe_batch -> okb_match_map3d_batch -> okb_match_stereo_batch,
    with pageable and with page-locked caller memory (the latter is copied by the DMA engines directly)."""
    import torch
    from okvis2_b200.synth import map_scene
    B = 3
    fe = Frontend(2, 752, 480, max_batch=B)
    fe.configure(threshold=30, octaves=3, max_keypoints=1000)
    for c in range(2):
        fe.setCameraModel(c, **EUROC[c])
    L_ = okl.lib()
    cap = fe._capacity(0) + 24      # caller capacity differs from the device row stride on purpose

    def buf(shape, dtype):
        if not pinned:
            return np.zeros(shape, dtype)
        t = torch.zeros(int(np.prod(shape)) * np.dtype(dtype).itemsize, dtype=torch.uint8).pin_memory()
        return t.numpy().view(dtype).reshape(shape)

    imgs = []
    for c in range(2):
        a = buf((B, 480, 752), np.uint8)
        for t in range(B):
            a[t] = synth_stereo(900 + t, 752, 480)[c]
        imgs.append(a)
    kps, descs, ns = [], [], []
    for c in range(2):
        kp = buf((B, cap), okl.KP_DTYPE); d = buf((B, cap, 64), np.uint8); n = np.zeros(B, np.int32)
        okl.check(L_.okb_detect_describe_batch(fe.ctx, c, B, imgs[c].ctypes.data, 752, kp.ctypes.data, d.ctypes.data, cap, n.ctypes.data))
        kps.append(kp); descs.append(d); ns.append(n)
    feats = [[oracle.Brisk(30, 3).detect_and_compute(imgs[c][b], 1000) for b in range(B)] for c in range(2)]
    for c in range(2):
        for b in range(B):
            rk, rd = feats[c][b]
            assert ns[c][b] == len(rk) and kps[c][b, :len(rk)].tobytes() == rk.tobytes() and np.array_equal(descs[c][b, :len(rk)], rd)
    # M1: one landmark pool, one projection table per frame (shifted a little from frame to frame)
    rk, rd = feats[0][0]
    m = map_scene(5, np.stack([rk["x"], rk["y"]], 1).astype(np.float64), rd, 2000, W=752, H=480)
    proj = buf((B,) + m["lm_proj"].shape, np.float64)
    for b in range(B):
        proj[b] = m["lm_proj"] + 1.5 * b
    dist = buf((B, cap), np.uint32); lm = buf((B, cap), np.int32)
    fe.matchToMapBatch(0, B, m["cand_desc"], m["cand_lm"], proj, m["lm_is3d"], out=(dist, lm))
    hits = 0
    for b in range(B):
        rk, rd = feats[0][b]
        ref = oracle.match_map3d(rd, np.stack([rk["x"], rk["y"]], 1).astype(np.float64), None, m["cand_desc"], m["cand_lm"], proj[b],
                                 m["lm_is3d"], 20.0, 60)
        assert np.array_equal(dist[b, :len(rk)], ref[0]) and np.array_equal(lm[b, :len(rk)], ref[1])
        hits += (ref[1] >= 0).sum()
    assert hits > 50
    # M4
    C0 = np.eye(3); r0 = np.zeros(3); C1 = np.eye(3); r1 = np.array([0.11, 0.0, 0.0])
    out = (buf((B, cap), np.int32), buf((B, cap), np.uint32), buf((B, cap, 4), np.float64), buf((B, cap), np.uint8))
    k1, sdist, hp, init = fe.matchStereoBatch(0, 1, B, C0, r0, C1, r1, out=out)
    for b in range(B):
        (kp0, d0), (kp1, d1) = feats[0][b], feats[1][b]
        rays0, v0 = oracle_bp(EUROC[0], kp0); rays1, v1 = oracle_bp(EUROC[1], kp1)
        f0 = 0.5 * sum(EUROC[0]["focal_length"]); f1 = 0.5 * sum(EUROC[1]["focal_length"])
        ref = oracle.match_stereo(d0, v0, world_rays(C0, rays0), kp0["size"].astype(np.float64) / f0, d1, v1, world_rays(C1, rays1),
                                  kp1["size"].astype(np.float64) / f1, r0, r1, T_CW(C0, r0), T_CW(C1, r1), 60)
        n0 = len(kp0)
        assert np.array_equal(k1[b, :n0], ref[0]) and np.array_equal(sdist[b, :n0], ref[1])
        assert np.array_equal(hp[b, :n0].view(np.uint64), ref[2].view(np.uint64)) and np.array_equal(init[b, :n0], ref[3])
    fe.close()


def test_stereo_scan_gate_split_with_dense_hits_and_overflow(scan_form):
    """M4 device form on crafted feature blocks: frame 0 holds near-duplicate descriptors (every pair is below the matching
    threshold: the hit list overflows and the sequential-replay kernel redoes the frame), frame 1 a normal mix with many
    hits per query, frame 2 is empty on the query side. All three must equal the oracle's transcription of the loop."""
    matched = _stereo_ptr_case(3, 320, [[300, 250, 0], [310, 280, 100]], dense_div=3)
    assert matched[0] > 100 and matched[1] > 30 and matched[2] == 0


def test_stereo_device_form_batch_of_eight_large_frames(scan_form):
    """the batched forms of the scan (tcgen05: one CTA per candidate tile streams all query tiles) at TUM-VI sizes: 8 frames of
    ~2 000 keypoints in blocks of capacity 2 496 (20 candidate tiles, 16 query tiles)"""
    counts = [[2000, 1900, 2496, 1, 1777, 2048, 2100, 1500], [1950, 2010, 2496, 300, 1800, 2047, 1, 1600]]
    matched = _stereo_ptr_case(8, 2496, counts, dense_div=25, all_near_first=False)
    assert sum(matched) > 400


def _stereo_ptr_case(B, cap, counts, dense_div, all_near_first=True):
    import torch
    rng = np.random.default_rng(42)
    fe = Frontend(0)
    L_ = okl.lib()
    models = []
    for c in range(2):
        m = okl.CameraModel(); m.model = 1
        m.fu, m.fv = EUROC[c]["focal_length"]; m.cu, m.cv = EUROC[c]["principal_point"]
        for i in range(4):
            m.k[i] = EUROC[c]["distortion_coefficients"][i]
        models.append(m)
    base = rng.integers(0, 256, 64, dtype=np.uint8)

    def near(n, flips):
        d = np.tile(base, (n, 1))
        for i in range(n):
            for b in rng.choice(512, flips, replace=False):
                d[i, b // 8] ^= 1 << (b % 8)
        return d

    kps = [np.zeros((B, cap), okl.KP_DTYPE) for _ in range(2)]
    descs = [np.zeros((B, cap, 64), np.uint8) for _ in range(2)]
    for c in range(2):
        for b in range(B):
            n = counts[c][b]
            kps[c][b]["x"][:n] = rng.uniform(40, 700, n); kps[c][b]["y"][:n] = rng.uniform(40, 440, n)
            kps[c][b]["size"][:n] = rng.choice([12.0, 18.0, 24.0, 36.0], n)
            if b == 0 and all_near_first:
                descs[c][b, :n] = near(n, 8)                                          # all pairs < 60
            else:
                descs[c][b, :n] = rng.integers(0, 256, (n, 64), dtype=np.uint8)
                descs[c][b, : n // dense_div] = near(n // dense_div, 20)               # a part of them: dense hits
    # stereo geometry: points in front of both cameras so that many gates pass: camera 1 keypoints = camera 0 shifted
    for b in range(B):
        n = min(counts[0][b], counts[1][b])
        kps[1][b]["x"][:n] = kps[0][b]["x"][:n] - rng.uniform(2, 30, n).astype(np.float32)
        kps[1][b]["y"][:n] = kps[0][b]["y"][:n] + rng.uniform(-0.3, 0.3, n).astype(np.float32)
    d_kp = [torch.from_numpy(k.view(np.uint8).reshape(B, cap * 28)).cuda() for k in kps]
    d_desc = [torch.from_numpy(d).cuda() for d in descs]
    d_cnt = [torch.tensor(c, dtype=torch.int32).cuda() for c in counts]
    C0 = np.eye(3); r0 = np.zeros(3); C1 = np.eye(3); r1 = np.array([0.11, 0.0, 0.0])
    k1 = torch.zeros((B, cap), dtype=torch.int32, device="cuda"); dist = torch.zeros((B, cap), dtype=torch.int32, device="cuda")
    hp = torch.zeros((B, cap, 4), dtype=torch.float64, device="cuda"); init = torch.zeros((B, cap), dtype=torch.uint8, device="cuda")
    okl.check(L_.okb_match_stereo_device_ptr(fe.ctx, B, cap, d_kp[0].data_ptr(), d_desc[0].data_ptr(), d_cnt[0].data_ptr(),
                                             C.addressof(models[0]), C0.ctypes.data, r0.ctypes.data, cap, d_kp[1].data_ptr(),
                                             d_desc[1].data_ptr(), d_cnt[1].data_ptr(), C.addressof(models[1]), C1.ctypes.data,
                                             r1.ctypes.data, 60, None, k1.data_ptr(), dist.data_ptr(), hp.data_ptr(), init.data_ptr()))
    okl.check(L_.okb_sync(fe.ctx)); torch.cuda.synchronize()
    k1, dist, hp, init = k1.cpu().numpy(), dist.cpu().numpy().view(np.uint32), hp.cpu().numpy(), init.cpu().numpy()
    matched = []
    for b in range(B):
        n0, n1 = counts[0][b], counts[1][b]
        kp0, kp1 = kps[0][b][:n0], kps[1][b][:n1]
        rays0, v0 = oracle_bp(EUROC[0], kp0); rays1, v1 = oracle_bp(EUROC[1], kp1)
        f0 = 0.5 * sum(EUROC[0]["focal_length"]); f1 = 0.5 * sum(EUROC[1]["focal_length"])
        ref = oracle.match_stereo(descs[0][b, :n0], v0, world_rays(C0, rays0), kp0["size"].astype(np.float64) / f0, descs[1][b, :n1], v1,
                                  world_rays(C1, rays1), kp1["size"].astype(np.float64) / f1, r0, r1, T_CW(C0, r0), T_CW(C1, r1), 60)
        assert np.array_equal(k1[b, :n0], ref[0]) and np.array_equal(dist[b, :n0], ref[1]), b
        assert np.array_equal(hp[b, :n0].view(np.uint64), ref[2].view(np.uint64)) and np.array_equal(init[b, :n0], ref[3]), b
        assert (k1[b, n0:] == -1).all()
        matched.append(int((ref[0] >= 0).sum()))
    fe.close()
    return matched


def test_gate_cos_is_libm_and_both_m4_forms_agree():
    """The gate constants cos(2.6 sigma) / cos(6 sigma) come from ONE function on host and device (csrc/okb_gatecos.h) that
    okb_create verified against this machine's libm; the host-buffer M4 (okb_match_stereo) and the device-resident M4
    (okb_match_stereo_device_ptr) must therefore return identical bits on the same features."""
    import torch
    rng = np.random.default_rng(7)
    fe = Frontend(0)
    L_ = okl.lib()
    assert L_.okb_gate_cos_exact(fe.ctx) == 1
    B, cap, n0, n1 = 1, 512, 480, 500
    models = []
    for c in range(2):
        m = okl.CameraModel(); m.model = 1
        m.fu, m.fv = EUROC[c]["focal_length"]; m.cu, m.cv = EUROC[c]["principal_point"]
        for i in range(4):
            m.k[i] = EUROC[c]["distortion_coefficients"][i]
        models.append(m)
    kps = [np.zeros((B, cap), okl.KP_DTYPE) for _ in range(2)]
    descs = [np.zeros((B, cap, 64), np.uint8) for _ in range(2)]
    base = rng.integers(0, 256, (n0, 64), dtype=np.uint8)
    for c, n in enumerate((n0, n1)):
        kps[c][0]["x"][:n] = rng.uniform(40, 700, n); kps[c][0]["y"][:n] = rng.uniform(40, 440, n)
        kps[c][0]["size"][:n] = (rng.uniform(8.4, 108.0, n)).astype(np.float32)     # continuous sizes: many distinct sigmas
        descs[c][0, :n] = rng.integers(0, 256, (n, 64), dtype=np.uint8)
    m = min(n0, n1)
    kps[1][0]["x"][:m] = kps[0][0]["x"][:m] - rng.uniform(2, 30, m).astype(np.float32)
    kps[1][0]["y"][:m] = kps[0][0]["y"][:m] + rng.uniform(-0.3, 0.3, m).astype(np.float32)
    flips = rng.random((m, 512)) < 0.03
    descs[0][0, :m] = base[:m]; descs[1][0, :m] = base[:m] ^ np.packbits(flips, axis=1)
    counts = [[n0], [n1]]
    d_kp = [torch.from_numpy(k.view(np.uint8).reshape(B, cap * 28)).cuda() for k in kps]
    d_desc = [torch.from_numpy(d).cuda() for d in descs]
    d_cnt = [torch.tensor(c, dtype=torch.int32).cuda() for c in counts]
    C0 = np.eye(3); r0 = np.zeros(3); C1 = np.eye(3); r1 = np.array([0.11, 0.0, 0.0])
    k1 = torch.zeros((B, cap), dtype=torch.int32, device="cuda"); dist = torch.zeros((B, cap), dtype=torch.int32, device="cuda")
    hp = torch.zeros((B, cap, 4), dtype=torch.float64, device="cuda"); init = torch.zeros((B, cap), dtype=torch.uint8, device="cuda")
    okl.check(L_.okb_match_stereo_device_ptr(fe.ctx, B, cap, d_kp[0].data_ptr(), d_desc[0].data_ptr(), d_cnt[0].data_ptr(),
                                             C.addressof(models[0]), C0.ctypes.data, r0.ctypes.data, cap, d_kp[1].data_ptr(),
                                             d_desc[1].data_ptr(), d_cnt[1].data_ptr(), C.addressof(models[1]), C1.ctypes.data,
                                             r1.ctypes.data, 60, None, k1.data_ptr(), dist.data_ptr(), hp.data_ptr(), init.data_ptr()))
    okl.check(L_.okb_sync(fe.ctx)); torch.cuda.synchronize()
    kp0, kp1 = kps[0][0][:n0], kps[1][0][:n1]
    rays0, v0 = oracle_bp(EUROC[0], kp0); rays1, v1 = oracle_bp(EUROC[1], kp1)
    f0 = 0.5 * sum(EUROC[0]["focal_length"]); f1 = 0.5 * sum(EUROC[1]["focal_length"])
    host = fe.matchStereo(descs[0][0, :n0], v0, world_rays(C0, rays0), kp0["size"].astype(np.float64) / f0, descs[1][0, :n1], v1,
                          world_rays(C1, rays1), kp1["size"].astype(np.float64) / f1, r0, r1, T_CW(C0, r0), T_CW(C1, r1))
    ref = oracle.match_stereo(descs[0][0, :n0], v0, world_rays(C0, rays0), kp0["size"].astype(np.float64) / f0, descs[1][0, :n1], v1,
                              world_rays(C1, rays1), kp1["size"].astype(np.float64) / f1, r0, r1, T_CW(C0, r0), T_CW(C1, r1), 60)
    dev = (k1.cpu().numpy()[0, :n0], dist.cpu().numpy().view(np.uint32)[0, :n0], hp.cpu().numpy()[0, :n0], init.cpu().numpy()[0, :n0])
    for a, b, r in zip(dev, host, ref):
        assert np.array_equal(np.asarray(a).view(np.uint8), np.asarray(b).view(np.uint8)), "device-resident and host-buffer M4 differ"
        assert np.array_equal(np.asarray(b).view(np.uint8), np.asarray(r).view(np.uint8)), "M4 differs from the oracle (libm cos)"
    assert (ref[0] >= 0).sum() > 100
    fe.close()


def test_export_features_block_layout():
    """okb_export_features: [counts | keypoints | descriptors] of a batch, the unit the camera-sharded mode all-gathers."""
    import torch
    from okvis2_b200 import sharding as sh
    B = 2
    fe = Frontend(1, 752, 480, max_batch=B)
    fe.configure(threshold=30, octaves=3, max_keypoints=600)
    L_ = okl.lib()
    imgs = np.stack([synth_stereo(40 + t, 752, 480)[0] for t in range(B)])
    feats = fe.detectAndDescribeBatch(0, imgs)
    cap = C.c_int(0); L_.okb_device_features(fe.ctx, 0, None, None, None, C.byref(cap)); cap = cap.value
    nbytes = L_.okb_feature_block_bytes(B, cap)
    assert nbytes == sh.block_layout(B, cap)[3]
    blk = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    okl.check(L_.okb_export_features(fe.ctx, 0, B, blk.data_ptr()))
    okl.check(L_.okb_sync(fe.ctx)); torch.cuda.synchronize()
    counts, kps, descs = sh.unpack_block(blk.cpu().numpy(), B, cap, okl.KP_DTYPE)
    for b in range(B):
        assert counts[b] == len(feats[b][0]) and kps[b].tobytes() == feats[b][0].tobytes() and np.array_equal(descs[b], feats[b][1])
    fe.close()


def test_extraction_direction_is_gravity_in_the_camera_frame():
    """D1 (Frontend.cpp:245-251): T_WC.inverse().C() * (0, 0, -1) as floats, per camera; BRISK-512 output does not depend on it"""
    fe = Frontend(2, 752, 480)
    fe.configure(threshold=30, octaves=3, max_keypoints=600)
    try:
        img = synth_stereo(77, 752, 480)[0]
        a, b = 0.3, -0.2
        Rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
        Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
        T = np.eye(4); T[:3, :3] = Rx @ Ry; T[:3, 3] = [1.0, -2.0, 0.5]
        mf = MultiFrame(2); mf.setImage(0, img); mf.setImage(1, img)
        fe.detectAndDescribe(0, mf, T, None)
        fe.detectAndDescribe(1, mf, None, None)
        want = (T[:3, :3].T @ np.array([0.0, 0.0, -1.0])).astype(np.float32)
        assert np.array_equal(mf.frames[0].extractionDirection, want)
        assert mf.frames[1].extractionDirection is None
        d = np.zeros(3, np.float32)
        okl.check(okl.lib().okb_get_extraction_direction(fe.ctx, 1, okl.ptr(d)))
        assert np.array_equal(d, np.array([0, 0, -1], np.float32))      # default: camera looking along world x/y, z up
        assert np.array_equal(mf.frames[0].descriptors, mf.frames[1].descriptors)
    finally:
        fe.close()


@pytest.mark.parametrize("loop", [False, True])
def test_device_resident_m2_equals_oracle(loop):
    """M2 (matchToMapByThreadUnitialised) on the features that stay on the device: e1_W = T_WC1.C() * e1_C.normalized() per frame, the
    use mask from the back-projection validity and the caller's mask, per-frame pose and early-break counter; against the oracle"""
    import torch
    B = 3
    fe = Frontend(1, 752, 480, max_batch=B)
    fe.configure(threshold=30, octaves=3, max_keypoints=800)
    fe.setCameraModel(0, **EUROC[0])
    L_ = okl.lib()
    try:
        imgs = np.stack([synth_stereo(900 + t, 752, 480)[0] for t in range(B)])
        d_img = torch.from_numpy(imgs).cuda()
        okl.check(L_.okb_detect_describe_batch_device(fe.ctx, 0, B, d_img.data_ptr()))
        cap = C.c_int(0)
        L_.okb_device_features(fe.ctx, 0, None, None, None, C.byref(cap)); cap = cap.value
        feats = []
        for b in range(B):
            kp = np.zeros(cap, okl.KP_DTYPE); d = np.zeros((cap, 64), np.uint8); n = C.c_int(0)
            okl.check(L_.okb_fetch_features(fe.ctx, 0, b, kp.ctypes.data, d.ctypes.data, cap, C.byref(n)))
            feats.append((kp[:n.value], d[:n.value]))
        rng = np.random.default_rng(11)
        # pool from frame 0's descriptors; copied candidates get geometry that triangulates with their source keypoint
        kp0, d0 = feats[0]
        rays0, _ = oracle_bp(EUROC[0], kp0)
        m = map_scene(23, np.zeros((len(d0), 2)), d0, 1500, frac_3d=0.4, flip_p=0.05)
        poses = []
        for b in range(B):
            a = 0.02 * b
            Cm = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
            poses.append((Cm, np.array([0.05 + 0.01 * b, 0.0, 0.002 * b])))
        e0n = rays0 / np.sqrt((rays0[:, 0] ** 2 + rays0[:, 1] ** 2) + rays0[:, 2] ** 2)[:, None]
        ke0 = world_rays(poses[0][0], e0n)
        P = ke0[m["src"][m["cand_lm"]]] * rng.uniform(0.3, 30.0, (len(m["cand_lm"]), 1)) + poses[0][1]
        r = rng.normal(0, 0.4, P.shape)
        e = P - r; e /= np.linalg.norm(e, axis=1, keepdims=True)
        cp = m["is_copy"][m["cand_lm"]] & (rng.random(len(P)) < 0.8)
        m["cand_e_W"][cp] = e[cp]; m["cand_r_W"][cp] = r[cp]
        use = np.zeros((B, cap), np.uint8); prev = np.full((B, cap), -1, np.int32)
        for b in range(B):
            nb = len(feats[b][0])
            use[b, :nb] = rng.random(nb) > 0.05
            if loop:
                prev[b, :nb] = np.where(rng.random(nb) < 0.5, rng.integers(0, len(m["lm_is3d"]), nb), -1)
        T = np.stack([np.concatenate([Cm.reshape(9), rr]) for Cm, rr in poses])
        dev = {k: torch.from_numpy(np.ascontiguousarray(m[k])).cuda() for k in ("cand_desc", "cand_lm", "cand_e_W", "cand_r_W", "lm_is3d")}
        d_use = torch.from_numpy(use).cuda(); d_prev = torch.from_numpy(prev).cuda()
        dist = torch.zeros((B, cap), dtype=torch.int32, device="cuda"); lm = torch.zeros((B, cap), dtype=torch.int32, device="cuda")
        hp = torch.zeros((B, cap, 4), dtype=torch.float64, device="cuda"); ctr = torch.zeros(B, dtype=torch.int32, device="cuda")
        okl.check(L_.okb_match_map_uninit_device(fe.ctx, 0, B, len(m["cand_lm"]), dev["cand_desc"].data_ptr(), dev["cand_lm"].data_ptr(),
                                                 dev["cand_e_W"].data_ptr(), dev["cand_r_W"].data_ptr(), len(m["lm_is3d"]), dev["lm_is3d"].data_ptr(),
                                                 np.ascontiguousarray(T).ctypes.data, 1.0 / 458.0, 60, d_use.data_ptr(),
                                                 d_prev.data_ptr() if loop else None, dist.data_ptr(), lm.data_ptr(), hp.data_ptr(), ctr.data_ptr()))
        okl.check(L_.okb_sync(fe.ctx)); torch.cuda.synchronize()
        dist, lm, hp, ctr = dist.cpu().numpy().view(np.uint32), lm.cpu().numpy(), hp.cpu().numpy(), ctr.cpu().numpy()
        matched = 0
        for b in range(B):
            kp, d = feats[b]; nb = len(kp)
            rays, valid = oracle_bp(EUROC[0], kp)
            en = rays / np.sqrt((rays[:, 0] * rays[:, 0] + rays[:, 1] * rays[:, 1]) + rays[:, 2] * rays[:, 2])[:, None]
            Cm, rr = poses[b]
            ke = np.ascontiguousarray(np.stack([(Cm[i, 0] * en[:, 0] + Cm[i, 1] * en[:, 1]) + Cm[i, 2] * en[:, 2] for i in range(3)], 1))
            ref = oracle.match_map_uninit(d, ke, (use[b, :nb] & valid).astype(np.uint8), prev[b, :nb] if loop else None, m["cand_desc"], m["cand_lm"],
                                          m["cand_e_W"], m["cand_r_W"], m["lm_is3d"], rr, 1.0 / 458.0, 60)
            assert np.array_equal(dist[b, :nb], ref[0]) and np.array_equal(lm[b, :nb], ref[1]), b
            assert np.array_equal(hp[b, :nb].view(np.uint64), ref[2].view(np.uint64)), b
            assert ctr[b] == ref[3], b
            matched += int((ref[1] >= 0).sum())
        assert matched > 20
    finally:
        fe.close()
