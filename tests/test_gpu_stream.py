"""GPU parity of the live path: okb_process_multiframe (one call per stereo frame; direct submission for the first call, capture at the second,
CUDA-graph replay afterwards) against the oracle for every stage -- detect / describe / back-projection, M1, the M3 sequence with
its compact match lists, M4 -- on consecutive frames with changing images, projections and poses."""
import ctypes as C

import numpy as np
import pytest

import oracle
from okvis2_b200 import lib as okl
from okvis2_b200.frontend import Frontend
from okvis2_b200.synth import map_scene, pose12, rot, synth_stereo
from test_gpu_camera import EUROC, T_CW, oracle_bp, world_rays

pytestmark = pytest.mark.gpu
W, H, MAXKP, CAP_M, N_OLDER = 752, 480, 900, 256, 3


def make_views(kp, desc, rays, valid, seed, T_WC1):
    """older views geometrically consistent with the given frame (same construction as bench.py)"""
    rng = np.random.default_rng(seed)
    n = len(kp); r1 = T_WC1[9:]
    e = rays / np.linalg.norm(rays, axis=1, keepdims=True)
    P = r1 + e * np.exp(rng.uniform(np.log(1.5), np.log(25.0), n))[:, None]
    views = []
    for v in range(N_OLDER):
        Cv = rot((0, 1, 0), 0.012 * (v + 1)); rv = r1 + np.array([-0.06 * (v + 1), 0.01 * v, -0.02 * (v + 1)])
        pc = (P - rv) @ Cv
        idx = np.nonzero((valid != 0) & (pc[:, 2] > 0.3) & (rng.random(n) < 0.4))[0]
        bits = np.unpackbits(desc[idx], axis=1)
        d = np.packbits(bits ^ (rng.random(bits.shape) < 0.04).astype(np.uint8), axis=1)
        ry = np.stack([pc[idx, 0] / pc[idx, 2], pc[idx, 1] / pc[idx, 2], np.ones(len(idx))], 1)
        n_out = 300
        d = np.concatenate([d, rng.integers(0, 256, (n_out, 64), dtype=np.uint8)])
        ry = np.concatenate([ry, np.stack([rng.uniform(-0.7, 0.7, n_out), rng.uniform(-0.45, 0.45, n_out), np.ones(n_out)], 1)])
        Tw, Tc = pose12(Cv, rv)
        views.append(dict(desc=np.ascontiguousarray(d), rays=np.ascontiguousarray(ry), valid=(rng.random(len(d)) > 0.02).astype(np.uint8),
                          size=(rng.choice([12.0, 18.0, 24.0], len(d)) * rng.uniform(0.9, 1.1, len(d))).astype(np.float32),
                          use=(rng.random(len(d)) > 0.2).astype(np.uint8), T_WC=Tw, T_CW=Tc))
    return views


def pose_at(c, t):
    """the current camera pose of call t (moves a little every frame: the tables the graph copies must follow)"""
    return pose12(np.eye(3), np.array([0.11 * c + 0.003 * t, 0.001 * t, 0.0]))


def view_order(t):
    """the older keyframes of call t: the window slides (their order changes every frame)"""
    return [(v + t) % N_OLDER for v in range(N_OLDER)]


def run(use_graph, n_calls=6):
    import torch
    fe = Frontend(2, W, H)
    fe.configure(threshold=30, octaves=3, max_keypoints=MAXKP)
    for c in range(2):
        fe.setCameraModel(c, **EUROC[c])
    L_ = okl.lib()
    okl.check(L_.okb_stream_use_graph(fe.ctx, 1 if use_graph else 0))
    frames = [synth_stereo(500 + t % 2, W, H) for t in range(n_calls)]   # two scenes alternate: the graph must follow the inputs
    o = oracle.Brisk(30, 3)
    poses = [pose12(np.eye(3), np.array([0.11 * c, 0.0, 0.0])) for c in range(2)]
    intr = [np.array(list(EUROC[c]["focal_length"]) + list(EUROC[c]["principal_point"]) + list(EUROC[c]["distortion_coefficients"])) for c in range(2)]
    # pools and views from frame 0
    maps, views, dviews, keep = [], [], [], []
    cap0 = 0
    for c in range(2):
        kp, d = o.detect_and_compute(frames[0][c], MAXKP)
        rays, valid = oracle_bp(EUROC[c], kp)
        maps.append(map_scene(7 + c, np.stack([kp["x"], kp["y"]], 1).astype(np.float64), d, 1500, W=W, H=H, frac_near=0.3))
        views.append(make_views(kp, d, rays, valid, 30 + c, poses[c][0]))
        cap0 = max(cap0, max(len(v["desc"]) for v in views[c]))
    cap0 = (cap0 + 63) // 64 * 64
    for c in range(2):
        dviews.append([{k: torch.from_numpy(np.ascontiguousarray(v[k])).cuda() for k in ("desc", "rays", "valid", "size", "use")} for v in views[c]])
    cap = MAXKP
    results = []
    try:
        for t in range(n_calls):
            io = (okl.MultiframeCam * 2)(); bufs = []; tabs = []
            for c in range(2):
                q = io[c]
                tab = (okl.OlderView * N_OLDER)()
                for vi, src in enumerate(view_order(t)):
                    v, dv = views[c][src], dviews[c][src]
                    e = tab[vi]
                    e.d_desc, e.d_rays, e.d_valid, e.d_size, e.d_use = (dv[k].data_ptr() for k in ("desc", "rays", "valid", "size", "use"))
                    e.n = len(v["desc"]); e.T_WC[:] = list(v["T_WC"]); e.T_CW[:] = list(v["T_CW"])
                tabs.append(tab)
                img = np.ascontiguousarray(frames[t][c]); m = maps[c]
                proj = np.ascontiguousarray(m["lm_proj"] + 0.7 * t)    # the camera moves: projections change every frame
                if t % 3 != 0:
                    # page-locked inputs are read in place (the graph's upload nodes are re-pointed); pageable ones are staged: mix them
                    pinned = [torch.from_numpy(img).pin_memory(), torch.from_numpy(proj).pin_memory()]
                    keep.append(pinned)
                    img, proj = pinned[0].numpy(), pinned[1].numpy()
                b = dict(img=img, proj=proj, kp=np.zeros(cap, okl.KP_DTYPE), desc=np.zeros((cap, 64), np.uint8), rays=np.zeros((cap, 3)),
                         valid=np.zeros(cap, np.uint8), m1d=np.zeros(cap, np.uint32), m1l=np.zeros(cap, np.int32), m3n=np.zeros(N_OLDER, np.int32),
                         k0=np.zeros((N_OLDER, CAP_M), np.int32), k1=np.zeros((N_OLDER, CAP_M), np.int32), fl=np.zeros((N_OLDER, CAP_M), np.uint8),
                         hp=np.zeros((N_OLDER, CAP_M, 4)), Tw=np.ascontiguousarray(pose_at(c, t)[0]), Tc=np.ascontiguousarray(pose_at(c, t)[1]),
                         cd=np.ascontiguousarray(m["cand_desc"]), cl=np.ascontiguousarray(m["cand_lm"]), c3=np.ascontiguousarray(m["lm_is3d"]))
                bufs.append(b)
                q.image = img.ctypes.data; q.stride_bytes = W
                q.n_cand = len(b["cl"]); q.n_lm = len(b["c3"]); q.pool_changed = 1 if t == 0 else 0
                q.cand_desc, q.cand_lm, q.lm_is3d, q.lm_proj = b["cd"].ctypes.data, b["cl"].ctypes.data, b["c3"].ctypes.data, proj.ctypes.data
                q.T_WC1, q.T_CW1 = b["Tw"].ctypes.data, b["Tc"].ctypes.data
                q.n_older, q.cap0, q.older = N_OLDER, cap0, C.addressof(tab)
                q.cap = cap; q.kp, q.desc, q.rays, q.rays_valid = (b[k].ctypes.data for k in ("kp", "desc", "rays", "valid"))
                q.m1_dist, q.m1_lm = b["m1d"].ctypes.data, b["m1l"].ctypes.data
                q.cap_m = CAP_M; q.m3_n, q.m3_k0, q.m3_k1, q.m3_flags, q.m3_hp_W = (b[k].ctypes.data for k in ("m3n", "k0", "k1", "fl", "hp"))
            st = okl.MultiframeStereo(); st.cam0, st.cam1 = 0, 1
            st.C_WC0[:] = [1, 0, 0, 0, 1, 0, 0, 0, 1]; st.C_WC1[:] = [1, 0, 0, 0, 1, 0, 0, 0, 1]; st.r_WC0[:] = [0, 0, 0]; st.r_WC1[:] = [0.11, 0, 0]
            sb = dict(k1=np.zeros(cap, np.int32), dist=np.zeros(cap, np.uint32), hp=np.zeros((cap, 4)), init=np.zeros(cap, np.uint8))
            st.k1, st.dist, st.hp_W, st.initialisable = (sb[k].ctypes.data for k in ("k1", "dist", "hp", "init"))
            okl.check(L_.okb_process_multiframe(fe.ctx, 2, io, 1, C.byref(st), 20.0, 60))
            results.append(([dict(b, n=io[c].n) for c, b in enumerate(bufs)], sb))
        g, d = C.c_longlong(), C.c_longlong()
        okl.check(L_.okb_stream_stats(fe.ctx, C.byref(g), C.byref(d)))
    finally:
        fe.close()
    return results, frames, maps, views, poses, intr, (g.value, d.value)


@pytest.mark.parametrize("use_graph", [False, True])
def test_process_multiframe_equals_oracle(use_graph):
    results, frames, maps, views, poses, intr, (n_graph, n_direct) = run(use_graph)
    assert (n_graph, n_direct) == ((5, 1) if use_graph else (0, 6))   # captured at the second call
    o = oracle.Brisk(30, 3)
    C0 = np.eye(3); r0 = np.zeros(3); r1 = np.array([0.11, 0.0, 0.0])
    m3_total = m4_total = m1_total = 0
    for t, (cams, sb) in enumerate(results):
        feats = []
        for c in range(2):
            b = cams[c]; n = b["n"]
            rk, rd = o.detect_and_compute(frames[t][c], MAXKP)
            assert n == len(rk) and b["kp"][:n].tobytes() == rk.tobytes() and np.array_equal(b["desc"][:n], rd), (t, c)
            rays, valid = oracle_bp(EUROC[c], rk)
            assert np.array_equal(b["rays"][:n].view(np.uint64), rays.view(np.uint64)) and np.array_equal(b["valid"][:n], valid)
            m = maps[c]
            xy = np.stack([rk["x"], rk["y"]], 1).astype(np.float64)
            rdist, rlm = oracle.match_map3d(rd, xy, None, m["cand_desc"], m["cand_lm"], b["proj"], m["lm_is3d"], 20.0, 60)
            assert np.array_equal(b["m1d"][:n], rdist.astype(np.uint32)) and np.array_equal(b["m1l"][:n], rlm), (t, c)
            m1_total += int((rlm >= 0).sum())
            ov = [dict(views[c][src]) for src in view_order(t)]
            ref, _ = oracle.match_motion_stereo_sequence(ov, rd, rays, valid, np.stack([rk["x"], rk["y"]], 1), pose_at(c, t)[0], pose_at(c, t)[1], 1, intr[c], W, H, 60,
                                                         (rlm >= 0).astype(np.uint8))
            for v, (k1, dist, hp, fl) in enumerate(ref):
                k0s = np.nonzero(fl & 1)[0]
                assert b["m3n"][v] == len(k0s), (t, c, v)
                assert np.array_equal(b["k0"][v, :len(k0s)], k0s) and np.array_equal(b["k1"][v, :len(k0s)], k1[k0s])
                assert np.array_equal(b["fl"][v, :len(k0s)], fl[k0s]) and np.array_equal(b["hp"][v, :len(k0s)].view(np.uint64), hp[k0s].view(np.uint64))
                m3_total += int(((fl & 4) != 0).sum())
            f = 0.5 * sum(EUROC[c]["focal_length"])
            feats.append((rd, valid, world_rays(C0, rays), rk["size"].astype(np.float64) / f))
        ref = oracle.match_stereo(*feats[0], *feats[1], r0, r1, T_CW(C0, r0), T_CW(C0, r1), 60)
        n0 = cams[0]["n"]
        assert np.array_equal(sb["k1"][:n0], ref[0]) and np.array_equal(sb["dist"][:n0], ref[1])
        assert np.array_equal(sb["hp"][:n0].view(np.uint64), ref[2].view(np.uint64)) and np.array_equal(sb["init"][:n0], ref[3])
        m4_total += int((ref[0] >= 0).sum())
    assert m1_total > 100 and m3_total > 100 and m4_total > 10


def test_process_multiframe_mono_detect_only():
    """the smallest multiframe: one camera, no landmark pool, no older keyframes, no stereo pair (graph captured at the second call)"""
    fe = Frontend(1, W, H)
    fe.configure(threshold=30, octaves=3, max_keypoints=MAXKP)
    L_ = okl.lib()
    o = oracle.Brisk(30, 3)
    try:
        for t in range(4):
            img = np.ascontiguousarray(synth_stereo(520 + t, W, H)[0])
            io = (okl.MultiframeCam * 1)()
            kp = np.zeros(MAXKP, okl.KP_DTYPE); desc = np.zeros((MAXKP, 64), np.uint8)
            q = io[0]
            q.image = img.ctypes.data; q.stride_bytes = W; q.cap = MAXKP; q.kp = kp.ctypes.data; q.desc = desc.ctypes.data
            okl.check(L_.okb_process_multiframe(fe.ctx, 1, io, 0, None, 20.0, 60))
            rk, rd = o.detect_and_compute(img, MAXKP)
            assert io[0].n == len(rk) and kp[:len(rk)].tobytes() == rk.tobytes() and np.array_equal(desc[:len(rk)], rd), t
        g, d = C.c_longlong(), C.c_longlong()
        okl.check(L_.okb_stream_stats(fe.ctx, C.byref(g), C.byref(d)))
        assert (g.value, d.value) == (3, 1)
    finally:
        fe.close()
