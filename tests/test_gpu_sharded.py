"""GPU parity of the camera-sharded path (BASELINE config 4): detect + describe per camera on its own GPU,
okb_export_features -> okb_allgather_features (NCCL behind the C ABI) -> okb_match_stereo_device_ptr on the gathered blocks must
equal (a) the single-GPU okb_match_stereo_device result and (b) the oracle. Two forms: one process driving all visible GPUs
(okb_comm_init_all; runs with a single GPU too: the all-gather is then over one rank) and one process per GPU
(okb_comm_init_rank, world_size 2; skipped with fewer than two GPUs)."""
import ctypes as C
import os
import tempfile

import numpy as np
import pytest

import oracle
from okvis2_b200 import lib as okl, sharding as sh
from okvis2_b200.frontend import Frontend
from okvis2_b200.synth import synth_stereo
from test_gpu_camera import EUROC, T_CW, oracle_bp, world_rays

pytestmark = pytest.mark.gpu
B, W, H, MAXKP = 2, 752, 480, 800
C0 = np.eye(3); R0 = np.zeros(3); C1 = np.eye(3); R1 = np.array([0.11, 0.0, 0.0])


def camera_model(c):
    m = okl.CameraModel(); m.model = 1
    m.fu, m.fv = EUROC[c]["focal_length"]; m.cu, m.cv = EUROC[c]["principal_point"]
    for i in range(4):
        m.k[i] = EUROC[c]["distortion_coefficients"][i]
    return m


def images():
    return [np.stack([synth_stereo(300 + t, W, H)[c] for t in range(B)]) for c in range(2)]


def oracle_stereo(imgs):
    out = []
    o = oracle.Brisk(30, 3)
    for b in range(B):
        (kp0, d0), (kp1, d1) = (o.detect_and_compute(imgs[c][b], MAXKP) for c in range(2))
        rays0, v0 = oracle_bp(EUROC[0], kp0); rays1, v1 = oracle_bp(EUROC[1], kp1)
        f0 = 0.5 * sum(EUROC[0]["focal_length"]); f1 = 0.5 * sum(EUROC[1]["focal_length"])
        out.append((len(kp0),) + tuple(oracle.match_stereo(d0, v0, world_rays(C0, rays0), kp0["size"].astype(np.float64) / f0, d1, v1,
                                                           world_rays(C1, rays1), kp1["size"].astype(np.float64) / f1, R0, R1, T_CW(C0, R0), T_CW(C1, R1), 60)))
    return out


def stereo_from_gathered(L_, ctx, gathered, slots, world, kp_cap, stream):
    """M4 camera 0 -> camera 1 straight from the gathered buffer (block of camera c at [c % world][c // world])."""
    import torch
    o_c, o_k, o_d, blk = sh.block_layout(B, kp_cap)
    base = gathered.data_ptr()
    bi = base + (sh.slot_of(0, world)[0] * slots + sh.slot_of(0, world)[1]) * blk
    bj = base + (sh.slot_of(1, world)[0] * slots + sh.slot_of(1, world)[1]) * blk
    k1 = torch.zeros((B, kp_cap), dtype=torch.int32, device=gathered.device); dist = torch.zeros((B, kp_cap), dtype=torch.int32, device=gathered.device)
    hp = torch.zeros((B, kp_cap, 4), dtype=torch.float64, device=gathered.device); init = torch.zeros((B, kp_cap), dtype=torch.uint8, device=gathered.device)
    m0, m1 = camera_model(0), camera_model(1)
    okl.check(L_.okb_match_stereo_device_ptr(ctx, B, kp_cap, bi + o_k, bi + o_d, bi + o_c, C.addressof(m0), C0.ctypes.data, R0.ctypes.data, kp_cap,
                                             bj + o_k, bj + o_d, bj + o_c, C.addressof(m1), C1.ctypes.data, R1.ctypes.data, 60, stream, k1.data_ptr(),
                                             dist.data_ptr(), hp.data_ptr(), init.data_ptr()))
    return k1, dist, hp, init


def check_against(ref, k1, dist, hp, init):
    total = 0
    for b, (n0, rk1, rdist, rhp, rinit) in enumerate(ref):
        assert np.array_equal(k1[b, :n0], rk1) and np.array_equal(dist[b, :n0].view(np.uint32), rdist)
        assert np.array_equal(hp[b, :n0].view(np.uint64), rhp.view(np.uint64)) and np.array_equal(init[b, :n0], rinit)
        total += int((rk1 >= 0).sum())
    assert total > 5


def test_single_process_all_gpus():
    import torch
    L_ = okl.lib()
    n_dev = min(torch.cuda.device_count(), 2)
    world = n_dev
    devs = (C.c_int * n_dev)(*range(n_dev))
    comm = C.c_void_p()
    okl.check(L_.okb_comm_init_all(n_dev, devs, C.byref(comm)))
    imgs = images()
    slots = sh.slots_per_rank(world, 2)
    fes, local, gathered, kp_cap = [], [], [], None
    try:
        for r in range(world):
            mine = sh.cameras_of(r, world, 2)
            fe = Frontend(len(mine), W, H, device=r, max_batch=B)
            fe.configure(threshold=30, octaves=3, max_keypoints=MAXKP)
            for li, c in enumerate(mine):
                fe.setCameraModel(li, **EUROC[c])
            fes.append((fe, mine))
            cap = C.c_int(0); L_.okb_device_features(fe.ctx, 0, None, None, None, C.byref(cap)); kp_cap = cap.value
            blk = L_.okb_feature_block_bytes(B, kp_cap)
            with torch.cuda.device(r):
                local.append(torch.zeros((slots, blk), dtype=torch.uint8, device=f"cuda:{r}"))
                gathered.append(torch.zeros((world, slots, blk), dtype=torch.uint8, device=f"cuda:{r}"))
        # per rank: detect + export of its cameras (device resident), then ONE all-gather call for all local ranks
        d_imgs = []
        for r, (fe, mine) in enumerate(fes):
            with torch.cuda.device(r):
                for li, c in enumerate(mine):
                    t = torch.from_numpy(imgs[c]).to(f"cuda:{r}"); d_imgs.append(t)
                    okl.check(L_.okb_detect_describe_batch_device(fe.ctx, li, B, t.data_ptr()))
                    okl.check(L_.okb_export_features(fe.ctx, li, B, local[r][c // world].data_ptr()))
                # the gather rides on the LAST local camera's stream: it must see the exports of the other local cameras
                last = len(mine) - 1
                for li in range(last):
                    torch.cuda.ExternalStream(L_.okb_stream(fe.ctx, last)).wait_stream(torch.cuda.ExternalStream(L_.okb_stream(fe.ctx, li)))
        ctxs = (C.c_void_p * world)(*[fe.ctx for fe, _ in fes]); cams = (C.c_int * world)(*[len(m) - 1 for _, m in fes])
        send = (C.c_void_p * world)(*[t.data_ptr() for t in local]); recv = (C.c_void_p * world)(*[t.data_ptr() for t in gathered])
        okl.check(L_.okb_allgather_features(comm, world, ctxs, cams, send, recv, local[0].numel()))
        # M4 on the owner of camera 0 (rank 0) from the gathered blocks
        fe0 = fes[0][0]
        with torch.cuda.device(0):
            st = L_.okb_stream(fe0.ctx, -1)   # the context's match stream
            okl.check(L_.okb_comm_wait(comm, 0, st))
            k1, dist, hp, init = stereo_from_gathered(L_, fe0.ctx, gathered[0], slots, world, kp_cap, st)
            okl.check(L_.okb_sync(fe0.ctx)); torch.cuda.synchronize()
        for r in range(world):   # every rank holds the same gathered bytes
            with torch.cuda.device(r):
                torch.cuda.synchronize()
            assert torch.equal(gathered[r].cpu(), gathered[0].cpu())
        check_against(oracle_stereo(imgs), k1.cpu().numpy(), dist.cpu().numpy(), hp.cpu().numpy(), init.cpu().numpy())
        # the same pair on ONE GPU through okb_match_stereo_device
        fe = Frontend(2, W, H, device=0, max_batch=B)
        try:
            fe.configure(threshold=30, octaves=3, max_keypoints=MAXKP)
            with torch.cuda.device(0):
                ts = []
                for c in range(2):
                    fe.setCameraModel(c, **EUROC[c])
                    ts.append(torch.from_numpy(imgs[c]).cuda())
                    okl.check(L_.okb_detect_describe_batch_device(fe.ctx, c, B, ts[-1].data_ptr()))
                z = lambda s, dt: torch.zeros(s, dtype=dt, device="cuda:0")
                a = (z((B, kp_cap), torch.int32), z((B, kp_cap), torch.int32), z((B, kp_cap, 4), torch.float64), z((B, kp_cap), torch.uint8))
                okl.check(L_.okb_match_stereo_device(fe.ctx, 0, 1, B, C0.ctypes.data, R0.ctypes.data, C1.ctypes.data, R1.ctypes.data, 60,
                                                     *[t.data_ptr() for t in a]))
                okl.check(L_.okb_sync(fe.ctx)); torch.cuda.synchronize()
            for x, y in zip(a, (k1, dist, hp, init)):
                assert torch.equal(x.cpu(), y.cpu()), "sharded result differs from the single-GPU result"
        finally:
            fe.close()
    finally:
        for fe, _ in fes:
            fe.close()
        L_.okb_comm_destroy(comm)


def _rank_main(rank, world, id_path, q):
    import torch
    try:
        torch.cuda.set_device(rank)
        L_ = okl.lib()
        ident = (C.c_uint8 * 128)()
        if rank == 0:
            okl.check(L_.okb_comm_unique_id(ident))
            with open(id_path + ".tmp", "wb") as f:
                f.write(bytes(ident))
            os.replace(id_path + ".tmp", id_path)
        else:
            import time
            for _ in range(600):
                if os.path.exists(id_path):
                    break
                time.sleep(0.05)
            C.memmove(ident, open(id_path, "rb").read(), 128)
        comm = C.c_void_p()
        okl.check(L_.okb_comm_init_rank(world, rank, ident, rank, C.byref(comm)))
        imgs = images()
        fe = Frontend(1, W, H, device=rank, max_batch=B)
        fe.configure(threshold=30, octaves=3, max_keypoints=MAXKP)
        fe.setCameraModel(0, **EUROC[rank])
        cap = C.c_int(0); L_.okb_device_features(fe.ctx, 0, None, None, None, C.byref(cap)); kp_cap = cap.value
        blk = L_.okb_feature_block_bytes(B, kp_cap)
        local = torch.zeros((1, blk), dtype=torch.uint8, device="cuda"); gathered = torch.zeros((world, 1, blk), dtype=torch.uint8, device="cuda")
        t = torch.from_numpy(imgs[rank]).cuda()
        okl.check(L_.okb_detect_describe_batch_device(fe.ctx, 0, B, t.data_ptr()))
        okl.check(L_.okb_export_features(fe.ctx, 0, B, local.data_ptr()))
        ctxs = (C.c_void_p * 1)(fe.ctx); cams = (C.c_int * 1)(0); send = (C.c_void_p * 1)(local.data_ptr()); recv = (C.c_void_p * 1)(gathered.data_ptr())
        okl.check(L_.okb_allgather_features(comm, 1, ctxs, cams, send, recv, blk))
        if rank == 0:
            st = L_.okb_stream(fe.ctx, -1)
            okl.check(L_.okb_comm_wait(comm, 0, st))
            k1, dist, hp, init = stereo_from_gathered(L_, fe.ctx, gathered, 1, world, kp_cap, st)
            okl.check(L_.okb_sync(fe.ctx)); torch.cuda.synchronize()
            check_against(oracle_stereo(imgs), k1.cpu().numpy(), dist.cpu().numpy(), hp.cpu().numpy(), init.cpu().numpy())
        okl.check(L_.okb_sync(fe.ctx)); torch.cuda.synchronize()
        fe.close(); L_.okb_comm_destroy(comm)
        q.put((rank, "ok"))
    except Exception as e:   # report instead of hanging the peer
        import traceback
        q.put((rank, traceback.format_exc()))


def test_one_process_per_gpu_world_size_2():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    with tempfile.TemporaryDirectory() as d:
        ps = [ctx.Process(target=_rank_main, args=(r, 2, os.path.join(d, "nccl_id"), q)) for r in range(2)]
        for p in ps:
            p.start()
        res = [q.get(timeout=300) for _ in ps]
        for p in ps:
            p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
