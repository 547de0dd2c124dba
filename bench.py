#!/usr/bin/env python
"""bench.py -- stereo frames/s for detect + describe + match on N B200s, beside the CPU path on the same box.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W [--impl reference]` prints ONE JSON line.
A step = one pass of the hot path over one batch of synthetic stereo frames:
  detect+describe+back-project of both cameras (batched launches), M1 match-to-map of every frame of both cameras,
  M4 stereo match camera 0 -> camera 1 of every frame.
  * value : inputs (images, landmark pool) already resident in HBM, results left in HBM, device-timed (CUDA events).
  * e2e   : the same work through the reference-facing calls with HOST buffers, per stereo frame (streaming use):
            Frontend.detectAndDescribe per camera (one host thread per camera, like ThreadedSlam.cpp:432-448), then
            Frontend.matchStereo (M4) and Frontend.matchToMapByThread (M1) -- H2D/D2H copies inside the timed region.
  * cpu_baseline / --impl reference : the CPU oracle port (oracle/, restatement of OpenCV-BRISK + the reference's match
            loops) on the host cores, bounded sample. It is the only place this file executes oracle/ code.
N > 1: one process per GPU (torchrun), each rank replays its own independent sequences (BASELINE config 5: replicas, no
data-path collective); time = max over ranks.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # BASELINE.json configs[1]: EuRoC stereo 752x480 stream, 1000 kpts/frame
    "euroc": dict(W=752, H=480, max_kp=1000, threshold=30, octaves=3, n_lm=5000, batch=32, ring=6, f=458.0),
    # BASELINE.json configs[2]: TUM-VI 1024x1024 stereo, 2000 kpts/frame, 50-keyframe landmark set
    "tumvi": dict(W=1024, H=1024, max_kp=2000, threshold=30, octaves=3, n_lm=50000, batch=16, ring=5, f=190.0),
    # BASELINE.json configs[3]: Hilti-2022 5-camera rig, cameras sharded over the GPUs, NCCL all-gather for stereo matching
    "hilti": dict(W=720, H=540, max_kp=700, threshold=30, octaves=3, n_lm=5000, batch=16, ring=8, f=351.0, rig="HILTI_2022"),
}


def make_frames(cfg, n, seed0, base=8):
    """n stereo pairs: `base` rendered scenes + integer-shifted variants (distinct pixels, same statistics)."""
    from okvis2_b200.synth import synth_stereo
    W, H = cfg["W"], cfg["H"]
    scenes = [synth_stereo(seed0 + i, W, H) for i in range(min(base, n))]
    L = np.empty((n, H, W), np.uint8); R = np.empty((n, H, W), np.uint8)
    for i in range(n):
        l, r = scenes[i % len(scenes)]
        s = i // len(scenes)
        L[i] = np.roll(l, (3 * s, 5 * s), (0, 1)); R[i] = np.roll(r, (3 * s, 5 * s), (0, 1))
    return L, R


def make_map(cfg, kp, desc, seed):
    from okvis2_b200.synth import map_scene
    xy = np.stack([kp["x"], kp["y"]], 1).astype(np.float64)
    return map_scene(seed, xy, desc, cfg["n_lm"], W=cfg["W"], H=cfg["H"])


def pinhole_rays(kp, cfg, R_WC, dx=0.0):
    """world-frame unit rays of an ideal pinhole camera (host-side stand-in for Frame::computeBackProjections)."""
    x = (kp["x"].astype(np.float64) - cfg["W"] / 2 - dx) / cfg["f"]
    y = (kp["y"].astype(np.float64) - cfg["H"] / 2) / cfg["f"]
    e = np.stack([x, y, np.ones_like(x)], 1)
    e /= np.linalg.norm(e, axis=1, keepdims=True)
    return np.ascontiguousarray(e @ R_WC.T)


class ClockSampler:
    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            p = [x.strip() for x in line.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
def cpu_detector_name(use_cv2=True):
    if use_cv2:
        try:
            import cv2
            return f"cv2.BRISK {cv2.__version__} (1 OpenCV thread per job)"
        except Exception:
            pass
    return "C restatement of OpenCV-BRISK (oracle/brisk_oracle.c)"


def cpu_run(cfg, L, R, maps, n_threads, use_cv2=True):
    """CPU path over the given stereo frames; returns seconds. One worker per (frame, camera) job, then one per stereo
    frame. Detect/describe by OpenCV's BRISK when cv2 is importable (else the C restatement of it), matchers by the
    oracle's transcription of the reference loops."""
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    cv2 = None
    if use_cv2:
        try:
            import cv2
            cv2.setNumThreads(1)
        except Exception:
            cv2 = None
    n = len(L)
    jobs = [(i, c) for i in range(n) for c in range(2)]
    local = threading.local()

    feats = {}

    def work(job):
        i, c = job
        if cv2 is not None:
            # OpenCV's own BRISK (the implementation the oracle restates): detect, keep the N strongest, compute
            if not hasattr(local, "brisk"):
                local.brisk = cv2.BRISK_create(cfg["threshold"], cfg["octaves"], 1.0)
            img = (L, R)[c][i]
            kps = local.brisk.detect(img, None)
            if len(kps) > cfg["max_kp"]:
                kps = sorted(kps, key=lambda k: -k.response)[:cfg["max_kp"]]
            kps, d = local.brisk.compute(img, kps)
            kp = np.zeros(len(kps), oracle.KP_DTYPE)
            kp["x"] = [k.pt[0] for k in kps]; kp["y"] = [k.pt[1] for k in kps]; kp["size"] = [k.size for k in kps]
            if d is None:
                d = np.zeros((0, 64), np.uint8)
        else:
            if not hasattr(local, "brisk"):
                local.brisk = oracle.Brisk(cfg["threshold"], cfg["octaves"])
            kp, d = local.brisk.detect_and_compute((L, R)[c][i], cfg["max_kp"])
        m = maps[c]
        xy = np.stack([kp["x"], kp["y"]], 1).astype(np.float64)
        oracle.match_map3d(d, xy, None, m["cand_desc"], m["cand_lm"], m["lm_proj"], m["lm_is3d"], 20.0, 60, 1)
        feats[job] = (kp, d)
        return len(kp)

    W, H, f = cfg["W"], cfg["H"], cfg["f"]
    r = [np.zeros(3), np.array([0.11, 0.0, 0.0])]
    T = [np.array([1, 0, 0, -r[c][0], 0, 1, 0, -r[c][1], 0, 0, 1, -r[c][2]], np.float64) for c in range(2)]

    def stereo(i):   # Frame::computeBackProjections + Frontend::matchStereo of stereo frame i (same camera models as the GPU arm)
        side = []
        for c in range(2):
            kp, d = feats[(i, c)]
            rays, valid = oracle.back_project(1, f, f * 0.997, W / 2 - 8.8 + 12 * c, H / 2 + 8.4 + 7 * c, [-0.2834, 0.0740, 0.00019, 1.76e-05], kp)
            e = rays / np.sqrt((rays[:, 0] * rays[:, 0] + rays[:, 1] * rays[:, 1]) + rays[:, 2] * rays[:, 2])[:, None]
            side.append((d, valid, np.ascontiguousarray(e), kp["size"].astype(np.float64) / f))
        (d0, v0, e0, s0), (d1, v1, e1, s1) = side
        oracle.match_stereo(d0, v0, e0, s0, d1, v1, e1, s1, r[0], r[1], T[0], T[1], 60)

    t0 = time.perf_counter()
    with ThreadPoolExecutor(n_threads) as ex:
        list(ex.map(work, jobs))
        list(ex.map(stereo, range(n)))
    return time.perf_counter() - t0


def run_reference(args, cfg, rank, world):
    """--impl reference: the CPU path (oracle port; the real reference front-end cannot be built here)."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = max(2, min(16, cores // 2))   # bounded sample: stereo frames per step
    L, R = make_frames(cfg, n, 1000)
    import oracle
    o = oracle.Brisk(cfg["threshold"], cfg["octaves"])
    maps = []
    for c, img in enumerate((L[0], R[0])):
        kp, d = o.detect_and_compute(img, cfg["max_kp"])
        maps.append(make_map(cfg, kp, d, 40 + c))
    for _ in range(max(args.warmup, 1)):
        cpu_run(cfg, L[:2], R[:2], maps, cores)
    t = 0.0
    for _ in range(args.steps):
        t += cpu_run(cfg, L, R, maps, cores)
    value = n * args.steps / t
    t_port = cpu_run(cfg, L, R, maps, cores, use_cv2=False)
    line = {"impl": "reference", "metric": "stereo frames/sec detect+describe+match", "value": value, "unit": "stereo frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": args.config, "sample": f"{n} stereo frames per step", **{k: cfg[k] for k in ("W", "H", "max_kp", "threshold", "octaves", "n_lm")}},
            "cpu_baseline": {"value": value, "unit": "stereo frames/s", "cores": cores, "kind": "port",
                             "sample": f"{n} stereo frames x {args.steps} steps on {cores} threads: detect+describe by {cpu_detector_name()}, M1 per camera + back-projection + M4 by the oracle's transcription of the reference loops",
                             "port_only_value": n / t_port},
            "e2e": {"value": value, "unit": "stereo frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
def run_sharded(args, cfg, rank, world, local_rank):
    """Camera-sharded multiframe pipeline (BASELINE config 4): camera c on rank c % world; per step every rank runs
    detect+describe+back-project+M1 for its cameras on a batch of B multiframes, contributes its fixed-capacity feature
    blocks to ONE NCCL all-gather, then stereo-matches (M4) the overlapping pairs it owns from the gathered device buffer."""
    import torch
    import torch.distributed as dist
    from okvis2_b200 import lib as okl, rigs, sharding as sh
    from okvis2_b200.frontend import Frontend, MultiFrame
    from okvis2_b200.synth import synth_frame
    L_ = okl.lib()
    rig = getattr(rigs, cfg["rig"])
    n_cams, W, H, B, ring = len(rig), cfg["W"], cfg["H"], cfg["batch"], cfg["ring"]
    mine = sh.cameras_of(rank, world, n_cams)
    overlaps = sh.rig_overlaps(rig)
    pairs = sh.pairs_of(rank, world, overlaps)
    warm = max(args.warmup, 3)
    fe = Frontend(max(len(mine), 1), W, H, device=local_rank, max_batch=B)
    fe.configure(threshold=cfg["threshold"], octaves=cfg["octaves"], max_keypoints=cfg["max_kp"])
    ctx = fe.ctx
    models = []
    for c in range(n_cams):
        m = okl.CameraModel(); r = rig[c]
        m.model = Frontend.MODELS[r["distortion_type"]]; m.fu, m.fv = r["focal_length"]; m.cu, m.cv = r["principal_point"]
        for i in range(4):
            m.k[i] = r["distortion_coefficients"][i]
        models.append(m)
    for li, c in enumerate(mine):
        fe.setCameraModel(li, rig[c]["distortion_type"], rig[c]["focal_length"], rig[c]["principal_point"], rig[c]["distortion_coefficients"])
    T = [np.array(r["T_SC"]).reshape(4, 4) for r in rig]
    C_WC = [np.ascontiguousarray(t[:3, :3]) for t in T]; r_WC = [np.ascontiguousarray(t[:3, 3]) for t in T]
    cap = C.c_int(0); L_.okb_device_features(ctx, 0, None, None, None, C.byref(cap)); kp_cap = cap.value
    # inputs: ring * B frames per local camera (distinct synthetic views), resident in HBM
    n_frames = ring * B
    d_img, maps, d_maps, d_m1 = [], [], [], []
    for li, c in enumerate(mine):
        # every camera sees the same four scenes, displaced horizontally by 14 px per camera index (so that overlapping
        # pairs share content and the stereo matcher's gates run on real candidates), then drifting frame to frame
        base = [np.roll(synth_frame(3000 + i, W, H), 14 * c, 1) for i in range(4)]
        frames = np.stack([np.roll(base[i % 4], (3 * (i // 4), 5 * (i // 4)), (0, 1)) for i in range(n_frames)])
        d_img.append(torch.from_numpy(frames).cuda())
        mf = MultiFrame(len(mine)); mf.setImage(li, frames[0]); fe.detectAndDescribe(li, mf)
        m = make_map(cfg, mf.frames[li].keypoints, mf.frames[li].descriptors, 70 + c)
        proj = np.broadcast_to(m["lm_proj"], (B,) + m["lm_proj"].shape).copy()
        d_maps.append(dict(desc=torch.from_numpy(m["cand_desc"]).cuda(), lm=torch.from_numpy(m["cand_lm"]).cuda(),
                           proj=torch.from_numpy(proj).cuda(), is3d=torch.from_numpy(m["lm_is3d"]).cuda()))
        d_m1.append(dict(dist=torch.zeros((B, kp_cap), dtype=torch.int32, device="cuda"), lm=torch.zeros((B, kp_cap), dtype=torch.int32, device="cuda")))
    slots = sh.slots_per_rank(world, n_cams)
    o_c, o_k, o_d, blk = sh.block_layout(B, kp_cap)
    assert blk == L_.okb_feature_block_bytes(B, kp_cap)
    local = torch.zeros((slots, blk), dtype=torch.uint8, device="cuda")
    gathered = torch.zeros((world, slots, blk), dtype=torch.uint8, device="cuda")
    d_st = [dict(k1=torch.zeros((B, kp_cap), dtype=torch.int32, device="cuda"), dist=torch.zeros((B, kp_cap), dtype=torch.int32, device="cuda"),
                 hp=torch.zeros((B, kp_cap, 4), dtype=torch.float64, device="cuda"), init=torch.zeros((B, kp_cap), dtype=torch.uint8, device="cuda"))
            for _ in pairs]
    cur = torch.cuda.current_stream()
    cam_streams = [torch.cuda.ExternalStream(L_.okb_stream(ctx, li)) for li in range(len(mine))]

    def step(s):
        for li, c in enumerate(mine):
            cam_streams[li].wait_stream(cur)      # the previous step's matchers have finished reading the features
            frames = d_img[li][(s % ring) * B:(s % ring + 1) * B]
            okl.check(L_.okb_detect_describe_batch_device(ctx, li, B, frames.data_ptr()))
            dm = d_maps[li]
            okl.check(L_.okb_match_map3d_device(ctx, li, 64, B, len(dm["lm"]), dm["desc"].data_ptr(), dm["lm"].data_ptr(), len(dm["is3d"]),
                                                dm["proj"].data_ptr(), dm["is3d"].data_ptr(), 20.0, 60, d_m1[li]["dist"].data_ptr(),
                                                d_m1[li]["lm"].data_ptr()))
            okl.check(L_.okb_export_features(ctx, li, B, local[c // world].data_ptr()))
            cur.wait_stream(cam_streams[li])
        if world > 1:
            dist.all_gather_into_tensor(gathered.view(-1), local.view(-1))      # the one collective of the path
        else:
            gathered[0].copy_(local)
        for pi, (i, j) in enumerate(pairs):
            bi = gathered.data_ptr() + (sh.slot_of(i, world)[0] * slots + sh.slot_of(i, world)[1]) * blk
            bj = gathered.data_ptr() + (sh.slot_of(j, world)[0] * slots + sh.slot_of(j, world)[1]) * blk
            o = d_st[pi]
            okl.check(L_.okb_match_stereo_device_ptr(ctx, B, kp_cap, bi + o_k, bi + o_d, bi + o_c, C.byref(models[i]), C_WC[i].ctypes.data,
                                                     r_WC[i].ctypes.data, kp_cap, bj + o_k, bj + o_d, bj + o_c, C.byref(models[j]),
                                                     C_WC[j].ctypes.data, r_WC[j].ctypes.data, 60, cur.cuda_stream, o["k1"].data_ptr(),
                                                     o["dist"].data_ptr(), o["hp"].data_ptr(), o["init"].data_ptr()))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank); sampler.start()
    for s in range(warm):
        step(s)
    barrier()
    launches0 = L_.okb_launch_count(ctx)
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record(cur)
    for s in range(args.steps):
        step(warm + s)
    for st in cam_streams:
        cur.wait_stream(st)
    ev1.record(cur)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = L_.okb_launch_count(ctx) - launches0
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    n_match = int(sum((o["k1"] >= 0).sum().item() for o in d_st))
    if rank == 0:
        line = {"metric": "multiframes/sec detect+describe+match (5-camera rig)", "value": B * args.steps / (ms * 1e-3), "unit": "multiframes/s",
                "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": args.config, "multiframes_per_step": B, "cameras": n_cams, "overlapping_pairs": overlaps,
                           "parallelism": f"camera c on GPU c % {world}; one NCCL all-gather of {slots} x {blk} B feature blocks per rank per step",
                           "l2_policy": f"inputs larger than L2: ring of {ring} batches",
                           **{k: cfg[k] for k in ("W", "H", "max_kp", "threshold", "octaves", "n_lm")}},
                "roofline": None, "cpu_baseline": None, "e2e": None, "gpu_launches": int(launches), "clocks": clocks,
                "stereo_matches_rank0_last_step": n_match}
        print(json.dumps(line))
    fe.close()


# ---------------------------------------------------------------------------------------------------------------
def bench_next_rows(fe, cfg):
    """P1 (Frontend.cpp:1196-1360) on a 50 000-landmark map: the C-ABI call (H2D of the map, kernels, D2H of the pool)
    against the oracle transcription on one host core; outputs compared."""
    import oracle
    from okvis2_b200.synth import landmark_scene
    s = landmark_scene(77, n_lm=50000, n_slots=50, n_cams=2, n_kp=cfg["max_kp"], W=cfg["W"], H=cfg["H"], f=cfg["f"], step=0.05)
    fe.configureFeatureStore(s["n_slots"], 64)
    for t in range(s["n_slots"] * 2):
        fe.storeFrame(t // 2, t % 2, s["desc_tab"][t], s["ray_tab"][t])
    call = lambda: fe.prepareLandmarksToMatch(0, s["T_WC1"], s["T_CW1"], s["W"], s["H"], s["hp_W"], s["quality"], s["obs_begin"],
                                              s["obs"], s["T_WC_old"])
    got = call()
    t0 = time.perf_counter()
    for _ in range(5):
        got = call()
    gpu_ms = (time.perf_counter() - t0) / 5 * 1e3
    intr = np.array([cfg["f"], cfg["f"] * 0.997, s["W"] / 2 - 8.8, s["H"] / 2 + 8.4, -0.2834, 0.0740, 0.00019, 1.76e-05])
    t0 = time.perf_counter()
    ref = oracle.prepare_landmarks(s["hp_W"], s["quality"], s["obs_begin"], s["obs"], 2, s["T_WC_old"], s["desc_tab"], s["ray_tab"], 64,
                                   s["T_WC1"], s["T_CW1"], 1, intr, s["W"], s["H"])
    cpu_ms = (time.perf_counter() - t0) * 1e3
    same = all(np.array_equal(got[k], ref[k]) for k in ("lm", "lm_is3d", "desc_begin", "cand_desc", "kid", "lm_proj"))
    rows = {"P1_prepare_landmarks": {"landmarks": 50000, "observations": int(len(s["obs"])), "kept": int(len(got["lm"])),
                                     "pool_rows": int(len(got["cand_desc"])), "gpu_ms_host_buffers": gpu_ms, "cpu_port_ms_1_core": cpu_ms,
                                     "identical_to_oracle": bool(same)}}
    # K1 (Frontend.cpp:1058-1167): the masks of the current frame + 10 keyframes, 2 cameras each, max_kp keypoints per view
    rng = np.random.default_rng(3)
    views = []
    for _ in range(22):
        xy = np.stack([rng.uniform(0, cfg["W"], cfg["max_kp"]), rng.uniform(0, cfg["H"], cfg["max_kp"])], 1).astype(np.float32)
        views.append((cfg["H"], cfg["W"], xy, rng.random(cfg["max_kp"]) < 0.5))
    inter, uni = fe._overlap_counts(views)
    t0 = time.perf_counter()
    for _ in range(10):
        inter, uni = fe._overlap_counts(views)
    k1_gpu = (time.perf_counter() - t0) / 10 * 1e3
    t0 = time.perf_counter()
    ref = [oracle.overlap_counts(r, c, xy, m) for (r, c, xy, m) in views]
    k1_cpu = (time.perf_counter() - t0) * 1e3
    rows["K1_keyframe_overlap"] = {"views": len(views), "keypoints_per_view": cfg["max_kp"], "gpu_ms_host_buffers": k1_gpu,
                                   "cpu_port_ms_1_core": k1_cpu,
                                   "identical_to_oracle": bool(all((inter[i], uni[i]) == ref[i] for i in range(len(views))))}
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="euroc", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-lanes", "--lanes", dest="e2e_lanes", type=int, default=4,
                    help="independent sequences in flight per GPU (own library handle each) in the value and e2e legs")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from okvis2_b200 import lib as okl
    from okvis2_b200.frontend import Frontend, MultiFrame
    L_ = okl.lib()
    if "rig" in cfg:
        run_sharded(args, cfg, rank, world, local_rank)
        if world > 1:
            dist.destroy_process_group()
        return

    W, H, B, ring = cfg["W"], cfg["H"], cfg["batch"], cfg["ring"]
    warm = max(args.warmup, 3)
    fe = Frontend(2, W, H, device=local_rank, max_batch=B)
    fe.configure(threshold=cfg["threshold"], octaves=cfg["octaves"], max_keypoints=cfg["max_kp"])
    for c in range(2):   # radial-tangential pinhole cameras (EuRoC-like intrinsics scaled to the image size)
        fe.setCameraModel(c, "radialtangential", (cfg["f"], cfg["f"] * 0.997), (W / 2 - 8.8 + 12 * c, H / 2 + 8.4 + 7 * c),
                          [-0.2834, 0.0740, 0.00019, 1.76e-05])
    ctx = fe.ctx
    # further sequences in flight on this GPU: one library handle (own streams and workspaces) each. The single-CTA-per-frame
    # kernels of one sequence (tie resolution, selection: 200 us on 32 SMs) leave most SMs to the others.
    lanes = max(1, args.e2e_lanes)
    fes = [fe]
    for l in range(1, lanes):
        f2 = Frontend(2, W, H, device=local_rank, max_batch=B)
        f2.configure(threshold=cfg["threshold"], octaves=cfg["octaves"], max_keypoints=cfg["max_kp"])
        for c in range(2):
            f2.setCameraModel(c, "radialtangential", (cfg["f"], cfg["f"] * 0.997), (W / 2 - 8.8 + 12 * c, H / 2 + 8.4 + 7 * c),
                              [-0.2834, 0.0740, 0.00019, 1.76e-05])
        fes.append(f2)
    C_WC = [np.eye(3), np.eye(3)]; r_WC = [np.zeros(3), np.array([0.11, 0.0, 0.0])]
    # ---- synthetic inputs: ring * B stereo frames per rank (ring * B * 2 * W * H bytes > L2 so steps do not hit in L2)
    n_frames = ring * B
    Lh, Rh = make_frames(cfg, n_frames, 1000 + 100 * rank)
    in_bytes = 2 * n_frames * W * H
    d_img = [torch.from_numpy(Lh).cuda(), torch.from_numpy(Rh).cuda()]
    # landmark pools (one per camera) built from the features of frame 0, resident in HBM
    mf = MultiFrame(2); maps = []; d_maps = []
    for c, img in enumerate((Lh[0], Rh[0])):
        mf.setImage(c, img); fe.detectAndDescribe(c, mf)
        fr = mf.frames[c]
        m = make_map(cfg, fr.keypoints, fr.descriptors, 40 + c); maps.append(m)
        proj = np.broadcast_to(m["lm_proj"], (B,) + m["lm_proj"].shape).copy()
        d_maps.append(dict(desc=torch.from_numpy(m["cand_desc"]).cuda(), lm=torch.from_numpy(m["cand_lm"]).cuda(),
                           proj=torch.from_numpy(proj).cuda(), is3d=torch.from_numpy(m["lm_is3d"]).cuda()))
    cap = C.c_int(0)
    L_.okb_device_features(ctx, 0, None, None, None, C.byref(cap))
    kp_cap = cap.value
    mk_out = lambda: [dict(dist=torch.zeros((B, kp_cap), dtype=torch.int32, device="cuda"), lm=torch.zeros((B, kp_cap), dtype=torch.int32, device="cuda")) for _ in range(2)]
    mk_st = lambda: dict(k1=torch.zeros((B, kp_cap), dtype=torch.int32, device="cuda"), dist=torch.zeros((B, kp_cap), dtype=torch.int32, device="cuda"),
                         hp=torch.zeros((B, kp_cap, 4), dtype=torch.float64, device="cuda"), init=torch.zeros((B, kp_cap), dtype=torch.uint8, device="cuda"))
    lane_out = [mk_out() for _ in range(lanes)]; lane_st = [mk_st() for _ in range(lanes)]
    lane_streams = [[torch.cuda.ExternalStream(L_.okb_stream(f.ctx, c)) for c in range(2)] for f in fes]
    streams = lane_streams[0]

    chain = [torch.cuda.Event() for _ in range(2)]
    serialize = [False]

    def device_step(s, lane=0):
        cx = fes[lane].ctx; d_out = lane_out[lane]; d_st = lane_st[lane]; st = lane_streams[lane]
        for c in range(2):
            # the two camera streams run concurrently: the latency-bound single-CTA-per-frame kernels of one camera
            # (tie resolution, selection) leave SMs free for the other camera's wide kernels
            if serialize[0]:
                st[c].wait_event(chain[1 - c])
            frames = d_img[c][(s % ring) * B:(s % ring + 1) * B]
            okl.check(L_.okb_detect_describe_batch_device(cx, c, B, frames.data_ptr()))
            dm = d_maps[c]
            okl.check(L_.okb_match_map3d_device(cx, c, 64, B, len(dm["lm"]), dm["desc"].data_ptr(), dm["lm"].data_ptr(),
                                                len(dm["is3d"]), dm["proj"].data_ptr(), dm["is3d"].data_ptr(), 20.0, 60,
                                                d_out[c]["dist"].data_ptr(), d_out[c]["lm"].data_ptr()))
            chain[c].record(st[c])
        # M4: stereo matching camera 0 -> camera 1 of every frame of the batch (back-projection on the device)
        okl.check(L_.okb_match_stereo_device(cx, 0, 1, B, C_WC[0].ctypes.data, r_WC[0].ctypes.data, C_WC[1].ctypes.data,
                                             r_WC[1].ctypes.data, 60, d_st["k1"].data_ptr(), d_st["dist"].data_ptr(),
                                             d_st["hp"].data_ptr(), d_st["init"].data_ptr()))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident, device-timed
    sampler = ClockSampler(local_rank); sampler.start()   # samples clocks through the warm-up, value, roofline and e2e legs
    for s in range(warm * lanes):
        device_step(s, s % lanes)
    for f in fes:
        okl.check(L_.okb_sync(f.ctx))
    barrier()
    launches0 = sum(L_.okb_launch_count(f.ctx) for f in fes)
    # all streams are idle here: the event on lane 0 precedes every kernel of the timed region; step s is the next batch of
    # sequence s % lanes (exactly args.steps steps in total)
    ev0 = torch.cuda.Event(enable_timing=True); ev0.record(streams[0])
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(2 * lanes)]
    for s in range(args.steps):
        device_step(warm * lanes + s, s % lanes)
    for l in range(lanes):
        for c in range(2):
            ends[2 * l + c].record(lane_streams[l][c])
    barrier()
    dev_ms = max(ev0.elapsed_time(e) for e in ends)
    launches = sum(L_.okb_launch_count(f.ctx) for f in fes) - launches0
    # ---- kernel-level timing for the roofline: a few extra steps with the two camera streams serialized, so that the
    #      CUDA events around the pyramid+score launches (recorded on the launching stream inside the library) time those
    #      kernels alone and not whatever the other camera's stream runs next to them
    roof_steps = 5
    serialize[0] = True
    device_step(warm + args.steps); okl.check(L_.okb_sync(ctx))
    L_.okb_enable_timers(ctx, 1); L_.okb_reset_timers(ctx)
    for s in range(roof_steps):
        device_step(warm + args.steps + 1 + s)
    okl.check(L_.okb_sync(ctx))
    serialize[0] = False
    ps_ms = C.c_double(); ps_l = C.c_int64(); tot = C.c_double()
    ps_total_ms, ps_total_launches, score_total_ms = 0.0, 0, 0.0
    sc_ms = C.c_double()
    for c in range(2):
        L_.okb_get_timers(ctx, c, C.byref(ps_ms), C.byref(ps_l), C.byref(tot))
        L_.okb_get_score_kernel_ms(ctx, c, C.byref(sc_ms))
        ps_total_ms += ps_ms.value; ps_total_launches += ps_l.value; score_total_ms += sc_ms.value
    L_.okb_enable_timers(ctx, 0)
    if world > 1:
        t = torch.tensor([dev_ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); dev_ms = float(t.item())
    value = world * B * args.steps / (dev_ms * 1e-3)

    # ---- roofline of the pyramid+score pass (all its launches: resize x3 + score), device time from CUDA events
    ps_bytes = L_.okb_pyramid_score_bytes(ctx, 0)       # algorithmic bytes per image (actual layer sizes)
    passes = 2 * roof_steps                            # one pass per camera per step, B images each
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    ach = ps_bytes * B * passes / (ps_total_ms * 1e-3) / 1e9 if ps_total_ms > 0 else 0.0
    # DRAM traffic of the dominant kernel (k_score_nms, one launch = B images) from the committed ncu --set full capture
    # profiles/r01_ncu_full_v10_score_nms_raw.csv (dram__bytes_read.sum 22.218 MB + dram__bytes_write.sum 0.077 MB), euroc config
    traffic = 22.30e6 if args.config == "euroc" and B == 32 else None
    ach_k = ps_bytes * B * passes / (score_total_ms * 1e-3) / 1e9 if score_total_ms > 0 else 0.0
    # the limiter that actually binds k_score_nms: 81 VIMNMX3.U16x2 per pixel pair on the ALU pipe, which issues one warp
    # instruction per 2 cycles per SM sub-partition (bench/ubench_pipes.cu) -> floor time per launch at the sampled SM clock
    sm_hz = 1.965e9
    pairs_per_launch = (ps_bytes / 2) / 2 * B           # scored pixels of all layers / 2
    alu_floor_ms = 81 * (pairs_per_launch / 32) * 2 / (148 * 4) / sm_hz * 1e3
    roofline = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "kernel": "pyramid+score pass (k_resize launches + k_score_nms, TMA-staged tiles)", "bytes_per_image": int(ps_bytes),
                "dominant_kernel": {"name": "k_score_nms", "ms_per_launch": score_total_ms / passes, "achieved_GBps": ach_k,
                                    "frac": ach_k / peak, "alu_floor_ms": alu_floor_ms,
                                    "alu_frac": (alu_floor_ms / (score_total_ms / passes)) if score_total_ms > 0 else None, "limiter": "ALU pipe (79% busy, 64 lanes/clk/SM: 81 VIMNMX3.U16x2 per pixel pair = 49 us floor per 32 frames), not HBM"},
                "images_per_pass": B, "ms_per_pass": ps_total_ms / passes, "launches_per_pass": ps_total_launches / passes,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                "frac_of_nominal_8000_GBps": ach / 8000.0,
                "measured": f"CUDA events on the launching stream, {roof_steps} extra steps right after the timed region with the two camera streams serialized (they overlap in the timed region)"}

    # more waiting host threads than cores (many ranks per box): sleep instead of spinning while a batch is on the device
    host_threads = world * lanes * 3
    blocking = host_threads > (os.cpu_count() or 1)
    for f in fes:
        L_.okb_set_blocking_sync(f.ctx, 1 if blocking else 0)
    # ---- e2e: HOST buffers through the C ABI, per stereo frame (streaming use), driven by the C++ host loop of
    #      bench/e2e_driver.cpp (what an integrator of the library writes; one host thread per camera for detection,
    #      ThreadedSlam.cpp:432-448; then okb_match_stereo and okb_match_map3d). All H2D/D2H copies are inside.
    e2e_frames = min(n_frames, max(16, 4 * B))
    drv = C.CDLL(os.path.join(ROOT, "bench", "libokb_e2e.so"))
    PP = C.POINTER(C.c_void_p)
    arr_i = lambda v: (C.c_int * 2)(*v)
    arr_p = lambda v: (C.c_void_p * 2)(*[x.ctypes.data for x in v])
    keep = [[np.ascontiguousarray(m[k]) for m in maps] for k in ("cand_desc", "cand_lm", "lm_proj", "lm_is3d")]
    sec = C.c_double(); h2d = C.c_longlong(); d2h = C.c_longlong(); nkp = C.c_longlong(); nm = C.c_longlong()
    drv.okb_e2e_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_double,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    barrier()
    rc = drv.okb_e2e_run(ctx, e2e_frames, 4, W, H, Lh.ctypes.data, Rh.ctypes.data, kp_cap, cfg["f"],
                         arr_i([len(x) for x in keep[1]]), arr_p(keep[0]), arr_p(keep[1]), arr_i([len(x) for x in keep[3]]),
                         arr_p(keep[2]), arr_p(keep[3]), C.byref(sec), C.byref(h2d), C.byref(d2h), C.byref(nkp), C.byref(nm))
    okl.check(rc)
    e2e_s = sec.value
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); e2e_s = float(t.item())
    streaming = {"value": world * e2e_frames / e2e_s, "unit": "stereo frames/s", "h2d_bytes_per_frame": int(h2d.value / e2e_frames),
                 "d2h_bytes_per_frame": int(d2h.value / e2e_frames), "ms_per_stereo_frame": 1e3 * e2e_s / e2e_frames,
                 "step": "one stereo frame per call (live use): 2x okb_detect_describe (one host thread per camera) + okb_match_stereo + 2x okb_match_map3d, host buffers",
                 "frames": e2e_frames, "keypoints_per_frame": nkp.value / e2e_frames / 2, "matches_per_frame": nm.value / e2e_frames}

    # ---- e2e proper: the SAME step as `value` (a batch of B stereo frames: detect+describe both cameras, M1 per camera,
    #      M4) through the host-buffer C ABI (okb_detect_describe_batch / okb_match_map3d_batch / okb_match_stereo_batch)
    #      from page-locked host memory; every H2D / D2H copy of the step is inside the timed region
    class ReplayIO(C.Structure):
        _fields_ = [("n_steps", C.c_int32), ("warmup", C.c_int32), ("batch", C.c_int32), ("ring", C.c_int32), ("W", C.c_int32),
                    ("H", C.c_int32), ("cap", C.c_int32), ("pad_", C.c_int32),
                    ("img", C.c_void_p * 2), ("kp", C.c_void_p * 2), ("desc", C.c_void_p * 2), ("n", C.c_void_p * 2),
                    ("n_cand", C.c_int32 * 2), ("n_lm", C.c_int32 * 2),
                    ("cand_desc", C.c_void_p * 2), ("cand_lm", C.c_void_p * 2), ("lm_proj", C.c_void_p * 2), ("lm_is3d", C.c_void_p * 2),
                    ("m1_dist", C.c_void_p * 2), ("m1_lm", C.c_void_p * 2),
                    ("k1", C.c_void_p), ("sdist", C.c_void_p), ("hp", C.c_void_p), ("init", C.c_void_p),
                    ("seconds", C.c_double), ("h2d", C.c_longlong), ("d2h", C.c_longlong), ("nkp", C.c_longlong), ("nm", C.c_longlong),
                    ("lane", C.c_int32), ("lanes", C.c_int32)]
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    pz = lambda n, dt: torch.zeros(n, dtype=dt).pin_memory()
    h_img = [pin(Lh), pin(Rh)]
    h_map = dict(cand_desc=[pin(m["cand_desc"]) for m in maps], cand_lm=[pin(m["cand_lm"]) for m in maps],
                 lm_proj=[pin(np.broadcast_to(m["lm_proj"], (B,) + m["lm_proj"].shape)) for m in maps],
                 lm_is3d=[pin(m["lm_is3d"]) for m in maps])

    def make_io(n_steps, lane, lanes):
        hold = dict(img=h_img, kp=[pz(B * kp_cap * 28, torch.uint8) for _ in range(2)],
                    desc=[pz(B * kp_cap * 64, torch.uint8) for _ in range(2)], n=[pz(B, torch.int32) for _ in range(2)],
                    m1_dist=[pz(B * kp_cap, torch.int32) for _ in range(2)], m1_lm=[pz(B * kp_cap, torch.int32) for _ in range(2)],
                    k1=pz(B * kp_cap, torch.int32), sdist=pz(B * kp_cap, torch.int32), hp=pz(B * kp_cap * 4, torch.float64),
                    init=pz(B * kp_cap, torch.uint8), **h_map)
        io = ReplayIO(n_steps=n_steps, warmup=warm, batch=B, ring=ring, W=W, H=H, cap=kp_cap, lane=lane, lanes=lanes)
        for k in ("img", "kp", "desc", "n", "cand_desc", "cand_lm", "lm_proj", "lm_is3d", "m1_dist", "m1_lm"):
            for c in range(2):
                getattr(io, k)[c] = hold[k][c].data_ptr()
        for c in range(2):
            io.n_cand[c] = len(maps[c]["cand_lm"]); io.n_lm[c] = len(maps[c]["lm_is3d"])
        io.k1, io.sdist, io.hp, io.init = (hold[k].data_ptr() for k in ("k1", "sdist", "hp", "init"))
        return io, hold

    # (i) one replay alone: every call returns its results before the next batch is submitted
    io, hold = make_io(args.steps, 0, 0)
    drv.okb_e2e_replay.argtypes = [C.c_void_p, C.c_void_p]
    barrier()
    okl.check(drv.okb_e2e_replay(ctx, C.byref(io)))
    rep_s = io.seconds
    # (ii) `lanes` independent sequences replayed concurrently on this GPU (BASELINE config 5 interleaves sequences), each
    #      through its own library handle and host-thread pair: exactly args.steps steps in total, split over the lanes.
    #      One lane's H2D / D2H copies overlap the other lane's kernels.
    per_lane = [args.steps // lanes + (1 if l < args.steps % lanes else 0) for l in range(lanes)]
    ios = [make_io(per_lane[l], l, lanes) for l in range(lanes)]
    ctx_arr = (C.c_void_p * lanes)(*[f.ctx for f in fes])
    io_arr = (C.c_void_p * lanes)(*[C.addressof(x[0]) for x in ios])
    lane_s = C.c_double()
    drv.okb_e2e_replay_lanes.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    barrier()
    okl.check(drv.okb_e2e_replay_lanes(ctx_arr, io_arr, lanes, C.byref(lane_s)))
    lanes_s = lane_s.value
    if world > 1:
        t = torch.tensor([rep_s, lanes_s], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); rep_s, lanes_s = (float(x) for x in t.tolist())
    for f2 in fes[1:]:
        f2.close()
    e2e = {"value": world * B * args.steps / lanes_s, "unit": "stereo frames/s", "h2d_bytes_per_step": int(io.h2d),
           "d2h_bytes_per_step": int(io.d2h), "ms_per_step": 1e3 * lanes_s / args.steps,
           "step": f"the value leg's step ({B} stereo frames) from page-locked HOST buffers: per camera (one host thread each) "
                   "okb_detect_describe_batch + okb_match_map3d_batch, then okb_match_stereo_batch; results in host memory; "
                   f"{lanes} independent sequences in flight on the GPU (one library handle + host-thread pair each), "
                   f"{args.steps} steps in total",
           "lanes": lanes, "host_wait": "blocking event" if blocking else "spin",
           "one_sequence_alone": {"value": world * B * args.steps / rep_s, "ms_per_step": 1e3 * rep_s / args.steps},
           "keypoints_per_frame": io.nkp / (2 * B), "matches_per_stereo_frame": io.nm / B, "streaming": streaming}

    # ---- next rows of the scope table, measured beside the headline: P1 landmark-candidate preparation
    #      (host buffers in, packed pool out; TUM-VI-sized map of 50 000 landmarks) against its oracle on one host core
    next_rows = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            next_rows = bench_next_rows(fe, cfg)
        except Exception as e:   # never lose the headline line to an auxiliary measurement
            next_rows = {"error": repr(e)}

    clocks = sampler.stop()
    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload on the host cores
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n = max(2, min(8, cores // 2))
        cpu_run(cfg, Lh[:2], Rh[:2], maps, cores)
        t = min(cpu_run(cfg, Lh[:n], Rh[:n], maps, cores) for _ in range(3))
        t_port = cpu_run(cfg, Lh[:n], Rh[:n], maps, cores, use_cv2=False)
        cpu = {"value": n / t, "unit": "stereo frames/s", "cores": cores, "kind": "port", "port_only_value": n / t_port,
               "sample": f"{n} stereo frames on {cores} host threads (best of 3): detect+describe by {cpu_detector_name()}, M1 per camera + back-projection + M4 by the oracle's transcription of the reference loops; port_only_value = the same with the C restatement of BRISK"}

    if rank == 0:
        line = {"metric": "stereo frames/sec detect+describe+match", "value": value, "unit": "stereo frames/s", "n_gpus": world,
                "steps": args.steps, "warmup": warm, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": args.config, "stereo_frames_per_step_per_gpu": B, "l2_policy": f"inputs larger than L2: ring of {ring} batches = {in_bytes >> 20} MiB per GPU",
                           "parallelism": "replicas (independent sequences per GPU)" if world > 1 else "single GPU",
                           "sequences_in_flight_per_gpu": lanes,
                           **{k: cfg[k] for k in ("W", "H", "max_kp", "threshold", "octaves", "n_lm")}},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                "next_rows": next_rows}
        print(json.dumps(line))
    fe.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
