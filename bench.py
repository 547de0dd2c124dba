#!/usr/bin/env python
"""bench.py -- stereo frames/s for detect + describe + match on N B200s, beside the CPU path on the same box.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W [--impl reference]` prints ONE JSON line.
A step = one pass of the hot path over one batch of synthetic stereo frames; per stereo frame (reference call sites):
  detect + describe + back-projection of both cameras   Frontend::detectAndDescribe            Frontend.cpp:221-269
  M1 match-to-map per camera                            Frontend::matchToMapByThread           Frontend.cpp:1515-1590
  M3 motion stereo per camera vs 5 older keyframes      Frontend::matchMotionStereo            Frontend.cpp:1775-1958
  M4 stereo match camera 0 -> camera 1                  Frontend::matchStereo                  Frontend.cpp:2016-2074
  (the reference runs M4 on keyframes only; here both arms run it on every frame)
  * value : inputs (images, landmark pool, older keyframe features) already resident in HBM, results left in HBM, CUDA events.
  * e2e   : the same step through the host-buffer C ABI from page-locked HOST memory (bench/e2e_driver.cpp), every H2D / D2H
            copy inside the timed region; e2e.streaming = one stereo frame per call (live use).
  * cpu_baseline / --impl reference : the CPU arm (oracle/: C++ std::thread driver over the C restatement of OpenCV-BRISK and the
            transcribed match loops; cv2.BRISK when importable) on the host cores, bounded sample. The only place this file
            executes oracle/ code.
  * configs : sub-records of the other BASELINE.json workloads measured in the same run: euroc_okvis48 (Harris + BRISK2-48, the
    reference's own detector / extractor pair, parity unpinned), euroc_octaves0 (the shipped okvis
            setting), tumvi (1024x1024, 2000 keypoints, 50 000 landmarks = config 3 and, run per GPU, config 5) and
            hilti_sharded (5-camera rig, camera c on GPU c % N, NCCL all-gather of the feature blocks = config 4).
N > 1: one process per GPU (torchrun), each rank replays its own independent sequences (replicas, no data-path collective);
time = max over ranks.
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

# kpts = keypoints per frame the workload names; max_kp = FrontendParameters::max_num_keypoints given to the detector so
# that `kpts` keypoints survive the extractor's border removal (Frame::describe may erase keypoints,
# implementation/Frame.hpp:146: the strongest corners sit on the coarse layers, whose sampling pattern is wide). threshold is
# chosen so that the raw detections are >= 1.5 x max_kp, i.e. the cap binds.
CONFIGS = {
    # BASELINE.json configs[1]: EuRoC stereo 752x480 stream, 1000 kpts/frame
    "euroc": dict(W=752, H=480, kpts=1000, max_kp=1472, threshold=30, octaves=3, n_lm=5000, batch=32, ring=6, f=458.0, n_older=5),
    # the same with octaves = 0, the setting of every shipped okvis yaml (config/euroc.yaml:66)
    "euroc_octaves0": dict(W=752, H=480, kpts=1000, max_kp=1152, threshold=6, octaves=0, n_lm=5000, batch=32, ring=6, f=458.0, n_older=5),
    # BASELINE.json configs[2] / configs[4]: TUM-VI 1024x1024 stereo, 2000 kpts/frame, 50-keyframe landmark set
    "tumvi": dict(W=1024, H=1024, kpts=2000, max_kp=2496, threshold=30, octaves=3, n_lm=50000, batch=16, ring=5, f=190.0, n_older=5),
    # BASELINE.json configs[3]: Hilti-2022 5-camera rig, cameras sharded over the GPUs, NCCL all-gather for stereo matching
    "hilti": dict(W=720, H=540, kpts=700, max_kp=1024, threshold=30, octaves=3, n_lm=5000, batch=16, ring=8, f=351.0, n_older=0, rig="HILTI_2022"),
}
DIST = [-0.2834, 0.0740, 0.00019, 1.76e-05]
STEP_TEXT = ("per stereo frame: detect+describe+back-project both cameras, M1 match-to-map per camera, M3 motion stereo per camera against "
             "the older keyframes (sequential, matched-mask update between views), M4 stereo match on every frame")
ELIGIBLE = 0.5   # M3: fraction of an older keyframe's keypoints that are still without an initialised landmark (Frontend.cpp:1813-1821)
CAP_M = 512   # M3 host-buffer form: capacity of the compact match list per (frame, view)


def config_dict(name, cfg):
    """the workload, identical in both arms (the driver compares the two dicts)"""
    in_mib = (2 * cfg["ring"] * cfg["batch"] * cfg["W"] * cfg["H"]) >> 20
    return {"workload": name, "W": cfg["W"], "H": cfg["H"], "keypoints_per_frame": cfg["kpts"], "detector_max_keypoints": cfg["max_kp"],
            "threshold": cfg["threshold"], "octaves": cfg["octaves"], "n_lm": cfg["n_lm"], "older_keyframes": cfg["n_older"], "older_keypoints_eligible": ELIGIBLE,
            "step": STEP_TEXT, "frames": FRAMES_TEXT, "parallelism": "replicas (independent sequences per GPU)",
            "l2_policy": f"GPU arm: inputs larger than L2, a ring of {cfg['ring']} batches = {in_mib} MiB per GPU"}


def intrinsics(cfg, c):
    """radial-tangential pinhole cameras (EuRoC-like intrinsics scaled to the image size): fu fv cu cv"""
    return cfg["f"], cfg["f"] * 0.997, cfg["W"] / 2 - 8.8 + 12 * c, cfg["H"] / 2 + 8.4 + 7 * c


COHERENT = bool(os.environ.get("OKB_BENCH_COHERENT"))   # experimental workload, see make_frames_coherent
FRAMES_TEXT = ("8 rendered stereo scenes + integer-shifted copies (distinct pixels, same statistics); the landmark pool and the older keyframes are "
               "built from frame 0, so frame 0 is the frame the map fits (workload_stats.frame0 gives its match yields); the other frames load the "
               "Hamming scans but rarely pass a gate") if not COHERENT else (
               "EXPERIMENTAL (OKB_BENCH_COHERENT=1): one rendered stereo scene in every frame with photometric jitter (gain within +-3 %, Gaussian "
               "noise of 1.5 grey levels): ~94 % of the keypoints repeat, the pool and the older keyframes fit every frame")


def make_frames(cfg, n, seed0, base=8):
    """n stereo pairs: `base` rendered scenes + integer-shifted variants (distinct pixels, same statistics)."""
    if COHERENT:
        return make_frames_coherent(cfg, n, seed0)
    from okvis2_b200.synth import synth_stereo
    W, H = cfg["W"], cfg["H"]
    scenes = [synth_stereo(seed0 + i, W, H) for i in range(min(base, n))]
    L = np.empty((n, H, W), np.uint8); R = np.empty((n, H, W), np.uint8)
    for i in range(n):
        l, r = scenes[i % len(scenes)]
        s = i // len(scenes)
        L[i] = np.roll(l, (3 * s, 5 * s), (0, 1)); R[i] = np.roll(r, (3 * s, 5 * s), (0, 1))
    return L, R


def make_frames_coherent(cfg, n, seed0):
    """EXPERIMENTAL workload (OKB_BENCH_COHERENT=1): n stereo pairs of ONE scene, frame 0 as rendered, frame i > 0 with its own gain and
    noise, so that the landmark pool and the older keyframes (built from frame 0) fit EVERY frame and every frame passes through the
    gate / triangulate / insert stages (EuRoC: 167 M1 matches and 426 M3 insertions per frame instead of ~0 outside frame 0; 23.4 k instead
    of 25.3 k stereo frames/s). Not the default: at the TUM-VI size (2 000 keypoints, 50 000 landmarks) this workload ended 2 of 12 runs
    in a device fault and one in a hang that the default workload has never shown and that memcheck / initcheck runs of the same step do
    not reproduce; not root-caused yet."""
    from okvis2_b200.synth import synth_stereo
    W, H = cfg["W"], cfg["H"]
    l, r = synth_stereo(seed0, W, H)
    L = np.empty((n, H, W), np.uint8); R = np.empty((n, H, W), np.uint8)
    lf, rf = l.astype(np.float32), r.astype(np.float32)
    for i in range(n):
        if i == 0:
            L[i], R[i] = l, r
            continue
        rng = np.random.default_rng(7000 + 131 * seed0 + i)
        g = np.float32(1.0 + 0.03 * np.sin(0.7 * i))
        L[i] = np.clip(np.rint(lf * g + rng.normal(0, 1.5, lf.shape).astype(np.float32)), 0, 255).astype(np.uint8)
        R[i] = np.clip(np.rint(rf * g + rng.normal(0, 1.5, rf.shape).astype(np.float32)), 0, 255).astype(np.uint8)
    return L, R


def make_map(cfg, kp, desc, seed):
    from okvis2_b200.synth import map_scene
    xy = np.stack([kp["x"], kp["y"]], 1).astype(np.float64)
    return map_scene(seed, xy, desc, cfg["n_lm"], W=cfg["W"], H=cfg["H"])


def cam_pose(c):
    from okvis2_b200.synth import pose12
    return pose12(np.eye(3), np.array([0.11 * c, 0.0, 0.0]))


def make_older_views(cfg, c, kp, desc, rays, valid, seed):
    """cfg['n_older'] older keyframe views of camera c for M3, geometrically consistent with frame 0: the keypoints of frame 0
    get a depth, the 3-D points are seen from displaced poses (back-projections exact, descriptors noisy copies), the rest
    of every view are outliers. dicts(desc, rays, valid, size, use, T_WC, T_CW)."""
    from okvis2_b200.synth import pose12, random_descriptors, rot
    rng = np.random.default_rng(seed)
    n = len(kp)
    T_WC1, _ = cam_pose(c)
    r1 = T_WC1[9:]
    e = rays / np.linalg.norm(rays, axis=1, keepdims=True)
    P = r1 + e * np.exp(rng.uniform(np.log(1.5), np.log(25.0), n))[:, None]
    views = []
    for v in range(cfg["n_older"]):
        Cv = rot((0, 1, 0), 0.012 * (v + 1)) @ rot((1, 0, 0), -0.004 * v)
        rv = r1 + np.array([-0.06 * (v + 1), 0.01 * v, -0.02 * (v + 1)])
        pc = (P - rv) @ Cv
        ok = (valid != 0) & (pc[:, 2] > 0.3) & (np.abs(pc[:, 0] / pc[:, 2]) < 0.75) & (np.abs(pc[:, 1] / pc[:, 2]) < 0.5) & (rng.random(n) < 0.35)
        idx = np.nonzero(ok)[0]
        bits = np.unpackbits(desc[idx], axis=1)
        d = np.packbits(bits ^ (rng.random(bits.shape) < 0.04).astype(np.uint8), axis=1)
        ry = np.stack([pc[idx, 0] / pc[idx, 2], pc[idx, 1] / pc[idx, 2], np.ones(len(idx))], 1)
        ry[:, :2] += rng.normal(0, 3e-4, (len(idx), 2))
        n_out = max(cfg["kpts"] - len(idx), 0)
        d = np.concatenate([d, random_descriptors(rng, n_out, desc.shape[1])])
        ry = np.concatenate([ry, np.stack([rng.uniform(-0.7, 0.7, n_out), rng.uniform(-0.45, 0.45, n_out), np.ones(n_out)], 1)])
        perm = rng.permutation(len(d))
        Tw, Tc = pose12(Cv, rv)
        views.append(dict(desc=np.ascontiguousarray(d[perm]), rays=np.ascontiguousarray(ry[perm]), valid=(rng.random(len(d)) > 0.02).astype(np.uint8),
                          size=(rng.choice([12.0, 18.0, 24.0, 36.0], len(d)) * rng.uniform(0.9, 1.1, len(d))).astype(np.float32),
                          use=(rng.random(len(d)) < ELIGIBLE).astype(np.uint8), T_WC=Tw, T_CW=Tc))
    return views


def build_workload(cfg, n_frames, seed0, detect0):
    """frames + landmark pools + older views. detect0(img, cam) -> (kp, desc, rays, valid) of frame 0 (GPU arm: the library; CPU
    arm: the oracle -- the two are bit-identical, so both arms match against the same pools)."""
    L, R = make_frames(cfg, n_frames, seed0)
    maps, older = [], []
    for c, img in enumerate((L[0], R[0])):
        kp, desc, rays, valid = detect0(img, c)
        maps.append(make_map(cfg, kp, desc, 40 + c))
        older.append(make_older_views(cfg, c, kp, desc, rays, valid, 90 + c) if cfg["n_older"] else [])
    return dict(L=L, R=R, maps=maps, older=older)


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
# CPU arm
class CamWork(C.Structure):
    """okvo_cam_work_t (oracle/cpu_frontend.cpp)"""
    _fields_ = [("images", C.c_void_p), ("model", C.c_int32), ("pad_", C.c_int32), ("intr", C.c_double * 8), ("T_WC", C.c_double * 12),
                ("T_CW", C.c_double * 12), ("n_cand", C.c_int32), ("n_lm", C.c_int32), ("cand_desc", C.c_void_p), ("cand_lm", C.c_void_p),
                ("lm_proj", C.c_void_p), ("lm_is3d", C.c_void_p), ("n_views", C.c_int32), ("pad2_", C.c_int32), ("v_n", C.c_void_p),
                ("v_off", C.c_void_p), ("v_desc", C.c_void_p), ("v_rays", C.c_void_p), ("v_valid", C.c_void_p), ("v_size", C.c_void_p),
                ("v_use", C.c_void_p), ("v_T_WC", C.c_void_p), ("v_T_CW", C.c_void_p)]


class FrontendCfg(C.Structure):
    """okvo_frontend_cfg_t"""
    _fields_ = [(k, C.c_int32) for k in ("W", "H", "threshold", "octaves", "max_kp", "n_cams", "n_frames", "warmup",
                                         "detect_threads_per_frame", "match_threads", "workers", "stereo")]


def cpu_port_run(cfg, wl, n_frames, workers, detect_threads, match_threads, warmup=1):
    """The C++ std::thread driver (oracle/cpu_frontend.cpp) over frames [0, n_frames) of the workload."""
    import oracle
    L_ = oracle.lib()
    keep = []

    def arr(a, t):
        a = np.ascontiguousarray(a, t); keep.append(a); return a.ctypes.data

    cams = (CamWork * 2)()
    for c in range(2):
        w = cams[c]
        img = np.ascontiguousarray((wl["L"], wl["R"])[c][:n_frames]); keep.append(img)
        w.images = img.ctypes.data; w.model = 1
        w.intr[:] = list(intrinsics(cfg, c)) + DIST
        Tw, Tc = cam_pose(c); w.T_WC[:] = list(Tw); w.T_CW[:] = list(Tc)
        m = wl["maps"][c]
        w.n_cand, w.n_lm = len(m["cand_lm"]), len(m["lm_is3d"])
        w.cand_desc, w.cand_lm, w.lm_proj, w.lm_is3d = arr(m["cand_desc"], np.uint8), arr(m["cand_lm"], np.int32), arr(m["lm_proj"], np.float64), arr(m["lm_is3d"], np.uint8)
        vs = wl["older"][c]
        w.n_views = len(vs)
        if vs:
            n0 = np.array([len(v["desc"]) for v in vs], np.int32)
            w.v_n = arr(n0, np.int32); w.v_off = arr(np.concatenate([[0], np.cumsum(n0)[:-1]]), np.int32)
            w.v_desc = arr(np.concatenate([v["desc"] for v in vs]), np.uint8); w.v_rays = arr(np.concatenate([v["rays"] for v in vs]), np.float64)
            w.v_valid = arr(np.concatenate([v["valid"] for v in vs]), np.uint8); w.v_size = arr(np.concatenate([v["size"] for v in vs]), np.float32)
            w.v_use = arr(np.concatenate([v["use"] for v in vs]), np.uint8)
            w.v_T_WC = arr(np.stack([v["T_WC"] for v in vs]), np.float64); w.v_T_CW = arr(np.stack([v["T_CW"] for v in vs]), np.float64)
    fc = FrontendCfg(cfg["W"], cfg["H"], cfg["threshold"], cfg["octaves"], cfg["max_kp"], 2, n_frames, warmup, detect_threads, match_threads, workers, 1)
    per = np.zeros(n_frames); tot = C.c_double(); nkp = C.c_long(); nm = C.c_long()
    f = L_.okvo_frontend_run
    f.argtypes = [C.c_void_p] * 6; f.restype = C.c_int
    f(C.byref(fc), cams, per.ctypes.data, C.byref(tot), C.byref(nkp), C.byref(nm))
    return {"total_s": tot.value, "stereo_frames_per_s": n_frames / tot.value, "ms_mean": float(per.mean()), "ms_min": float(per.min()),
            "ms_max": float(per.max()), "multiframes": n_frames, "keypoints_per_frame": nkp.value / (2 * n_frames),
            "matches_per_stereo_frame": nm.value / n_frames}


def cv2_version():
    try:
        import cv2
        return cv2.__version__
    except Exception:
        return None


def cpu_cv2_run(cfg, wl, n_frames, n_threads):
    """The same step with OpenCV's own BRISK (the implementation the oracle restates, the fastest CPU BRISK available here) for
    detect + describe: one job per stereo frame on a thread pool (cv2 and the ctypes matchers release the GIL). Seconds."""
    import cv2
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    cv2.setNumThreads(1)
    local = threading.local()
    imgs = (wl["L"], wl["R"])
    T3 = lambda T: np.concatenate([np.asarray(T[:9]).reshape(3, 3), np.asarray(T[9:]).reshape(3, 1)], 1).reshape(12)
    poses = [cam_pose(c) for c in range(2)]

    def job(i):
        if not hasattr(local, "brisk"):
            local.brisk = cv2.BRISK_create(cfg["threshold"], cfg["octaves"], 1.0)
        side = []
        for c in range(2):
            img = imgs[c][i]
            kps = local.brisk.detect(img, None)
            if len(kps) > cfg["max_kp"]:
                resp = np.fromiter((k.response for k in kps), np.float32, len(kps))
                keep = np.sort(np.argsort(-resp, kind="stable")[:cfg["max_kp"]])
                kps = [kps[j] for j in keep]
            kps, d = local.brisk.compute(img, kps)
            n = len(kps)
            kp = np.zeros(n, oracle.KP_DTYPE)
            if n:
                pts = cv2.KeyPoint_convert(kps)
                kp["x"], kp["y"] = pts[:, 0], pts[:, 1]
                kp["size"] = np.fromiter((k.size for k in kps), np.float32, n)
            else:
                d = np.zeros((0, 64), np.uint8)
            fu, fv, cu, cv = intrinsics(cfg, c)
            rays, valid = oracle.back_project(1, fu, fv, cu, cv, DIST, kp)
            m = wl["maps"][c]
            xy = np.stack([kp["x"], kp["y"]], 1).astype(np.float64)
            _, lm = oracle.match_map3d(d, xy, None, m["cand_desc"], m["cand_lm"], m["lm_proj"], m["lm_is3d"], 20.0, 60, 1)
            if wl["older"][c]:
                oracle.match_motion_stereo_sequence(wl["older"][c], d, rays, valid, np.stack([kp["x"], kp["y"]], 1), poses[c][0], poses[c][1], 1,
                                                    np.array([fu, fv, cu, cv] + DIST), cfg["W"], cfg["H"], 60, (lm >= 0).astype(np.uint8), 1)
            e = rays / np.sqrt((rays[:, 0] * rays[:, 0] + rays[:, 1] * rays[:, 1]) + rays[:, 2] * rays[:, 2])[:, None]
            side.append((d, valid, np.ascontiguousarray(e), kp["size"].astype(np.float64) / (0.5 * (fu + fv))))
        (d0, v0, e0, s0), (d1, v1, e1, s1) = side
        oracle.match_stereo(d0, v0, e0, s0, d1, v1, e1, s1, poses[0][0][9:], poses[1][0][9:], T3(poses[0][1]), T3(poses[1][1]), 60)

    t0 = time.perf_counter()
    with ThreadPoolExecutor(n_threads) as ex:
        list(ex.map(job, range(n_frames)))
    return time.perf_counter() - t0


def oracle_detect0(cfg):
    import oracle
    o = oracle.Brisk(cfg["threshold"], cfg["octaves"])

    def f(img, c):
        kp, d = o.detect_and_compute(img, cfg["max_kp"])
        fu, fv, cu, cv = intrinsics(cfg, c)
        rays, valid = oracle.back_project(1, fu, fv, cu, cv, DIST, kp)
        return kp, d, rays, valid
    return f


def cpu_model():
    try:
        return [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        return None


def cpu_arm(cfg, wl, n_sample, n_latency):
    """Both schedules of the CPU arm on the first frames of the workload; `value` = the best all-cores throughput."""
    cores = os.cpu_count() or 1
    allc = cpu_port_run(cfg, wl, n_sample, workers=cores, detect_threads=1, match_threads=1)
    refc = cpu_port_run(cfg, wl, n_latency, workers=1, detect_threads=2, match_threads=4)
    out = {"cores": cores, "kind": "port", "unit": "stereo frames/s", "cpu_model": cpu_model(),
           "all_cores": {"value": allc["stereo_frames_per_s"], "multiframes": allc["multiframes"], "schedule": f"{cores} multiframes in flight, one std::thread each",
                         "keypoints_per_frame": allc["keypoints_per_frame"], "matches_per_stereo_frame": allc["matches_per_stereo_frame"]},
           "reference_configuration": {"value": refc["stereo_frames_per_s"], "ms_per_multiframe_mean": refc["ms_mean"], "ms_per_multiframe_min": refc["ms_min"],
                                       "ms_per_multiframe_max": refc["ms_max"], "multiframes": refc["multiframes"],
                                       "schedule": "one multiframe at a time, 1 std::thread per camera for detect+describe, num_matching_threads = 4 (config/euroc.yaml:71)"}}
    value, detector = allc["stereo_frames_per_s"], "C restatement of OpenCV-BRISK (oracle/brisk_oracle.c)"
    ver = cv2_version()
    if ver:
        n2 = max(4, min(n_sample, 2 * cores))
        cpu_cv2_run(cfg, wl, min(4, n2), cores)
        t = cpu_cv2_run(cfg, wl, n2, cores)
        out["all_cores_cv2"] = {"value": n2 / t, "multiframes": n2,
                                "schedule": f"Python thread pool of {cores}, one job per stereo frame, detect+describe by cv2.BRISK {ver} (GIL released inside cv2 and the ctypes matchers)"}
        if n2 / t > value:
            value, detector = n2 / t, f"cv2.BRISK {ver}"
    out["value"] = value
    out["sample"] = (f"{allc['multiframes']} stereo frames on all {cores} host threads (value = the faster of the C++ std::thread driver over the C restatement "
                     f"and the cv2.BRISK thread pool: {detector}) + {refc['multiframes']} multiframes in the reference's own schedule; M1 / M3 / M4 by the "
                     "oracle's transcription of the reference loops (oracle/match_oracle.cpp, -O3 -msse4.2 -mpopcnt)")
    return out


def run_reference(args, name, cfg, rank, world):
    """--impl reference: the CPU arm on rank 0 (the real reference front-end cannot be built here)."""
    if rank != 0:
        return
    if "rig" in cfg:
        cpu = cpu_sharded_baseline(cfg)
        print(json.dumps({"impl": "reference", "metric": "multiframes/sec detect+describe+match (5-camera rig)", "value": cpu["value"], "unit": "multiframes/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / cpu["value"], "higher_is_better": True,
                          "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": {"workload": name}, "cpu_baseline": cpu,
                          "e2e": {"value": cpu["value"], "unit": "multiframes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    cores = os.cpu_count() or 1
    n = max(4, min(32, 2 * cores))      # bounded sample: stereo frames per step
    wl = build_workload(cfg, n, 1000, oracle_detect0(cfg))
    cpu_port_run(cfg, wl, min(n, cores), workers=cores, detect_threads=1, match_threads=1, warmup=0)
    use_cv2 = cv2_version() is not None
    if use_cv2:
        cpu_cv2_run(cfg, wl, min(4, n), cores)
    t = 0.0; t_cv2 = 0.0
    steps = max(1, min(args.steps, 5))  # every step is the bounded sample; a handful keeps the run within minutes
    for _ in range(steps):
        t += cpu_port_run(cfg, wl, n, workers=cores, detect_threads=1, match_threads=1, warmup=0)["total_s"]
        if use_cv2:
            t_cv2 += cpu_cv2_run(cfg, wl, n, cores)
    best = min(t, t_cv2) if use_cv2 else t
    value = n * steps / best
    lat = cpu_port_run(cfg, wl, min(n, 16), workers=1, detect_threads=2, match_threads=4)
    line = {"impl": "reference", "metric": "stereo frames/sec detect+describe+match", "value": value, "unit": "stereo frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * best / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": config_dict(name, cfg),
            "cpu_baseline": {"value": value, "unit": "stereo frames/s", "cores": cores, "kind": "port", "cpu_model": cpu_model(),
                             "sample": f"{n} stereo frames per step x {steps} steps (of the {args.steps} requested: each step is a bounded sample) on {cores} host threads; "
                                       f"value = the faster of the C++ std::thread driver over the C restatement of OpenCV-BRISK ({n * steps / t:.1f}/s) and the "
                                       f"cv2.BRISK thread pool ({(n * steps / t_cv2) if use_cv2 else 0:.1f}/s); matchers by the oracle's transcription of the reference loops",
                             "reference_configuration": {"ms_per_multiframe_mean": lat["ms_mean"], "ms_per_multiframe_min": lat["ms_min"],
                                                         "ms_per_multiframe_max": lat["ms_max"], "multiframes": lat["multiframes"],
                                                         "schedule": "1 std::thread per camera for detect+describe, num_matching_threads = 4, one multiframe at a time"}},
            "e2e": {"value": value, "unit": "stereo frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
# GPU arm: replicas
class ReplayIO(C.Structure):
    """okb_replay_io (bench/e2e_driver.cpp)"""
    _fields_ = [("n_steps", C.c_int32), ("warmup", C.c_int32), ("batch", C.c_int32), ("ring", C.c_int32), ("W", C.c_int32),
                ("H", C.c_int32), ("cap", C.c_int32), ("pad_", C.c_int32),
                ("img", C.c_void_p * 2), ("kp", C.c_void_p * 2), ("desc", C.c_void_p * 2), ("n", C.c_void_p * 2),
                ("n_cand", C.c_int32 * 2), ("n_lm", C.c_int32 * 2),
                ("cand_desc", C.c_void_p * 2), ("cand_lm", C.c_void_p * 2), ("lm_proj", C.c_void_p * 2), ("lm_is3d", C.c_void_p * 2),
                ("m1_dist", C.c_void_p * 2), ("m1_lm", C.c_void_p * 2),
                ("k1", C.c_void_p), ("sdist", C.c_void_p), ("hp", C.c_void_p), ("init", C.c_void_p),
                ("seconds", C.c_double), ("h2d", C.c_longlong), ("d2h", C.c_longlong), ("nkp", C.c_longlong), ("nm", C.c_longlong),
                ("lane", C.c_int32), ("lanes", C.c_int32),
                ("n_older", C.c_int32), ("cap0", C.c_int32), ("cap_m", C.c_int32), ("pad2_", C.c_int32),
                ("older", C.c_void_p * 2), ("T_WC1", C.c_void_p * 2), ("T_CW1", C.c_void_p * 2), ("matched", C.c_void_p * 2),
                ("n_match", C.c_void_p * 2), ("m_k0", C.c_void_p * 2), ("m_k1", C.c_void_p * 2), ("m_flags", C.c_void_p * 2), ("m_hp", C.c_void_p * 2),
                ("n_m3", C.c_longlong)]


class StreamM3(C.Structure):
    _fields_ = [("n_older", C.c_int32), ("cap0", C.c_int32), ("cap_m", C.c_int32), ("pad_", C.c_int32), ("older", C.c_void_p * 2),
                ("T_WC1", C.c_void_p * 2), ("T_CW1", C.c_void_p * 2)]


class Replica:
    """One GPU's replica of a stereo workload: library handles (one per sequence in flight), device-resident inputs, the step."""

    def __init__(self, name, cfg, args, rank, world, local_rank, lanes):
        import torch
        from okvis2_b200 import lib as okl
        from okvis2_b200.frontend import Frontend, MultiFrame
        self.torch, self.okl, self.L_ = torch, okl, okl.lib()
        self.name, self.cfg, self.args, self.rank, self.world, self.lanes = name, cfg, args, rank, world, lanes
        W, H, B, ring = cfg["W"], cfg["H"], cfg["batch"], cfg["ring"]
        self.B, self.ring = B, ring

        def mk():
            f = Frontend(2, W, H, device=local_rank, max_batch=B)
            f.configure(threshold=cfg["threshold"], octaves=cfg["octaves"], max_keypoints=cfg["max_kp"])
            for c in range(2):
                fu, fv, cu, cv = intrinsics(cfg, c)
                f.setCameraModel(c, "radialtangential", (fu, fv), (cu, cv), DIST)
            return f
        self.fes = [mk() for _ in range(lanes)]
        fe = self.fes[0]

        def detect0(img, c):
            mf = MultiFrame(2); mf.setImage(c, img); fe.detectAndDescribe(c, mf); fe.computeBackProjections(mf, c)
            fr = mf.frames[c]
            return fr.keypoints, fr.descriptors, fr.backProjections, fr.backProjectionsValid
        self.wl = build_workload(cfg, ring * B, 1000 + 100 * rank, detect0)
        wl = self.wl
        self.d_img = [torch.from_numpy(wl["L"]).cuda(), torch.from_numpy(wl["R"]).cuda()]
        self.d_maps = []
        for m in wl["maps"]:
            proj = np.broadcast_to(m["lm_proj"], (B,) + m["lm_proj"].shape).copy()
            self.d_maps.append(dict(desc=torch.from_numpy(m["cand_desc"]).cuda(), lm=torch.from_numpy(m["cand_lm"]).cuda(),
                                    proj=torch.from_numpy(proj).cuda(), is3d=torch.from_numpy(m["lm_is3d"]).cuda()))
        cap = C.c_int(0)
        self.L_.okb_device_features(fe.ctx, 0, None, None, None, C.byref(cap))
        self.kp_cap = kp_cap = cap.value
        # older keyframe views: device blocks (the keyframe feature store of an integration), one view table per camera
        self.n_older = cfg["n_older"]
        self.cap0 = (max([len(v["desc"]) for vs in wl["older"] for v in vs] + [64]) + 63) // 64 * 64
        self._keep = []
        self.views, self.Tw1, self.Tc1 = [], [], []
        no = max(self.n_older, 1)
        for c in range(2):
            tab = (okl.OlderView * (B * no))()
            blocks = []
            for v in wl["older"][c]:
                t = {k: torch.from_numpy(np.ascontiguousarray(v[k])).cuda() for k in ("desc", "rays", "valid", "size", "use")}
                self._keep.append(t); blocks.append((t, v))
            for b in range(B):
                for vi, (t, v) in enumerate(blocks):
                    e = tab[b * self.n_older + vi]
                    e.d_desc, e.d_rays, e.d_valid, e.d_size, e.d_use = (t[k].data_ptr() for k in ("desc", "rays", "valid", "size", "use"))
                    e.n = len(v["desc"]); e.T_WC[:] = list(v["T_WC"]); e.T_CW[:] = list(v["T_CW"])
            self.views.append(tab)
            Tw, Tc = cam_pose(c)
            self.Tw1.append(np.ascontiguousarray(np.broadcast_to(Tw, (B, 12)))); self.Tc1.append(np.ascontiguousarray(np.broadcast_to(Tc, (B, 12))))
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device="cuda")
        self.out = [dict(m1=[dict(dist=z((B, kp_cap), torch.int32), lm=z((B, kp_cap), torch.int32), mask=z((B, kp_cap), torch.uint8),
                                  k1=z((B, no, self.cap0), torch.int32), d3=z((B, no, self.cap0), torch.int32),
                                  hp3=z((B, no, self.cap0, 4), torch.float64), fl3=z((B, no, self.cap0), torch.uint8)) for _ in range(2)],
                         st=dict(k1=z((B, kp_cap), torch.int32), dist=z((B, kp_cap), torch.int32), hp=z((B, kp_cap, 4), torch.float64),
                                 init=z((B, kp_cap), torch.uint8))) for _ in range(lanes)]
        self.streams = [[torch.cuda.ExternalStream(self.L_.okb_stream(f.ctx, c)) for c in range(2)] for f in self.fes]
        self.chain = [torch.cuda.Event() for _ in range(2)]
        self.serialize = False
        self.C_WC = [np.eye(3), np.eye(3)]; self.r_WC = [np.zeros(3), np.array([0.11, 0.0, 0.0])]

    def close(self):
        for f in self.fes:
            f.close()

    def match_stage(self, cx, o, c):
        """M1 + matched mask + M3 sequence of camera c on the features of its last detect call (device resident)"""
        L_, okl, B = self.L_, self.okl, self.B
        dm = self.d_maps[c]; m1 = o["m1"][c]
        okl.check(L_.okb_match_map3d_device(cx, c, 64, B, len(dm["lm"]), dm["desc"].data_ptr(), dm["lm"].data_ptr(), len(dm["is3d"]),
                                            dm["proj"].data_ptr(), dm["is3d"].data_ptr(), 20.0, 60, m1["dist"].data_ptr(), m1["lm"].data_ptr()))
        if self.n_older:
            okl.check(L_.okb_matched_mask_device(cx, c, B, m1["lm"].data_ptr(), m1["mask"].data_ptr()))
            okl.check(L_.okb_match_motion_stereo_device(cx, c, B, self.Tw1[c].ctypes.data, self.Tc1[c].ctypes.data, self.n_older, self.views[c],
                                                        self.cap0, 60, m1["mask"].data_ptr(), m1["k1"].data_ptr(), m1["d3"].data_ptr(),
                                                        m1["hp3"].data_ptr(), m1["fl3"].data_ptr()))

    def stereo_stage(self, cx, o):
        d_st = o["st"]
        self.okl.check(self.L_.okb_match_stereo_device(cx, 0, 1, self.B, self.C_WC[0].ctypes.data, self.r_WC[0].ctypes.data, self.C_WC[1].ctypes.data,
                                                       self.r_WC[1].ctypes.data, 60, d_st["k1"].data_ptr(), d_st["dist"].data_ptr(),
                                                       d_st["hp"].data_ptr(), d_st["init"].data_ptr()))

    def device_step(self, s, lane=0):
        B = self.B
        cx = self.fes[lane].ctx; o = self.out[lane]; st = self.streams[lane]
        for c in range(2):
            # the two camera streams run concurrently
            if self.serialize:
                st[c].wait_event(self.chain[1 - c])
            frames = self.d_img[c][(s % self.ring) * B:(s % self.ring + 1) * B]
            self.okl.check(self.L_.okb_detect_describe_batch_device(cx, c, B, frames.data_ptr()))
            self.match_stage(cx, o, c)
            self.chain[c].record(st[c])
        # M4: stereo matching camera 0 -> camera 1 of every frame of the batch (back-projection on the device)
        self.stereo_stage(cx, o)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize()

    def allmax(self, vals):
        if self.world == 1:
            return vals
        import torch.distributed as dist
        t = self.torch.tensor(vals, device="cuda", dtype=self.torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def measure_value(self, steps, warm):
        torch, L_, okl, lanes = self.torch, self.L_, self.okl, self.lanes
        for s in range(warm * lanes):
            self.device_step(s, s % lanes)
        for f in self.fes:
            okl.check(L_.okb_sync(f.ctx))
        self.barrier()
        launches0 = sum(L_.okb_launch_count(f.ctx) for f in self.fes)
        # all streams are idle here: the event on lane 0 precedes every kernel of the timed region; step s is the next batch of
        # sequence s % lanes (exactly `steps` steps in total)
        ev0 = torch.cuda.Event(enable_timing=True); ev0.record(self.streams[0][0])
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(2 * lanes)]
        for s in range(steps):
            self.device_step(warm * lanes + s, s % lanes)
        for l in range(lanes):
            for c in range(2):
                ends[2 * l + c].record(self.streams[l][c])
        self.barrier()
        dev_ms = max(ev0.elapsed_time(e) for e in ends)
        launches = sum(L_.okb_launch_count(f.ctx) for f in self.fes) - launches0
        dev_ms = self.allmax([dev_ms])[0]
        # sanity of the workload: keypoints per frame of the last batch of lane 0, matches of its last step
        fe = self.fes[0]; cnt = []
        for c in range(2):
            for b in range(self.B):
                n = C.c_int(0)
                okl.check(L_.okb_fetch_features(fe.ctx, c, b, None, None, self.kp_cap, C.byref(n)))
                cnt.append(n.value)
        o = self.out[0]
        stats = {"keypoints_per_frame": float(np.mean(cnt)), "keypoints_per_frame_min": int(min(cnt)),
                 "m1_matches_per_frame": float(sum((o["m1"][c]["lm"] >= 0).sum().item() for c in range(2)) / (2 * self.B)),
                 "m3_inserted_per_frame": float(sum(((o["m1"][c]["fl3"] & 4) != 0).sum().item() for c in range(2)) / (2 * self.B)) if self.n_older else 0.0,
                 "m4_matches_per_stereo_frame": float((o["st"]["k1"] >= 0).sum().item() / self.B)}
        # the landmark pool and the older views are built from frame 0 of the ring (the other frames are other scenes / shifted copies whose
        # geometry the pool does not fit, so they load the Hamming scans but rarely pass a gate): one more untimed step on the batch that
        # holds frame 0 shows what the gate / triangulate / insert stages do on a frame the map fits
        self.device_step(0, 0)
        okl.check(L_.okb_sync(fe.ctx)); torch.cuda.synchronize()
        stats["frame0"] = {"m1_matches": int((o["m1"][0]["lm"][0] >= 0).sum().item()),
                           "m3_matching": int(((o["m1"][0]["fl3"][0] & 1) != 0).sum().item()) if self.n_older else 0,
                           "m3_inserted": int(((o["m1"][0]["fl3"][0] & 4) != 0).sum().item()) if self.n_older else 0,
                           "m4_matches": int((o["st"]["k1"][0] >= 0).sum().item())}
        return dev_ms, launches, stats

    def measure_roofline(self, warm, steps):
        """kernel-level timing for the roofline: extra steps with the two camera streams serialized, so that the CUDA events
        around the pyramid+score launches (recorded on the launching stream inside the library) time those kernels alone."""
        L_, okl = self.L_, self.okl
        ctx = self.fes[0].ctx
        roof_steps = 5
        self.serialize = True
        self.device_step(warm + steps); okl.check(L_.okb_sync(ctx))
        L_.okb_enable_timers(ctx, 1); L_.okb_reset_timers(ctx)
        for s in range(roof_steps):
            self.device_step(warm + steps + 1 + s)
        okl.check(L_.okb_sync(ctx))
        self.serialize = False
        ps_ms = C.c_double(); ps_l = C.c_int64(); tot = C.c_double(); sc_ms = C.c_double()
        ps_total_ms, ps_total_launches, score_total_ms = 0.0, 0, 0.0
        for c in range(2):
            L_.okb_get_timers(ctx, c, C.byref(ps_ms), C.byref(ps_l), C.byref(tot))
            L_.okb_get_score_kernel_ms(ctx, c, C.byref(sc_ms))
            ps_total_ms += ps_ms.value; ps_total_launches += ps_l.value; score_total_ms += sc_ms.value
        L_.okb_enable_timers(ctx, 0)
        ps_bytes = L_.okb_pyramid_score_bytes(ctx, 0)       # algorithmic bytes per image (actual layer sizes)
        passes = 2 * roof_steps                            # one pass per camera per step, B images each
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        B = self.B
        ach = ps_bytes * B * passes / (ps_total_ms * 1e-3) / 1e9 if ps_total_ms > 0 else 0.0
        ach_k = ps_bytes * B * passes / (score_total_ms * 1e-3) / 1e9 if score_total_ms > 0 else 0.0
        # DRAM traffic of the dominant kernel per launch, from the ncu --set full capture of THIS code committed under profiles/
        traffic, traffic_src = None, None
        try:
            e = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(f"{self.name}_b{B}")
            if e:
                traffic, traffic_src = e["dram_bytes_per_launch"], e["source"]
        except Exception:
            pass
        # the limiter that binds k_score_nms: 80 VIMNMX3.U16x2 + 4 packed add/max per pixel pair on the ALU pipe, which issues one
        # warp instruction per 2 cycles per SM sub-partition (bench/ubench_pipes.cu) -> floor per launch at the max SM clock
        n_layers = 2 * self.cfg["octaves"] if self.cfg["octaves"] else 1
        scored_px = (ps_bytes / 2) if n_layers > 1 else (ps_bytes / 2)    # every layer pixel is read/written once and scored once
        alu_floor_ms = 84 * (scored_px / 2 * B / 32) * 2 / (148 * 4) / 1.965e9 * 1e3
        return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_src,
                "kernel": ("pyramid+score pass (k_pyramid: every reduced layer in one launch; k_score_nms: persistent, TMA-staged tiles)" if n_layers > 1
                           else "score pass (k_score_nms: persistent, TMA-staged tiles; single scale, no pyramid)"),
                "bytes_per_image": int(ps_bytes),
                "dominant_kernel": {"name": "k_score_nms", "ms_per_launch": score_total_ms / passes, "achieved_GBps": ach_k, "frac": ach_k / peak,
                                    "alu_floor_ms": alu_floor_ms, "alu_frac": (alu_floor_ms / (score_total_ms / passes)) if score_total_ms > 0 else None,
                                    "limiter": "ALU pipe (64 lanes/clk/SM): 84 packed 16x2 min/max/add instructions per pixel pair, not HBM"},
                "images_per_pass": B, "ms_per_pass": ps_total_ms / passes, "launches_per_pass": ps_total_launches / passes,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650", "frac_of_nominal_8000_GBps": ach / 8000.0,
                "measured": f"CUDA events on the launching stream, {roof_steps} extra steps right after the timed region with the two camera streams "
                            "serialized (they overlap in the timed region); the integral image runs after the pass"}

    def time_matchers(self, reps=6):
        """device time of the match stages alone on the features of the last step (CUDA events, one sequence)"""
        torch, L_, okl = self.torch, self.L_, self.okl
        cx = self.fes[0].ctx; o = self.out[0]; st = self.streams[0]
        okl.check(L_.okb_sync(cx))
        res = {}

        def run(which):
            if which == "m1+m3":
                for c in range(2):
                    self.match_stage(cx, o, c)
            elif which == "m1":
                for c in range(2):
                    dm = self.d_maps[c]; m1 = o["m1"][c]
                    okl.check(L_.okb_match_map3d_device(cx, c, 64, self.B, len(dm["lm"]), dm["desc"].data_ptr(), dm["lm"].data_ptr(), len(dm["is3d"]),
                                                        dm["proj"].data_ptr(), dm["is3d"].data_ptr(), 20.0, 60, m1["dist"].data_ptr(), m1["lm"].data_ptr()))
            else:
                self.stereo_stage(cx, o)
        for which in ("m1", "m1+m3", "m4"):
            run(which); okl.check(L_.okb_sync(cx))
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            st[1].wait_stream(st[0]); st[0].wait_stream(st[1])
            e0.record(st[0]); st[1].wait_stream(st[0])
            for _ in range(reps):
                run(which)
            st[0].wait_stream(st[1])
            e1.record(st[0])
            okl.check(L_.okb_sync(cx)); torch.cuda.synchronize()
            res[which] = e0.elapsed_time(e1) / reps
        return res

    def measure_e2e(self, steps, warm, streaming):
        torch, L_, okl, lanes, B, cfg = self.torch, self.L_, self.okl, self.lanes, self.B, self.cfg
        W, H, kp_cap, ring = cfg["W"], cfg["H"], self.kp_cap, self.ring
        wl = self.wl
        host_threads = self.world * lanes * 3
        blocking = host_threads > (os.cpu_count() or 1)
        for f in self.fes:
            L_.okb_set_blocking_sync(f.ctx, 1 if blocking else 0)
        drv = C.CDLL(os.path.join(ROOT, "bench", "libokb_e2e.so"))
        maps = wl["maps"]
        keep = [[np.ascontiguousarray(m[k]) for m in maps] for k in ("cand_desc", "cand_lm", "lm_proj", "lm_is3d")]
        arr_i = lambda v: (C.c_int * 2)(*v)
        arr_p = lambda v: (C.c_void_p * 2)(*[x.ctypes.data for x in v])
        out = {}
        if streaming:
            e2e_frames = min(ring * B, max(16, 4 * B))
            only = bool(os.environ.get("OKB_BENCH_STREAM_ONLY"))
            if only:
                e2e_frames = 8
            sec = C.c_double(); h2d = C.c_longlong(); d2h = C.c_longlong(); nkp = C.c_longlong(); nm = C.c_longlong()
            sm3 = StreamM3(self.n_older, self.cap0, CAP_M, 0)
            for c in range(2):
                sm3.older[c] = C.addressof(self.views[c]); sm3.T_WC1[c] = self.Tw1[c].ctypes.data; sm3.T_CW1[c] = self.Tc1[c].ctypes.data
            drv.okb_e2e_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_double] + [C.c_void_p] * 12
            self.barrier()
            rc = 0 if only else drv.okb_e2e_run(self.fes[0].ctx, e2e_frames, 4, W, H, wl["L"].ctypes.data, wl["R"].ctypes.data, kp_cap, cfg["f"],
                                 arr_i([len(x) for x in keep[1]]), arr_p(keep[0]), arr_p(keep[1]), arr_i([len(x) for x in keep[3]]),
                                 arr_p(keep[2]), arr_p(keep[3]), C.byref(sm3) if self.n_older else None, C.byref(sec), C.byref(h2d), C.byref(d2h),
                                 C.byref(nkp), C.byref(nm))
            okl.check(rc)
            e2e_s = self.allmax([sec.value])[0] if not only else 1.0
            # the same live use through ONE call per stereo frame (okb_process_multiframe), replayed as a CUDA graph
            sec2 = C.c_double(); worst = C.c_double(); h2 = C.c_longlong(); d2 = C.c_longlong(); nk2 = C.c_longlong(); nm2 = C.c_longlong()
            drv.okb_e2e_multiframe.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 13
            # inputs in page-locked host memory (the e2e contract's "pinned host memory"): the library reads them in place
            pin_ = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
            p_img = [pin_(wl["L"][:e2e_frames]), pin_(wl["R"][:e2e_frames])]
            p_proj = [pin_(m["lm_proj"]) for m in maps]
            proj1 = [x.numpy() for x in p_proj]
            self.barrier()
            rc = drv.okb_e2e_multiframe(self.fes[0].ctx, e2e_frames, 6, W, H, p_img[0].data_ptr(), p_img[1].data_ptr(), kp_cap,
                                        arr_i([len(x) for x in keep[1]]), arr_p(keep[0]), arr_p(keep[1]), arr_i([len(x) for x in keep[3]]),
                                        arr_p(proj1), arr_p(keep[3]), C.byref(sm3) if self.n_older else None, C.byref(sec2), C.byref(worst), C.byref(h2),
                                        C.byref(d2), C.byref(nk2), C.byref(nm2))
            okl.check(rc)
            g_l = C.c_longlong(); d_l = C.c_longlong(); L_.okb_stream_stats(self.fes[0].ctx, C.byref(g_l), C.byref(d_l))
            ph = (C.c_double * 4)(); L_.okb_stream_timing(self.fes[0].ctx, ph, 1)
            calls = e2e_frames
            mf_s = self.allmax([sec2.value])[0]
            del p_img, p_proj
            separate = {"value": self.world * e2e_frames / e2e_s, "ms_per_stereo_frame": 1e3 * e2e_s / e2e_frames,
                        "step": "the same frame as separate host-buffer calls: 2x okb_detect_describe (one host thread per camera) + okb_match_stereo + "
                                "2x okb_match_map3d + 2x okb_match_motion_stereo_batch"}
            out["streaming"] = {"value": self.world * e2e_frames / mf_s, "unit": "stereo frames/s", "h2d_bytes_per_frame": int(h2.value / e2e_frames),
                                "d2h_bytes_per_frame": int(d2.value / e2e_frames), "ms_per_stereo_frame": 1e3 * mf_s / e2e_frames,
                                "ms_worst_frame": worst.value,
                                "host_ms_per_frame": {"stage_inputs": 1e3 * ph[0] / calls, "submit": 1e3 * ph[1] / calls, "wait_device": 1e3 * ph[2] / calls,
                                                      "copy_results": 1e3 * ph[3] / calls},
                                "step": "one stereo frame per call (live use, ThreadedSlam::processFrame): okb_process_multiframe = detect+describe both cameras, "
                                        "M1, M3 sequence, M4 enqueued at once from page-locked HOST buffers (results into pageable host buffers) and replayed "
                                        "as a CUDA graph, one synchronisation per frame",
                                "frames": e2e_frames, "cuda_graph_launches": int(g_l.value), "direct_submissions": int(d_l.value),
                                "keypoints_per_frame": nk2.value / e2e_frames / 2, "matches_per_frame": nm2.value / e2e_frames,
                                "separate_calls": separate}
            if os.environ.get("OKB_BENCH_STREAM_ONLY"):      # profiling hook: only the per-frame calls above run (ncu launch lists of one frame)
                return out
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        pz = lambda n, dt: torch.zeros(n, dtype=dt).pin_memory()
        h_img = [pin(wl["L"]), pin(wl["R"])]
        h_map = dict(cand_desc=[pin(m["cand_desc"]) for m in maps], cand_lm=[pin(m["cand_lm"]) for m in maps],
                     lm_proj=[pin(np.broadcast_to(m["lm_proj"], (B,) + m["lm_proj"].shape)) for m in maps],
                     lm_is3d=[pin(m["lm_is3d"]) for m in maps])
        no, cm = max(self.n_older, 1), CAP_M

        def make_io(n_steps, lane, lanes_):
            hold = dict(img=h_img, kp=[pz(B * kp_cap * 28, torch.uint8) for _ in range(2)],
                        desc=[pz(B * kp_cap * 64, torch.uint8) for _ in range(2)], n=[pz(B, torch.int32) for _ in range(2)],
                        m1_dist=[pz(B * kp_cap, torch.int32) for _ in range(2)], m1_lm=[pz(B * kp_cap, torch.int32) for _ in range(2)],
                        k1=pz(B * kp_cap, torch.int32), sdist=pz(B * kp_cap, torch.int32), hp=pz(B * kp_cap * 4, torch.float64),
                        init=pz(B * kp_cap, torch.uint8),
                        matched=[pz(B * kp_cap, torch.uint8) for _ in range(2)], n_match=[pz(B * no, torch.int32) for _ in range(2)],
                        m_k0=[pz(B * no * cm, torch.int32) for _ in range(2)], m_k1=[pz(B * no * cm, torch.int32) for _ in range(2)],
                        m_flags=[pz(B * no * cm, torch.uint8) for _ in range(2)], m_hp=[pz(B * no * cm * 4, torch.float64) for _ in range(2)], **h_map)
            io = ReplayIO(n_steps=n_steps, warmup=warm, batch=B, ring=ring, W=W, H=H, cap=kp_cap, lane=lane, lanes=lanes_,
                          n_older=self.n_older, cap0=self.cap0, cap_m=cm)
            for k in ("img", "kp", "desc", "n", "cand_desc", "cand_lm", "lm_proj", "lm_is3d", "m1_dist", "m1_lm", "matched", "n_match", "m_k0", "m_k1",
                      "m_flags", "m_hp"):
                for c in range(2):
                    getattr(io, k)[c] = hold[k][c].data_ptr()
            for c in range(2):
                io.n_cand[c] = len(maps[c]["cand_lm"]); io.n_lm[c] = len(maps[c]["lm_is3d"])
                io.older[c] = C.addressof(self.views[c]); io.T_WC1[c] = self.Tw1[c].ctypes.data; io.T_CW1[c] = self.Tc1[c].ctypes.data
            io.k1, io.sdist, io.hp, io.init = (hold[k].data_ptr() for k in ("k1", "sdist", "hp", "init"))
            return io, hold
        # (i) one replay alone: every call returns its results before the next batch is submitted
        io, hold = make_io(steps, 0, 0)
        drv.okb_e2e_replay.argtypes = [C.c_void_p, C.c_void_p]
        self.barrier()
        okl.check(drv.okb_e2e_replay(self.fes[0].ctx, C.byref(io)))
        rep_s = io.seconds
        # (ii) `lanes` independent sequences replayed concurrently on this GPU, each through its own library handle and host-thread
        #      pair: exactly `steps` steps in total. One lane's H2D / D2H copies overlap the other lanes' kernels.
        per_lane = [steps // lanes + (1 if l < steps % lanes else 0) for l in range(lanes)]
        ios = [make_io(per_lane[l], l, lanes) for l in range(lanes)]
        ctx_arr = (C.c_void_p * lanes)(*[f.ctx for f in self.fes])
        io_arr = (C.c_void_p * lanes)(*[C.addressof(x[0]) for x in ios])
        lane_s = C.c_double()
        drv.okb_e2e_replay_lanes.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        self.barrier()
        okl.check(drv.okb_e2e_replay_lanes(ctx_arr, io_arr, lanes, C.byref(lane_s)))
        rep_s, lanes_s = self.allmax([rep_s, lane_s.value])
        out.update({"value": self.world * B * steps / lanes_s, "unit": "stereo frames/s", "h2d_bytes_per_step": int(io.h2d),
                    "d2h_bytes_per_step": int(io.d2h), "ms_per_step": 1e3 * lanes_s / steps,
                    "step": f"the value leg's step ({B} stereo frames) from page-locked HOST buffers: per camera (one host thread each) okb_detect_describe_batch + "
                            "okb_match_map3d_batch + okb_match_motion_stereo_batch, then okb_match_stereo_batch; results in host memory; "
                            f"{lanes} independent sequences in flight on the GPU (one library handle + host-thread pair each), {steps} steps in total",
                    "lanes": lanes, "host_wait": "blocking event" if blocking else "spin",
                    "one_sequence_alone": {"value": self.world * B * steps / rep_s, "ms_per_step": 1e3 * rep_s / steps},
                    "keypoints_per_frame": io.nkp / (2 * B), "matches_per_stereo_frame": io.nm / B, "m3_inserted_per_stereo_frame": io.n_m3 / B})
        return out


def matcher_record(rep):
    """pairs/s of the Hamming scans of one step against the measured POPC and IMMA issue rates (bench/ubench_pipes.cu, bench/ubench_imma.cu;
    profiles/popc_rate.json, profiles/imma_rate.json)."""
    cfg, B = rep.cfg, rep.B
    t = rep.time_matchers()
    n = rep.wl_stats["keypoints_per_frame"]
    n0 = float(np.mean([int((v["use"] & v["valid"]).sum()) for vs in rep.wl["older"] for v in vs])) if rep.n_older else 0.0
    m3_pairs = 2 * B * rep.n_older * n0 * n           # two cameras: every (eligible older keypoint, current keypoint) pair is scanned
    m4_pairs = B * n * n
    m1_ms, m3_ms, m4_ms = t["m1"], max(t["m1+m3"] - t["m1"], 0.0), t["m4"]
    pairs = {"m3": m3_pairs, "m4": m4_pairs}
    rate = imma = None
    try:
        rate = json.load(open(os.path.join(ROOT, "profiles", "popc_rate.json")))["popc_b32_lanes_per_clk_per_sm"]
        imma = json.load(open(os.path.join(ROOT, "profiles", "imma_rate.json")))["dot512_per_s"]
    except Exception:
        pass
    rec = {"m1_ms_per_step": m1_ms, "m3_ms_per_step": m3_ms, "m4_ms_per_step": m4_ms,
           "scan": "tensor cores (tcgen05): Hamming = popc(a) + popc(b) - 2 popc(a & b), popc(a & b) as a dot product of u8 bit planes: "
                   "k_scan_umma = tcgen05.mma kind::i8, 128 x 128 x 512 per tile, operands expanded into shared memory by producer warps, "
                   "accumulators in TMEM, hit test in the tcgen05.ld epilogue",
           "note": "device time of the match stages alone on the features of the last step (scan + gate + outputs + checks); M1 is gated by the "
                   "re-projection radius before any Hamming distance (row binning), M3 and M4 scan every eligible pair. Two ceilings are "
                   "quoted for the scan: the measured POPC issue rate (16 popc.b32 per pair; what the round-1 form of the scan ran at) and the "
                   "measured legacy IMMA.16832 issue rate (bench/ubench_imma.cu, 16 per 128 pairs; the mid-round form); the tcgen05 form's own "
                   "ceiling is 83 cycles per 128 x 128 x 32 MMA (bench/umma_probe.cu) = 1.8e12 pairs/s",
           "measured_popc_b32_lanes_per_clk_per_sm": rate, "measured_imma_dot512_per_s": imma}
    for key, ms in (("m3", m3_ms), ("m4", m4_ms)):
        if ms > 0:
            pps = pairs[key] / (ms * 1e-3)
            rec[key] = {"pairs_per_step": pairs[key], "pairs_per_s": pps}
            if rate:
                peak_pairs = rate * 148 * 1.965e9 / 16.0     # 16 popc.b32 per 512-bit pair
                rec[key]["frac_of_popc_issue_rate"] = pps / peak_pairs
            if imma:
                rec[key]["frac_of_imma_issue_rate"] = pps / imma
    return rec


def run_replica(name, cfg, args, rank, world, local_rank, lanes, steps, full):
    """value (+ roofline, e2e) of one stereo workload as replicas; `full` adds streaming and the matcher record and keeps the
    replica open (returned under '_rep')."""
    warm = max(args.warmup, 3)
    rep = Replica(name, cfg, args, rank, world, local_rank, lanes)
    ok = False
    try:
        if os.environ.get("OKB_BENCH_STREAM_ONLY"):
            print(json.dumps(rep.measure_e2e(steps, warm, streaming=True)))
            raise SystemExit(0)
        dev_ms, launches, stats = rep.measure_value(steps, warm)
        rep.wl_stats = stats
        if os.environ.get("OKB_BENCH_VALUE_ONLY"):      # profiling hook: launch lists of the device-resident step
            print(json.dumps({"ms_per_step": dev_ms / steps, "gpu_launches": int(launches)}))
            raise SystemExit(0)
        B = rep.B
        value = world * B * steps / (dev_ms * 1e-3)
        if stats["keypoints_per_frame"] < 0.95 * cfg["kpts"]:
            raise SystemExit(f"bench.py: workload {name}: {stats['keypoints_per_frame']:.0f} keypoints per frame < 0.95 x {cfg['kpts']} "
                             "(the detector cap does not deliver the stated workload)")
        rec = {"value": value, "unit": "stereo frames/s", "ms_per_step": dev_ms / steps, "stereo_frames_per_step_per_gpu": B,
               "sequences_in_flight_per_gpu": lanes, "gpu_launches": int(launches), "workload_stats": stats,
               "roofline": rep.measure_roofline(warm, steps), "e2e": rep.measure_e2e(steps, warm, streaming=full), "config": config_dict(name, cfg)}
        if full:
            rec["matcher"] = matcher_record(rep)
            rec["_rep"] = rep
        ok = True
        return rec
    finally:
        if not (full and ok):
            rep.close()


# ---------------------------------------------------------------------------------------------------------------
# BASELINE configs[1] in the detector / extractor pair OKVIS2 itself constructs: Harris + uniformity enforcement,
# 48-byte camera-aware, gravity-aligned BRISK2, octaves = 0 (config/euroc.yaml:63-67). The uniformity radius / threshold are chosen so
# that the detector cap binds on the synthetic frames and >= 700 keypoints (euroc.yaml:67) survive the extractor's border removal.
OKVIS48 = dict(W=752, H=480, kpts=700, max_kp=864, radius=8.0, abs_threshold=20, n_lm=5000, batch=32, ring=6, f=458.0, n_older=5)


def run_okvis48(args, rank, world, local_rank, steps, with_cpu):
    """value: device-resident step = per camera okb_detect_describe_batch_device (Harris, uniformity, BRISK2-48 with the camera-awareness
    maps and the extraction direction, D4), okb_match_map3d_device on the 48-byte rows, the M3 sequence against 5 older keyframes, then M4
    (the headline step in the D = 48 mode); e2e: okb_detect_describe_batch from host buffers (+ the same matchers on the device);
    cpu_baseline: the oracle's detect + describe on one core."""
    import torch
    from okvis2_b200 import lib as okl
    from okvis2_b200.frontend import Frontend
    cfg = OKVIS48
    W, H, B, ring = cfg["W"], cfg["H"], cfg["batch"], cfg["ring"]
    L_ = okl.lib()
    fe = Frontend(2, W, H, device=local_rank, max_batch=B, descriptor_bytes=48)
    fe.configure(threshold=cfg["radius"], absolute_threshold=cfg["abs_threshold"], octaves=0, max_keypoints=cfg["max_kp"])
    T_WC = np.eye(4); T_WC[:3, :3] = np.array([[1, 0, 0], [0, 0, 1], [0, -1, 0.]])   # camera looking horizontally: gravity = image +y
    maps = []
    for c in range(2):
        fu, fv, cu, cv = intrinsics(cfg, c)
        fe.setCameraModel(c, "radialtangential", (fu, fv), (cu, cv), DIST)
        maps.append(fe.cameraAwarenessMaps(c))
        okl.check(L_.okb_set_extraction_direction(fe.ctx, c, np.ascontiguousarray(T_WC[:3, :3]).ctypes.data))
    Lf, Rf = make_frames(cfg, ring * B, 1000 + 100 * rank)
    d_img = [torch.from_numpy(Lf).cuda(), torch.from_numpy(Rf).cuda()]
    from okvis2_b200.frontend import MultiFrame
    pools, older = [], []
    for c, img in enumerate((Lf[0], Rf[0])):
        mf = MultiFrame(2); mf.setImage(c, img); fe.detectAndDescribe(c, mf); fe.computeBackProjections(mf, c)
        fr = mf.frames[c]
        pools.append(make_map(cfg, fr.keypoints, fr.descriptors, 40 + c))
        older.append(make_older_views(cfg, c, fr.keypoints, fr.descriptors, fr.backProjections, fr.backProjectionsValid, 90 + c))
    cap = C.c_int(0); L_.okb_device_features(fe.ctx, 0, None, None, None, C.byref(cap))
    kp_cap = cap.value
    n_older = cfg["n_older"]
    cap0 = (max(len(v["desc"]) for vs in older for v in vs) + 63) // 64 * 64
    z = lambda shape, dt: torch.zeros(shape, dtype=dt, device="cuda")
    d_maps, outs, views, keep = [], [], [], []
    for c, m in enumerate(pools):
        proj = np.broadcast_to(m["lm_proj"], (B,) + m["lm_proj"].shape).copy()
        d_maps.append(dict(desc=torch.from_numpy(m["cand_desc"]).cuda(), lm=torch.from_numpy(m["cand_lm"]).cuda(), proj=torch.from_numpy(proj).cuda(),
                           is3d=torch.from_numpy(m["lm_is3d"]).cuda()))
        outs.append(dict(dist=z((B, kp_cap), torch.int32), lm=z((B, kp_cap), torch.int32), mask=z((B, kp_cap), torch.uint8),
                         k1=z((B, n_older, cap0), torch.int32), d3=z((B, n_older, cap0), torch.int32), hp3=z((B, n_older, cap0, 4), torch.float64),
                         fl3=z((B, n_older, cap0), torch.uint8)))
        # older keyframe views as device blocks: descriptor rows in 64-byte slots with a zero tail (what the tensor-core scans read)
        tab = (okl.OlderView * (B * n_older))()
        blocks = []
        for v in older[c]:
            slots = np.zeros((len(v["desc"]), 64), np.uint8); slots[:, :48] = v["desc"]
            t = dict(desc=torch.from_numpy(slots).cuda(), **{k: torch.from_numpy(np.ascontiguousarray(v[k])).cuda() for k in ("rays", "valid", "size", "use")})
            keep.append(t); blocks.append((t, v))
        for b in range(B):
            for vi, (t, v) in enumerate(blocks):
                e = tab[b * n_older + vi]
                e.d_desc, e.d_rays, e.d_valid, e.d_size, e.d_use = (t[k].data_ptr() for k in ("desc", "rays", "valid", "size", "use"))
                e.n = len(v["desc"]); e.T_WC[:] = list(v["T_WC"]); e.T_CW[:] = list(v["T_CW"])
        views.append(tab)
    Tw1 = [np.ascontiguousarray(np.broadcast_to(cam_pose(c)[0], (B, 12))) for c in range(2)]
    Tc1 = [np.ascontiguousarray(np.broadcast_to(cam_pose(c)[1], (B, 12))) for c in range(2)]
    st4 = dict(k1=z((B, kp_cap), torch.int32), dist=z((B, kp_cap), torch.int32), hp=z((B, kp_cap, 4), torch.float64), init=z((B, kp_cap), torch.uint8))
    C_WC = [np.eye(3), np.eye(3)]; r_WC = [np.zeros(3), np.array([0.11, 0.0, 0.0])]
    streams = [torch.cuda.ExternalStream(L_.okb_stream(fe.ctx, c)) for c in range(2)]

    def m1(c):
        dm, o = d_maps[c], outs[c]
        okl.check(L_.okb_match_map3d_device(fe.ctx, c, 48, B, len(dm["lm"]), dm["desc"].data_ptr(), dm["lm"].data_ptr(), len(dm["is3d"]),
                                            dm["proj"].data_ptr(), dm["is3d"].data_ptr(), 20.0, 60, o["dist"].data_ptr(), o["lm"].data_ptr()))

    def m3(c):
        o = outs[c]
        okl.check(L_.okb_matched_mask_device(fe.ctx, c, B, o["lm"].data_ptr(), o["mask"].data_ptr()))
        okl.check(L_.okb_match_motion_stereo_device(fe.ctx, c, B, Tw1[c].ctypes.data, Tc1[c].ctypes.data, n_older, views[c], cap0, 60, o["mask"].data_ptr(),
                                                    o["k1"].data_ptr(), o["d3"].data_ptr(), o["hp3"].data_ptr(), o["fl3"].data_ptr()))

    def m4():
        okl.check(L_.okb_match_stereo_device(fe.ctx, 0, 1, B, C_WC[0].ctypes.data, r_WC[0].ctypes.data, C_WC[1].ctypes.data, r_WC[1].ctypes.data, 60,
                                             st4["k1"].data_ptr(), st4["dist"].data_ptr(), st4["hp"].data_ptr(), st4["init"].data_ptr()))

    def step(s):
        for c in range(2):
            okl.check(L_.okb_detect_describe_batch_device(fe.ctx, c, B, d_img[c][(s % ring) * B:(s % ring + 1) * B].data_ptr()))
            m1(c); m3(c)
        m4()
    warm = max(args.warmup, 3)
    for s in range(warm):
        step(s)
    okl.check(L_.okb_sync(fe.ctx)); torch.cuda.synchronize()
    l0 = L_.okb_launch_count(fe.ctx)
    ev0 = torch.cuda.Event(enable_timing=True); ev0.record(streams[0])
    streams[1].wait_event(ev0)
    for s in range(steps):
        step(warm + s)
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for c in range(2):
        ends[c].record(streams[c])
    okl.check(L_.okb_sync(fe.ctx)); torch.cuda.synchronize()
    dev_ms = max(ev0.elapsed_time(e) for e in ends)
    launches = L_.okb_launch_count(fe.ctx) - l0
    cnt = []
    for c in range(2):
        for b in range(B):
            n = C.c_int(0); okl.check(L_.okb_fetch_features(fe.ctx, c, b, None, None, kp_cap, C.byref(n))); cnt.append(n.value)
    if np.mean(cnt) < 0.95 * cfg["kpts"]:
        raise RuntimeError(f"okvis48 workload: {np.mean(cnt):.0f} keypoints per frame < 0.95 x {cfg['kpts']}")
    m1_matches = float(sum((o["lm"] >= 0).sum().item() for o in outs) / (2 * B))
    m3_inserted = float(sum(((o["fl3"] & 4) != 0).sum().item() for o in outs) / (2 * B))
    m4_matches = float((st4["k1"] >= 0).sum().item() / B)
    step(0)   # untimed: the batch that holds frame 0, which the pool and the older views were built from
    okl.check(L_.okb_sync(fe.ctx)); torch.cuda.synchronize()
    frame0 = {"m1_matches": int((outs[0]["lm"][0] >= 0).sum().item()), "m3_matching": int(((outs[0]["fl3"][0] & 1) != 0).sum().item()),
              "m3_inserted": int(((outs[0]["fl3"][0] & 4) != 0).sum().item()), "m4_matches": int((st4["k1"][0] >= 0).sum().item())}
    # ---- end to end: host buffers (page-locked) through okb_detect_describe_batch, every copy inside the timed region
    h_img = [torch.from_numpy(x).pin_memory() for x in (Lf, Rf)]
    h_kp = [torch.zeros((B, kp_cap, 28), dtype=torch.uint8).pin_memory() for _ in range(2)]
    h_desc = [torch.zeros((B, kp_cap, 48), dtype=torch.uint8).pin_memory() for _ in range(2)]
    n_out = np.zeros((2, B), np.int32)

    def cam_job(c, s):
        okl.check(L_.okb_detect_describe_batch(fe.ctx, c, B, h_img[c][(s % ring) * B:(s % ring + 1) * B].data_ptr(), W, h_kp[c].data_ptr(),
                                               h_desc[c].data_ptr(), kp_cap, n_out[c].ctypes.data))
        m1(c); m3(c)

    def host_step(s):
        ts = [threading.Thread(target=cam_job, args=(c, s)) for c in range(2)]
        for t in ts: t.start()
        for t in ts: t.join()
        m4()
        okl.check(L_.okb_sync(fe.ctx))
    for s in range(2):
        host_step(s)
    okl.check(L_.okb_sync(fe.ctx)); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(steps):
        host_step(2 + s)
    okl.check(L_.okb_sync(fe.ctx)); torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    rec = {"value": B * steps / (dev_ms * 1e-3), "unit": "stereo frames/s", "ms_per_step": dev_ms / steps, "steps": steps,
           "stereo_frames_per_step_per_gpu": B, "gpu_launches": int(launches),
           "workload_stats": {"keypoints_per_frame": float(np.mean(cnt)), "keypoints_per_frame_min": int(min(cnt)), "m1_matches_per_frame": m1_matches,
                              "m3_inserted_per_frame": m3_inserted, "m4_matches_per_stereo_frame": m4_matches, "frame0": frame0},
           "e2e": {"value": B * steps / e2e_s, "unit": "stereo frames/s", "h2d_bytes_per_step": 2 * B * W * H,
                   "d2h_bytes_per_step": int(n_out.max(1).sum() * B * (28 + 48 + 25)),
                   "api": "okb_detect_describe_batch (one host thread per camera, page-locked buffers) + the device-resident matchers"},
           "config": {"workload": "euroc_okvis48", "W": W, "H": H, "keypoints_per_frame": cfg["kpts"], "detector_max_keypoints": cfg["max_kp"],
                      "uniformity_radius": cfg["radius"], "absolute_threshold": cfg["abs_threshold"], "octaves": 0, "descriptor_bytes": 48, "n_lm": cfg["n_lm"],
                      "older_keyframes": n_older, "older_keypoints_eligible": ELIGIBLE,
                      "camera_aware": True, "frames": FRAMES_TEXT, "step": "per stereo frame: Harris + uniformity detect, camera-aware gravity-aligned BRISK2-48 describe, "
                      "back-project, M1 match-to-map per camera, M3 motion stereo per camera against the older keyframes, M4 stereo match "
                      "(the headline step in the D = 48 mode)", "l2_policy": f"ring of {ring} batches"},
           "parity": "bit-exact vs oracle/brisk_oracle.c section 6; PARITY UNPINNED vs smartroboticslab/brisk@1ef8b42a (source absent)"}
    # ---- live use: one stereo frame per call (one host thread per camera, okb_detect_describe on a page-locked frame; D4 rides along)
    fe.close()
    fe = Frontend(2, W, H, device=local_rank, max_batch=1, descriptor_bytes=48)
    fe.configure(threshold=cfg["radius"], absolute_threshold=cfg["abs_threshold"], octaves=0, max_keypoints=cfg["max_kp"])
    for c in range(2):
        fu, fv, cu, cv = intrinsics(cfg, c)
        fe.setCameraModel(c, "radialtangential", (fu, fv), (cu, cv), DIST)
        okl.check(L_.okb_camera_awareness_maps(fe.ctx, c, None, None))
        okl.check(L_.okb_set_extraction_direction(fe.ctx, c, np.ascontiguousarray(T_WC[:3, :3]).ctypes.data))
    drv = C.CDLL(os.path.join(ROOT, "bench", "libokb_e2e.so"))
    drv.okb_e2e_multiframe.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 13
    sm3 = StreamM3(n_older, cap0, CAP_M, 0)
    for c in range(2):
        sm3.older[c] = C.addressof(views[c]); sm3.T_WC1[c] = Tw1[c].ctypes.data; sm3.T_CW1[c] = Tc1[c].ctypes.data
    arr_i = lambda v: (C.c_int * 2)(*v)
    arr_p = lambda v: (C.c_void_p * 2)(*[x.ctypes.data for x in v])
    keepm = [[np.ascontiguousarray(m[k]) for m in pools] for k in ("cand_desc", "cand_lm", "lm_is3d")]
    p_proj = [torch.from_numpy(np.ascontiguousarray(m["lm_proj"])).pin_memory() for m in pools]
    n_live = min(128, ring * B)
    sec2 = C.c_double(); worst = C.c_double(); h2 = C.c_longlong(); d2 = C.c_longlong(); nk2 = C.c_longlong(); nm2 = C.c_longlong()
    okl.check(drv.okb_e2e_multiframe(fe.ctx, n_live, 6, W, H, h_img[0].data_ptr(), h_img[1].data_ptr(), kp_cap, arr_i([len(x) for x in keepm[1]]),
                                     arr_p(keepm[0]), arr_p(keepm[1]), arr_i([len(x) for x in keepm[2]]), arr_p([x.numpy() for x in p_proj]), arr_p(keepm[2]),
                                     C.byref(sm3), C.byref(sec2), C.byref(worst), C.byref(h2), C.byref(d2), C.byref(nk2), C.byref(nm2)))
    g_l = C.c_longlong(); d_l = C.c_longlong(); L_.okb_stream_stats(fe.ctx, C.byref(g_l), C.byref(d_l))
    rec["e2e"]["streaming"] = {"ms_per_stereo_frame": 1e3 * sec2.value / n_live, "value": n_live / sec2.value, "unit": "stereo frames/s",
                               "ms_worst_frame": worst.value, "frames": n_live, "cuda_graph_launches": int(g_l.value), "direct_submissions": int(d_l.value),
                               "keypoints_per_frame": nk2.value / n_live / 2, "matches_per_frame": nm2.value / n_live,
                               "step": "one stereo frame per call (okb_process_multiframe: detect + describe both cameras, M1, M3 sequence, M4, replayed "
                                       "as a CUDA graph) from page-locked host buffers, D = 48"}
    if with_cpu:
        import oracle
        o = oracle.HarrisBrisk2(cfg["radius"], cfg["abs_threshold"], cfg["max_kp"])
        d = np.zeros(3, np.float32); okl.check(L_.okb_get_extraction_direction(fe.ctx, 0, d.ctypes.data))
        fu = float(np.float32(intrinsics(cfg, 0)[0]))
        n_s = 6
        t0 = time.perf_counter()
        for i in range(n_s):
            for c, img in enumerate((Lf, Rf)):
                o.detect_and_compute(img[i], maps[c][0], maps[c][1], fu, d)
        dt = time.perf_counter() - t0
        rec["cpu_baseline"] = {"value": n_s / dt, "unit": "stereo frames/s", "cores": 1, "kind": "port",
                               "sample": f"{n_s} stereo frames, detect + describe only (no M1), oracle restatement on one core"}
    fe.close()
    return rec


# ---------------------------------------------------------------------------------------------------------------
def run_sharded(args, name, cfg, rank, world, local_rank, steps):
    """Camera-sharded multiframe pipeline (BASELINE config 4): camera c on rank c % world; per step every rank runs
    detect+describe+back-project+M1 for its cameras on a batch of B multiframes, contributes its fixed-capacity feature
    blocks to ONE NCCL all-gather, then stereo-matches (M4) the overlapping pairs it owns from the gathered device buffer.
    value: images resident in HBM; e2e: images from page-locked host memory every step, results copied back to the host."""
    import torch
    import torch.distributed as dist
    from okvis2_b200 import lib as okl, rigs, sharding as sh
    from okvis2_b200.frontend import Frontend, MultiFrame
    from okvis2_b200.synth import synth_frame
    L_ = okl.lib()
    rig = getattr(rigs, cfg["rig"])
    n_cams, W, H, B, ring = len(rig), cfg["W"], cfg["H"], cfg["batch"], cfg["ring"]
    mine = sh.cameras_of(rank, world, n_cams)
    warm = max(args.warmup, 3)
    fe = Frontend(max(len(mine), 1), W, H, device=local_rank, max_batch=B)
    fe.configure(threshold=cfg["threshold"], octaves=cfg["octaves"], max_keypoints=cfg["max_kp"])
    ctx = fe.ctx
    # NCameraSystem::computeOverlaps on the device (exact, NCameraSystem.cpp:48-118): which pairs Frontend::matchStereo visits
    overlaps = rig_overlaps_exact(rig, fe.computeOverlaps)
    pairs = sh.pairs_of(rank, world, overlaps)
    # the collective behind the C ABI: one process per GPU -> okb_comm_init_rank, the id travels through torch.distributed
    comm = C.c_void_p()
    ident = (C.c_uint8 * 128)()
    if rank == 0:
        okl.check(L_.okb_comm_unique_id(ident))
    if world > 1:
        obj = [bytes(ident)]
        dist.broadcast_object_list(obj, src=0)
        C.memmove(ident, obj[0], 128)
    okl.check(L_.okb_comm_init_rank(world, rank, ident, local_rank, C.byref(comm)))
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
    n_frames = ring * B
    h_img, d_img, d_stage, d_maps, d_m1, kpf = [], [], [], [], [], []
    for li, c in enumerate(mine):
        # every camera sees the same four scenes, displaced horizontally by 14 px per camera index (so that overlapping
        # pairs share content and the stereo matcher's gates run on real candidates), then drifting frame to frame
        base = [np.roll(synth_frame(3000 + i, W, H), 14 * c, 1) for i in range(4)]
        frames = np.stack([np.roll(base[i % 4], (3 * (i // 4), 5 * (i // 4)), (0, 1)) for i in range(n_frames)])
        h_img.append(torch.from_numpy(frames).pin_memory()); d_img.append(torch.from_numpy(frames).cuda())
        d_stage.append(torch.zeros((B, H, W), dtype=torch.uint8, device="cuda"))
        mf = MultiFrame(len(mine)); mf.setImage(li, frames[0]); fe.detectAndDescribe(li, mf)
        kpf.append(len(mf.frames[li].keypoints))
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
    mk_st = lambda dev: dict(k1=torch.zeros((B, kp_cap), dtype=torch.int32, device=dev), dist=torch.zeros((B, kp_cap), dtype=torch.int32, device=dev),
                             hp=torch.zeros((B, kp_cap, 4), dtype=torch.float64, device=dev), init=torch.zeros((B, kp_cap), dtype=torch.uint8, device=dev))
    d_st = [mk_st("cuda") for _ in pairs]
    # host mirrors of the results (e2e leg)
    h_blk = torch.zeros((slots, blk), dtype=torch.uint8).pin_memory()
    h_m1 = [dict(dist=torch.zeros((B, kp_cap), dtype=torch.int32).pin_memory(), lm=torch.zeros((B, kp_cap), dtype=torch.int32).pin_memory()) for _ in mine]
    h_st = [{k: v.pin_memory() for k, v in mk_st("cpu").items()} for _ in pairs]
    cur = torch.cuda.current_stream()
    cam_streams = [torch.cuda.ExternalStream(L_.okb_stream(ctx, li)) for li in range(len(mine))]
    bytes_h2d = len(mine) * B * W * H
    bytes_d2h = len(mine) * (blk + 2 * B * kp_cap * 4) + len(pairs) * B * kp_cap * 41

    def step(s, host):
        for li, c in enumerate(mine):
            sl = slice((s % ring) * B, (s % ring + 1) * B)
            if host:   # H2D on torch's stream (the page-locked tensors must never be tied to a library stream that okb_destroy ends)
                d_stage[li].copy_(h_img[li][sl], non_blocking=True)
                frames = d_stage[li]
            else:
                frames = d_img[li][sl]
            cam_streams[li].wait_stream(cur)      # the copy is done and the previous step's matchers have finished reading the features
            okl.check(L_.okb_detect_describe_batch_device(ctx, li, B, frames.data_ptr()))
            dm = d_maps[li]
            okl.check(L_.okb_match_map3d_device(ctx, li, 64, B, len(dm["lm"]), dm["desc"].data_ptr(), dm["lm"].data_ptr(), len(dm["is3d"]),
                                                dm["proj"].data_ptr(), dm["is3d"].data_ptr(), 20.0, 60, d_m1[li]["dist"].data_ptr(),
                                                d_m1[li]["lm"].data_ptr()))
            okl.check(L_.okb_export_features(ctx, li, B, local[c // world].data_ptr()))
        # the one collective of the path: NCCL all-gather behind the C ABI, enqueued on the LAST local camera's stream (which first
        # waits for the exports of the other local cameras); the stereo matchers' stream waits on its event
        if mine:
            last = len(mine) - 1
            for li in range(last):
                cam_streams[last].wait_stream(cam_streams[li])
            okl.check(L_.okb_allgather_features(comm, 1, (C.c_void_p * 1)(ctx), (C.c_int * 1)(last), (C.c_void_p * 1)(local.data_ptr()),
                                                (C.c_void_p * 1)(gathered.data_ptr()), local.numel()))
            okl.check(L_.okb_comm_wait(comm, 0, cur.cuda_stream))
            for li in range(len(mine)):
                cur.wait_stream(cam_streams[li])      # M1 of every local camera is part of the step
        for pi, (i, j) in enumerate(pairs):
            bi = gathered.data_ptr() + (sh.slot_of(i, world)[0] * slots + sh.slot_of(i, world)[1]) * blk
            bj = gathered.data_ptr() + (sh.slot_of(j, world)[0] * slots + sh.slot_of(j, world)[1]) * blk
            o = d_st[pi]
            okl.check(L_.okb_match_stereo_device_ptr(ctx, B, kp_cap, bi + o_k, bi + o_d, bi + o_c, C.byref(models[i]), C_WC[i].ctypes.data,
                                                     r_WC[i].ctypes.data, kp_cap, bj + o_k, bj + o_d, bj + o_c, C.byref(models[j]),
                                                     C_WC[j].ctypes.data, r_WC[j].ctypes.data, 60, cur.cuda_stream, o["k1"].data_ptr(),
                                                     o["dist"].data_ptr(), o["hp"].data_ptr(), o["init"].data_ptr()))
        if host:   # results back to the host: feature blocks of the own cameras, M1 and M4 outputs
            h_blk.copy_(local, non_blocking=True)
            for li in range(len(mine)):
                for k in ("dist", "lm"):
                    h_m1[li][k].copy_(d_m1[li][k], non_blocking=True)
            for pi in range(len(pairs)):
                for k in ("k1", "dist", "hp", "init"):
                    h_st[pi][k].copy_(d_st[pi][k], non_blocking=True)
            cur.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        if world == 1:
            return v
        t = torch.tensor([v], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.item())

    for s in range(warm):
        step(s, False)
    barrier()
    launches0 = L_.okb_launch_count(ctx)
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record(cur)
    for s in range(steps):
        step(warm + s, False)
    for st in cam_streams:
        cur.wait_stream(st)
    ev1.record(cur)
    barrier()
    ms = allmax(ev0.elapsed_time(ev1))
    launches = L_.okb_launch_count(ctx) - launches0
    n_match = int(sum((o["k1"] >= 0).sum().item() for o in d_st))
    # e2e: the same step with the images coming from host memory and the results going back, wall clock
    for s in range(2):
        step(s, True)
    barrier()
    t0 = time.perf_counter()
    for s in range(steps):
        step(warm + s, True)
    torch.cuda.synchronize()
    e2e_s = allmax(time.perf_counter() - t0)
    # roofline of the pyramid+score pass of this rank's first camera
    roofline = None
    if mine:
        L_.okb_enable_timers(ctx, 1); L_.okb_reset_timers(ctx)
        for s in range(3):
            step(s, False)
        okl.check(L_.okb_sync(ctx)); torch.cuda.synchronize()
        ps_ms = C.c_double(); ps_l = C.c_int64(); tot = C.c_double()
        L_.okb_get_timers(ctx, 0, C.byref(ps_ms), C.byref(ps_l), C.byref(tot))
        L_.okb_enable_timers(ctx, 0)
        ps_bytes = L_.okb_pyramid_score_bytes(ctx, 0)
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0))
        except Exception:
            peak = 6650.0
        ach = ps_bytes * B * 3 / (ps_ms.value * 1e-3) / 1e9 if ps_ms.value > 0 else 0.0
        roofline = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                    "kernel": "pyramid+score pass of one camera of rank 0 (the other cameras' streams run beside it)", "bytes_per_image": int(ps_bytes),
                    "images_per_pass": B, "ms_per_pass": ps_ms.value / 3}
    rec = None
    if rank == 0:
        rec = {"metric": "multiframes/sec detect+describe+match (5-camera rig)", "value": B * steps / (ms * 1e-3), "unit": "multiframes/s",
               "n_gpus": world, "steps": steps, "ms_per_step": ms / steps, "scaling": "strong",
               "config": {"workload": name, "multiframes_per_step": B, "cameras": n_cams, "overlapping_pairs": [list(p) for p in overlaps],
                          "parallelism": f"camera c on GPU c % {world}; one NCCL all-gather (okb_allgather_features, C ABI) of {slots} x {blk} B feature blocks per rank per step",
                          "l2_policy": f"inputs larger than L2: ring of {ring} batches",
                          "W": W, "H": H, "keypoints_per_frame": cfg["kpts"], "detector_max_keypoints": cfg["max_kp"], "threshold": cfg["threshold"],
                          "octaves": cfg["octaves"], "n_lm": cfg["n_lm"]},
               "keypoints_per_frame_rank0": float(np.mean(kpf)) if kpf else None,
               "roofline": roofline,
               "e2e": {"value": B * steps / e2e_s, "unit": "multiframes/s", "h2d_bytes_per_step": int(bytes_h2d), "d2h_bytes_per_step": int(bytes_d2h),
                       "ms_per_step": 1e3 * e2e_s / steps,
                       "note": "per rank: images of its cameras H2D from page-locked memory, feature blocks + M1 + M4 results D2H, every step"},
               "gpu_launches": int(launches), "stereo_matches_rank0_last_step": n_match}
    torch.cuda.synchronize()
    del h_img, h_blk, h_m1, h_st, cam_streams
    L_.okb_comm_destroy(comm)
    fe.close()
    return rec


def rig_overlaps_exact(rig, compute):
    """the overlapping camera pairs (i < j) of a rig by NCameraSystem::computeOverlaps; `compute` = Frontend.computeOverlaps (device)
    or oracle.compute_overlaps (CPU arm)"""
    from okvis2_b200.frontend import Frontend
    n = len(rig)
    models = [Frontend.MODELS[r["distortion_type"]] for r in rig]
    intr = [list(r["focal_length"]) + list(r["principal_point"]) + (list(r["distortion_coefficients"]) + [0.0] * 4)[:4] for r in rig]
    Wd = [r["image_dimension"][0] for r in rig]; Hd = [r["image_dimension"][1] for r in rig]
    Cs = [np.array(r["T_SC"]).reshape(4, 4)[:3, :3] for r in rig]
    C_rel = np.array([[Cs[s].T @ Cs[c] for c in range(n)] for s in range(n)])
    ov = compute(models, intr, Wd, Hd, C_rel)
    return [(i, j) for i in range(n) for j in range(i + 1, n) if ov[i][j]]


def cpu_sharded_baseline(cfg):
    """CPU arm of the 5-camera workload on a bounded sample: per multiframe detect+describe+M1 of the 5 cameras and M4 of the
    overlapping pairs (oracle), all host cores over the multiframes."""
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    from okvis2_b200 import rigs, sharding as sh
    from okvis2_b200.synth import synth_frame
    rig = getattr(rigs, cfg["rig"])
    W, H = cfg["W"], cfg["H"]
    overlaps = rig_overlaps_exact(rig, oracle.compute_overlaps)
    cores = os.cpu_count() or 1
    n = max(2, min(8, cores // 2))
    local = threading.local()
    frames = [[np.roll(synth_frame(3000 + i % 4, W, H), 14 * c, 1) for c in range(len(rig))] for i in range(n)]
    o0 = oracle.Brisk(cfg["threshold"], cfg["octaves"])
    maps = []
    for c in range(len(rig)):
        kp, d = o0.detect_and_compute(frames[0][c], cfg["max_kp"])
        maps.append(make_map(cfg, kp, d, 70 + c))
    T = [np.array(r["T_SC"]).reshape(4, 4) for r in rig]
    modelid = {"none": 0, "radialtangential": 1, "equidistant": 2}

    def job(i):
        if not hasattr(local, "o"):
            local.o = oracle.Brisk(cfg["threshold"], cfg["octaves"])
        feats = []
        for c, r in enumerate(rig):
            kp, d = local.o.detect_and_compute(frames[i][c], cfg["max_kp"])
            m = maps[c]
            oracle.match_map3d(d, np.stack([kp["x"], kp["y"]], 1).astype(np.float64), None, m["cand_desc"], m["cand_lm"], m["lm_proj"], m["lm_is3d"], 20.0, 60, 1)
            k = list(r["distortion_coefficients"]) + [0.0] * (4 - len(r["distortion_coefficients"]))
            rays, valid = oracle.back_project(modelid[r["distortion_type"]], *r["focal_length"], *r["principal_point"], k, kp)
            w = rays @ T[c][:3, :3].T
            e = np.ascontiguousarray(w / np.linalg.norm(w, axis=1, keepdims=True))
            feats.append((d, valid, e, kp["size"].astype(np.float64) / (0.5 * sum(r["focal_length"]))))
        for (a, b) in overlaps:
            Ta = np.concatenate([T[a][:3, :3].T, (-(T[a][:3, :3].T @ T[a][:3, 3]))[:, None]], 1).reshape(12)
            Tb = np.concatenate([T[b][:3, :3].T, (-(T[b][:3, :3].T @ T[b][:3, 3]))[:, None]], 1).reshape(12)
            oracle.match_stereo(*feats[a], *feats[b], T[a][:3, 3].copy(), T[b][:3, 3].copy(), Ta, Tb, 60)
    with ThreadPoolExecutor(cores) as ex:
        list(ex.map(job, range(min(2, n))))
        t0 = time.perf_counter()
        list(ex.map(job, range(n)))
        t = time.perf_counter() - t0
    return {"value": n / t, "unit": "multiframes/s", "cores": cores, "kind": "port",
            "sample": f"{n} multiframes (5 cameras each) on {cores} host threads: detect+describe by the C restatement of OpenCV-BRISK, M1 per camera, M4 for the overlapping pairs"}


# ---------------------------------------------------------------------------------------------------------------
def bench_next_rows(fe, cfg):
    """P1 (Frontend.cpp:1196-1360) on a 50 000-landmark map: the C-ABI call (H2D of the map, kernels, D2H of the pool)
    against the oracle transcription on one host core; outputs compared. K1 likewise."""
    import oracle
    from okvis2_b200.synth import landmark_scene
    s = landmark_scene(77, n_lm=50000, n_slots=50, n_cams=2, n_kp=cfg["kpts"], W=cfg["W"], H=cfg["H"], f=cfg["f"], step=0.05)
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
    intr = np.array([cfg["f"], cfg["f"] * 0.997, s["W"] / 2 - 8.8, s["H"] / 2 + 8.4] + DIST)
    t0 = time.perf_counter()
    ref = oracle.prepare_landmarks(s["hp_W"], s["quality"], s["obs_begin"], s["obs"], 2, s["T_WC_old"], s["desc_tab"], s["ray_tab"], 64,
                                   s["T_WC1"], s["T_CW1"], 1, intr, s["W"], s["H"])
    cpu_ms = (time.perf_counter() - t0) * 1e3
    same = all(np.array_equal(got[k], ref[k]) for k in ("lm", "lm_is3d", "desc_begin", "cand_desc", "kid", "lm_proj"))
    rows = {"P1_prepare_landmarks": {"landmarks": 50000, "observations": int(len(s["obs"])), "kept": int(len(got["lm"])),
                                     "pool_rows": int(len(got["cand_desc"])), "gpu_ms_host_buffers": gpu_ms, "cpu_port_ms_1_core": cpu_ms,
                                     "identical_to_oracle": bool(same)}}
    rng = np.random.default_rng(3)
    views = []
    for _ in range(22):
        xy = np.stack([rng.uniform(0, cfg["W"], cfg["kpts"]), rng.uniform(0, cfg["H"], cfg["kpts"])], 1).astype(np.float32)
        views.append((cfg["H"], cfg["W"], xy, rng.random(cfg["kpts"]) < 0.5))
    inter, uni = fe._overlap_counts(views)
    t0 = time.perf_counter()
    for _ in range(10):
        inter, uni = fe._overlap_counts(views)
    k1_gpu = (time.perf_counter() - t0) / 10 * 1e3
    t0 = time.perf_counter()
    ref = [oracle.overlap_counts(r, c, xy, m) for (r, c, xy, m) in views]
    k1_cpu = (time.perf_counter() - t0) * 1e3
    rows["K1_keyframe_overlap"] = {"views": len(views), "keypoints_per_view": cfg["kpts"], "gpu_ms_host_buffers": k1_gpu,
                                   "cpu_port_ms_1_core": k1_cpu,
                                   "identical_to_oracle": bool(all((inter[i], uni[i]) == ref[i] for i in range(len(views))))}
    return rows


def bind_to_gpu_numa_node(index):
    """One process per GPU: run its host threads on the CPUs next to that GPU (NVML's ideal CPU set), so that the page-locked
    buffers of the e2e legs (first touched by these threads) and the copy-engine traffic stay on the GPU's NUMA node. Round 1's
    end-to-end figure at 8 GPUs was bound by the host memory system. Returns the number of CPUs bound to, or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if len(cpus) >= 4:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def main():
    # stdout carries ONE JSON line: NCCL's version banner / debug output (torch's process group, okb_comm_init_*) goes to stderr
    # (NCCL honours NCCL_DEBUG_FILE only above the VERSION level, and prints the banner at VERSION and WARN)
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="euroc", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the sub-records of the other workloads")
    ap.add_argument("--e2e-lanes", "--lanes", dest="e2e_lanes", type=int, default=0,
                    help="independent sequences in flight per GPU (own library handle each) in the value and e2e legs; default: 4, fewer "
                         "when world x lanes x 3 host threads would outnumber the box's cores (never below 2)")
    args = ap.parse_args()
    name, cfg = args.config, CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, name, cfg, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None   # before any page-locked allocation
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sampler = ClockSampler(local_rank); sampler.start()   # samples clocks through every leg
    lanes = args.e2e_lanes
    if lanes <= 0:
        # every sequence costs three host threads in the e2e leg (two cameras + the submitting thread); an oversubscribed host was what
        # limited the end-to-end figure at 8 GPUs in round 1, and one sequence alone now reaches 94 % of the device throughput
        lanes = max(2, min(4, (os.cpu_count() or 8) // (3 * world)))
    if "rig" in cfg:
        rec = run_sharded(args, name, cfg, rank, world, local_rank, args.steps)
        clocks = sampler.stop()
        if rank == 0:
            rec.update({"warmup": max(args.warmup, 3), "higher_is_better": True, "vs_baseline": None, "dtype": "u8", "data": "synthetic", "clocks": clocks,
                        "cpu_baseline": None if (args.no_cpu_baseline or world > 1) else cpu_sharded_baseline(cfg)})
            print(json.dumps(rec))
        if world > 1:
            dist.destroy_process_group()
        return

    if os.environ.get("OKB_BENCH_OKVIS48_ONLY"):   # development hook: the D = 48 sub-record alone
        print(json.dumps(run_okvis48(args, rank, world, local_rank, args.steps, with_cpu=not args.no_cpu_baseline)))
        return
    head = run_replica(name, cfg, args, rank, world, local_rank, lanes, args.steps, full=True)
    rep = head.pop("_rep")
    gate_exact = bool(rep.L_.okb_gate_cos_exact(rep.fes[0].ctx))
    # ---- sub-records: the other BASELINE.json workloads in the same run (fewer steps each)
    configs = {}
    if not args.no_configs:
        sub_steps = max(4, min(args.steps, 8))
        for sub in ("euroc_octaves0", "tumvi"):
            if sub == name or (os.environ.get("OKB_BENCH_SUBS") and sub not in os.environ["OKB_BENCH_SUBS"].split(",")):   # development hook
                continue
            try:
                r = run_replica(sub, CONFIGS[sub], args, rank, world, local_rank, lanes if sub != "tumvi" else min(lanes, 2), sub_steps, full=False)
                r["steps"] = sub_steps
                configs[sub] = r
            except SystemExit:
                raise
            except Exception as e:   # never lose the headline line to a sub-record
                configs[sub] = {"error": repr(e)}
        if rank == 0:   # one GPU's number at any N (the other ranks wait in the next collective)
            try:
                configs["euroc_okvis48"] = run_okvis48(args, rank, world, local_rank, sub_steps, with_cpu=(world == 1 and not args.no_cpu_baseline))
                configs["euroc_okvis48"]["n_gpus"] = 1
            except SystemExit:
                raise
            except Exception as e:
                configs["euroc_okvis48"] = {"error": repr(e)}
        try:
            configs["hilti_sharded"] = run_sharded(args, "hilti", CONFIGS["hilti"], rank, world, local_rank, max(4, min(args.steps, 8)))
        except Exception as e:
            configs["hilti_sharded"] = {"error": repr(e)}
    next_rows = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            next_rows = bench_next_rows(rep.fes[0], cfg)
        except Exception as e:   # never lose the headline line to an auxiliary measurement
            next_rows = {"error": repr(e)}
    clocks = sampler.stop()
    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload on the host cores
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        cpu = cpu_arm(cfg, rep.wl, n_sample=max(8, min(32, 2 * cores)), n_latency=12)
        if isinstance(configs.get("hilti_sharded"), dict) and "error" not in configs["hilti_sharded"]:
            try:
                configs["hilti_sharded"]["cpu_baseline"] = cpu_sharded_baseline(CONFIGS["hilti"])
            except Exception as e:
                configs["hilti_sharded"]["cpu_baseline"] = {"error": repr(e)}
    rep.close()
    if rank == 0:
        line = {"metric": "stereo frames/sec detect+describe+match", "value": head["value"], "unit": "stereo frames/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": head["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": head["config"],
                "run": {"stereo_frames_per_step_per_gpu": head["stereo_frames_per_step_per_gpu"], "sequences_in_flight_per_gpu": lanes,
                        "parallelism": "replicas (independent sequences per GPU)" if world > 1 else "single GPU",
                        "gate_cos_equals_libm": gate_exact,
                        "orientation_atan2": "device fp64 atan2 (<= 2 ulp before rounding to fp32): a keypoint angle can differ from cv2 with probability ~2^-27"},
                "workload_stats": head["workload_stats"], "roofline": head["roofline"], "matcher": head.get("matcher"), "cpu_baseline": cpu,
                "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "clocks": clocks, "configs": configs, "next_rows": next_rows,
                "host": {"cpus": os.cpu_count(), "sequences_in_flight_per_gpu": lanes,
                         "cpus_bound_per_rank": numa, "note": "N > 1: every rank runs on the CPUs NVML lists as local to its GPU"}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
