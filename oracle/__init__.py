"""CPU oracle (TEST INFRASTRUCTURE ONLY -- never imported by okvis2_b200/).

ctypes bindings of oracle/liboracle.so (C restatement, see brisk_oracle.c / match_oracle.c headers).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28


def build(force=False):
    """Compile the oracle with gcc (oracle/Makefile)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".cpp", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        u8p = C.POINTER(C.c_uint8)
        L.okvo_brisk_create.restype = C.c_void_p
        L.okvo_brisk_create.argtypes = [C.c_int, C.c_int, C.c_float]
        L.okvo_brisk_destroy.argtypes = [C.c_void_p]
        L.okvo_brisk_descriptor_bytes.argtypes = [C.c_void_p]
        L.okvo_brisk_pattern.restype = C.POINTER(C.c_float)
        L.okvo_brisk_pattern.argtypes = [C.c_void_p]
        L.okvo_brisk_detect_raw.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.okvo_brisk_compute.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.okvo_brisk_detect_and_compute.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                                    C.c_void_p, C.c_int, C.c_void_p]
        L.okvo_resize_area.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.okvo_brisk_layer.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                       C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(u8p), C.POINTER(u8p)]
        L.okvo_brisk_num_layers.argtypes = [C.c_void_p]
        L.okvo_cap_strongest.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.okvo_integral.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.okvo_brisk_size_list.argtypes = [C.c_void_p, C.c_void_p]
        L.okvo_brisk_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.okvo_brisk_num_pairs.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.okvo_brisk_kscale.argtypes = [C.c_float]
        L.okvo_oast916_bstar.argtypes = [C.c_void_p, C.c_int]
        L.okvo_agast58_bstar.argtypes = [C.c_void_p, C.c_int]
    return _LIB


def resize_area(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((dh, dw), np.uint8)
    lib().okvo_resize_area(src.ctypes.data, src.shape[1], src.shape[0], src.shape[1], dst.ctypes.data, dw, dh, dw)
    return dst


def dense_b0(img):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.zeros_like(img)
    f = lib().okvo_dense_b0
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    f(img.ctypes.data, img.shape[1], img.shape[0], out.ctypes.data)
    return out


def integral(img):
    img = np.ascontiguousarray(img, np.uint8)
    H, W = img.shape
    out = np.empty((H + 1, W + 1), np.int32)
    lib().okvo_integral(img.ctypes.data, W, H, W, out.ctypes.data)
    return out


class Brisk:
    """Oracle detector/extractor object (OpenCV-BRISK semantics + okvis max_num_keypoints cap)."""

    def __init__(self, threshold=30, octaves=3, pattern_scale=1.0):
        self.h = lib().okvo_brisk_create(threshold, octaves, pattern_scale)
        self.D = lib().okvo_brisk_descriptor_bytes(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().okvo_brisk_destroy(self.h)
            self.h = None

    def detect_raw(self, img, cap=1 << 17):
        """getKeypoints only: no cap, no border removal, angle = -1."""
        img = np.ascontiguousarray(img, np.uint8)
        kp = np.zeros(cap, KP_DTYPE)
        n = lib().okvo_brisk_detect_raw(self.h, img.ctypes.data, img.shape[1], img.shape[0], img.shape[1],
                                        kp.ctypes.data, cap)
        assert n <= cap
        return kp[:n].copy()

    def compute(self, img, kp):
        img = np.ascontiguousarray(img, np.uint8)
        kp = np.ascontiguousarray(kp, KP_DTYPE).copy()
        desc = np.zeros((max(len(kp), 1), self.D), np.uint8)
        m = lib().okvo_brisk_compute(self.h, img.ctypes.data, img.shape[1], img.shape[0], kp.ctypes.data, len(kp),
                                     desc.ctypes.data)
        return kp[:m].copy(), desc[:m].copy()

    def detect_and_compute(self, img, max_kp=0, cap=1 << 17):
        img = np.ascontiguousarray(img, np.uint8)
        kp = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, self.D), np.uint8)
        n = lib().okvo_brisk_detect_and_compute(self.h, img.ctypes.data, img.shape[1], img.shape[0], img.shape[1],
                                                max_kp, kp.ctypes.data, cap, desc.ctypes.data)
        assert n >= 0
        return kp[:n].copy(), desc[:n].copy()

    def layers(self):
        """(img, scores, scale, offset) of every layer of the last pyramid (scores = lazily filled cache)."""
        out = []
        u8p = C.POINTER(C.c_uint8)
        for i in range(lib().okvo_brisk_num_layers(self.h)):
            w, h, s, o = C.c_int(), C.c_int(), C.c_float(), C.c_float()
            pi, ps = u8p(), u8p()
            lib().okvo_brisk_layer(self.h, i, C.byref(w), C.byref(h), C.byref(s), C.byref(o), C.byref(pi), C.byref(ps))
            img = np.ctypeslib.as_array(pi, (h.value, w.value)).copy()
            sc = np.ctypeslib.as_array(ps, (h.value, w.value)).copy()
            out.append((img, sc, s.value, o.value))
        return out

    def pattern(self):
        p = lib().okvo_brisk_pattern(self.h)
        return np.ctypeslib.as_array(p, (64, 1024, 60, 3)).copy()

    def size_list(self):
        out = np.zeros(64, np.uint32)
        lib().okvo_brisk_size_list(self.h, out.ctypes.data)
        return out

    def pairs(self):
        ns, nl = C.c_int(), C.c_int()
        lib().okvo_brisk_num_pairs(self.h, C.byref(ns), C.byref(nl))
        sp = np.zeros((ns.value, 2), np.uint32)
        lp = np.zeros((nl.value, 4), np.int32)
        lib().okvo_brisk_pairs(self.h, sp.ctypes.data, lp.ctypes.data)
        return sp, lp


def cv_keypoints_to_array(kps):
    a = np.zeros(len(kps), KP_DTYPE)
    for i, k in enumerate(kps):
        a[i] = (k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave, k.class_id)
    return a


# ---- matchers (match_oracle.cpp) ------------------------------------------------------------------------------
def _p(a):
    return None if a is None else a.ctypes.data


def _c(a, t):
    return None if a is None else np.ascontiguousarray(a, t)


def hamming_matrix(a, b):
    a, b = _c(a, np.uint8), _c(b, np.uint8)
    out = np.zeros((len(a), len(b)), np.uint16)
    lib().okvo_hamming_matrix(C.c_int(a.shape[1]), C.c_int(len(a)), C.c_void_p(_p(a)), C.c_int(len(b)), C.c_void_p(_p(b)),
                              C.c_void_p(_p(out)))
    return out


def triangulate_fast(p1, e1, p2, e2, sigma):
    a = [np.ascontiguousarray(v, np.float64) for v in (p1, e1, p2, e2)]
    hp = np.zeros(4); v, p = C.c_int(), C.c_int()
    lib().okvo_triangulate_fast(*[C.c_void_p(x.ctypes.data) for x in a], C.c_double(sigma), C.c_void_p(hp.ctypes.data),
                                C.byref(v), C.byref(p))
    return hp, bool(v.value), bool(p.value)


def match_map3d(kp_desc, kp_xy, kp_use, cand_desc, cand_lm, lm_proj, lm_is3d, reproj_thr, match_thr, n_threads=1):
    kp_desc, cand_desc = _c(kp_desc, np.uint8), _c(cand_desc, np.uint8)
    n, D = kp_desc.shape
    kp_xy, kp_use = _c(kp_xy, np.float64), _c(kp_use, np.uint8)
    cand_lm, lm_proj, lm_is3d = _c(cand_lm, np.int32), _c(lm_proj, np.float64), _c(lm_is3d, np.uint8)
    dist = np.zeros(n, np.uint32); lm = np.zeros(n, np.int32)
    f = lib().okvo_match_map3d
    f.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                  C.c_void_p, C.c_void_p, C.c_double, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
    f(D, n, _p(kp_desc), _p(kp_xy), _p(kp_use), len(cand_desc), _p(cand_desc), _p(cand_lm), len(lm_is3d), _p(lm_proj),
      _p(lm_is3d), reproj_thr, match_thr, _p(dist), _p(lm), n_threads)
    return dist, lm


def match_map_uninit(kp_desc, kp_e_W, kp_use, kp_prev_lm, cand_desc, cand_lm, cand_e_W, cand_r_W, lm_is3d, r_WC1, sigma,
                     match_thr, n_threads=1):
    kp_desc, cand_desc = _c(kp_desc, np.uint8), _c(cand_desc, np.uint8)
    n, D = kp_desc.shape
    kp_e_W, kp_use, kp_prev_lm = _c(kp_e_W, np.float64), _c(kp_use, np.uint8), _c(kp_prev_lm, np.int32)
    cand_lm, cand_e_W, cand_r_W = _c(cand_lm, np.int32), _c(cand_e_W, np.float64), _c(cand_r_W, np.float64)
    lm_is3d, r = _c(lm_is3d, np.uint8), _c(r_WC1, np.float64)
    dist = np.zeros(n, np.uint32); lm = np.zeros(n, np.int32); hp = np.zeros((n, 4)); ctr = C.c_int32(0)
    f = lib().okvo_match_map_uninit
    f.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_void_p,
                  C.c_double, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    f(D, n, _p(kp_desc), _p(kp_e_W), _p(kp_use), _p(kp_prev_lm), len(cand_desc), _p(cand_desc), _p(cand_lm), _p(cand_e_W),
      _p(cand_r_W), len(lm_is3d), _p(lm_is3d), _p(r), sigma, match_thr, _p(dist), _p(lm), _p(hp), C.addressof(ctr), n_threads)
    return dist, lm, hp, ctr.value


def _stereo(name, desc0, use0, e0, sof0, desc1, valid1, e1, sof1, r0, r1, T0, T1, match_thr, n_threads):
    desc0, desc1 = _c(desc0, np.uint8), _c(desc1, np.uint8)
    n0, D = desc0.shape
    use0, valid1 = _c(use0, np.uint8), _c(valid1, np.uint8)
    e0, e1, sof0, sof1 = _c(e0, np.float64), _c(e1, np.float64), _c(sof0, np.float64), _c(sof1, np.float64)
    r0, r1, T0, T1 = _c(r0, np.float64), _c(r1, np.float64), _c(T0, np.float64), _c(T1, np.float64)
    k1 = np.zeros(n0, np.int32); dist = np.zeros(n0, np.uint32); hp = np.zeros((n0, 4)); init = np.zeros(n0, np.uint8)
    f = getattr(lib(), name)
    if name == "okvo_match_motion_stereo":
        f.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 7 + [C.c_uint32] + [C.c_void_p] * 4 + [C.c_int]
        f(D, n0, _p(desc0), _p(use0), _p(e0), _p(sof0), len(desc1), _p(desc1), _p(valid1), _p(e1), _p(r0), _p(r1), _p(T0),
          _p(T1), match_thr, _p(k1), _p(dist), _p(hp), _p(init), n_threads)
    else:
        f.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 8 + [C.c_uint32] + [C.c_void_p] * 4 + [C.c_int]
        f(D, n0, _p(desc0), _p(use0), _p(e0), _p(sof0), len(desc1), _p(desc1), _p(valid1), _p(e1), _p(sof1), _p(r0), _p(r1),
          _p(T0), _p(T1), match_thr, _p(k1), _p(dist), _p(hp), _p(init), n_threads)
    return k1, dist, hp, init


def match_motion_stereo(desc0, use0, e0, sof0, desc1, valid1, e1, r0, r1, T0, T1, match_thr, n_threads=1):
    return _stereo("okvo_match_motion_stereo", desc0, use0, e0, sof0, desc1, valid1, e1, None, r0, r1, T0, T1, match_thr, n_threads)


def match_motion_stereo_sequence(views, desc1, rays1, valid1, xy1, T_WC1, T_CW1, model, intr, width, height, match_thr,
                                 matched1, n_threads=1):
    """Frontend::matchMotionStereo over the older keyframes of one camera (Frontend.cpp:1775-1958).
    views: list of dicts(desc, rays, valid, size, use (or None), T_WC (12), T_CW (12)). Returns (list of per-view
    (k1, dist, hp_W, flags), matched1 after the sequence). flags: 1 matching, 2 initialisable, 4 inserted."""
    D = desc1.shape[1]
    n0 = np.array([len(v["desc"]) for v in views], np.int32)
    off = np.concatenate([[0], np.cumsum(n0)[:-1]]).astype(np.int32) if len(views) else np.zeros(0, np.int32)
    tot = int(n0.sum())
    cat = lambda k, t, shape: (np.ascontiguousarray(np.concatenate([np.asarray(v[k], t).reshape((-1,) + shape) for v in views]), t)
                               if views else np.zeros((0,) + shape, t))
    d0, r0, v0, s0 = cat("desc", np.uint8, (D,)), cat("rays", np.float64, (3,)), cat("valid", np.uint8, ()), cat("size", np.float32, ())
    use = None
    if any(v.get("use") is not None for v in views):
        use = np.ascontiguousarray(np.concatenate([np.ones(len(v["desc"]), np.uint8) if v.get("use") is None else np.asarray(v["use"], np.uint8)
                                                   for v in views]))
    Tw = np.ascontiguousarray(np.stack([np.asarray(v["T_WC"], np.float64).reshape(12) for v in views])) if views else np.zeros((0, 12))
    Tc = np.ascontiguousarray(np.stack([np.asarray(v["T_CW"], np.float64).reshape(12) for v in views])) if views else np.zeros((0, 12))
    desc1 = _c(desc1, np.uint8); rays1 = _c(rays1, np.float64); valid1 = _c(valid1, np.uint8); xy1 = _c(xy1, np.float32)
    T_WC1 = _c(np.asarray(T_WC1).reshape(12), np.float64); T_CW1 = _c(np.asarray(T_CW1).reshape(12), np.float64)
    intr = _c(intr, np.float64)
    m1 = np.ascontiguousarray(matched1, np.uint8).copy()
    k1 = np.zeros(tot, np.int32); dist = np.zeros(tot, np.uint32); hp = np.zeros((tot, 4), np.float64); fl = np.zeros(tot, np.uint8)
    f = lib().okvo_match_motion_stereo_sequence
    f.restype = None
    f.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 9 + [C.c_int] + [C.c_void_p] * 6 + [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_uint32] + \
                 [C.c_void_p] * 5 + [C.c_int]
    f(D, len(views), _p(n0), _p(off), _p(d0), _p(r0), _p(v0), _p(s0), _p(use), _p(Tw), _p(Tc), len(desc1), _p(desc1), _p(rays1), _p(valid1),
      _p(xy1), _p(T_WC1), _p(T_CW1), int(model), _p(intr), int(width), int(height), int(match_thr), _p(m1), _p(k1), _p(dist), _p(hp), _p(fl),
      n_threads)
    out = [(k1[o:o + n], dist[o:o + n], hp[o:o + n], fl[o:o + n]) for o, n in zip(off, n0)]
    return out, m1


def match_stereo(desc0, valid0, e0, sof0, desc1, valid1, e1, sof1, r0, r1, T0, T1, match_thr, n_threads=1):
    return _stereo("okvo_match_stereo", desc0, valid0, e0, sof0, desc1, valid1, e1, sof1, r0, r1, T0, T1, match_thr, n_threads)


def match_place(lm_offsets, lm_desc, kp_desc, match_thr):
    lm_offsets, lm_desc, kp_desc = _c(lm_offsets, np.int32), _c(lm_desc, np.uint8), _c(kp_desc, np.uint8)
    n_lm = len(lm_offsets) - 1
    k = np.zeros(n_lm, np.int32); dist = np.zeros(n_lm, np.uint32)
    f = lib().okvo_match_place
    f.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    f(kp_desc.shape[1], n_lm, _p(lm_offsets), _p(lm_desc), len(kp_desc), _p(kp_desc), match_thr, _p(k), _p(dist))
    return k, dist


def back_project(model, fu, fv, cu, cv, k, kp):
    """Frame::computeBackProjections restatement. kp: structured keypoint array. Returns (rays n x 3, valid n)."""
    kp = np.ascontiguousarray(kp, KP_DTYPE)
    n = len(kp)
    rays = np.zeros((n, 3)); valid = np.zeros(n, np.uint8)
    kk = np.ascontiguousarray(k, np.float64)
    f = lib().okvo_back_project
    f.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    f(model, fu, fv, cu, cv, kk.ctypes.data, n, kp.ctypes.data, 7, rays.ctypes.data, valid.ctypes.data)
    return rays, valid


def prepare_landmarks(hp_W, quality, obs_begin, obs, n_cams, T_WC_old, desc_tab, ray_tab, D, T_WC1, T_CW1, model, intr,
                      width, height, repr_thr=20.0, exclusive=False):
    """P1 oracle (prepare_oracle.cpp = Frontend.cpp:1196-1360). desc_tab / ray_tab: lists indexed slot * n_cams + cam."""
    L = lib()
    hp_W = _c(hp_W, np.float64); quality = _c(quality, np.float64); obs_begin = _c(obs_begin, np.int32)
    obs = _c(np.asarray(obs).reshape(-1, 3), np.int32); T_WC_old = _c(T_WC_old, np.float64)
    descs = [_c(d, np.uint8) for d in desc_tab]; rays = [_c(r, np.float64) for r in ray_tab]
    dt = (C.c_void_p * len(descs))(*[d.ctypes.data for d in descs]); rt = (C.c_void_p * len(rays))(*[r.ctypes.data for r in rays])
    n = len(quality)
    out = dict(lm=np.zeros(n, np.int32), lm_proj=np.zeros((n, 2)), lm_is3d=np.zeros(n, np.uint8), p_W=np.zeros((n, 3)),
               desc_begin=np.zeros(n + 1, np.int32), cand_desc=np.zeros((3 * n + 3, D), np.uint8), e_W=np.zeros((3 * n + 3, 3)),
               r_W=np.zeros((3 * n + 3, 3)), kid=np.zeros((3 * n + 3, 3), np.int32))
    n_rows = C.c_int32()
    T1 = _c(T_WC1, np.float64); T2 = _c(T_CW1, np.float64); intr = _c(intr, np.float64)
    L.okvo_prepare_landmarks.restype = C.c_int
    L.okvo_prepare_landmarks.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 2 + \
        [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int] + [C.c_void_p] * 10
    nl = L.okvo_prepare_landmarks(n, _p(hp_W), _p(quality), _p(obs_begin), _p(obs), n_cams, _p(T_WC_old), dt, rt, D, _p(T1), _p(T2),
                                  int(model), _p(intr), int(width), int(height), float(repr_thr), 1 if exclusive else 0,
                                  _p(out["lm"]), _p(out["lm_proj"]), _p(out["lm_is3d"]), _p(out["p_W"]), _p(out["desc_begin"]),
                                  _p(out["cand_desc"]), _p(out["e_W"]), _p(out["r_W"]), _p(out["kid"]), C.byref(n_rows))
    nr = n_rows.value
    for k in ("lm", "lm_proj", "lm_is3d", "p_W"):
        out[k] = out[k][:nl]
    out["desc_begin"] = out["desc_begin"][:nl + 1]
    for k in ("cand_desc", "e_W", "r_W", "kid"):
        out[k] = out[k][:nr]
    out["cand_lm"] = np.repeat(np.arange(nl, dtype=np.int32), np.diff(out["desc_begin"]))
    return out


def circle_filled(img, cx, cy, radius):
    """cv::circle(img, (cx, cy), radius, 255, FILLED) restated (overlap_oracle.cpp); img: u8 2-D, modified in place."""
    assert img.dtype == np.uint8 and img.flags.c_contiguous
    L = lib()
    L.okvo_circle_filled.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.okvo_circle_filled(img.ctypes.data, img.shape[1], img.shape[0], int(cx), int(cy), int(radius))
    return img


def overlap_counts(image_rows, image_cols, xy, matched, kptrad=0.09, masks=False):
    """K1 oracle: (intersection, union) pixel counts of the matches / detections masks of one camera image."""
    L = lib()
    xy = _c(np.asarray(xy, np.float32).reshape(-1, 2), np.float32); matched = _c(matched, np.uint8)
    rows, cols = image_rows // 10, image_cols // 10
    det = np.zeros((rows, cols), np.uint8); mat = np.zeros((rows, cols), np.uint8)
    i = C.c_int32(); u = C.c_int32()
    L.okvo_overlap_counts.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]
    L.okvo_overlap_counts(image_rows, image_cols, len(xy), _p(xy), _p(matched), float(kptrad), C.byref(i), C.byref(u), _p(det), _p(mat))
    return (i.value, u.value, det, mat) if masks else (i.value, u.value)


def camera_awareness_maps(model, intr, width, height):
    """PinholeCamera::initialiseCameraAwarenessMaps (PinholeCamera.hpp:179-208): (rays H x W x 3 f32, jacobians H x W x 6 f32)."""
    intr = _c(intr, np.float64)
    rays = np.zeros((height, width, 3), np.float32); jac = np.zeros((height, width, 6), np.float32)
    f = lib().okvo_camera_awareness_maps
    f.restype = None; f.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    f(int(model), _p(intr), int(width), int(height), _p(rays), _p(jac))
    return rays, jac


def compute_overlaps(models, intr, widths, heights, C_rel, masks=False):
    """NCameraSystem::computeOverlaps (NCameraSystem.cpp:48-118). C_rel: n x n x 3 x 3 relative rotations
    (T_SC[seenBy]^-1 * T_SC[cam]).C(). Returns the n x n boolean matrix (and the per-pair masks when asked)."""
    n = len(models)
    models = _c(models, np.int32); intr = _c(intr, np.float64); widths = _c(widths, np.int32); heights = _c(heights, np.int32)
    C_rel = _c(np.asarray(C_rel).reshape(n, n, 9), np.float64)
    out = np.zeros((n, n), np.uint8)
    offs = data = None
    if masks:
        sizes = np.array([[int(widths[c]) * int(heights[c]) for c in range(n)] for _ in range(n)], np.int64).reshape(-1)
        offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
        data = np.zeros(int(sizes.sum()), np.uint8)
    f = lib().okvo_compute_overlaps
    f.restype = None; f.argtypes = [C.c_int] + [C.c_void_p] * 8
    f(n, _p(models), _p(intr), _p(widths), _p(heights), _p(C_rel), _p(out), _p(offs), _p(data))
    if not masks:
        return out.astype(bool)
    mats = [[data[offs[s * n + c]:offs[s * n + c] + int(widths[c]) * int(heights[c])].reshape(int(heights[c]), int(widths[c])) for c in range(n)] for s in range(n)]
    return out.astype(bool), mats


class HarrisBrisk2:
    """The detector / extractor pair OKVIS2 constructs (Frontend.cpp:2406-2412): Harris score + uniformity enforcement, 48-byte BRISK2
    at one pattern scale, optionally camera-aware and aligned with the extraction direction (Frontend.cpp:232-251). octaves = 0.
    PARITY UNPINNED vs smartroboticslab/brisk@1ef8b42a (oracle/brisk_oracle.c section 6)."""

    def __init__(self, uniformity_radius=38.0, absolute_threshold=150, max_keypoints=700, pattern_scale=1.0):
        L = lib()
        L.okvo_brisk_create_dmax.restype = C.c_void_p
        L.okvo_brisk_create_dmax.argtypes = [C.c_int, C.c_int, C.c_float, C.c_double]
        L.okvo_harris_scores.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.okvo_harris_maxima.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.okvo_harris_lut.restype = C.c_float
        L.okvo_harris_lut.argtypes = [C.c_float, C.c_int, C.c_int]
        L.okvo_harris_detect.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.okvo_brisk2_compute.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_float, C.c_void_p]
        L.okvo_brisk2_warp.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
        self.h = L.okvo_brisk_create_dmax(0, 0, pattern_scale, 5.1)
        self.D = L.okvo_brisk_descriptor_bytes(self.h)
        self.radius, self.threshold, self.max_kp = float(uniformity_radius), int(absolute_threshold), int(max_keypoints)

    def __del__(self):
        if getattr(self, "h", None):
            lib().okvo_brisk_destroy(self.h)
            self.h = None

    @staticmethod
    def scores(img):
        img = np.ascontiguousarray(img, np.uint8)
        out = np.zeros(img.shape, np.int32)
        lib().okvo_harris_scores(_p(img), img.shape[1], img.shape[0], img.shape[1], _p(out))
        return out

    def maxima(self, score):
        score = np.ascontiguousarray(score, np.int32)
        n = lib().okvo_harris_maxima(_p(score), score.shape[1], score.shape[0], self.threshold, None, 0)
        xy = np.zeros((max(n, 1), 2), np.int32)
        lib().okvo_harris_maxima(_p(score), score.shape[1], score.shape[0], self.threshold, _p(xy), n)
        return xy[:n]

    def detect(self, img, cap=1 << 16):
        img = np.ascontiguousarray(img, np.uint8)
        kp = np.zeros(cap, KP_DTYPE)
        n = lib().okvo_harris_detect(_p(img), img.shape[1], img.shape[0], img.shape[1], self.radius, self.threshold, self.max_kp, _p(kp), cap)
        assert 0 <= n <= cap
        return kp[:n].copy()

    def compute(self, img, kp, rays=None, jac=None, fu=0.0, direction=None):
        img = np.ascontiguousarray(img, np.uint8)
        kp = np.ascontiguousarray(kp, KP_DTYPE).copy()
        desc = np.zeros((max(len(kp), 1), self.D), np.uint8)
        if rays is not None:
            rays = _c(rays, np.float32); jac = _c(jac, np.float32); direction = _c(direction, np.float32)
        m = lib().okvo_brisk2_compute(self.h, _p(img), img.shape[1], img.shape[0], _p(kp), len(kp), _p(desc),
                                      None if rays is None else _p(rays), None if rays is None else _p(jac), float(fu),
                                      None if rays is None else _p(direction))
        return kp[:m].copy(), desc[:m].copy()

    def detect_and_compute(self, img, rays=None, jac=None, fu=0.0, direction=None):
        return self.compute(img, self.detect(img), rays, jac, fu, direction)
