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
