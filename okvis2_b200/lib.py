"""ctypes binding of libokvis_b200.so (the C ABI declared in include/okvis_b200.h).

The library is the product; this module only loads it and declares prototypes. There is no fallback:
if the shared object is missing, or no CUDA device is present, the calls fail loudly.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libokvis_b200.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])

OKB_OK, OKB_ERR_NO_DEVICE, OKB_ERR_CUDA, OKB_ERR_ARGUMENT, OKB_ERR_CAPACITY, OKB_ERR_UNSUPPORTED, OKB_ERR_NCCL = 0, -1, -2, -3, -4, -5, -6


class OkbError(RuntimeError):
    """Mirror of okvis::Frontend::Exception (reference okvis_frontend/include/okvis/Frontend.hpp:60)."""

    def __init__(self, status, msg):
        super().__init__(f"okvis_b200 status {status}: {msg}")
        self.status = status


class CameraConfig(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("threshold", C.c_int32), ("octaves", C.c_int32),
                ("max_keypoints", C.c_int32), ("descriptor_bytes", C.c_int32), ("max_batch", C.c_int32),
                ("pattern_scale", C.c_float), ("uniformity_radius", C.c_float)]


class CameraModel(C.Structure):
    """okb_camera_model_t: model 0 none / 1 radial-tangential / 2 equidistant"""
    _fields_ = [("model", C.c_int32), ("reserved", C.c_int32), ("fu", C.c_double), ("fv", C.c_double), ("cu", C.c_double),
                ("cv", C.c_double), ("k", C.c_double * 4)]


def build(force=False):
    """Compile libokvis_b200.so for sm_100a with nvcc (okvis2_b200/csrc/Makefile)."""
    src = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(src, f) for f in os.listdir(src) if f.endswith((".cu", ".h"))]
    srcs.append(os.path.join(_HERE, "..", "include", "okvis_b200.h"))
    stale = not os.path.exists(SO_PATH) or any(os.path.getmtime(s) > os.path.getmtime(SO_PATH) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-s", "-j4", "-C", src])
    return SO_PATH


_LIB = None
vp, i32, u32, f64 = C.c_void_p, C.c_int, C.c_uint32, C.c_double

_PROTOS = {
    "okb_create": (i32, [i32, i32, vp, C.POINTER(vp)]),
    "okb_destroy": (None, [vp]),
    "okb_last_error": (C.c_char_p, []),
    "okb_version": (C.c_char_p, []),
    "okb_launch_count": (C.c_int64, [vp]),
    "okb_gate_cos_exact": (i32, [vp]),
    "okb_stream": (vp, [vp, i32]),
    "okb_sync": (i32, [vp]),
    "okb_set_blocking_sync": (i32, [vp, i32]),
    "okb_detect_describe": (i32, [vp, i32, vp, C.c_size_t, vp, vp, i32, vp]),
    "okb_detect_describe_batch": (i32, [vp, i32, i32, vp, C.c_size_t, vp, vp, i32, vp]),
    "okb_detect_describe_batch_device": (i32, [vp, i32, i32, vp]),
    "okb_fetch_features": (i32, [vp, i32, i32, vp, vp, i32, vp]),
    "okb_device_features": (i32, [vp, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(i32)]),
    "okb_feature_block_bytes": (C.c_size_t, [i32, i32]),
    "okb_export_features": (i32, [vp, i32, i32, vp]),
    "okb_comm_unique_id": (i32, [vp]),
    "okb_comm_init_all": (i32, [i32, vp, C.POINTER(vp)]),
    "okb_comm_init_rank": (i32, [i32, i32, vp, i32, C.POINTER(vp)]),
    "okb_comm_destroy": (None, [vp]),
    "okb_comm_world": (i32, [vp]),
    "okb_comm_local_ranks": (i32, [vp]),
    "okb_allgather_features": (i32, [vp, i32, vp, vp, vp, vp, C.c_size_t]),
    "okb_comm_wait": (i32, [vp, i32, vp]),
    "okb_num_layers": (i32, [vp, i32]),
    "okb_layer_info": (i32, [vp, i32, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "okb_fetch_layer": (i32, [vp, i32, i32, i32, vp, vp]),
    "okb_debug_stamps": (i32, [vp, i32, i32, vp]),
    "okb_pyramid_score_bytes": (C.c_int64, [vp, i32]),
    "okb_enable_timers": (i32, [vp, i32]),
    "okb_reset_timers": (i32, [vp]),
    "okb_get_score_kernel_ms": (i32, [vp, i32, C.POINTER(f64)]),
    "okb_get_timers": (i32, [vp, i32, C.POINTER(f64), C.POINTER(C.c_int64), C.POINTER(f64)]),
    "okb_match_map3d": (i32, [vp, i32, i32, vp, vp, vp, i32, vp, vp, i32, vp, vp, f64, u32, vp, vp]),
    "okb_match_map_uninit": (i32, [vp, i32, i32, vp, vp, vp, vp, i32, vp, vp, vp, vp, i32, vp, vp, f64, u32, vp, vp, vp, vp]),
    "okb_match_motion_stereo": (i32, [vp, i32, i32, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, u32, vp, vp, vp, vp]),
    "okb_match_stereo": (i32, [vp, i32, i32, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, u32, vp, vp, vp, vp]),
    "okb_match_place": (i32, [vp, i32, i32, vp, vp, i32, vp, u32, vp, vp]),
    "okb_hamming_matrix": (i32, [vp, i32, i32, vp, i32, vp, vp]),
    "okb_set_camera_model": (i32, [vp, i32, vp]),
    "okb_last_back_projections": (i32, [vp, i32, i32, i32, vp, vp, vp]),
    "okb_camera_awareness_maps": (i32, [vp, i32, vp, vp]),
    "okb_compute_overlaps": (i32, [vp, i32, vp, vp, vp, vp, vp, vp]),
    "okb_back_project": (i32, [vp, i32, i32, vp, vp, vp]),
    "okb_match_stereo_device": (i32, [vp, i32, i32, i32, vp, vp, vp, vp, u32, vp, vp, vp, vp]),
    "okb_match_stereo_device_ptr": (i32, [vp, i32, i32, vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, u32, vp, vp, vp, vp, vp]),
    "okb_process_multiframe": (i32, [vp, i32, vp, i32, vp, f64, u32]),
    "okb_stream_use_graph": (i32, [vp, i32]),
    "okb_match_map_uninit_device": (i32, [vp, i32, i32, i32, vp, vp, vp, vp, i32, vp, vp, C.c_double, u32, vp, vp, vp, vp, vp, vp]),
    "okb_set_extraction_direction": (i32, [vp, i32, vp]),
    "okb_get_extraction_direction": (i32, [vp, i32, vp]),
    "okb_m3_set_fused": (None, [i32]),
    "okb_scan_set_mma": (None, [i32]),
    "okb_stream_timing": (i32, [vp, C.POINTER(C.c_double), i32]),
    "okb_stream_stats": (i32, [vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "okb_device_back_projections": (i32, [vp, i32, C.POINTER(vp), C.POINTER(vp)]),
    "okb_matched_mask_device": (i32, [vp, i32, i32, vp, vp]),
    "okb_match_motion_stereo_batch": (i32, [vp, i32, i32, vp, vp, i32, vp, i32, u32, i32, vp, i32, vp, vp, vp, vp, vp]),
    "okb_match_motion_stereo_device": (i32, [vp, i32, i32, vp, vp, i32, vp, i32, u32, vp, vp, vp, vp, vp]),
    "okb_match_motion_stereo_device_ptr": (i32, [vp, i32, i32, vp, vp, vp, vp, i32, i32, vp, vp, i32, vp, i32, u32, vp, vp, vp, vp, vp, vp]),
    "okb_match_map3d_device": (i32, [vp, i32, i32, i32, i32, vp, vp, i32, vp, vp, f64, u32, vp, vp]),
    "okb_match_map3d_batch": (i32, [vp, i32, i32, i32, i32, vp, vp, i32, vp, vp, f64, u32, i32, vp, vp]),
    "okb_match_stereo_batch": (i32, [vp, i32, i32, i32, vp, vp, vp, vp, u32, i32, vp, vp, vp, vp]),
    "okb_store_configure": (i32, [vp, i32, i32, i32]),
    "okb_store_frame": (i32, [vp, i32, i32, i32, vp, vp]),
    "okb_store_frame_from_last": (i32, [vp, i32, i32, i32]),
    "okb_prepare_landmarks": (i32, [vp, vp, i32, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "okb_bow_load": (i32, [vp, i32, i32, i32, i32, vp, vp, vp, vp, i32, vp, vp]),
    "okb_bow_transform": (i32, [vp, i32, vp, i32, vp, vp, vp]),
    "okb_overlap_counts": (i32, [vp, i32, vp, i32, vp, vp, f64, vp, vp]),
    "okb_prepared_device": (i32, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(i32), C.POINTER(i32)]),
}


class OlderView(C.Structure):
    """okb_older_view_t: one older keyframe view of the M3 sequence (device pointers)"""
    _fields_ = [("d_desc", C.c_void_p), ("d_rays", C.c_void_p), ("d_valid", C.c_void_p), ("d_size", C.c_void_p), ("d_use", C.c_void_p),
                ("n", C.c_int32), ("reserved", C.c_int32), ("T_WC", C.c_double * 12), ("T_CW", C.c_double * 12)]


class MultiframeCam(C.Structure):
    """okb_multiframe_cam_t"""
    _fields_ = [("image", C.c_void_p), ("stride_bytes", C.c_size_t), ("n_cand", C.c_int32), ("n_lm", C.c_int32), ("pool_changed", C.c_int32),
                ("reserved", C.c_int32), ("cand_desc", C.c_void_p), ("cand_lm", C.c_void_p), ("lm_is3d", C.c_void_p), ("lm_proj", C.c_void_p),
                ("T_WC1", C.c_void_p), ("T_CW1", C.c_void_p), ("n_older", C.c_int32), ("cap0", C.c_int32), ("older", C.c_void_p),
                ("cap", C.c_int32), ("n", C.c_int32), ("kp", C.c_void_p), ("desc", C.c_void_p), ("rays", C.c_void_p), ("rays_valid", C.c_void_p),
                ("m1_dist", C.c_void_p), ("m1_lm", C.c_void_p), ("cap_m", C.c_int32), ("reserved2", C.c_int32), ("m3_n", C.c_void_p),
                ("m3_k0", C.c_void_p), ("m3_k1", C.c_void_p), ("m3_flags", C.c_void_p), ("m3_hp_W", C.c_void_p)]


class MultiframeStereo(C.Structure):
    """okb_multiframe_stereo_t"""
    _fields_ = [("cam0", C.c_int32), ("cam1", C.c_int32), ("C_WC0", C.c_double * 9), ("r_WC0", C.c_double * 3), ("C_WC1", C.c_double * 9),
                ("r_WC1", C.c_double * 3), ("k1", C.c_void_p), ("dist", C.c_void_p), ("hp_W", C.c_void_p), ("initialisable", C.c_void_p)]


class OverlapView(C.Structure):
    """okb_overlap_view_t"""
    _fields_ = [("image_rows", C.c_int32), ("image_cols", C.c_int32), ("first_keypoint", C.c_int32), ("n_keypoints", C.c_int32)]


class PrepareView(C.Structure):
    """okb_prepare_view_t"""
    _fields_ = [("T_WC1", C.c_double * 12), ("T_CW1", C.c_double * 12), ("model", CameraModel), ("width", C.c_int32),
                ("height", C.c_int32), ("repr_threshold", C.c_double), ("exclusive", C.c_int32), ("reserved", C.c_int32)]


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise OkbError(OKB_ERR_UNSUPPORTED, f"{SO_PATH} is missing: build it with __graft_entry__.build() "
                           "(nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(status):
    if status != 0:
        raise OkbError(status, lib().okb_last_error().decode(errors="replace"))


def ptr(a):
    return None if a is None else a.ctypes.data
