"""The cos() behind the matchers' gate constants (cos(2.6 sigma), cos(6 sigma) of triangulateFast, reference
okvis_frontend/src/stereo_triangulation.cpp:82-127): okb::gate_cos (csrc/okb_gatecos.h) is the one function the host-buffer
and the device-resident matcher forms both call; it restates the libm algorithm and must return this machine's libm bits."""
import ctypes as C
import math
import struct

import numpy as np
import pytest

from conftest import build_emul


@pytest.fixture(scope="module")
def emul():
    lib = C.CDLL(build_emul())
    lib.okb_emul_gate_cos.restype = C.c_double; lib.okb_emul_gate_cos.argtypes = [C.c_double]
    lib.okb_emul_gate_cos_mismatches.restype = C.c_long; lib.okb_emul_gate_cos_mismatches.argtypes = [C.c_long, C.c_ulonglong, C.c_double]
    lib.okb_emul_gate_cos_sizes.restype = C.c_long; lib.okb_emul_gate_cos_sizes.argtypes = [C.c_double, C.c_uint, C.c_uint, C.c_uint]
    return lib


def test_gate_cos_equals_libm_on_random_arguments(emul):
    assert emul.okb_emul_gate_cos_mismatches(12_000_000, 1, 0.5) == 0      # 6 sigma of any real camera
    assert emul.okb_emul_gate_cos_mismatches(8_000_000, 2, 0.9) == 0       # up to and beyond the table range (0.85546875)


def test_gate_cos_equals_libm_on_keypoint_sizes(emul):
    bits = lambda f: struct.unpack("<I", struct.pack("<f", f))[0]
    # focal lengths of config/euroc.yaml, tumvi_slam_1024.yaml, hilti_challenge_2022.yaml; float sizes 6 .. 130 px, strided
    for f in (458.654, 457.296, 190.97847715128717, 351.31400364193297):
        assert emul.okb_emul_gate_cos_sizes(f, bits(6.0), bits(130.0), 5) == 0


def test_gate_cos_special_values(emul):
    g = emul.okb_emul_gate_cos
    assert g(0.0) == 1.0 and g(1e-9) == 1.0 and g(-0.3) == g(0.3) == math.cos(0.3)
    for x in (0.85546875, np.nextafter(0.85546875, 0), 1.0, 3.0, 100.0):   # boundary of the table range and the libm fall-through
        assert g(float(x)) == math.cos(float(x))
    assert math.isnan(g(float("nan")))
