"""CPU suite: the C-ABI library loads and exports every symbol include/okvis_b200.h declares; without a CUDA device
the product path fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, _has_gpu
from okvis2_b200 import lib as L


def declared_symbols():
    h = open(os.path.join(ROOT, "include", "okvis_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(okb_[a-z0-9_]+)\s*\(", h)))


def test_library_exports_every_declared_symbol():
    L.build()
    dll = C.CDLL(L.SO_PATH)
    syms = declared_symbols()
    assert len(syms) >= 25
    missing = [s for s in syms if not hasattr(dll, s)]
    assert not missing, missing
    # and the Python binding declares a prototype for each of them
    assert sorted(L._PROTOS) == syms


def test_keypoint_record_is_cv_keypoint_layout():
    assert L.KP_DTYPE.itemsize == 28
    assert L.KP_DTYPE.names == ("x", "y", "size", "angle", "response", "octave", "class_id")
    assert C.sizeof(L.CameraConfig) == 36


@pytest.mark.skipif(_has_gpu(), reason="CPU-only behaviour")
def test_no_device_is_a_loud_error():
    from okvis2_b200.frontend import Frontend
    with pytest.raises(L.OkbError) as e:
        Frontend(1, 752, 480)
    assert e.value.status == L.OKB_ERR_NO_DEVICE
    assert "no CPU path" in str(e.value)


def test_product_package_does_not_import_the_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "okvis2_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                txt = open(os.path.join(root, f)).read()
                assert "import oracle" not in txt and "liboracle" not in txt and "okvo_" not in txt, f
