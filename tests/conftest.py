import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def build_emul():
    """tests/emul/libokb_emul.so: the kernels' per-element code compiled for the host (test infrastructure)."""
    import subprocess
    so = os.path.join(ROOT, "tests", "emul", "libokb_emul.so")
    src = os.path.join(ROOT, "tests", "emul", "okb_emul.cpp")
    csrc = os.path.join(ROOT, "okvis2_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, src])
    return so


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "brisk_cv2_4_13.npz"))


@pytest.fixture(scope="session")
def voc_desc():
    return np.load(os.path.join(ROOT, "tests", "golden", "voc_descriptors.npy"))


def kp_struct(a):
    """golden (n,7) float32 -> structured keypoint array"""
    from okvis2_b200.lib import KP_DTYPE
    out = np.zeros(len(a), KP_DTYPE)
    for i, f in enumerate(KP_DTYPE.names):
        out[f] = a[:, i].astype(KP_DTYPE[f])
    return out


def assert_same_features(kp, desc, ref_kp, ref_desc, what=""):
    """bit-exact comparison of keypoint records and descriptor rows"""
    assert len(kp) == len(ref_kp), f"{what}: {len(kp)} keypoints vs {len(ref_kp)}"
    for f in ref_kp.dtype.names:
        a, b = kp[f], ref_kp[f]
        bad = np.nonzero(a.view(np.uint32 if a.dtype.itemsize == 4 else a.dtype) != b.view(np.uint32 if b.dtype.itemsize == 4 else b.dtype))[0]
        assert len(bad) == 0, f"{what}: field {f} differs at {bad[:5]}: {a[bad[:5]]} vs {b[bad[:5]]}"
    assert desc.shape == ref_desc.shape, what
    bad = np.nonzero((desc != ref_desc).any(1))[0]
    assert len(bad) == 0, f"{what}: {len(bad)} descriptor rows differ (first {bad[:5]})"
