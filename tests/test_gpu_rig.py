"""GPU parity of D5 (okb_camera_awareness_maps) and okb_compute_overlaps against the oracle transcriptions
(PinholeCamera.hpp:179-208, NCameraSystem.cpp:48-118), incl. the reference test's rig and the Hilti-2022 5-camera rig."""
import ctypes as C

import numpy as np
import pytest

import oracle
from okvis2_b200 import lib as okl, rigs
from okvis2_b200.frontend import Frontend
from test_oracle_rig import TEST_INTR, TEST_MODELS, rig_C_rel

pytestmark = pytest.mark.gpu
NAMES = {0: "none", 1: "radialtangential", 2: "equidistant"}


def test_awareness_maps_equal_oracle():
    for model, intr in zip(TEST_MODELS, TEST_INTR):
        fe = Frontend(1, 752, 480)
        try:
            fe.setCameraModel(0, NAMES[model], intr[:2], intr[2:4], intr[4:])
            rays, jac = fe.cameraAwarenessMaps(0)
            r_ref, j_ref = oracle.camera_awareness_maps(model, intr, 752, 480)
            if model == 2:   # device atan: last-bit differences before the rounding to float
                assert np.allclose(rays, r_ref, atol=1e-6, rtol=0)
                # row 0 / column 0 re-project onto the image border itself (v = 0 -> ky = +-1e-14): whether that counts as
                # inside is a last-ulp matter in the reference too; the interior must agree
                assert np.allclose(jac[1:, 1:], j_ref[1:, 1:], atol=2e-3, rtol=1e-6)
                assert (rays.view(np.uint32) != r_ref.view(np.uint32)).mean() < 0.01
            else:
                assert np.array_equal(rays.view(np.uint32), r_ref.view(np.uint32))
                assert np.array_equal(jac.view(np.uint32), j_ref.view(np.uint32))
        finally:
            fe.close()


def test_compute_overlaps_equals_oracle_on_the_reference_rig():
    fe = Frontend(0)
    try:
        ov, mats = fe.computeOverlaps(TEST_MODELS, TEST_INTR, [752] * 3, [480] * 3, rig_C_rel(), masks=True)
        rov, rmats = oracle.compute_overlaps(TEST_MODELS, TEST_INTR, [752] * 3, [480] * 3, rig_C_rel(), masks=True)
        assert np.array_equal(ov, rov)
        assert ov[0, 1] and ov[1, 0] and not ov[1, 2] and not ov[2, 1] and not ov[0, 2] and not ov[2, 0]   # TestNCameraSystem.cpp:96-110
        for s in range(3):
            for c in range(3):
                diff = (mats[s][c] != rmats[s][c]).mean()
                # radial-tangential / no distortion: bit-exact; pairs that involve the equidistant camera go through the device atan
                assert diff == 0.0 if 2 not in (s, c) else diff < 1e-3, (s, c, diff)
    finally:
        fe.close()


def test_compute_overlaps_hilti_rig():
    rig = rigs.HILTI_2022
    models = [Frontend.MODELS[r["distortion_type"]] for r in rig]
    intr = [list(r["focal_length"]) + list(r["principal_point"]) + list(r["distortion_coefficients"])[:4] for r in rig]
    W = [r["image_dimension"][0] for r in rig]; H = [r["image_dimension"][1] for r in rig]
    Cs = [np.array(r["T_SC"]).reshape(4, 4)[:3, :3] for r in rig]
    C_rel = np.array([[Cs[s].T @ Cs[c] for c in range(len(rig))] for s in range(len(rig))])
    fe = Frontend(0)
    try:
        ov = fe.computeOverlaps(models, intr, W, H, C_rel)
        rov = oracle.compute_overlaps(models, intr, W, H, C_rel)
        assert np.array_equal(ov, rov) and ov.trace() == len(rig) and np.array_equal(ov, ov.T)
    finally:
        fe.close()
