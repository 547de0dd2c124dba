"""CPU suite: the camera-model restatements of the oracle (PinholeCamera::project with radial-tangential / equidistant
distortion, okvis_cv/include/okvis/cameras/implementation/PinholeCamera.hpp:257-292, RadialTangentialDistortion.hpp:90-109,
EquidistantDistortion.hpp:86-106; backProject = Gauss-Newton undistortion, :574-592) cross-checked against OpenCV's
independent implementations of the same models (cv2.projectPoints, cv2.fisheye.projectPoints, cv2.undistortPoints).
Different operation order, so the comparison is to a stated tolerance, not bit-exact: 1e-9 px for projections,
1e-7 (normalised image units) for back-projections."""
import numpy as np
import pytest

import oracle
from okvis2_b200.lib import KP_DTYPE
from okvis2_b200.synth import landmark_scene

cv2 = pytest.importorskip("cv2")


def kept_projections(model, intr, s):
    r = oracle.prepare_landmarks(s["hp_W"], s["quality"], s["obs_begin"], s["obs"], s["n_cams"], s["T_WC_old"], s["desc_tab"],
                                 s["ray_tab"], s["D"], s["T_WC1"], s["T_CW1"], model, intr, s["W"], s["H"])
    C_CW = s["T_CW1"][:9].reshape(3, 3); r_CW = s["T_CW1"][9:]
    p_C = r["p_W"] @ C_CW.T + r_CW
    return r, p_C


def test_radtan_projection_equals_cv2_projectpoints():
    s = landmark_scene(21, n_lm=1500)
    intr = s["intr"]
    r, p_C = kept_projections(1, intr, s)
    K = np.array([[intr[0], 0, intr[2]], [0, intr[1], intr[3]], [0, 0, 1.0]])
    front = p_C[:, 2] > 0.1
    ref, _ = cv2.projectPoints(p_C[front].reshape(-1, 1, 3), np.zeros(3), np.zeros(3), K, intr[4:8])
    assert front.sum() > 300
    assert np.abs(ref.reshape(-1, 2) - r["lm_proj"][front]).max() < 1e-9


def test_equidistant_projection_equals_cv2_fisheye():
    s = landmark_scene(22, n_lm=1500)
    intr = s["intr"].copy(); intr[4:] = [-0.0369, -0.0089, 0.0089, -0.0038]
    r, p_C = kept_projections(2, intr, s)
    K = np.array([[intr[0], 0, intr[2]], [0, intr[1], intr[3]], [0, 0, 1.0]])
    front = p_C[:, 2] > 0.1
    ref, _ = cv2.fisheye.projectPoints(p_C[front].reshape(-1, 1, 3), np.zeros(3), np.zeros(3), K, intr[4:8])
    assert front.sum() > 300
    assert np.abs(ref.reshape(-1, 2) - r["lm_proj"][front]).max() < 1e-9


def test_radtan_back_projection_equals_cv2_undistortpoints():
    rng = np.random.default_rng(4)
    fu, fv, cu, cv_ = 458.654880721, 457.296696463, 367.215803962, 248.37534061
    k = [-0.28340811217, 0.0739590738929, 0.000193595028569, 1.76187114545e-05]
    kp = np.zeros(2000, KP_DTYPE)
    kp["x"] = rng.uniform(0, 752, 2000).astype(np.float32); kp["y"] = rng.uniform(0, 480, 2000).astype(np.float32)
    rays, valid = oracle.back_project(1, fu, fv, cu, cv_, k, kp)
    K = np.array([[fu, 0, cu], [0, fv, cv_], [0, 0, 1.0]])
    pts = np.stack([kp["x"], kp["y"]], 1).astype(np.float64).reshape(-1, 1, 2)
    ref = cv2.undistortPointsIter(pts, K, np.array(k), None, None, (cv2.TERM_CRITERIA_COUNT | cv2.TERM_CRITERIA_EPS, 100, 1e-14))
    ok = valid.astype(bool)
    assert ok.sum() > 1900
    assert np.abs(ref.reshape(-1, 2)[ok] - rays[ok, :2]).max() < 1e-7 and np.all(rays[:, 2] == 1.0)
