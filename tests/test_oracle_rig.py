"""CPU suite: the oracle's transcriptions of PinholeCamera::initialiseCameraAwarenessMaps (PinholeCamera.hpp:179-208) and
NCameraSystem::computeOverlaps (NCameraSystem.cpp:48-118) against the reference's own known answers: the booleans asserted in
okvis_cv/test/TestNCameraSystem.cpp:60-110 (test cameras of PinholeCamera::createTestObject and the three distortion
testObjects) and analytic properties of the maps."""
import numpy as np

import oracle

# PinholeCamera(752, 480, 350, 360, 378, 238, distortion_t::testObject())  (PinholeCamera.hpp:389-398); distortion test objects:
# NoDistortion, RadialTangentialDistortion(-0.16, 0.15, 0.0003, 0.0002), EquidistantDistortion(-0.21, 0.14, 0.0006, 0.0003)
TEST_INTR = [[350, 360, 378, 238, 0, 0, 0, 0], [350, 360, 378, 238, -0.16, 0.15, 0.0003, 0.0002], [350, 360, 378, 238, -0.21, 0.14, 0.0006, 0.0003]]
TEST_MODELS = [0, 1, 2]


def quat_C(w, x, y, z):
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], np.float64)


def rig_C_rel():
    # T_SC of TestNCameraSystem.cpp:74-86: identity, identity, quaternion (w, x, y, z) = (0, 0, 1, 0): the third looks backwards
    Cs = [quat_C(1, 0, 0, 0), quat_C(1, 0, 0, 0), quat_C(0, 0, 1, 0)]
    return np.array([[Cs[s].T @ Cs[c] for c in range(3)] for s in range(3)])


def test_compute_overlaps_known_answers_of_the_reference_test():
    ov, mats = oracle.compute_overlaps(TEST_MODELS, TEST_INTR, [752] * 3, [480] * 3, rig_C_rel(), masks=True)
    assert ov[0, 0] and ov[1, 1] and ov[2, 2]                 # self overlaps (TestNCameraSystem.cpp:96-98)
    assert ov[0, 1] and ov[1, 0]                              # 0 and 1 overlap (:101-102)
    assert not ov[1, 2] and not ov[2, 1]                      # :105-106
    assert not ov[0, 2] and not ov[2, 0]                      # :109-110
    assert mats[0][0].all() and not mats[2][0].any()
    # identical orientation: nearly every pixel of camera 1 lands in camera 0 (same intrinsics, mild distortion)
    assert 0.9 < mats[0][1].mean() <= 1.0 and mats[0][1][240, 378] == 1


def test_awareness_maps_properties():
    for model, intr in zip(TEST_MODELS, TEST_INTR):
        rays, jac = oracle.camera_awareness_maps(model, intr, 752, 480)
        n = np.linalg.norm(rays.astype(np.float64), axis=2)
        ok = n > 0
        assert ok.mean() > 0.99 and np.allclose(n[ok], 1.0, atol=1e-6)
        # the principal point looks along +z and its Jacobian is diag(fu, fv) (first order in the distortion)
        r, j = rays[238, 378], jac[238, 378]
        assert abs(r[0]) < 1e-6 and abs(r[1]) < 1e-6 and abs(r[2] - 1) < 1e-6
        assert abs(j[0] - 350) < 0.5 and abs(j[4] - 360) < 0.5 and abs(j[1]) < 0.5 and abs(j[3]) < 0.5
        # finite differences: J * d(ray) reproduces the pixel step between horizontal neighbours
        v, u = 200, 300
        d = (rays[v, u + 1].astype(np.float64) - rays[v, u].astype(np.float64))
        step = jac[v, u].astype(np.float64).reshape(2, 3) @ d
        assert abs(step[0] - 1.0) < 0.02 and abs(step[1]) < 0.02
