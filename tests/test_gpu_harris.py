"""GPU suite of the D = 48 mode: Harris + uniformity-enforcement detector and 48-byte BRISK2 extractor through the C ABI (k_harris_score,
k_harris_maxima, k_uniformity, k_describe48) against oracle/brisk_oracle.c section 6. Bit-exact for positions, responses and descriptor
rows; the reported angle (double atan2 on the device, rounded to float) within 1e-4 degrees. PARITY UNPINNED vs smartroboticslab/brisk."""
import ctypes as C

import numpy as np
import pytest

import oracle
from okvis2_b200 import lib as _l
from okvis2_b200.frontend import Frontend, MultiFrame
from okvis2_b200.lib import OkbError
from okvis2_b200.synth import map_scene, synth_frame

pytestmark = pytest.mark.gpu

EUROC = [dict(distortion_type="radialtangential", focal_length=(458.654880721, 457.296696463), principal_point=(367.215803962, 248.37534061),
              distortion_coefficients=[-0.28340811217, 0.0739590738929, 0.000193595028569, 1.76187114545e-05]),
         dict(distortion_type="equidistant", focal_length=(380.81, 380.81), principal_point=(510.29, 514.33),
              distortion_coefficients=[0.0103, -0.0046, 0.0024, -0.0008])]


def same48(kp, d, rk, rd, what=""):
    assert len(kp) == len(rk), f"{what}: {len(kp)} keypoints vs {len(rk)}"
    for f in ("x", "y", "size", "response", "octave", "class_id"):
        a, b = kp[f].view(np.uint32), rk[f].view(np.uint32)
        bad = np.nonzero(a != b)[0]
        assert len(bad) == 0, f"{what}: field {f} differs at {bad[:5]}: {kp[f][bad[:5]]} vs {rk[f][bad[:5]]}"
    da = np.abs(kp["angle"] - rk["angle"]); da = np.minimum(da, 360 - da)
    assert da.max(initial=0) <= 1e-4, f"{what}: angle differs by {da.max()}"
    assert d.shape == rd.shape == (len(rk), 48), what
    bad = np.nonzero((d != rd).any(1))[0]
    assert len(bad) == 0, f"{what}: {len(bad)} descriptor rows differ (first {bad[:5]})"


def make(W, H, radius, thr, max_kp, max_batch=1, n_cams=1):
    fe = Frontend(n_cams, W, H, max_batch=max_batch, descriptor_bytes=48)
    fe.configure(threshold=radius, absolute_threshold=thr, octaves=0, max_keypoints=max_kp)
    return fe


def run(fe, img, cam=0, T_WC=None):
    mf = MultiFrame(fe.numCameras)
    mf.setImage(cam, img)
    assert fe.detectAndDescribe(cam, mf, T_WC, None) is True
    fr = mf.frames[cam]
    assert fr.descriptors.flags["C_CONTIGUOUS"] and fr.descriptors.shape[1] == 48 and (fr.landmarkIds == 0).all()
    return fr


@pytest.mark.parametrize("seed,W,H,radius,thr,max_kp", [(31, 752, 480, 38.0, 150, 700), (32, 752, 480, 12.0, 20, 0), (33, 1024, 1024, 20.0, 100, 2000),
                                                        (34, 341, 255, 6.0, 5, 0), (35, 720, 540, 38.0, 150, 100)])
def test_plain_mode_equals_oracle(seed, W, H, radius, thr, max_kp):
    img = synth_frame(seed, W, H)
    fe = make(W, H, radius, thr, max_kp)
    fr = run(fe, img)
    rk, rd = oracle.HarrisBrisk2(radius, thr, max_kp).detect_and_compute(img)
    assert len(rk) > 20
    same48(fr.keypoints, fr.descriptors, rk, rd, f"seed {seed}")
    fe.close()


def test_real_image_equals_oracle(golden):
    img = golden["real752_img"]
    fe = make(752, 480, 38.0, 150, 700)
    fr = run(fe, img)
    rk, rd = oracle.HarrisBrisk2(38.0, 150, 700).detect_and_compute(img)
    same48(fr.keypoints, fr.descriptors, rk, rd, "real752")
    fe.close()


@pytest.mark.parametrize("cam_model,W,H", [(0, 752, 480), (1, 1024, 1024)])
def test_camera_aware_mode_equals_oracle(cam_model, W, H):
    """Frontend::detectAndDescribe with the camera-awareness maps and the gravity direction (Frontend.cpp:232-251)."""
    m = EUROC[cam_model]
    fe = make(W, H, 25.0, 80, 800)
    fe.setCameraModel(0, **m)
    rays, jac = fe.cameraAwarenessMaps(0)       # D5 on the device; the maps stay there for the extractor
    a = 0.3; Cx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    b = -0.2; Cz = np.array([[np.cos(b), -np.sin(b), 0], [np.sin(b), np.cos(b), 0], [0, 0, 1]])
    T_WC = np.eye(4); T_WC[:3, :3] = Cz @ Cx @ np.array([[1, 0, 0], [0, 0, 1], [0, -1, 0.]])   # camera looking roughly horizontally
    for seed in (41, 42):
        img = synth_frame(seed, W, H)
        fr = run(fe, img, 0, T_WC)
        d = fr.extractionDirection
        assert abs(np.linalg.norm(d) - 1) < 1e-5
        intr = [*m["focal_length"], *m["principal_point"], *m["distortion_coefficients"]]
        orays, ojac = oracle.camera_awareness_maps(Frontend.MODELS[m["distortion_type"]], intr, W, H)
        # the oracle is fed the DEVICE maps (D5's own parity, incl. the equidistant model's 1e-13 interior tolerance, is test_gpu_rig's)
        rk, rd = oracle.HarrisBrisk2(25.0, 80, 800).detect_and_compute(img, rays, jac, float(np.float32(m["focal_length"][0])), d)
        assert len(rk) > 50
        same48(fr.keypoints, fr.descriptors, rk, rd, f"aware seed {seed}")
        if cam_model == 0:   # radtan maps are bit-exact, so the oracle's own maps give the same features
            rk2, rd2 = oracle.HarrisBrisk2(25.0, 80, 800).detect_and_compute(img, orays, ojac, float(np.float32(m["focal_length"][0])), d)
            same48(fr.keypoints, fr.descriptors, rk2, rd2, f"aware/oracle maps seed {seed}")
        # the warp differs from the plain mode: the descriptors are not the plain ones
        pk, pd = oracle.HarrisBrisk2(25.0, 80, 800).detect_and_compute(img)
        assert len(pk) != len(rk) or not np.array_equal(pd, rd)
    fe.close()


def test_batch_and_m1_with_48_byte_rows():
    W, H, B = 752, 480, 6
    imgs = np.stack([synth_frame(50 + i, W, H) for i in range(B)])
    imgs[3] = 17                                   # a frame without a single corner
    fe = make(W, H, 30.0, 100, 500, max_batch=B)
    out = fe.detectAndDescribeBatch(0, imgs)
    o = oracle.HarrisBrisk2(30.0, 100, 500)
    for b in range(B):
        rk, rd = o.detect_and_compute(imgs[b])
        same48(out[b][0], out[b][1], rk, rd, f"batch frame {b}")
    assert len(out[3][0]) == 0
    # M1 (host-buffer form, D = 48) on the rows this mode produced
    L = _l.lib()
    kp, d = out[0]
    xy = np.stack([kp["x"], kp["y"]], 1).astype(np.float64)
    m = map_scene(5, xy, d, 1500, W=W, H=H, frac_near=0.3)
    assert m["cand_desc"].shape[1] == 48
    rdist, rlm = oracle.match_map3d(d, xy, None, m["cand_desc"], m["cand_lm"], m["lm_proj"], m["lm_is3d"], 20.0, 60)
    dist, lm = fe.matchToMapByThread(d, xy, None, m["cand_desc"], m["cand_lm"], m["lm_proj"], m["lm_is3d"])
    assert np.array_equal(dist, rdist) and np.array_equal(lm, rlm) and (lm >= 0).sum() > 20
    fe.close()


def test_rejections():
    with pytest.raises(OkbError) as e:
        fe = Frontend(1, 752, 480, descriptor_bytes=48)
        fe.configure(threshold=38.0, absolute_threshold=150, octaves=2, max_keypoints=100)
    assert e.value.status == _l.OKB_ERR_UNSUPPORTED
    with pytest.raises(OkbError) as e:
        fe = Frontend(1, 752, 480, descriptor_bytes=48)
        fe.configure(threshold=0.0, absolute_threshold=150, octaves=0, max_keypoints=100)
    assert e.value.status == _l.OKB_ERR_ARGUMENT
    with pytest.raises(OkbError) as e:
        Frontend(1, 752, 480, descriptor_bytes=32)
    assert e.value.status == _l.OKB_ERR_ARGUMENT
