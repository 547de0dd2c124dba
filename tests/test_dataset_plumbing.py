"""BASELINE config 1 substitute: a synthetic stereo sequence written in the EuRoC directory layout that
okvis::DatasetReader consumes (DatasetReader.cpp:97,151-183), read back, and pushed through the front-end -- the CPU oracle
in the CPU suite, the CUDA path (which must equal it) in the GPU suite."""
import os

import numpy as np
import pytest

import oracle
from okvis2_b200 import dataset
from okvis2_b200.synth import synth_stereo


@pytest.fixture(scope="module")
def sequence(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("euroc_synth"))
    n = 4
    frames = [synth_stereo(300 + t, 752, 480) for t in range(n)]
    cams = [np.stack([f[c] for f in frames]) for c in range(2)]
    ts = [1403636579763555584 + 50_000_000 * t for t in range(n)]           # 20 Hz, EuRoC-style nanosecond stamps
    imu = [[ts[0] + 5_000_000 * k, 0.01 * k, -0.02, 0.003, 9.7, 0.1 * k, -0.3] for k in range(10 * n)]
    dataset.write_euroc(root, cams, ts, imu)
    return root, cams, ts, imu


def test_layout_and_round_trip(sequence):
    root, cams, ts, imu = sequence
    for c in range(2):
        lines = open(os.path.join(root, f"cam{c}", "data.csv")).read().split("\n")
        assert lines[0] == "#timestamp [ns],filename" and lines[1] == f"{ts[0]},{ts[0]}.png"
        assert os.path.isfile(os.path.join(root, f"cam{c}", "data", f"{ts[0]}.png"))
    r = dataset.EurocReader(root, 2)
    assert len(r) == len(ts)
    for k, (stamps, imgs) in enumerate(r):
        assert stamps == [ts[k], ts[k]]
        for c in range(2):
            assert np.array_equal(imgs[c], cams[c][k])
    assert np.allclose(r.imu(), np.array(imu, np.float64))
    with pytest.raises(FileNotFoundError):
        dataset.EurocReader(root, 3)


def test_csv_quirks_of_the_reference_reader(tmp_path):
    d = tmp_path / "cam0"; (d / "data").mkdir(parents=True)
    (d / "data.csv").write_bytes(b"#timestamp [ns],filename\n10, a.png\r\n20,b.png\r\n30,c.png\n")
    names = dataset.read_camera_image_csv(str(tmp_path), "cam", 0)
    assert [n[0] for n in names] == ["10", "20", "30"]
    assert [os.path.basename(n[1]) for n in names] == ["a.png", "b.png", "c.png"]     # leading blank and '\r' stripped
    assert dataset.read_camera_image_csv(str(tmp_path), "cam", 1) is None


def test_png_decoder_handles_all_filter_types(tmp_path):
    cv2 = pytest.importorskip("cv2")              # cv2's encoder picks filters adaptively: exercises types 1-4
    img = synth_stereo(5, 320, 200)[0]
    p = str(tmp_path / "x.png")
    cv2.imwrite(p, img)
    assert np.array_equal(dataset.read_png_gray8(p), img)
    dataset.write_png_gray8(p, img)
    assert np.array_equal(cv2.imread(p, cv2.IMREAD_GRAYSCALE), img)


def test_sequence_through_the_cpu_oracle(sequence):
    root, cams, _, _ = sequence
    o = oracle.Brisk(30, 3)
    total = 0
    for k, (_, imgs) in enumerate(dataset.EurocReader(root, 2)):
        for c in range(2):
            kp, d = o.detect_and_compute(imgs[c], 1000)
            rk, rd = o.detect_and_compute(cams[c][k], 1000)
            assert kp.tobytes() == rk.tobytes() and np.array_equal(d, rd)
            total += len(kp)
    assert total > 4000


@pytest.mark.gpu
def test_sequence_through_the_cuda_frontend(sequence):
    from okvis2_b200.frontend import Frontend, MultiFrame
    root, _, _, _ = sequence
    fe = Frontend(2, 752, 480)
    fe.configure(threshold=30, octaves=3, max_keypoints=1000)
    o = oracle.Brisk(30, 3)
    try:
        for stamps, imgs in dataset.EurocReader(root, 2):
            mf = MultiFrame(2, timestamp=stamps[0] * 1e-9)
            for c in range(2):
                mf.setImage(c, imgs[c]); fe.detectAndDescribe(c, mf, None, None)
                rk, rd = o.detect_and_compute(imgs[c], 1000)
                assert mf.frames[c].keypoints.tobytes() == rk.tobytes() and np.array_equal(mf.frames[c].descriptors, rd)
    finally:
        fe.close()
