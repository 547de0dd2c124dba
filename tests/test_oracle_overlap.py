"""CPU suite: the K1 oracle (keyframe-overlap masks; Frontend.cpp:1058-1167, ViSlamBackend.cpp:2341-2426). Its circle
rasteriser is pinned against the committed cv2 4.13.0 masks (tests/golden/circle_cv2_4_13.npz, made by
tests/golden/make_golden_circle.py)."""
import os

import numpy as np
import pytest

import oracle
from conftest import ROOT


@pytest.fixture(scope="module")
def circles():
    return np.load(os.path.join(ROOT, "tests", "golden", "circle_cv2_4_13.npz"))


def test_circle_equals_cv2_golden(circles):
    cases = circles["cases"]
    assert len(cases) >= 500
    for i, (rows, cols, radius, cx, cy) in enumerate(cases):
        ref = np.unpackbits(circles[f"m{i}"])[:rows * cols].reshape(rows, cols) * 255
        got = oracle.circle_filled(np.zeros((rows, cols), np.uint8), cx, cy, radius)
        assert np.array_equal(got, ref), (rows, cols, radius, cx, cy)


def test_circle_equals_live_cv2_when_importable():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    for _ in range(300):
        rows, cols = int(rng.integers(5, 110)), int(rng.integers(5, 110))
        r = int(rng.integers(0, 12)); cx, cy = int(rng.integers(-15, cols + 15)), int(rng.integers(-15, rows + 15))
        ref = np.zeros((rows, cols), np.uint8); cv2.circle(ref, (cx, cy), r, 255, cv2.FILLED)
        assert np.array_equal(oracle.circle_filled(np.zeros((rows, cols), np.uint8), cx, cy, r), ref)


def test_overlap_counts_known_answers():
    # 480 x 752 image -> 48 x 75 mask, radius int(48 * 0.09) = 4: one isolated disc of the midpoint circle has 49 pixels
    i, u, det, mat = oracle.overlap_counts(480, 752, [[300.0, 200.0]], [0], masks=True)
    assert (i, u) == (0, 49) and det[20, 30] == 255 and mat.sum() == 0
    i, u = oracle.overlap_counts(480, 752, [[300.0, 200.0], [600.0, 100.0]], [1, 0])
    assert (i, u) == (49, 98)
    # centre rounding is half-to-even on the float product: 25.0 -> 2.5 -> 2, 35.0 -> 3.5 -> 4
    _, _, det, _ = oracle.overlap_counts(480, 752, [[25.0, 35.0]], [0], masks=True)
    ys, xs = np.nonzero(det)
    assert xs.max() == 2 + 4 and ys.max() == 4 + 4
    # empty frame: 0 / 0 (the reference divides them: NaN, handled by the caller exactly like the reference does)
    assert oracle.overlap_counts(480, 752, np.zeros((0, 2)), np.zeros(0)) == (0, 0)
