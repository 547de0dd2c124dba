"""CPU suite: the Hamming core of the matcher oracle against OpenCV. With the geometric gate disabled (reprojection
threshold larger than the image, every landmark 3-D, one descriptor per landmark, match threshold above 8*D) the M1 loop
(Frontend.cpp:1515-1590) degenerates to "nearest train descriptor, first index on ties" -- what cv2.BFMatcher(NORM_HAMMING)
returns. Golden: tests/golden/bfmatcher_cv2_4_13.npz (tests/golden/make_golden_bfmatcher.py, OpenCV 4.13.0). The GPU suite
repeats it through the CUDA matchers."""
import os

import numpy as np
import pytest

import oracle
from conftest import ROOT


@pytest.fixture(scope="module")
def bf():
    return np.load(os.path.join(ROOT, "tests", "golden", "bfmatcher_cv2_4_13.npz"))


def ungated_m1(match, query, train):
    n_q, n_t = len(query), len(train)
    xy = np.zeros((n_q, 2)); proj = np.zeros((n_t, 2))
    return match(query, xy, None, train, np.arange(n_t, dtype=np.int32), proj, np.ones(n_t, np.uint8))


@pytest.mark.parametrize("D", [48, 64])
def test_oracle_hamming_and_ungated_m1_equal_bfmatcher(bf, D):
    query, train, idx, dist = bf[f"query{D}"], bf[f"train{D}"], bf[f"idx{D}"], bf[f"dist{D}"]
    h = oracle.hamming_matrix(query, train)
    assert np.array_equal(h.min(1), dist.astype(h.dtype))
    assert np.array_equal(h.argmin(1), idx)                                 # numpy and OpenCV both keep the first minimum
    d, lm = ungated_m1(lambda *a: oracle.match_map3d(*a, 1e5, 8 * D + 1), query, train)
    assert np.array_equal(d, dist.astype(np.uint32)) and np.array_equal(lm, idx)
    if D == 64:
        assert (np.sort(h, 1)[:, 0] == np.sort(h, 1)[:, 1]).sum() > 10       # the duplicate rows really produce ties


@pytest.mark.gpu
@pytest.mark.parametrize("D", [48, 64])
def test_cuda_hamming_and_ungated_m1_equal_bfmatcher(bf, D):
    from okvis2_b200.frontend import Frontend
    query, train, idx, dist = bf[f"query{D}"], bf[f"train{D}"], bf[f"idx{D}"], bf[f"dist{D}"]
    fe = Frontend(0)
    try:
        fe.setBriskMatchingThreshold(8 * D + 1)
        h = fe.hammingMatrix(query, train)
        assert np.array_equal(h.min(1), dist.astype(h.dtype)) and np.array_equal(h.argmin(1), idx)
        n_t = len(train)
        d, lm = fe.matchToMapByThread(query, np.zeros((len(query), 2)), None, train, np.arange(n_t, dtype=np.int32), np.zeros((n_t, 2)),
                                      np.ones(n_t, np.uint8), False)
        # use_imu False = 150 px gate: all projections and keypoints sit at the origin, so the gate passes for every pair
        assert np.array_equal(d, dist.astype(np.uint32)) and np.array_equal(lm, idx)
    finally:
        fe.close()
