"""GPU suite: K1 (keyframe-overlap masks) through the C ABI against the oracle (whose rasteriser is pinned to cv2), and
the two reference decisions built on it (Frontend::doWeNeedANewKeyframe, ViSlamBackend::overlapFraction)."""
import numpy as np
import pytest

import oracle
from okvis2_b200.frontend import Frontend, MultiFrame
from okvis2_b200.lib import KP_DTYPE

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fe():
    f = Frontend(0)
    yield f
    f.close()


def random_views(rng, shapes, n_max=1200):
    views = []
    for (r, c) in shapes:
        n = int(rng.integers(0, n_max))
        xy = np.stack([rng.uniform(-15, c + 15, n), rng.uniform(-15, r + 15, n)], 1).astype(np.float32)
        xy[: n // 10] = np.round(xy[: n // 10] / 5) * 5          # x.5 products: exercises the half-to-even rounding
        views.append((r, c, xy, rng.random(n) < 0.3))
    return views


@pytest.mark.parametrize("shapes", [[(480, 752)] * 7, [(1024, 1024)] * 4 + [(480, 752)] * 3, [(540, 720)] * 10, [(10, 10), (19, 25)]])
def test_overlap_counts_equal_oracle(fe, shapes):
    rng = np.random.default_rng(len(shapes))
    views = random_views(rng, shapes)
    inter, uni = fe._overlap_counts(views)
    for i, (r, c, xy, m) in enumerate(views):
        assert (inter[i], uni[i]) == oracle.overlap_counts(r, c, xy, m), i


def test_empty_views(fe):
    inter, uni = fe._overlap_counts([(480, 752, np.zeros((0, 2), np.float32), np.zeros(0, bool))])
    assert (inter[0], uni[0]) == (0, 0)
    inter, uni = fe._overlap_counts([])
    assert len(inter) == 0


def make_mf(rng, n_cams, shape, n, ids_pool, p_match):
    mf = MultiFrame(n_cams)
    for c in range(n_cams):
        mf.setImage(c, np.zeros(shape, np.uint8))
        kp = np.zeros(n, KP_DTYPE); kp["x"] = rng.uniform(0, shape[1], n); kp["y"] = rng.uniform(0, shape[0], n)
        mf.frames[c].resetKeypoints(kp)
        lm = np.where(rng.random(n) < p_match, rng.choice(ids_pool, n), 0).astype(np.uint64)
        mf.frames[c].landmarkIds = lm
    return mf


def ref_ratio(mf, matched_of, kptrad=0.09):
    i = u = 0
    for fr in mf.frames:
        xy = np.stack([fr.keypoints["x"], fr.keypoints["y"]], 1)
        a, b = oracle.overlap_counts(fr.image.shape[0], fr.image.shape[1], xy, matched_of(fr), kptrad)
        i += a; u += b
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.float64(i) / np.float64(u)


def test_keyframe_decision_and_overlap_fraction(fe):
    rng = np.random.default_rng(5)
    ids = np.arange(1, 400, dtype=np.uint64)
    for trial in range(6):
        cur = make_mf(rng, 2, (480, 752), 500, ids, [0.1, 0.5, 0.9, 0.3, 0.02, 0.7][trial])
        others = [make_mf(rng, 2, (480, 752), int(rng.integers(50, 600)), ids, float(rng.uniform(0.2, 0.9))) for _ in range(5)]
        # reference decision from oracle counts
        lm = set(int(x) for fr in cur.frames for x in fr.landmarkIds if x != 0)
        arr = np.array(sorted(lm), np.uint64)
        overlap = ref_ratio(cur, lambda fr: fr.landmarkIds != 0)
        oo = np.float64(0.0)
        for mf in others:
            x = ref_ratio(mf, lambda fr: (fr.landmarkIds != 0) & np.isin(fr.landmarkIds, arr))
            oo = x if oo < x else oo
        overlap = overlap if overlap < oo else oo
        want = not (np.float32(overlap) > np.float32(0.55))
        assert fe.doWeNeedANewKeyframe(10, cur, others) == want
        # overlapFraction
        common = np.array(sorted(lm & set(int(x) for fr in others[0].frames for x in fr.landmarkIds if x != 0)), np.uint64)
        o = [ref_ratio(f, lambda fr: np.isin(fr.landmarkIds, common)) for f in (cur, others[0])]
        assert fe.overlapFraction(cur, others[0]) == float(min(o[0], o[1]))
    assert fe.doWeNeedANewKeyframe(3, cur, others) is True
    assert fe.doWeNeedANewKeyframe(10, cur, others, isInitialized=False) is False
    few = make_mf(rng, 2, (480, 752), 6, ids, 0.5)
    assert fe.doWeNeedANewKeyframe(10, few, others) is False     # fewer than 7 keypoints per camera
