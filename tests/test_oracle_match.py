"""CPU suite: known answers for the matcher oracle (oracle/match_oracle.cpp)."""
import numpy as np

import oracle
from okvis2_b200.synth import map_scene, stereo_scene


def test_hamming_known_answer_real_descriptors(voc_desc):
    import os
    from conftest import ROOT
    rowsum = np.load(os.path.join(ROOT, "tests", "golden", "voc_hamming_rowsum.npy"))
    h = oracle.hamming_matrix(voc_desc, voc_desc)
    assert (np.diag(h) == 0).all() and (h == h.T).all()
    assert np.array_equal(h.sum(1).astype(np.int64), rowsum)


def test_m1_equals_numpy_bruteforce():
    rng = np.random.default_rng(3)
    kp_xy = rng.uniform(0, 752, (300, 2)); kd = rng.integers(0, 256, (300, 48), dtype=np.uint8)
    m = map_scene(9, kp_xy, kd, 800, W=752, H=480)
    dist, lm = oracle.match_map3d(kd, kp_xy, None, m["cand_desc"], m["cand_lm"], m["lm_proj"], m["lm_is3d"], 20.0, 60)
    ham = np.unpackbits(kd[:, None, :] ^ m["cand_desc"][None], axis=2).sum(2)
    d2 = ((m["lm_proj"][m["cand_lm"]][None] - kp_xy[:, None]) ** 2).sum(2)
    ok = (d2 <= 400.0) & (m["lm_is3d"][m["cand_lm"]][None] > 0)
    ham = np.where(ok, ham, 10000)
    best = ham.argmin(1)  # first minimum = ascending landmark, ascending descriptor
    exp_lm = np.where(ham.min(1) < 60, m["cand_lm"][best], -1)
    assert np.array_equal(lm, exp_lm)
    assert np.array_equal(dist, np.minimum(ham.min(1), 60).astype(np.uint32))
    assert (lm >= 0).sum() > 5


def test_thread_split_is_invisible():
    s = stereo_scene(2, 400, 500, D=48)
    a = oracle.match_motion_stereo(s["desc0"], s["valid0"], s["e0_W"], s["sof0"], s["desc1"], s["valid1"], s["e1_W"],
                                   s["r_WC0"], s["r_WC1"], s["T_CW0"], s["T_CW1"], 60, n_threads=1)
    b = oracle.match_motion_stereo(s["desc0"], s["valid0"], s["e0_W"], s["sof0"], s["desc1"], s["valid1"], s["e1_W"],
                                   s["r_WC0"], s["r_WC1"], s["T_CW0"], s["T_CW1"], 60, n_threads=4)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert (a[0] >= 0).sum() > 20


def test_triangulate_fast_cases():
    # intersecting rays -> midpoint at the intersection, valid, not parallel
    hp, v, p = oracle.triangulate_fast([0, 0, 0], [0, 0, 1], [1, 0, 0], [-np.sqrt(0.5), 0, np.sqrt(0.5)], 0.002)
    assert v and not p and np.allclose(hp[:3], [0, 0, 1], atol=1e-12)
    # parallel rays -> far point, flagged parallel, valid
    hp, v, p = oracle.triangulate_fast([0, 0, 0], [0, 0, 1], [0.1, 0, 0], [0, 0, 1], 0.005)
    assert v and p and hp[2] > 7.9
    # same, but a tighter sigma rejects the 40-baseline point (angle 0.1/2/8 rad > 2.6 sigma)
    hp, v, p = oracle.triangulate_fast([0, 0, 0], [0, 0, 1], [0.1, 0, 0], [0, 0, 1], 0.002)
    assert p and not v
    # diverging rays -> parallel branch, rejected by the 2.6 sigma cone
    hp, v, p = oracle.triangulate_fast([0, 0, 0], [-0.2, 0, 0.98], [0.1, 0, 0], [0.2, 0, 0.98], 0.002)
    assert p and not v


def test_m4_true_matches_and_gates():
    s = stereo_scene(5, 600, 600)
    k1, dist, hp, init = oracle.match_stereo(s["desc0"], s["valid0"], s["e0_W"], s["sof0"], s["desc1"], s["valid1"], s["e1_W"],
                                             s["sof1"], s["r_WC0"], s["r_WC1"], s["T_CW0"], s["T_CW1"], 60)
    m = k1 >= 0
    assert m.sum() > 100
    assert (s["idx1"][k1[m]] == s["idx0"][m]).mean() > 0.99
    assert (dist[~m] == 60).all() and (hp[~m] == 0).all()
    assert (s["valid0"][m] == 1).all() and (s["valid1"][k1[m]] == 1).all()


def test_triangulate_fast_equals_least_squares_midpoint():
    """Independent check of the G1 restatement (stereo_triangulation.cpp:50-132): in the regular branch the result is the
    midpoint of the common perpendicular of the two rays, which numpy's least-squares solver gives directly."""
    rng = np.random.default_rng(9)
    n_regular = 0
    for _ in range(300):
        p1 = rng.normal(0, 0.2, 3); p2 = p1 + rng.normal(0, 0.3, 3)
        X = rng.normal(0, 1.0, 3) + np.array([0, 0, 6.0])
        e1 = X - p1 + rng.normal(0, 0.01, 3); e1 /= np.linalg.norm(e1)
        e2 = X - p2 + rng.normal(0, 0.01, 3); e2 /= np.linalg.norm(e2)
        hp, valid, parallel = oracle.triangulate_fast(p1, e1, p2, e2, 0.01)
        lam = np.linalg.lstsq(np.stack([e1, -e2], 1), p2 - p1, rcond=None)[0]
        if lam[0] < 0.01 or lam[1] < 0.01:
            assert parallel
            continue
        mid = 0.5 * ((p1 + lam[0] * e1) + (p2 + lam[1] * e2))
        assert np.allclose(hp[:3], mid, rtol=0, atol=1e-9) and hp[3] == 1.0
        n_regular += 1
    assert n_regular > 250
