"""CPU suite: the oracle's M3 sequence (Frontend::matchMotionStereo over the older keyframes, Frontend.cpp:1775-1958) against
the single-view M3 transcription it is built from, plus the properties of the serial insertion."""
import numpy as np

import oracle
from okvis2_b200.synth import motion_scene


def world_rays(T_WC, rays):
    Cm = np.asarray(T_WC[:9]).reshape(3, 3)
    x, y, z = rays[:, 0], rays[:, 1], rays[:, 2]
    w = [(Cm[i, 0] * x + Cm[i, 1] * y) + Cm[i, 2] * z for i in range(3)]
    n = np.sqrt((w[0] * w[0] + w[1] * w[1]) + w[2] * w[2])
    return np.ascontiguousarray(np.stack([w[0] / n, w[1] / n, w[2] / n], 1))


def kp_of(xy, size):
    kp = np.zeros(len(xy), oracle.KP_DTYPE)
    kp["x"], kp["y"], kp["size"] = xy[:, 0], xy[:, 1], size
    return kp


def oracle_views(s):
    intr = s["intr"]
    out = []
    for v in s["views"]:
        rays, valid = oracle.back_project(1, intr[0], intr[1], intr[2], intr[3], list(intr[4:8]), kp_of(v["xy"], v["size"]))
        out.append(dict(desc=v["desc"], rays=rays, valid=valid, size=v["size"], use=v["use"], T_WC=v["T_WC"], T_CW=v["T_CW"]))
    return out


def test_sequence_first_view_equals_single_m3_and_insertion_rules():
    s = motion_scene(5, n_views=4, n0=300, n1=450)
    intr = s["intr"]; cur = s["cur"]
    rays1, valid1 = oracle.back_project(1, intr[0], intr[1], intr[2], intr[3], list(intr[4:8]), kp_of(cur["xy"], cur["size"]))
    views = oracle_views(s)
    res, m1 = oracle.match_motion_stereo_sequence(views, cur["desc"], rays1, valid1, cur["xy"], s["T_WC1"], s["T_CW1"], 1, intr,
                                                  s["W"], s["H"], 60, cur["matched"], n_threads=3)
    # view 0 = the single-view transcription on the compacted unmatched set
    v = views[0]
    k1s = np.nonzero(cur["matched"] == 0)[0]
    f0 = 0.5 * (intr[0] + intr[1])
    T3x4 = lambda T: np.concatenate([np.asarray(T[:9]).reshape(3, 3), np.asarray(T[9:]).reshape(3, 1)], 1).reshape(12)
    ref = oracle.match_motion_stereo(v["desc"], (v["use"] & v["valid"]).astype(np.uint8), world_rays(v["T_WC"], v["rays"]),
                                     v["size"].astype(np.float64) / f0, cur["desc"][k1s], valid1[k1s], world_rays(s["T_WC1"], rays1)[k1s],
                                     v["T_WC"][9:], s["T_WC1"][9:], T3x4(v["T_CW"]), T3x4(s["T_CW1"]), 60)
    k1, dist, hp, fl = res[0]
    hit = ref[0] >= 0
    assert hit.sum() > 30
    assert np.array_equal(dist, ref[1]) and np.array_equal(np.where(hit, k1s[np.clip(ref[0], 0, None)], -1), k1)
    assert np.array_equal(hp.view(np.uint64), ref[2].view(np.uint64)) and np.array_equal((fl & 2) >> 1, ref[3] & hit)
    # serial insertion: every inserted match owns its k1 exclusively, in ascending k0 order; the mask only grows
    claimed = cur["matched"].copy()
    n_ins = 0
    for (k1, dist, hp, fl) in res:
        for k0 in range(len(k1)):
            if fl[k0] & 1:
                if claimed[k1[k0]]:
                    assert not (fl[k0] & 4)
                else:
                    assert fl[k0] & 4
                    claimed[k1[k0]] = 1; n_ins += 1
            else:
                assert not (fl[k0] & 4)
    assert np.array_equal(claimed, m1) and n_ins > 50
    # later views never match a keypoint that an earlier view inserted
    for i in range(1, len(res)):
        before = cur["matched"].copy()
        for j in range(i):
            before[res[j][0][(res[j][3] & 4) != 0]] = 1
        k1 = res[i][0]
        assert not before[k1[k1 >= 0]].any()


def test_sequence_is_thread_count_invariant_and_handles_empty_views():
    s = motion_scene(8, n_views=3, n0=200, n1=260)
    intr = s["intr"]; cur = s["cur"]
    rays1, valid1 = oracle.back_project(1, intr[0], intr[1], intr[2], intr[3], list(intr[4:8]), kp_of(cur["xy"], cur["size"]))
    views = oracle_views(s)
    views[1] = dict(desc=np.zeros((0, 64), np.uint8), rays=np.zeros((0, 3)), valid=np.zeros(0, np.uint8), size=np.zeros(0, np.float32),
                    use=None, T_WC=views[1]["T_WC"], T_CW=views[1]["T_CW"])
    a, ma = oracle.match_motion_stereo_sequence(views, cur["desc"], rays1, valid1, cur["xy"], s["T_WC1"], s["T_CW1"], 1, intr, s["W"],
                                                s["H"], 60, cur["matched"], n_threads=1)
    b, mb = oracle.match_motion_stereo_sequence(views, cur["desc"], rays1, valid1, cur["xy"], s["T_WC1"], s["T_CW1"], 1, intr, s["W"],
                                                s["H"], 60, cur["matched"], n_threads=4)
    assert np.array_equal(ma, mb) and all(np.array_equal(x, y) for ra, rb in zip(a, b) for x, y in zip(ra, rb))
    assert len(a[1][0]) == 0
