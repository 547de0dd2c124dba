"""CPU suite: the P1 oracle (oracle/prepare_oracle.cpp = Frontend.cpp:1196-1360) on hand-built known-answer cases that
spell out what the reference loop leaves behind, plus structural invariants on a synthetic map."""
import numpy as np

import oracle
from okvis2_b200.synth import landmark_scene

I9 = np.eye(3).ravel()


def pose(r):
    return np.concatenate([I9, np.asarray(r, float)])


def run(hp, q, obs_per_lm, cam_pos, n_kp=8, model=0, exclusive=False, thr=20.0, W=752, H=480, f=458.0):
    """one camera per slot; slot s sits at cam_pos[s]; the current camera at the origin looking along +z"""
    n_slots = len(cam_pos)
    T_old = np.stack([pose(p) for p in cam_pos])[:, None, :]
    descs = [np.full((n_kp, 64), 16 * s, np.uint8) + np.arange(n_kp, dtype=np.uint8)[:, None] for s in range(n_slots)]
    rays = [np.tile(np.array([0.0, 0.0, 1.0]), (n_kp, 1)) for _ in range(n_slots)]
    ob = [0]; obs = []
    for o in obs_per_lm:
        obs.extend(sorted(o)); ob.append(len(obs))
    intr = np.array([f, f, W / 2, H / 2, 0, 0, 0, 0.0])
    return oracle.prepare_landmarks(np.asarray(hp, float), np.asarray(q, float), ob, np.array(obs, np.int32).reshape(-1, 3), 1, T_old,
                                    descs, rays, 64, pose([0, 0, 0]), pose([0, 0, 0]), model, intr, W, H, thr, exclusive)


def test_rows_are_second_and_third_accepted_observation_newest_first():
    # five co-located old cameras near the current one: every observation passes the viewpoint / scale gates
    cams = [[0.01 * s, 0, 0] for s in range(5)]
    hp = [[0, 0, 5, 1]] * 4
    obs = [[(0, 0, 1)],                                   # 1 accepted  -> o = 0 -> dropped
           [(0, 0, 1), (1, 0, 2)],                        # 2 accepted  -> 1 row  = second newest = slot 0
           [(0, 0, 1), (1, 0, 2), (2, 0, 3)],             # 3 accepted  -> 2 rows = slots 1, 0
           [(0, 0, 1), (1, 0, 2), (2, 0, 3), (4, 0, 5)]]  # 4 accepted  -> 2 rows = slots 2, 1 (newest is overwritten)
    r = run(hp, [0.5] * 4, obs, cams)
    assert list(r["lm"]) == [1, 2, 3]
    assert list(r["desc_begin"]) == [0, 1, 3, 5]
    assert [tuple(k) for k in r["kid"]] == [(0, 0, 1), (1, 0, 2), (0, 0, 1), (2, 0, 3), (1, 0, 2)]
    assert [int(d[0]) for d in r["cand_desc"]] == [1, 18, 1, 35, 18]      # 16 * slot + keypoint
    assert np.allclose(r["lm_proj"], [[376.0, 240.0]] * 3)
    assert np.array_equal(r["e_W"], np.tile([0.0, 0.0, 1.0], (5, 1)))
    assert np.array_equal(r["r_W"][:, 0], [0.0, 0.01, 0.0, 0.02, 0.01])


def test_gates():
    cams = [[0, 0, 0], [0.01, 0, 0], [3.6, 0, 0], [0, 0, -3.0]]
    two = [(0, 0, 0), (1, 0, 0)]
    hp = [[0, 0, 5, 1],        # fine
          [0, 0, -5, 1],       # inside the image but behind -> Behind -> skipped
          [-50, 0, -5, 1],     # behind AND projecting outside the image -> OutsideImage, not Behind; x = 4956 > maxU -> skipped
          [0.3, 0, -5, 1],     # behind, projects to u = 348.5: inside -> Behind
          [0, 0, 0, 1],        # singular
          [4.25, 0, 5, 1],     # u = 765.3 <= W + 20 -> kept although outside the image
          [4.4, 0, 5, 1],      # u = 779.0 > 772 -> dropped
          [0, 0, -5, -1]]      # w < 0: projectHomogeneous flips the head -> in front
    r = run(hp, [0.5] * 8, [two] * 8, cams)
    assert list(r["lm"]) == [0, 5, 7]
    # viewpoint gate: old camera 3.6 m to the side sees the point under 0.62 rad > 0.6; scale gate: old camera 3 m further
    # back (8 m instead of 5 m: 60 % > 50 %)
    r = run([[0, 0, 5, 1]] * 2, [0.5] * 2, [[(0, 0, 0), (1, 0, 0), (2, 0, 0)], [(0, 0, 0), (1, 0, 0), (3, 0, 0)]], cams)
    assert list(np.diff(r["desc_begin"])) == [1, 1]
    # exclusive (loop-closure) mode switches both gates off (the scores, 0.75 and 0.6, are still < 1): 3 accepted -> 2 rows
    r = run([[0, 0, 5, 1]] * 2, [0.5] * 2, [[(0, 0, 0), (1, 0, 0), (2, 0, 0)], [(0, 0, 0), (1, 0, 0), (3, 0, 0)]], cams, exclusive=True)
    assert list(np.diff(r["desc_begin"])) == [2, 2]


def test_is3d_uses_every_observation_until_set():
    cams = [[0, 0, 0], [0.01, 0, 0], [6.0, 0, 0]]
    # quality 1: r_close = r_W - 0.2/916 * r_W_old: tiny parallax -> cosA ~ 1 > cos(10/916) -> 3d
    r = run([[0, 0, 5, 1]], [1.0], [[(0, 0, 0), (1, 0, 0)]], cams)
    assert list(r["lm_is3d"]) == [1]
    # very low quality blows the offset up: with the sideways camera the angle exceeds 10/f rad for it, but the other
    # observations (parallel offset) still flip is3d
    r = run([[0, 0, 5, 1]], [1e-5], [[(2, 0, 0)], ], cams)
    assert len(r["lm"]) == 0


def test_synthetic_scene_invariants():
    s = landmark_scene(3)
    r = oracle.prepare_landmarks(s["hp_W"], s["quality"], s["obs_begin"], s["obs"], s["n_cams"], s["T_WC_old"], s["desc_tab"],
                                 s["ray_tab"], s["D"], s["T_WC1"], s["T_CW1"], 1, s["intr"], s["W"], s["H"])
    n = len(r["lm"])
    assert 300 < n < len(s["quality"])
    assert np.all(np.diff(r["lm"]) > 0)
    rows = np.diff(r["desc_begin"])
    assert rows.min() >= 1 and rows.max() <= 2 and r["desc_begin"][-1] == len(r["cand_desc"])
    assert np.all(r["lm_proj"][:, 0] >= -20) and np.all(r["lm_proj"][:, 0] <= s["W"] + 20)
    # every pool row is the descriptor of the observation it names, and that observation belongs to the landmark
    for j in range(n):
        i = r["lm"][j]
        own = {tuple(o) for o in s["obs"][s["obs_begin"][i]:s["obs_begin"][i + 1]]}
        for row in range(r["desc_begin"][j], r["desc_begin"][j + 1]):
            k = tuple(r["kid"][row])
            assert k in own
            assert np.array_equal(r["cand_desc"][row], s["desc_tab"][k[0] * s["n_cams"] + k[1]][k[2]])
    assert 0 < r["lm_is3d"].sum() < n
