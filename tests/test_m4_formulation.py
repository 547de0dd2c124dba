"""CPU suite: the scan / gate split of the device M4 matcher (k_m4_scan, k_m4_gate, k_m4_finish) rests on one claim: the
reference's sequential loop (Frontend.cpp:2016-2074: `if (dist < best && gate(k0, k1)) best = dist`) returns, per query,
the minimum of (distance, candidate index) over the candidates with distance < threshold that pass the gate -- the gate being
a pure function of the pair. Checked here with the oracle itself: the per-pair gate is obtained by running the oracle's loop
on single-candidate pools, the split result is assembled in numpy and compared with the oracle on the full pool."""
import numpy as np

import oracle
from okvis2_b200.synth import stereo_scene

PC = np.array([bin(i).count("1") for i in range(256)], np.int64)


def test_min_over_gated_hits_equals_sequential_loop():
    s = stereo_scene(5, 160, 150, flip_p=0.03)
    # near-duplicate descriptors so that queries have SEVERAL candidates below the threshold, in both index orders
    s["desc1"][40:80] = s["desc1"][0:40]
    s["desc1"][100:120] = s["desc1"][20:40]
    args = (s["r_WC0"], s["r_WC1"], s["T_CW0"], s["T_CW1"], 60)
    ref = oracle.match_stereo(s["desc0"], s["valid0"], s["e0_W"], s["sof0"], s["desc1"], s["valid1"], s["e1_W"], s["sof1"], *args)
    n0, n1 = len(s["desc0"]), len(s["desc1"])
    dist = PC[s["desc0"][:, None, :] ^ s["desc1"][None, :, :]].sum(2)
    hits = [(q, c) for q in range(n0) for c in range(n1) if dist[q, c] < 60]          # what k_m4_scan lists
    assert len(hits) > 40 and max(np.bincount([q for q, _ in hits])) >= 2
    best = {}
    for q, c in hits:                                                                    # k_m4_gate: one thread per hit
        one = oracle.match_stereo(s["desc0"][q:q + 1], s["valid0"][q:q + 1], s["e0_W"][q:q + 1], s["sof0"][q:q + 1],
                                  s["desc1"][c:c + 1], s["valid1"][c:c + 1], s["e1_W"][c:c + 1], s["sof1"][c:c + 1], *args)
        if one[0][0] == 0:                                                               # the pair passes the gate
            key = (int(dist[q, c]) << 32) | c
            if key < best.get(q, (60 << 32) | 0xFFFFFFFF):
                best[q] = key; best[(q, "hp")] = (one[2][0], one[3][0])
    k1 = np.full(n0, -1, np.int32); d = np.full(n0, 60, np.uint32); hp = np.zeros((n0, 4)); init = np.zeros(n0, np.uint8)
    for q in range(n0):                                                                  # k_m4_finish
        if q in best:
            k1[q] = best[q] & 0xFFFFFFFF; d[q] = best[q] >> 32; hp[q], init[q] = best[(q, "hp")]
    assert np.array_equal(k1, ref[0]) and np.array_equal(d, ref[1].astype(np.uint32))
    assert np.array_equal(hp.view(np.uint64), ref[2].view(np.uint64)) and np.array_equal(init, ref[3])
    assert (ref[0] >= 0).sum() > 20
