"""B1 (DBoW2 vocabulary descent with FBrisk distance) and the descriptor text formats. CPU part: oracle + formats on the
reference's own vocabulary (resources/small_voc.yml.gz -> tests/golden/voc_tree.npz); GPU part: device descent == oracle."""
import os

import numpy as np
import pytest

from conftest import ROOT
from okvis2_b200 import formats
from okvis2_b200.lib import KP_DTYPE
from oracle.bow_oracle import Vocabulary, score_l1, transform_image


@pytest.fixture(scope="module")
def voc(voc_desc):
    t = np.load(os.path.join(ROOT, "tests", "golden", "voc_tree.npz"))
    return t, Vocabulary(t["k"], t["L"], t["node_id"], t["parent_id"], t["weight"], voc_desc, t["word_id"], t["word_node"])


def test_oracle_descent_on_the_vocabulary_itself(voc, voc_desc):
    t, v = voc
    assert len(t["word_id"]) == 9 ** 3 and all(len(c) in (0, 9) for c in v.children)
    # a leaf's own descriptor descends to a leaf at distance 0 or to an earlier sibling at the same distance; its path
    # passes through nodes whose descriptors are the bitwise majority of their subtree, so most leaves find themselves
    hits = 0
    for i, nid in enumerate(t["node_id"]):
        if v.word[nid] >= 0:
            w, wt, n1 = v.transform(voc_desc[i], levelsup=1)
            hits += int(w == v.word[nid])
            assert 0 <= w < 729 and wt == v.weight[t["word_node"][list(t["word_id"]).index(w)]]
            assert v.children[n1] and any(v.word[c] == w for c in v.children[n1])       # nid is the leaf's parent
    assert hits > 600
    # levelsup = L: the root
    assert v.transform(voc_desc[20], levelsup=3)[2] == 0


def test_oracle_image_vectors(voc, voc_desc):
    t, v = voc
    rng = np.random.default_rng(2)
    a = voc_desc[rng.choice(819, 300)]; b = a.copy(); b[:150] = voc_desc[rng.choice(819, 150)]
    va, fa = transform_image(v, a, levelsup=2)
    vb, _ = transform_image(v, b, levelsup=2)
    assert abs(sum(w for _, w in va) - 1.0) < 1e-12 and all(w > 0 for _, w in va) and [k for k, _ in va] == sorted(k for k, _ in va)
    assert sorted(i for idx in fa.values() for i in idx) == [i for i in range(300) if v.transform(a[i])[1] > 0]
    assert all(n in range(1, 10) for n in fa)                       # levelsup = 2 of L = 3: the 9 first-level nodes
    assert abs(score_l1(va, va) - 1.0) < 1e-12 and 0.0 < score_l1(va, vb) < 1.0
    assert score_l1(va, []) == 0.0


def test_fbrisk_string_round_trip(voc_desc):
    s = formats.fbrisk_to_string(voc_desc[0])
    assert s.startswith("31 0 253 255 99 140 ") and s.endswith(" ") and len(s.split()) == 48   # first node of small_voc.yml.gz
    assert np.array_equal(formats.fbrisk_from_string(s), voc_desc[0])


def test_brisk2_record_round_trip(voc_desc):
    kp = np.zeros(3, KP_DTYPE)
    kp["x"] = [12.5, 700.123474, 0.000123456789]; kp["y"] = [3.0, 479.999969, 1e6]; kp["size"] = [12.0, 18.0, 8.48528]
    lines = formats.write_frame_keypoints(17, 1, kp, voc_desc[:3])
    assert lines[0] == "FRAME:KEYPOINT 17 1 12.5 3 12 BRISK2 " + "".join(f"{b:02x}" for b in voc_desc[0])
    # precision 17 (Component.cpp:407-411): the float32 value promoted to double, %.17g
    assert lines[1].split()[3:6] == ["700.12347412109375", "479.99996948242188", "18"]
    assert lines[2].split()[3:5] == ["0.00012345678987912834", "1000000"]
    k2, d2, n = formats.read_frame_keypoints(lines + ["FRAME 18 0"], 17, 1)
    assert n == 3 and np.array_equal(d2, voc_desc[:3])
    for f in ("x", "y", "size"):   # exact float32 round trip
        assert np.array_equal(k2[f].view(np.uint32), kp[f].view(np.uint32)), f
    with pytest.raises(ValueError):
        formats.read_frame_keypoints(lines, 18, 1)
    with pytest.raises(ValueError):
        formats.read_frame_keypoints([lines[0].replace("BRISK2", "ORB")], 17, 1)


@pytest.mark.gpu
def test_device_descent_equals_oracle(voc, voc_desc):
    from okvis2_b200.frontend import Frontend
    t, v = voc
    fe = Frontend(0)
    try:
        fe.loadVocabulary(t["k"], t["L"], t["node_id"], t["parent_id"], t["weight"], voc_desc, t["word_id"], t["word_node"])
        rng = np.random.default_rng(0)
        flips = (rng.random((819, 48, 8)) < 0.06)
        feats = np.concatenate([voc_desc, voc_desc ^ np.packbits(flips, axis=2)[:, :, 0], rng.integers(0, 256, (500, 48), dtype=np.uint8)])
        for levelsup in (0, 1, 2, 3, 4):
            word, weight, node = fe.bowTransform(feats, levelsup)
            for i in range(0, len(feats), 3):
                assert (word[i], weight[i], node[i]) == v.transform(feats[i], levelsup), (i, levelsup)
        w0, _, _ = fe.bowTransform(feats[:0])
        assert len(w0) == 0
        # image-level vectors (BowVector / FeatureVector) and the L1 score
        va, fa = fe.bowTransformImage(feats[:400], levelsup=2)
        ra, rfa = transform_image(v, feats[:400], levelsup=2)
        assert va == ra and fa == rfa
        vb, _ = fe.bowTransformImage(feats[819:1300], levelsup=2)
        rb, _ = transform_image(v, feats[819:1300], levelsup=2)
        assert fe.bowScoreL1(va, vb) == score_l1(ra, rb) and 0.0 < score_l1(ra, rb) < 1.0
    finally:
        fe.close()
