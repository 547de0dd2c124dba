"""oracle/bow_oracle.py -- TEST INFRASTRUCTURE ONLY (CPU oracle of B1; never on the product path).

Restatement of DBoW2::TemplatedVocabulary<FBrisk::TDescriptor, FBrisk>::transform(feature, word_id, weight, nid, levelsup)
with F::distance = FBrisk::distance (reference okvis_frontend/src/FBrisk.cpp:64-67: popcount of the XOR over L/16 128-bit
words). external/DBoW2 (dorian3d/DBoW2, patched by external/patches/DBoW2) is an EMPTY submodule in /root/reference; the
descent below follows the published TemplatedVocabulary.h: start at the root, `final_id = children[0]`, replace it by a
later child only when its distance is strictly smaller, repeat until a leaf; `nid` is the node at level L - levelsup.
Children are in the order the nodes are listed in the vocabulary file (load() appends them to their parent in that order).
PARITY UNPINNED against DBoW2 itself (no golden word ids exist in the reference tree); pinned facts: node descriptors,
parents, weights and word ids of resources/small_voc.yml.gz (tests/golden/voc_tree.npz, voc_descriptors.npy)."""
import numpy as np

_PC = np.array([bin(i).count("1") for i in range(256)], np.int32)


class Vocabulary:
    def __init__(self, k, L, node_id, parent_id, weight, desc, word_id, word_node):
        self.k, self.L = int(k), int(L)
        n = int(max(node_id))
        self.children = [[] for _ in range(n + 1)]
        self.desc = np.zeros((n + 1, desc.shape[1]), np.uint8)
        self.weight = np.zeros(n + 1)
        self.word = -np.ones(n + 1, np.int64)
        for i, (nid, pid) in enumerate(zip(node_id, parent_id)):
            self.children[int(pid)].append(int(nid))
            self.desc[int(nid)] = desc[i]; self.weight[int(nid)] = weight[i]
        for w, nid in zip(word_id, word_node):
            self.word[int(nid)] = int(w)

    def distance(self, a, b):
        return float(_PC[np.bitwise_xor(a, b)].sum())

    def transform(self, feature, levelsup=0):
        nid_level = self.L - levelsup
        nid = 0
        final_id = 0
        current_level = 0
        while True:
            current_level += 1
            nodes = self.children[final_id]
            final_id = nodes[0]
            best_d = self.distance(feature, self.desc[final_id])
            for cid in nodes[1:]:
                d = self.distance(feature, self.desc[cid])
                if d < best_d:
                    best_d = d
                    final_id = cid
            if current_level == nid_level:
                nid = final_id
            if not self.children[final_id]:
                break
        return int(self.word[final_id]), float(self.weight[final_id]), nid


def transform_image(voc, features, levelsup=0):
    """TemplatedVocabulary::transform(features, BowVector, FeatureVector, levelsup) for TF_IDF weighting and L1 scoring (the
    types stored in small_voc.yml.gz: weightingType 0, scoringType 0), restated from the published DBoW2 source: every
    feature adds its leaf's weight to its word (skipped when the weight is 0 = stopped word) and its index to the node
    `levelsup` levels above the leaves; L1 scoring normalises the vector by the sum of its weights (words in id order).
    Returns (sorted [(word id, weight)], {node id: [feature indices]})."""
    v, fv = {}, {}
    for i, f in enumerate(features):
        word, w, nid = voc.transform(f, levelsup)
        if w > 0:
            v[word] = v.get(word, 0.0) + w
            fv.setdefault(nid, []).append(i)
    items = sorted(v.items())
    norm = 0.0
    for _, w in items:
        norm += abs(w)
    if norm > 0.0:
        items = [(k, w / norm) for k, w in items]
    return items, fv


def score_l1(v1, v2):
    """DBoW2::L1Scoring::score on two normalised BowVectors (sorted (word, weight) lists): 1 - 0.5 * ||v1 - v2||_1, summed
    over the common words as |a - b| - |a| - |b| in word order."""
    score, i, j = 0.0, 0, 0
    while i < len(v1) and j < len(v2):
        if v1[i][0] == v2[j][0]:
            a, b = v1[i][1], v2[j][1]
            score += abs(a - b) - abs(a) - abs(b)
            i += 1; j += 1
        elif v1[i][0] < v2[j][0]:
            i += 1
        else:
            j += 1
    return -score / 2.0
