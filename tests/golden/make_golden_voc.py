"""Extracts the 819 real 48-byte BRISK2 descriptors stored as DBoW2 vocabulary nodes in the reference's
resources/small_voc.yml.gz (k=9, L=3) into tests/golden/voc_descriptors.npy, plus their full Hamming matrix checksum
computed with numpy (known-answer for brisk::Hamming::PopcntofXORed(a, b, 3), reference okvis_frontend/src/FBrisk.cpp:66).
Run in the authoring container:  python tests/golden/make_golden_voc.py
"""
import gzip
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
t = gzip.open("/root/reference/resources/small_voc.yml.gz", "rt").read()
rows = [np.array(m.split(), np.uint8) for m in re.findall(r'descriptor:"([^"]*)"', t)]
d = np.stack(rows)
assert d.shape == (819, 48), d.shape
ham = np.unpackbits(d[:, None, :] ^ d[None, :, :], axis=2).sum(2).astype(np.uint16)
np.save(os.path.join(HERE, "voc_descriptors.npy"), d)
np.save(os.path.join(HERE, "voc_hamming_rowsum.npy"), ham.sum(1).astype(np.int64))
print(d.shape, ham.mean(), ham[ham > 0].min())

# tree structure (file order matters: DBoW2 appends children to their parent in the order the nodes are listed)
nodes = re.findall(r'nodeId:(\d+), parentId:(\d+), weight:([0-9.eE+-]+),', t)
words = re.findall(r'wordId:(\d+), nodeId:(\d+)', t)
assert len(nodes) == 819
np.savez_compressed(os.path.join(HERE, "voc_tree.npz"), node_id=np.array([int(a) for a, _, _ in nodes], np.int32),
                    parent_id=np.array([int(b) for _, b, _ in nodes], np.int32), weight=np.array([float(c) for _, _, c in nodes]),
                    word_id=np.array([int(a) for a, _ in words], np.int32), word_node=np.array([int(b) for _, b in words], np.int32),
                    k=9, L=3)
print(len(nodes), "nodes", len(words), "words")
