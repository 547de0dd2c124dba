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
