"""Generates tests/golden/bfmatcher_cv2_4_13.npz: cv2.BFMatcher(cv2.NORM_HAMMING) best matches (OpenCV 4.13.0) for a fixed
query / train descriptor set -- the cross-check SURVEY §8c names for the Hamming core of the matchers when their geometric
gates are disabled. Run in the authoring container (cv2 importable); the fixture travels."""
import os

import cv2
import numpy as np

rng = np.random.default_rng(123)
voc = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "voc_descriptors.npy"))      # 819 real 48-byte descriptors
out = {}
for D in (48, 64):
    if D == 48:
        train = voc[:600].copy()
        bits = np.unpackbits(voc[rng.choice(600, 250)], axis=1)
        query = np.packbits(bits ^ (rng.random(bits.shape) < 0.1).astype(np.uint8), axis=1)
    else:
        train = rng.integers(0, 256, (700, 64), dtype=np.uint8)
        train[300:350] = train[100:150]                                   # exact duplicates: ties on the minimum distance
        bits = np.unpackbits(train[rng.choice(700, 300)], axis=1)
        query = np.packbits(bits ^ (rng.random(bits.shape) < 0.05).astype(np.uint8), axis=1)
    m = cv2.BFMatcher(cv2.NORM_HAMMING).match(query, train)
    assert [x.queryIdx for x in m] == list(range(len(query)))
    out[f"train{D}"] = train; out[f"query{D}"] = query
    out[f"idx{D}"] = np.array([x.trainIdx for x in m], np.int32)
    out[f"dist{D}"] = np.array([x.distance for x in m], np.float32)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bfmatcher_cv2_4_13.npz"), **out)
print({k: v.shape for k, v in out.items()})
