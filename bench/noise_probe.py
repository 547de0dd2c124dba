"""Robustness probe: frames of pure noise (far more corners and ties than any capacity) through the D = 64 and D = 48 detectors: a loud
capacity error or a valid result, never a device fault. usage: python bench/noise_probe.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from okvis2_b200 import lib as okl
from okvis2_b200.frontend import Frontend

rng = np.random.default_rng(0)
for (W, H, octv, B) in [(1024, 1024, 3, 4), (752, 480, 3, 8), (752, 480, 0, 2)]:
    for kind in ("uniform", "binary", "gradient+noise"):
        if kind == "uniform":
            imgs = rng.integers(0, 256, (B, H, W), dtype=np.uint8)
        elif kind == "binary":
            imgs = (rng.integers(0, 2, (B, H, W), dtype=np.uint8) * 255).astype(np.uint8)
        else:
            imgs = ((np.arange(W)[None, None, :] % 256) + rng.integers(0, 64, (B, H, W))).astype(np.uint8)
        fe = Frontend(1, W, H, max_batch=B)
        fe.configure(threshold=30, octaves=octv, max_keypoints=2496)
        try:
            out = fe.detectAndDescribeBatch(0, imgs)
            res = f"ok: {[len(o[0]) for o in out]}"
        except okl.OkbError as e:
            res = f"OkbError {e.status}: {str(e)[:110]}"
        rc = okl.lib().okb_sync(fe.ctx)
        print(W, H, octv, kind, "->", res, "| sync rc", rc, flush=True)
        fe.close()
