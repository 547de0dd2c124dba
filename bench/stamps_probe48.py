"""Cycle stamps of k_uniformity (D = 48 mode) on the bench frames: where a frame's latency goes.
usage: python bench/stamps_probe48.py [batch] [radius] [threshold] [max_kp]"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
from okvis2_b200 import lib as okl
from okvis2_b200.frontend import Frontend

cfg = dict(B.OKVIS48)
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1
if len(sys.argv) > 2: cfg["radius"] = float(sys.argv[2])
if len(sys.argv) > 3: cfg["abs_threshold"] = int(sys.argv[3])
if len(sys.argv) > 4: cfg["max_kp"] = int(sys.argv[4])
fe = Frontend(1, cfg["W"], cfg["H"], 0, max_batch=nb, descriptor_bytes=48)
fe.configure(threshold=cfg["radius"], absolute_threshold=cfg["abs_threshold"], octaves=0, max_keypoints=cfg["max_kp"])
L_, _ = B.make_frames(cfg, nb, 1000)
for rep in range(3):
    out = fe.detectAndDescribeBatch(0, L_)
lib = okl.lib()
st = (C.c_longlong * 16)()
for f in range(min(nb, 3)):
    okl.check(lib.okb_debug_stamps(fe.ctx, 0, f, st))
    s = list(st)
    us = lambda a, b: (s[b] - s[a]) / 1.9e3
    print(f"frame {f}: maxima {s[6]} decided {s[7]} rounds {s[5]} kept {len(out[f][0])} | sort {us(0,1):.1f} cells {us(1,2):.1f} rounds {us(2,3):.1f} output {us(3,4):.1f} total {us(0,4):.1f} us")
fe.close()
