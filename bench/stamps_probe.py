"""Cycle stamps of the single-CTA kernels (k_resolve, k_finalize) on bench frames: where a frame's latency goes.
usage: python bench/stamps_probe.py [workload] [batch]"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
from okvis2_b200 import lib as okl
from okvis2_b200.frontend import Frontend

name = sys.argv[1] if len(sys.argv) > 1 else "euroc"
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = B.CONFIGS[name]
fe = Frontend(1, cfg["W"], cfg["H"], 0, max_batch=nb)
fe.configure(threshold=cfg["threshold"], octaves=cfg["octaves"], max_keypoints=cfg["max_kp"])
L_, _ = B.make_frames(cfg, nb, 1000)
for rep in range(3):
    out = fe.detectAndDescribeBatch(0, L_)
lib = okl.lib()
st = (C.c_longlong * 16)()
for f in range(min(nb, 4)):
    okl.check(lib.okb_debug_stamps(fe.ctx, 0, f, st))
    s = list(st)
    mhz = 1.9
    us = lambda a, b: (s[b] - s[a]) / mhz / 1e3
    print(f"frame {f}: resolve T={s[6]} rounds={s[5]} overflowed={s[15]} load {us(0,1):.1f} lists {us(1,2):.1f} rounds {us(2,4):.1f} (r1 {us(2,3):.1f} r2 {us(3,7):.1f} r3 {us(7,9):.1f}) total {us(0,4):.1f} us | "
          f"finalize V={s[13]} n={s[14]} order {us(8,10):.1f} select {us(10,11):.1f} out {us(11,12):.1f} total {us(8,12):.1f} us")
fe.close()
