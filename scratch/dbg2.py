import sys, numpy as np, cv2, ctypes
sys.path.insert(0, '/root/repo'); sys.path.insert(0,'/root/repo/scratch')
from cmp_cv import *
im = cv2.resize(img0, (752,480), interpolation=cv2.INTER_AREA)
o = oracle.Brisk(30,0); sp, lp = o.pairs(); print(len(sp), len(lp), o.D)
for v in range(int(sys.argv[1]) if len(sys.argv)>1 else 2):
    ctypes.c_int.in_dll(oracle.lib(), 'okvo_dbg').value = v
    print('variant', v)
    cmp(im, 30, 0); cmp(im, 30, 3)
ctypes.c_int.in_dll(oracle.lib(), 'okvo_dbg').value = 3
ref, desc, kp, d = cmp(im, 30, 3)
bad = np.nonzero((d != desc).any(1))[0]
print('bad sizes', np.unique(np.round(ref['size'][bad],1))[:20], 'min bad size', ref['size'][bad].min() if len(bad) else None)
good = np.nonzero(~(d != desc).any(1))[0]
print('max good size', ref['size'][good].max())
ab = np.nonzero(ref['angle'] != kp['angle'])[0]
print('angle bad', len(ab), ref['angle'][ab][:5], kp['angle'][ab][:5])
