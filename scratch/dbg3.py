import sys, numpy as np, cv2, ctypes
sys.path.insert(0, '/root/repo'); sys.path.insert(0,'/root/repo/scratch')
from cmp_cv import *
im = cv2.resize(img0, (752,480), interpolation=cv2.INTER_AREA)
ctypes.c_int.in_dll(oracle.lib(), 'okvo_dbg').value = 3
ref, desc, kp, d = cmp(im, 30, 3)
ks = np.array([oracle.lib().okvo_brisk_kscale(float(s)) for s in ref['size']])
amis = ref['angle'] != kp['angle']
dmis = (d != desc).any(1)
for k in np.unique(ks):
    m = ks == k
    print(k, m.sum(), 'angle mis', amis[m].sum(), 'desc mis', dmis[m].sum())
order = np.argsort(ref['size'])
prev=None
runs=[]
for i in order:
    s = float(ref['size'][i]); f = bool(amis[i]); k = int(ks[i])
    if s < 30: runs.append((round(s,3), k, int(f)))
print(runs)
