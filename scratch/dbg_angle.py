import sys, numpy as np, cv2
sys.path.insert(0, '/root/repo'); sys.path.insert(0,'/root/repo/scratch')
from cmp_cv import *
im = cv2.resize(img0, (752,480), interpolation=cv2.INTER_AREA)
ref, desc, kp, d = cmp(im, 30, 0)
da = kp['angle'] - ref['angle']
print(np.c_[ref['angle'][:15], kp['angle'][:15], da[:15]])
print('absdiff stats', np.abs(da).min(), np.median(np.abs(da)), np.abs(da).max())
