import sys, time, numpy as np, cv2
sys.path.insert(0, '/root/repo'); sys.path.insert(0,'/root/repo/scratch')
from cmp_cv import cmp
import oracle
from okvis2_b200.synth import synth_frame, synth_stereo
t=time.time(); im = synth_frame(1000, 752, 480); print('synth time', time.time()-t, im.mean(), im.std())
cv2.imwrite('/tmp/synth.png', im)
for seed in (1000, 2003):
  for (w,h) in [(752,480),(1024,1024),(720,540)]:
    l, r = synth_stereo(seed, w, h)
    for thr, octv in [(30,3),(50,3),(30,0), (15, 2)]:
        cmp(l, thr, octv, verbose=False)
    cmp(r, 40, 3, verbose=False)
