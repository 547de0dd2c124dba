import sys, numpy as np, cv2
sys.path.insert(0, '/root/repo')
import oracle
img0 = cv2.imread('/root/reference/okvis_multisensor_processing/test/testImage.jpg', 0)
def check_resize():
    for (sw, sh) in [(752,480),(1024,1024),(720,540),(500,320),(341,341), (375, 241)]:
        im = cv2.resize(img0, (sw, sh), interpolation=cv2.INTER_AREA)
        for (dw, dh) in [(sw//2, sh//2), (2*(sw//3), 2*(sh//3))]:
            ref = cv2.resize(im, (dw, dh), interpolation=cv2.INTER_AREA)
            got = oracle.resize_area(im, dw, dh)
            print('resize', (sw,sh), '->', (dw,dh), 'mismatch', int((ref != got).sum()))
def cmp(im, thr, octv, verbose=True):
    b = cv2.BRISK_create(thr, octv, 1.0)
    kps, desc = b.detectAndCompute(im, None)
    ref = oracle.cv_keypoints_to_array(kps)
    o = oracle.Brisk(thr, octv)
    kp, d = o.detect_and_compute(im)
    ok = len(kp) == len(ref)
    print(f'thr={thr} oct={octv} cv={len(ref)} oracle={len(kp)}', end=' ')
    if ok:
        for f in ref.dtype.names:
            bad = (ref[f] != kp[f]).sum()
            if bad: print(f'{f}:{bad}', end=' ')
        print('desc rows differing:', int((d != desc).any(1).sum()))
    else:
        # match by (octave, rounded x,y)
        sr = {(int(k['octave']), round(float(k['x']),3), round(float(k['y']),3)) for k in ref}
        so = {(int(k['octave']), round(float(k['x']),3), round(float(k['y']),3)) for k in kp}
        print('only cv', len(sr-so), 'only oracle', len(so-sr))
        if verbose:
            print(sorted(sr-so)[:10]); print(sorted(so-sr)[:10])
    return ref, desc, kp, d
if __name__ == '__main__':
    check_resize()
    for (w,h) in [(752,480),(1024,1024)]:
        im = cv2.resize(img0, (w,h), interpolation=cv2.INTER_AREA)
        for thr, octv in [(30,0),(30,3),(60,3),(20,4),(10,1)]:
            cmp(im, thr, octv)
