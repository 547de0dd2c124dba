import numpy as np, cv2
rng = np.random.default_rng(0)
img = cv2.imread('/root/reference/okvis_multisensor_processing/test/testImage.jpg', 0)
img = cv2.resize(img, (752,480), interpolation=cv2.INTER_AREA)
def bstar(img, offs, n):
    H,W = img.shape
    m = max(max(abs(a),abs(b)) for a,b in offs)
    p = img.astype(np.int32)
    c = []
    for dx,dy in offs:
        c.append(np.roll(np.roll(p, -dy, 0), -dx, 1))
    c = np.stack(c)  # c[i][y,x] = img[y+dy, x+dx]
    d = c - p[None]
    N = len(offs)
    best = np.full(p.shape, -999, np.int32)
    for s in range(N):
        idx = [(s+k)%N for k in range(n)]
        mn = d[idx].min(0); best = np.maximum(best, mn-1)
        mx = (-d[idx]).min(0); best = np.maximum(best, mx-1)
    return best, m
offs16 = [(-3,0),(-3,-1),(-2,-2),(-1,-3),(0,-3),(1,-3),(2,-2),(3,-1),(3,0),(3,1),(2,2),(1,3),(0,3),(-1,3),(-2,2),(-3,1)]
offs8 = [(-1,0),(-1,-1),(0,-1),(1,-1),(1,0),(1,1),(0,1),(-1,1)]
for name,offs,n,typ in [('oast916',offs16,9,cv2.AgastFeatureDetector_OAST_9_16),('agast58',offs8,5,cv2.AgastFeatureDetector_AGAST_5_8)]:
    for t in (1,5,30):
        det = cv2.AgastFeatureDetector_create(t, False, typ)
        kps = det.detect(img)
        b, m = bstar(img, offs, n)
        H,W = img.shape
        got = {(int(k.pt[0]),int(k.pt[1])): k.response for k in kps}
        xs = [k.pt[0] for k in kps]; ys=[k.pt[1] for k in kps]
        print(name, t, len(kps), 'x range', min(xs), max(xs), 'y range', min(ys), max(ys))
        # expected set
        exp = {}
        mm = 3 if n==9 else None
        for y in range(H):
            for x in range(W):
                pass
        mask = b >= t
        # try margins
        ok=None
        for mg in (1,2,3,4):
            mk = np.zeros_like(mask); mk[mg:H-mg, mg:W-mg] = True
            e = np.argwhere(mask & mk)
            es = {(int(x),int(y)) for y,x in e}
            if es == set(got.keys()):
                ok = mg; break
        print('  margin match:', ok)
        if ok:
            bad = sum(1 for (x,y),r in got.items() if r != b[y,x])
            print('  response mismatches', bad, 'order raster?', all((ys[i],xs[i]) < (ys[i+1],xs[i+1]) for i in range(len(kps)-1)))
