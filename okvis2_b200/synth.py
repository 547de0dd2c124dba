"""Synthetic camera frames and landmark maps for the benchmark and the parity tests.

numpy only (no cv2), deterministic in `seed`, so the very same bytes are produced in the authoring container
and on the GPU box.
"""
import numpy as np


def _gauss_blur(img, sigma=1.0):
    r = int(3 * sigma + 0.5)
    xs = np.arange(-r, r + 1, dtype=np.float32)
    k = np.exp(-0.5 * (xs / sigma) ** 2).astype(np.float32)
    k /= k.sum()
    p = np.pad(img, ((r, r), (r, r)), mode="edge")
    tmp = np.zeros_like(img)
    H, W = img.shape
    for i, w in enumerate(k):
        tmp += w * p[r:r + H, i:i + W]
    p = np.pad(tmp, ((r, r), (0, 0)), mode="edge")
    out = np.zeros_like(img)
    for i, w in enumerate(k):
        out += w * p[i:i + H, :]
    return out


def _scene(rng, n_obj, W, H):
    """Random objects: (kind, cx, cy, a, b, angle, intensity, disparity)."""
    kind = rng.integers(0, 3, n_obj)            # 0 axis rect, 1 rotated rect, 2 disc
    cx = rng.uniform(-20, W + 20, n_obj)
    cy = rng.uniform(-20, H + 20, n_obj)
    a = rng.uniform(6, 80, n_obj)
    b = rng.uniform(6, 80, n_obj)
    ang = rng.uniform(0, np.pi, n_obj)
    inten = rng.uniform(0, 255, n_obj)
    disp = rng.uniform(4, 40, n_obj)
    return kind, cx, cy, a, b, ang, inten, disp


def _render(scene, W, H, shift=(0.0, 0.0), use_disparity=False, noise_rng=None):
    kind, cx, cy, a, b, ang, inten, disp = scene
    img = np.full((H, W), 128.0, np.float32)
    for i in range(len(kind)):
        x0 = cx[i] + shift[0] - (disp[i] if use_disparity else 0.0)
        y0 = cy[i] + shift[1]
        r = 0.5 * np.hypot(a[i], b[i]) + 1
        xa, xb = int(max(0, np.floor(x0 - r))), int(min(W, np.ceil(x0 + r) + 1))
        ya, yb = int(max(0, np.floor(y0 - r))), int(min(H, np.ceil(y0 + r) + 1))
        if xa >= xb or ya >= yb:
            continue
        yy, xx = np.mgrid[ya:yb, xa:xb].astype(np.float32)
        dx, dy = xx - x0, yy - y0
        if kind[i] == 2:
            m = dx * dx + dy * dy <= (0.5 * a[i]) ** 2
        else:
            t = ang[i] if kind[i] == 1 else 0.0
            c, s = np.float32(np.cos(t)), np.float32(np.sin(t))
            u, v = c * dx + s * dy, -s * dx + c * dy
            m = (np.abs(u) <= 0.5 * a[i]) & (np.abs(v) <= 0.5 * b[i])
        img[ya:yb, xa:xb][m] = inten[i]
    img = _gauss_blur(img, 1.0)
    if noise_rng is not None:
        img = img + noise_rng.normal(0.0, 2.0, img.shape).astype(np.float32)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def synth_frame(seed, W, H, n_obj=None, t=0, right=False):
    """One u8 frame. `t` translates the scene by (2t, t) px; `right=True` renders the stereo partner."""
    if n_obj is None:
        n_obj = int(round(600 * (W * H) / (752.0 * 480.0)))
    rng = np.random.default_rng(seed)
    scene = _scene(rng, n_obj, W, H)
    nrng = np.random.default_rng((seed * 2 + (1 if right else 0)) * 7919 + t)
    return _render(scene, W, H, shift=(2.0 * t, 1.0 * t), use_disparity=right, noise_rng=nrng)


def synth_stereo(seed, W, H, t=0, n_obj=None):
    return synth_frame(seed, W, H, n_obj, t, False), synth_frame(seed, W, H, n_obj, t, True)


def noisy_copies(desc, n, flip_p, rng):
    """n descriptors drawn from `desc` with each bit flipped with probability flip_p."""
    idx = rng.integers(0, len(desc), n)
    bits = np.unpackbits(desc[idx], axis=1)
    flips = (rng.random(bits.shape) < flip_p).astype(np.uint8)
    return np.packbits(bits ^ flips, axis=1), idx


# ---- synthetic matcher inputs -------------------------------------------------------------------
def random_descriptors(rng, n, D):
    return rng.integers(0, 256, (n, D), dtype=np.uint8)


def map_scene(seed, kp_xy, kp_desc, n_lm, frac_copy=0.6, frac_near=0.10, frac_3d=0.7, flip_p=0.08, W=1024, H=1024):
    """Landmark pool for M1/M2: n_lm landmarks with 1-3 descriptors each (ascending landmark slot).

    frac_copy of the landmarks carry noisy copies of query descriptors (true matches below threshold 60), the rest are
    random. frac_near of the landmarks project within 20 px of their source keypoint; the others land anywhere.
    Returns dict(cand_desc, cand_lm, lm_proj, lm_is3d, cand_e_W, cand_r_W).
    """
    rng = np.random.default_rng(seed)
    n_kp, D = kp_desc.shape
    n_desc = rng.integers(1, 4, n_lm)
    cand_lm = np.repeat(np.arange(n_lm, dtype=np.int32), n_desc)
    n_cand = len(cand_lm)
    src = rng.integers(0, max(n_kp, 1), n_lm)
    is_copy = rng.random(n_lm) < frac_copy
    cand_desc = random_descriptors(rng, n_cand, D)
    if n_kp:
        csrc = src[cand_lm]
        bits = np.unpackbits(kp_desc[csrc], axis=1)
        flips = (rng.random(bits.shape) < flip_p).astype(np.uint8)
        noisy = np.packbits(bits ^ flips, axis=1)
        m = is_copy[cand_lm]
        cand_desc[m] = noisy[m]
    near = rng.random(n_lm) < frac_near
    lm_proj = np.stack([rng.uniform(-20, W + 20, n_lm), rng.uniform(-20, H + 20, n_lm)], 1)
    if n_kp:
        off = rng.uniform(-14, 14, (n_lm, 2))
        lm_proj[near] = kp_xy[src[near]] + off[near]
    lm_is3d = (rng.random(n_lm) < frac_3d).astype(np.uint8)
    # observation rays of the pooled descriptors (for M2): unit vectors around +z, observer positions near the origin
    e = rng.normal(0, 0.25, (n_cand, 3)); e[:, 2] = 1.0
    e /= np.linalg.norm(e, axis=1, keepdims=True)
    r = rng.normal(0, 0.3, (n_cand, 3))
    return dict(cand_desc=cand_desc, cand_lm=cand_lm, lm_proj=lm_proj, lm_is3d=lm_is3d, cand_e_W=e, cand_r_W=r,
                src=src, is_copy=is_copy)


def rot(axis, a):
    c, s = np.cos(a), np.sin(a)
    x, y, z = axis
    return np.array([[c + x * x * (1 - c), x * y * (1 - c) - z * s, x * z * (1 - c) + y * s],
                     [y * x * (1 - c) + z * s, c + y * y * (1 - c), y * z * (1 - c) - x * s],
                     [z * x * (1 - c) - y * s, z * y * (1 - c) + x * s, c + z * z * (1 - c)]])


def stereo_scene(seed, n0, n1, D=64, f=458.0, baseline=0.11, flip_p=0.04, frac_match=0.7, ray_noise=2e-4):
    """Two views of random 3-D points for M2/M3/M4: world-frame unit rays, descriptors, keypoint sizes, poses."""
    rng = np.random.default_rng(seed)
    C0 = rot((0, 1, 0), 0.02) @ rot((1, 0, 0), -0.01)
    C1 = rot((0, 1, 0), -0.015) @ rot((0, 0, 1), 0.01)
    r0 = np.array([0.0, 0.0, 0.0]); r1 = np.array([baseline, 0.003, -0.002])
    n_pts = max(n0, n1)
    depth = np.exp(rng.uniform(np.log(0.15), np.log(60.0), n_pts))
    dirs = rng.normal(0, 0.35, (n_pts, 3)); dirs[:, 2] = 1.0
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    P = dirs * depth[:, None]
    desc_pts = random_descriptors(rng, n_pts, D)

    def view(n, r, noise_seed):
        g = np.random.default_rng(noise_seed)
        idx = g.permutation(n_pts)[:n]
        e = P[idx] - r
        e /= np.linalg.norm(e, axis=1, keepdims=True)
        e += g.normal(0, ray_noise, e.shape)
        e /= np.linalg.norm(e, axis=1, keepdims=True)
        bits = np.unpackbits(desc_pts[idx], axis=1)
        flips = (g.random(bits.shape) < flip_p).astype(np.uint8)
        d = np.packbits(bits ^ flips, axis=1)
        outl = g.random(n) > frac_match
        d[outl] = random_descriptors(g, int(outl.sum()), D)
        size = g.choice([12.0, 18.0, 24.0, 36.0, 48.0], n) * g.uniform(0.9, 1.1, n)
        valid = (g.random(n) > 0.03).astype(np.uint8)
        return idx, np.ascontiguousarray(e), d, size / f, valid

    i0, e0, d0, sof0, v0 = view(n0, r0, seed * 3 + 1)
    i1, e1, d1, sof1, v1 = view(n1, r1, seed * 3 + 2)

    def T_CW(Cm, r):
        R = Cm.T
        t = -(R @ r)
        return np.ascontiguousarray(np.concatenate([R, t[:, None]], 1).reshape(12))

    return dict(desc0=d0, e0_W=e0, sof0=sof0, valid0=v0, desc1=d1, e1_W=e1, sof1=sof1, valid1=v1, r_WC0=r0, r_WC1=r1,
                T_CW0=T_CW(C0, r0), T_CW1=T_CW(C1, r1), idx0=i0, idx1=i1, f=f)


def landmark_scene(seed, n_lm=2000, n_slots=10, n_cams=2, n_kp=600, D=64, W=752, H=480, f=458.0, step=0.35):
    """Synthetic map for the landmark-candidate preparation (P1): a camera rig moving along x through a cloud of
    landmarks, every landmark observed by a random subset of the (frame slot, camera) views. Poses are packed as 12
    doubles (C_WC row-major, r_WC). Returns the inputs of Frontend.prepareLandmarksToMatch / oracle.prepare_landmarks."""
    rng = np.random.default_rng(seed)

    def rot(rx, ry, rz):
        cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
        Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]); Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
        Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
        return Rz @ Ry @ Rx

    T_WC_old = np.zeros((n_slots, n_cams, 12))
    for s in range(n_slots):
        C = rot(*(0.05 * rng.standard_normal(3)))
        r = np.array([step * s, 0.0, 0.0]) + 0.05 * rng.standard_normal(3)
        for c in range(n_cams):
            Cc = C @ rot(0.0, 0.02 * c, 0.0)
            T_WC_old[s, c, :9] = Cc.ravel(); T_WC_old[s, c, 9:] = r + C @ np.array([0.11 * c, 0.0, 0.0])
    # current view: a little beyond the last slot
    C1 = rot(*(0.05 * rng.standard_normal(3)))
    r1 = np.array([step * n_slots, 0.02, -0.01])
    T_WC1 = np.concatenate([C1.ravel(), r1]); T_CW1 = np.concatenate([C1.T.ravel(), -(C1.T @ r1)])
    # landmarks: mostly in front (z = 2..15 m), some behind / far off axis / at the w < 0 branch / near the singularity
    p = np.stack([rng.uniform(-8, 12, n_lm), rng.uniform(-5, 5, n_lm), rng.uniform(1.0, 15.0, n_lm)], 1)
    behind = rng.random(n_lm) < 0.1
    p[behind, 2] = -rng.uniform(0.5, 10.0, behind.sum())
    w = np.where(rng.random(n_lm) < 0.1, -1.0, 1.0) * rng.uniform(0.5, 2.0, n_lm)
    hp_W = np.concatenate([p * w[:, None], w[:, None]], 1)
    quality = 10.0 ** rng.uniform(-4.0, 0.0, n_lm)   # low-quality landmarks with parallax stay non-3d
    # feature tables
    desc_tab = [rng.integers(0, 256, (n_kp, D), dtype=np.uint8) for _ in range(n_slots * n_cams)]
    ray_tab = [np.concatenate([rng.uniform(-0.8, 0.8, (n_kp, 2)), np.ones((n_kp, 1))], 1) for _ in range(n_slots * n_cams)]
    obs_begin = [0]; obs = []
    for i in range(n_lm):
        k = int(rng.integers(0, 9))
        views = rng.choice(n_slots * n_cams, size=min(k, n_slots * n_cams), replace=False)
        ids = sorted((int(v) // n_cams, int(v) % n_cams, int(rng.integers(0, n_kp))) for v in views)  # std::set order
        obs.extend(ids); obs_begin.append(len(obs))
    return dict(hp_W=hp_W, quality=quality, obs_begin=np.array(obs_begin, np.int32), obs=np.array(obs, np.int32).reshape(-1, 3),
                T_WC_old=T_WC_old, T_WC1=T_WC1, T_CW1=T_CW1, desc_tab=desc_tab, ray_tab=ray_tab, n_cams=n_cams, n_slots=n_slots,
                D=D, W=W, H=H, intr=np.array([f, f * 0.997, W / 2 - 8.8, H / 2 + 8.4, -0.2834, 0.0740, 0.00019, 1.76e-05]))


def radtan_project(p_C, fu, fv, cu, cv, k):
    """PinholeCamera<RadialTangentialDistortion>::project of camera-frame points (n x 3) -> pixel coordinates (n x 2), float64."""
    u0, u1 = p_C[:, 0] / p_C[:, 2], p_C[:, 1] / p_C[:, 2]
    k1, k2, p1, p2 = k
    mx, my, mxy = u0 * u0, u1 * u1, u0 * u1
    rho = mx + my
    rad = k1 * rho + k2 * rho * rho
    d0 = u0 + u0 * rad + 2.0 * p1 * mxy + p2 * (rho + 2.0 * mx)
    d1 = u1 + u1 * rad + 2.0 * p2 * mxy + p1 * (rho + 2.0 * my)
    return np.stack([fu * d0 + cu, fv * d1 + cv], 1)


def pose12(Cm, r):
    """(C row-major 9, r 3) packing of a pose and of its inverse, the layout of okb_prepare_view_t / okb_older_view_t."""
    Cm = np.asarray(Cm, np.float64); r = np.asarray(r, np.float64)
    return np.concatenate([Cm.ravel(), r]), np.concatenate([Cm.T.ravel(), -(Cm.T @ r)])


def motion_scene(seed, n_views=5, n0=600, n1=900, W=752, H=480, f=458.0, k=(-0.2834, 0.0740, 0.00019, 1.76e-05), flip_p=0.04,
                 frac_seen=0.7, premated=0.2):
    """A camera moving through a cloud of 3-D points for the M3 sequence (Frontend::matchMotionStereo): the current view and
    `n_views` older keyframe views, each with keypoints (pixels, float32), descriptors (noisy copies of the points' descriptors,
    so that several older views compete for the same current keypoint), keypoint sizes, eligibility flags and poses.
    Returns dict(cur=dict(kp xy, desc, size, matched), views=[dict(xy, desc, size, use, T_WC, T_CW)], T_WC1, T_CW1, intr)."""
    rng = np.random.default_rng(seed)
    fu, fv, cu, cv = f, f * 0.997, W / 2 - 8.8, H / 2 + 8.4
    n_pts = max(n0, n1) * 2
    P = np.stack([rng.uniform(-6, 6, n_pts), rng.uniform(-4, 4, n_pts), rng.uniform(0.15, 30.0, n_pts)], 1)
    desc_pts = random_descriptors(rng, n_pts, 64)

    def view(Cm, r, n, noise_seed):
        g = np.random.default_rng(noise_seed)
        pc = (P - r) @ Cm          # C_CW = C_WC^T applied to row vectors
        ok = pc[:, 2] > 0.05
        px = np.full((n_pts, 2), -1.0)
        px[ok] = radtan_project(pc[ok], fu, fv, cu, cv, k)
        inside = ok & (px[:, 0] > 20) & (px[:, 0] < W - 20) & (px[:, 1] > 20) & (px[:, 1] < H - 20)
        idx = np.nonzero(inside)[0]
        idx = idx[g.permutation(len(idx))[:int(n * frac_seen)]]
        xy = px[idx] + g.normal(0, 0.4, (len(idx), 2))
        bits = np.unpackbits(desc_pts[idx], axis=1)
        d = np.packbits(bits ^ (g.random(bits.shape) < flip_p).astype(np.uint8), axis=1)
        n_out = n - len(idx)            # outliers: random pixels, random descriptors
        xy = np.concatenate([xy, np.stack([g.uniform(20, W - 20, n_out), g.uniform(20, H - 20, n_out)], 1)])
        d = np.concatenate([d, random_descriptors(g, n_out, 64)])
        perm = g.permutation(n)
        size = (g.choice([12.0, 18.0, 24.0, 36.0], n) * g.uniform(0.9, 1.1, n)).astype(np.float32)
        return np.ascontiguousarray(xy[perm].astype(np.float32)), np.ascontiguousarray(d[perm]), size

    C1 = rot((0, 1, 0), 0.03) @ rot((1, 0, 0), -0.02); r1 = np.array([0.4, 0.02, 0.05])
    T_WC1, T_CW1 = pose12(C1, r1)
    xy1, d1, s1 = view(C1, r1, n1, seed * 11 + 1)
    views = []
    for v in range(n_views):
        Cv = rot((0, 1, 0), 0.03 - 0.015 * (v + 1)) @ rot((0, 0, 1), 0.004 * v)
        rv = r1 - np.array([0.12 * (v + 1), 0.01 * v, 0.02])
        xy, d, sz = view(Cv, rv, n0, seed * 11 + 2 + v)
        Tw, Tc = pose12(Cv, rv)
        views.append(dict(xy=xy, desc=d, size=sz, use=(rng.random(n0) > 0.1).astype(np.uint8), T_WC=Tw, T_CW=Tc))
    matched = (rng.random(n1) < premated).astype(np.uint8)
    return dict(cur=dict(xy=xy1, desc=d1, size=s1, matched=matched), views=views, T_WC1=T_WC1, T_CW1=T_CW1,
                intr=np.array([fu, fv, cu, cv, *k]), W=W, H=H)
