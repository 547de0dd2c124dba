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
