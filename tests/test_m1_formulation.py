"""CPU suite: the parallel formulation of M1 that k_m1_rowbin / k_m1_match implement (okvis2_b200/csrc/okb_match.cu),
executed here in numpy, must equal the oracle's transcription of the sequential loop (Frontend.cpp:1515-1590):

  * pool rows of 3-D landmarks are binned by their landmark's projection into cells >= the gate radius, clamped into the
    grid; rows projecting more than radius + 1 px outside the keypoint extent are dropped; NaN projections go to an extra
    cell that every keypoint visits (the reference's `> thr^2 -> skip` does not reject NaN);
  * a keypoint visits the 3x3 cells around its own cell, applies the exact gate, and keeps min (distance << 32 | row).
The arithmetic (gate expression, Hamming) is the oracle's; what is under test is the binning / clamping / pre-filter logic."""
import numpy as np
import pytest

import oracle
from okvis2_b200.synth import map_scene

PC = np.array([bin(i).count("1") for i in range(256)], np.int64)


def m1_formulation(kp_desc, kp_xy, use, cand_desc, cand_lm, lm_proj, lm_is3d, thr_px, thr):
    n_kp = len(kp_xy)
    ok = np.isfinite(kp_xy).all(1) if n_kp else np.zeros(0, bool)
    fin = kp_xy[ok]
    min_x, min_y = (fin[:, 0].min(), fin[:, 1].min()) if len(fin) else (0.0, 0.0)
    max_x, max_y = (fin[:, 0].max(), fin[:, 1].max()) if len(fin) else (0.0, 0.0)
    cell = max(int(np.ceil(thr_px)), 8)
    while True:
        gx, gy = int((max_x - min_x) / cell) + 1, int((max_y - min_y) / cell) + 1
        if gx * gy <= 4096:
            break
        cell *= 2

    def cell_of(x, y):
        fx, fy = np.floor((x - min_x) / cell), np.floor((y - min_y) / cell)
        return int(min(max(fx, 0), gx - 1)), int(min(max(fy, 0), gy - 1))

    bins = {}
    m = thr_px + 1.0
    for r, lm in enumerate(cand_lm):
        if not lm_is3d[lm]:
            continue
        px, py = lm_proj[lm]
        if np.isnan(px) or np.isnan(py):
            bins.setdefault("nan", []).append(r); continue
        if not (px >= min_x - m and px <= max_x + m and py >= min_y - m and py <= max_y + m):
            continue
        bins.setdefault(cell_of(px, py), []).append(r)
    dist = np.full(n_kp, thr, np.uint32); out = np.full(n_kp, -1, np.int32)
    for k in range(n_kp):
        if (use is not None and not use[k]) or not ok[k] or len(cand_lm) == 0:
            continue
        x, y = kp_xy[k]
        cx, cy = cell_of(x, y)
        rows = list(bins.get("nan", []))
        for cyy in range(cy - 1, cy + 2):
            for cxx in range(max(cx - 1, 0), min(cx + 1, gx - 1) + 1):
                if 0 <= cyy < gy:
                    rows += bins.get((cxx, cyy), [])
        best = (int(thr) << 32) | 0xFFFFFFFF
        for r in rows:
            px, py = lm_proj[cand_lm[r]]
            dx, dy = px - x, py - y
            d2 = dx * dx + dy * dy
            if d2 > thr_px * thr_px:
                continue
            d = int(PC[kp_desc[k] ^ cand_desc[r]].sum())
            best = min(best, (d << 32) | r)
        if (best >> 32) < thr:
            dist[k] = best >> 32; out[k] = cand_lm[best & 0xFFFFFFFF]
    return dist, out


@pytest.mark.parametrize("seed,n_kp,n_lm,thr_px", [(1, 300, 900, 20.0), (2, 200, 600, 150.0), (3, 50, 200, 3.0)])
def test_binned_formulation_equals_sequential_loop(seed, n_kp, n_lm, thr_px):
    rng = np.random.default_rng(seed)
    kp_xy = rng.uniform(0, 752, (n_kp, 2)); kp_xy[:, 1] *= 480 / 752
    kd = rng.integers(0, 256, (n_kp, 64), dtype=np.uint8)
    use = (rng.random(n_kp) > 0.1).astype(np.uint8)
    m = map_scene(seed, kp_xy, kd, n_lm, W=752, H=480, frac_near=0.3)
    proj = m["lm_proj"].copy()
    # degenerate projections: NaN (never gated out by the reference), infinities, far outside, exactly on the gate radius
    proj[0] = [np.nan, 10.0]; proj[1] = [np.inf, 5.0]; proj[2] = [-1e9, 3.0]; proj[3] = [5.0, np.nan]
    proj[4] = kp_xy[0] + [thr_px, 0.0]; proj[5] = kp_xy[1] + [0.0, -thr_px]; proj[6] = [kp_xy[:, 0].max() + thr_px, kp_xy[:, 1].max()]
    is3d = m["lm_is3d"].copy(); is3d[:7] = 1
    ref = oracle.match_map3d(kd, kp_xy, use, m["cand_desc"], m["cand_lm"], proj, is3d, thr_px, 60)
    got = m1_formulation(kd, kp_xy, use, m["cand_desc"], m["cand_lm"], proj, is3d, thr_px, 60)
    assert np.array_equal(got[0], ref[0].astype(np.uint32)) and np.array_equal(got[1], ref[1])
    assert (ref[1] >= 0).sum() > (5 if thr_px >= 20 else 0)
