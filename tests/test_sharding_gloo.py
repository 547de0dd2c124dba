"""CPU suite: the camera-sharded multi-GPU path with world_size 2 over gloo. Every rank "detects" its cameras (the oracle
stands in for the kernels here: this test is about the sharding, the block format and the all-gather), contributes its
feature blocks to one all-gather and matches the stereo pairs it owns; the union must equal the single-process result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

N_CAMS, B, CAP, W, H = 3, 2, 256, 320, 240
OVERLAPS = [(0, 1), (0, 2), (1, 2)]


def _features(cam, frame):
    import oracle
    from okvis2_b200.synth import synth_frame
    img = synth_frame(900 + frame, W, H, right=(cam == 1), t=cam)
    return oracle.Brisk(30, 1).detect_and_compute(img, CAP)


def _match(f0, f1):
    import oracle
    from okvis2_b200.synth import stereo_scene
    (kp0, d0), (kp1, d1) = f0, f1
    s = stereo_scene(1, len(kp0), max(len(kp1), 1))   # geometry only: rays/poses of the right sizes
    return oracle.match_stereo(d0, s["valid0"], s["e0_W"], s["sof0"], d1, s["valid1"][:len(kp1)], s["e1_W"][:len(kp1)],
                               s["sof1"][:len(kp1)], s["r_WC0"], s["r_WC1"], s["T_CW0"], s["T_CW1"], 60)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from okvis2_b200 import sharding as sh
    from okvis2_b200.lib import KP_DTYPE
    slots = sh.slots_per_rank(world, N_CAMS)
    total = sh.block_layout(B, CAP)[3]
    local = torch.zeros((slots, total), dtype=torch.uint8)
    for cam in sh.cameras_of(rank, world, N_CAMS):
        feats = [_features(cam, b) for b in range(B)]
        sh.pack_block(local[cam // world].numpy(), B, CAP, [len(f[0]) for f in feats], [f[0] for f in feats], [f[1] for f in feats])
    gathered = sh.all_gather_blocks(local, world)
    res = {}
    for (i, j) in sh.pairs_of(rank, world, OVERLAPS):
        ri, si = sh.slot_of(i, world); rj, sj = sh.slot_of(j, world)
        ci, ki, di = sh.unpack_block(gathered[ri, si].numpy(), B, CAP, KP_DTYPE)
        cj, kj, dj = sh.unpack_block(gathered[rj, sj].numpy(), B, CAP, KP_DTYPE)
        for b in range(B):
            res[f"{i}_{j}_{b}"] = _match((ki[b], di[b]), (kj[b], dj[b]))[0]
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **res)
    dist.barrier()
    dist.destroy_process_group()


def test_camera_sharding_world2(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from okvis2_b200 import sharding as sh
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = {}
    for r in range(2):
        z = np.load(tmp_path / f"rank{r}.npz")
        assert not (set(z.files) & set(got)), "a pair was matched on two ranks"
        got.update({k: z[k] for k in z.files})
    assert sorted(got) == sorted(f"{i}_{j}_{b}" for (i, j) in OVERLAPS for b in range(B))   # every pair exactly once
    feats = {(c, b): _features(c, b) for c in range(N_CAMS) for b in range(B)}
    for (i, j) in OVERLAPS:
        for b in range(B):
            assert np.array_equal(got[f"{i}_{j}_{b}"], _match(feats[(i, b)], feats[(j, b)])[0]), (i, j, b)
    # ownership rules
    assert sh.cameras_of(0, 2, 5) == [0, 2, 4] and sh.cameras_of(1, 2, 5) == [1, 3]
    assert sh.pairs_of(1, 2, [(0, 1), (1, 3), (1, 4), (0, 3)]) == [(1, 3), (1, 4)]
    assert sh.block_layout(4, 1024)[3] == 256 + 4 * 1024 * 92


def test_block_roundtrip():
    from okvis2_b200 import sharding as sh
    from okvis2_b200.lib import KP_DTYPE
    rng = np.random.default_rng(0)
    counts = [5, 0, 17]
    kps = [np.frombuffer(rng.bytes(28 * n), KP_DTYPE) for n in counts]
    descs = [rng.integers(0, 256, (n, 64), dtype=np.uint8) for n in counts]
    blk = np.zeros(sh.block_layout(3, 32)[3], np.uint8)
    sh.pack_block(blk, 3, 32, counts, kps, descs)
    c, k, d = sh.unpack_block(blk, 3, 32, KP_DTYPE)
    assert list(c) == counts
    for a, b in zip(kps, k):
        assert a.tobytes() == b.tobytes()
    for a, b in zip(descs, d):
        assert np.array_equal(a, b)
