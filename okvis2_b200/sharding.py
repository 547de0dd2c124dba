"""Camera-per-GPU sharding of an NCameraSystem (BASELINE config 4).

detect/describe and the map matchers are independent per camera (the reference already runs one thread per camera,
okvis_multisensor_processing/src/ThreadedSlam.cpp:432-448); only Frontend::matchStereo (Frontend.cpp:1990-2000) couples
cameras, pairwise and only where MultiFrame::hasOverlap(im0, im1). So: camera c lives on rank c % world; once per
multiframe batch every rank contributes the fixed-capacity feature blocks of its cameras to ONE all-gather
(torch.distributed: NCCL on GPUs, gloo in the CPU tests), and the pair (i, j), i < j, is matched on the rank that owns
camera i. Pure host-side bookkeeping: no feature arithmetic happens here.
"""
import numpy as np

KP_BYTES, DESC_BYTES = 28, 64


def camera_owner(cam, world):
    return cam % world


def cameras_of(rank, world, n_cams):
    return [c for c in range(n_cams) if camera_owner(c, world) == rank]


def slots_per_rank(world, n_cams):
    """every rank contributes the same number of camera slots to the all-gather (padded)"""
    return (n_cams + world - 1) // world


def slot_of(cam, world):
    """(rank, slot) of camera `cam` inside the gathered buffer"""
    return camera_owner(cam, world), cam // world


def pairs_of(rank, world, overlaps):
    """stereo pairs (i < j) this rank matches: the owner of the lower camera index"""
    return [(i, j) for (i, j) in sorted(overlaps) if i < j and camera_owner(i, world) == rank]


def block_layout(n_frames, capacity):
    """byte offsets (counts, keypoints, descriptors, total) of one feature block -- mirrors okb_feature_block_bytes"""
    counts = (n_frames * 4 + 255) // 256 * 256
    kp = n_frames * capacity * KP_BYTES
    desc = n_frames * capacity * DESC_BYTES
    return 0, counts, counts + kp, counts + kp + desc


def pack_block(block, n_frames, capacity, counts, kps, descs):
    """host-side packing (CPU tests): block is a uint8 array of block_layout(...)[3] bytes"""
    o_c, o_k, o_d, total = block_layout(n_frames, capacity)
    assert block.nbytes == total
    block[:] = 0
    block[o_c:o_c + 4 * n_frames] = np.asarray(counts, np.int32).view(np.uint8)
    for b in range(n_frames):
        n = int(counts[b])
        block[o_k + b * capacity * KP_BYTES:o_k + b * capacity * KP_BYTES + n * KP_BYTES] = np.ascontiguousarray(kps[b]).view(np.uint8).reshape(-1)
        block[o_d + b * capacity * DESC_BYTES:o_d + b * capacity * DESC_BYTES + n * DESC_BYTES] = np.ascontiguousarray(descs[b]).reshape(-1)
    return block


def unpack_block(block, n_frames, capacity, kp_dtype):
    o_c, o_k, o_d, total = block_layout(n_frames, capacity)
    counts = block[o_c:o_c + 4 * n_frames].view(np.int32).copy()
    kps = block[o_k:o_d].view(kp_dtype).reshape(n_frames, capacity)
    descs = block[o_d:total].reshape(n_frames, capacity, DESC_BYTES)
    return counts, [kps[b, :counts[b]].copy() for b in range(n_frames)], [descs[b, :counts[b]].copy() for b in range(n_frames)]


def all_gather_blocks(local, world, group=None):
    """local: torch uint8 tensor [slots_per_rank, block_bytes] (CPU -> gloo, CUDA -> NCCL). Returns [world, slots, bytes]."""
    import torch
    import torch.distributed as dist
    out = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
    if world == 1:
        out[0].copy_(local)
        return out
    dist.all_gather_into_tensor(out.view(-1), local.contiguous().view(-1), group=group)
    return out


def rig_overlaps(rig, step=16):
    """Approximate MultiFrame::hasOverlap graph of a rig (okvis_cv/src/NCameraSystem.cpp:48-118): camera j overlaps camera
    i if some pixel ray of j (ideal pinhole, rotation only) projects inside i's image. Used to define the bench workload."""
    def R(c):
        return np.array(c["T_SC"]).reshape(4, 4)[:3, :3]
    n = len(rig)
    pairs = set()
    for i in range(n):
        for j in range(n):
            if i == j:
                continue
            W, H = rig[j]["image_dimension"]
            u, v = np.meshgrid(np.arange(0, W, step), np.arange(0, H, step))
            fj, cj = rig[j]["focal_length"], rig[j]["principal_point"]
            rays = np.stack([(u - cj[0]) / fj[0], (v - cj[1]) / fj[1], np.ones_like(u, float)], -1).reshape(-1, 3)
            r_i = rays @ (R(rig[i]).T @ R(rig[j])).T
            fi, ci = rig[i]["focal_length"], rig[i]["principal_point"]
            Wi, Hi = rig[i]["image_dimension"]
            z = r_i[:, 2]
            ok = z > 1e-6
            x = r_i[:, 0] / np.where(ok, z, 1) * fi[0] + ci[0]
            y = r_i[:, 1] / np.where(ok, z, 1) * fi[1] + ci[1]
            if (ok & (x >= 0) & (x < Wi) & (y >= 0) & (y < Hi)).any():
                pairs.add((min(i, j), max(i, j)))
    return sorted(pairs)
