"""Text formats of the descriptor the map files and the vocabulary use (host-side; no arithmetic of the hot path).

* DBoW2::FBrisk::toString / fromString (reference okvis_frontend/src/FBrisk.cpp:71-95): the bytes as decimal integers,
  each followed by a space -- the `descriptor:"..."` strings of resources/small_voc.yml.gz.
* `FRAME:KEYPOINT <stateId> <cameraIdx> <pt.x> <pt.y> <size> BRISK2 <hex>` records of a saved map
  (writer okvis_ceres/src/Component.cpp:449-460, reader :235-258): two lower-case hex digits per descriptor byte. The
  stream was given std::setprecision(17) BEFORE init.copyfmt(file) (Component.cpp:407-411), so the "default format" that
  file.copyfmt(init) restores after every record still carries 17 significant digits: the float fields are promoted to
  double and printed as %.17g, which round-trips every float32 exactly.
"""
import numpy as np

from .lib import KP_DTYPE


def fbrisk_to_string(desc):
    return "".join(f"{int(b)} " for b in np.asarray(desc, np.uint8))


def fbrisk_from_string(s, L=48):
    v = s.split()
    return np.array([int(v[i]) & 0xFF for i in range(L)], np.uint8)


def _g(x):
    return format(float(np.float32(x)), ".17g")   # operator<<(float) -> double, precision 17


def write_frame_keypoints(state_id, camera_idx, keypoints, descriptors):
    """Lines (without newline) for one camera of one multiframe, as Component::save writes them."""
    out = []
    for kp, d in zip(keypoints, np.asarray(descriptors, np.uint8)):
        hexs = "".join(f"{int(b):02x}" for b in d)
        out.append(f"FRAME:KEYPOINT {int(state_id)} {int(camera_idx)} {_g(kp['x'])} {_g(kp['y'])} {_g(kp['size'])} BRISK2 {hexs}")
    return out


def read_frame_keypoints(lines, state_id, camera_idx, D=48):
    """Component::load's sub-loop: consumes consecutive FRAME:KEYPOINT lines, checks ids and the descriptor kind, returns
    (keypoints, N x D descriptors, number of lines consumed)."""
    kps, descs, n = [], [], 0
    for line in lines:
        tok = line.split()
        if not tok or tok[0] != "FRAME:KEYPOINT":
            break
        if tok[6] != "BRISK2":
            raise ValueError(f"descriptor {tok[6]} not supported, only BRISK 2")
        if int(tok[1]) != int(state_id) or int(tok[2]) != int(camera_idx):
            raise ValueError("mismatching keypoint stateId")
        kps.append((float(tok[3]), float(tok[4]), float(tok[5])))
        descs.append([int(tok[7][2 * c:2 * c + 2], 16) for c in range(D)])
        n += 1
    kp = np.zeros(len(kps), KP_DTYPE)
    if kps:
        a = np.array(kps, np.float32)
        kp["x"], kp["y"], kp["size"] = a[:, 0], a[:, 1], a[:, 2]
    return kp, np.array(descs, np.uint8).reshape(len(kps), D), n
