"""EuRoC-layout dataset plumbing (host side; no arithmetic of the hot path).

The layout okvis::DatasetReader consumes (reference okvis_multisensor_processing/src/DatasetReader.cpp:97 `imu0/data.csv`,
:151-183 `readCameraImageCsv`: `<folder><camIdx>/data.csv` with one header line, then `timestamp,filename` rows -- a leading
blank of the file name and a trailing '\\r' are tolerated -- and the images under `<folder><camIdx>/data/`):

    cam0/data.csv   cam0/data/<t_ns>.png   cam1/data.csv   cam1/data/<t_ns>.png   imu0/data.csv

BASELINE config 1 (EuRoC MH_01 through okvis_app_synchronous) cannot run here (no dataset, no app dependencies); its
substitute is a synthetic sequence written in this layout, read back by `EurocReader` and pushed through the front-end
(tests/test_dataset_plumbing.py). Images are 8-bit grayscale PNGs written / read with the standard library only
(zlib + struct; all five PNG filter types are decoded, non-interlaced).
"""
import os
import struct
import zlib

import numpy as np

_SIG = b"\x89PNG\r\n\x1a\n"


def _chunk(tag, data):
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def write_png_gray8(path, img):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    raw = np.zeros((h, w + 1), np.uint8); raw[:, 1:] = img          # filter type 0 on every row
    with open(path, "wb") as f:
        f.write(_SIG + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 0, 0, 0, 0)) +
                _chunk(b"IDAT", zlib.compress(raw.tobytes(), 6)) + _chunk(b"IEND", b""))


def read_png_gray8(path):
    data = open(path, "rb").read()
    if data[:8] != _SIG:
        raise ValueError(f"{path}: not a PNG file")
    pos, idat, w = 8, [], None
    while pos < len(data):
        n, tag = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]; pos += 12 + n
        if tag == b"IHDR":
            w, h, depth, ctype, _, _, interlace = struct.unpack(">IIBBBBB", body)
            if depth != 8 or ctype != 0 or interlace != 0:
                raise ValueError(f"{path}: only 8-bit grayscale non-interlaced PNGs are supported")
        elif tag == b"IDAT":
            idat.append(body)
        elif tag == b"IEND":
            break
    raw = np.frombuffer(zlib.decompress(b"".join(idat)), np.uint8).reshape(h, w + 1)
    out = np.zeros((h, w), np.uint8)
    prev = np.zeros(w, np.int32)
    for y in range(h):
        ft, line = int(raw[y, 0]), raw[y, 1:].astype(np.int32)
        if ft == 0:
            cur = line
        elif ft == 2:
            cur = (line + prev) & 255
        else:                                   # 1 (sub), 3 (average), 4 (Paeth): left neighbour dependent
            cur = np.zeros(w, np.int32)
            left = upleft = 0
            for x in range(w):
                up = int(prev[x])
                if ft == 1:
                    p = left
                elif ft == 3:
                    p = (left + up) >> 1
                elif ft == 4:
                    pa, pb, pc = abs(up - upleft), abs(left - upleft), abs(left + up - 2 * upleft)
                    p = left if (pa <= pb and pa <= pc) else (up if pb <= pc else upleft)
                else:
                    raise ValueError(f"{path}: bad filter type {ft}")
                left = (int(line[x]) + p) & 255; cur[x] = left; upleft = up
        out[y] = cur; prev = cur
    return out


CAM_HEADER = "#timestamp [ns],filename"
IMU_HEADER = ("#timestamp [ns],w_RS_S_x [rad s^-1],w_RS_S_y [rad s^-1],w_RS_S_z [rad s^-1],"
              "a_RS_S_x [m s^-2],a_RS_S_y [m s^-2],a_RS_S_z [m s^-2]")


def write_euroc(path, cams, timestamps_ns, imu=None):
    """cams: per camera an (n, H, W) u8 array; timestamps_ns: n integers; imu: (m, 7) rows [t_ns, gyro xyz, acc xyz]."""
    for c, frames in enumerate(cams):
        d = os.path.join(path, f"cam{c}", "data"); os.makedirs(d, exist_ok=True)
        with open(os.path.join(path, f"cam{c}", "data.csv"), "w") as f:
            f.write(CAM_HEADER + "\n")
            for t, img in zip(timestamps_ns, frames):
                write_png_gray8(os.path.join(d, f"{int(t)}.png"), img)
                f.write(f"{int(t)},{int(t)}.png\n")
    os.makedirs(os.path.join(path, "imu0"), exist_ok=True)
    with open(os.path.join(path, "imu0", "data.csv"), "w") as f:
        f.write(IMU_HEADER + "\n")
        for row in ([] if imu is None else imu):
            f.write(f"{int(row[0])}," + ",".join(repr(float(v)) for v in row[1:]) + "\n")


def read_camera_image_csv(path, folder, cam_idx):
    """DatasetReader::readCameraImageCsv: [(timestamp string, image path)], or None when the csv is missing."""
    filename = os.path.join(path, f"{folder}{cam_idx}", "data.csv")
    if not os.path.isfile(filename):
        return None
    names = []
    with open(filename, newline="") as f:
        lines = f.read().split("\n")
    for line in lines[1:]:
        if line == "":
            continue
        if "," not in line:
            break
        s0, s1 = line.split(",", 1)
        if s1 == "":
            break
        if s1[0] == " ":
            s1 = s1[1:]
        if s1.endswith("\r"):
            s1 = s1[:-1]
        names.append((s0, os.path.join(path, f"{folder}{cam_idx}", "data", s1)))
    return names


class EurocReader:
    """Synchronised multiframes of an EuRoC-layout directory, in timestamp order (gray `cam<i>` folders)."""

    def __init__(self, path, num_cameras):
        self.path, self.num_cameras = path, num_cameras
        if not os.path.isfile(os.path.join(path, "imu0", "data.csv")):
            raise FileNotFoundError(f"no imu file found at {path}/imu0/data.csv")
        self.names = []
        for c in range(num_cameras):
            n = read_camera_image_csv(path, "cam", c)
            if not n:
                raise FileNotFoundError(f"no images found for camera {c}")
            self.names.append(n)

    def imu(self):
        rows = [l.split(",") for l in open(os.path.join(self.path, "imu0", "data.csv")).read().split("\n")[1:] if l.strip()]
        return np.array([[float(v) for v in r] for r in rows], np.float64).reshape(-1, 7)

    def __len__(self):
        return min(len(n) for n in self.names)

    def __iter__(self):
        for k in range(len(self)):
            ts = [int(self.names[c][k][0]) for c in range(self.num_cameras)]
            yield ts, [read_png_gray8(self.names[c][k][1]) for c in range(self.num_cameras)]
