"""On-disk formats at the two edges of the hot path: what the unchanged data pipeline hands to
the path and what the unchanged depthfusion.py reads back.  The byte layouts are the reference's; the code is not.

  in    load_cam / load_cam_from_path     MVSNet camera text files                 preprocess.py:116-155
        camera_from_json                  session camera JSON -> cam [2,4,4]      mvs_cluster.py:91-127
        load_covisibility / Cluster       covisibility.json -> view index lists   cluster_generator.py:139-156,
                                                                                  mvs_cluster.py:129-140
        load_pfm                          PFM maps                                 preprocess.py:294-325
  out   write_cam, write_pfm              camera text, PFM                         preprocess.py:273-292, 327-356
        write_png16, write_depth_map, write_confidence_map   16-bit PNG           preprocess.py:253-270
        write_reference_image             <index>.jpg                              preprocess.py:208-212
        write_output_slice                everything predictlib.py:105-159 writes for one reference view
        probability_filter                depthfusion.py:172-191

Host code only (NumPy + the standard library; OpenCV for the JPEG, which the reference also uses): nothing here touches
the GPU.  The reference goes through tf.gfile (`file_io.FileIO`) and cv2.flip; plain files and a reversed row order
are byte-identical for local paths.
"""
from __future__ import annotations

import glob
import json
import os
import struct
import sys
import zlib

import numpy as np

_CAM_EXTRINSIC = slice(1, 17)        # words after the 'extrinsic' tag
_CAM_INTRINSIC = slice(18, 27)       # words after the 'intrinsic' tag


# ------------------------------------------------------------------------------------------------ cameras
def load_cam(file, interval_scale=1, max_d=None):
    """Camera text file -> cam [2,4,4] float64: cam[0] = extrinsic, cam[1][:3,:3] = K, cam[1][3] = (depth_min,
    depth_interval * interval_scale, depth_num, depth_max).  The tail of the file holds 2, 3 or 4 depth words
    (preprocess.py:132-152); with only (depth_min, interval) the plane count is `max_d` (FLAGS.max_d upstream)."""
    words = file.read().split()
    cam = np.zeros((2, 4, 4))
    cam[0] = np.array(words[_CAM_EXTRINSIC], dtype=np.float64).reshape(4, 4)
    cam[1, :3, :3] = np.array(words[_CAM_INTRINSIC], dtype=np.float64).reshape(3, 3)
    tail = [float(w) for w in words[27:]]
    if len(tail) not in (2, 3, 4):
        return cam                                     # the reference leaves the depth row at zero in this case
    if len(tail) == 2 and max_d is None:
        raise ValueError("load_cam: a camera file with only (depth_min, interval) needs max_d (FLAGS.max_d upstream)")
    depth_min, interval = tail[0], tail[1] * interval_scale
    depth_num = float(max_d) if len(tail) == 2 else tail[2]
    depth_max = tail[3] if len(tail) == 4 else depth_min + interval * depth_num
    cam[1, 3] = (depth_min, interval, depth_num, depth_max)
    return cam


def load_cam_from_path(path, interval_scale=1.0, max_d=None):
    with open(path) as f:
        return load_cam(f, interval_scale, max_d)


def write_cam(file, cam):
    """cam [2,4,4] -> text: an 'extrinsic' block of four rows, an 'intrinsic' block of three, then the depth row;
    every number is str(value) followed by one blank (preprocess.py:273-292)."""
    cam = np.asarray(cam)

    def rows(block):
        return "".join("".join(str(v) + " " for v in row) + "\n" for row in block)

    text = ("extrinsic\n" + rows(cam[0]) + "\n" + "intrinsic\n" + rows(cam[1][:3, :3]) + "\n" +
            " ".join(str(v) for v in cam[1][3]) + "\n")
    with open(file, "w") as f:
        f.write(text)


def camera_from_json(camera_data, min_depth, max_depth, depth_num, interval_scale=1.0):
    """One session camera JSON ({'pose': {'matrix': {'i,j': v}}, 'intrinsics': {fx, fy, px, py}}) -> cam [2,4,4] in the
    layout the path reads (mvs_cluster.py:91-127): the pose translation goes from metres to millimetres, the depth
    interval is (max - min) / (depth_num - 1) * interval_scale."""
    m = camera_data["pose"]["matrix"]
    k = camera_data["intrinsics"]
    cam = np.zeros((2, 4, 4))
    cam[0] = np.array([[m["{},{}".format(i, j)] for j in range(4)] for i in range(4)], dtype=np.float64)
    cam[0, :3, 3] *= 1000.0
    cam[1, :3, :3] = [[k["fx"], 0.0, k["px"]], [0.0, k["fy"], k["py"]], [0.0, 0.0, 1.0]]
    interval = (max_depth - min_depth) / (depth_num - 1) * interval_scale
    cam[1, 3] = (min_depth, interval, depth_num, max_depth)
    return cam


class Cluster:
    """One reference view and its covisible views (mvs_cluster.py:27-140): `indices` always has view_num entries, the
    reference first, padded with copies of the reference when too few views are covisible (:129-140)."""

    def __init__(self, session_dir, ref_index, views, min_depth, max_depth, view_num, depth_num=256, interval_scale=1.0):
        self.session_dir = session_dir
        self.ref_index = int(ref_index)
        self.views = [int(v) for v in views]
        self.min_depth, self.max_depth = min_depth, max_depth
        self.view_num, self.depth_num, self.interval_scale = int(view_num), int(depth_num), interval_scale
        picked = [self.ref_index] + self.views
        self.indices = (picked + [self.ref_index] * max(0, self.view_num - len(picked)))[:self.view_num]

    def camera_path(self, index):
        return os.path.join(self.session_dir, "cameras", "{}.json".format(index))

    def image_path(self, index):
        return os.path.join(self.session_dir, "images", "{}.jpg".format(index))

    def load_camera(self, index):
        with open(self.camera_path(index)) as f:
            return camera_from_json(json.load(f), self.min_depth, self.max_depth, self.depth_num, self.interval_scale)

    def cameras(self):
        """cams [view_num,2,4,4] float32, the reference view first: the `cams` argument of the path (before
        scale_camera, mvs_data_generation/utils.py:64-73)."""
        return np.stack([self.load_camera(i) for i in self.indices]).astype(np.float32)


def load_covisibility(session_dir, view_num, depth_num=256, interval_scale=1.0, include_empty=False, max_clusters=None):
    """<session_dir>/covisibility.json ({ref: {'views': [...], 'min_depth': .., 'max_depth': ..}}) -> [Cluster], in file
    order; reference views without covisible views are skipped unless include_empty (cluster_generator.py:139-156)."""
    with open(os.path.join(session_dir, "covisibility.json")) as f:
        data = json.load(f)
    limit = len(data) if max_clusters is None else max_clusters
    clusters = []
    for ref, entry in data.items():
        if len(clusters) >= limit:
            break
        if entry["views"] or include_empty:
            clusters.append(Cluster(session_dir, int(ref), entry["views"], entry["min_depth"], entry["max_depth"],
                                    view_num, depth_num, interval_scale))
    return clusters


# ------------------------------------------------------------------------------------------------ PFM
def load_pfm(file):
    """PFM (binary file object) -> float32 [H,W] or [H,W,3], first row on top.  Header: 'Pf' / 'PF', 'width height',
    scale (negative = little-endian); rows are stored bottom-up (preprocess.py:294-325)."""
    tag = file.readline().decode("latin-1").rstrip()
    if tag not in ("Pf", "PF"):
        raise Exception("Not a PFM file.")
    dims = file.readline().decode("latin-1").split()
    if len(dims) != 2 or not all(d.isdigit() for d in dims):
        raise Exception("Malformed PFM header.")
    width, height = int(dims[0]), int(dims[1])
    little = float(file.readline().decode("latin-1").rstrip()) < 0
    shape = (height, width, 3) if tag == "PF" else (height, width)
    data = np.frombuffer(file.read(), dtype="<f4" if little else ">f4").reshape(shape)
    return data[::-1].astype(np.float32)


def write_pfm(file, image, scale=1):
    """float32 [H,W], [H,W,1] or [H,W,3] -> PFM: rows bottom-up, native byte order, sign of the scale = byte order
    (preprocess.py:327-356)."""
    image = np.asarray(image)
    if image.dtype != np.float32:
        raise Exception("Image dtype must be float32.")
    if not (image.ndim == 2 or (image.ndim == 3 and image.shape[2] in (1, 3))):
        raise Exception("Image must have H x W x 3, H x W x 1 or H x W dimensions.")
    color = image.ndim == 3 and image.shape[2] == 3
    little = image.dtype.byteorder == "<" or (image.dtype.byteorder in "=|" and sys.byteorder == "little")
    header = "%s\n%d %d\n%f\n" % ("PF" if color else "Pf", image.shape[1], image.shape[0], -scale if little else scale)
    with open(file, "wb") as f:
        f.write(header.encode())
        f.write(np.ascontiguousarray(image[::-1]).tobytes())


# ------------------------------------------------------------------------------------------------ PNG / JPEG
def depth_map_to_uint16(image):
    """preprocess.py:255: clip to [0, 65535] and truncate."""
    return np.clip(image, 0, 65535).astype(np.uint16)


def confidence_map_to_uint16(image):
    """preprocess.py:267-269: probabilities [0,1] -> [0, 65535]."""
    return np.clip(np.asarray(image, dtype=np.float32) * 65535, 0, 65535).astype(np.uint16)


def write_png16(file_path, image):
    """uint16 [H,W] -> 16-bit greyscale PNG (what imageio.imsave writes for a uint16 array upstream; the pixel values are
    what depthfusion-side tools read back, the compression level is not part of the format)."""
    image = np.asarray(image)
    if image.dtype != np.uint16 or image.ndim != 2:
        raise ValueError("write_png16 takes a uint16 [H,W] array")
    h, w = image.shape
    rows = np.concatenate([np.zeros((h, 1), dtype=np.uint8), image.astype(">u2").view(np.uint8).reshape(h, 2 * w)], axis=1)

    def chunk(kind, payload):
        return struct.pack(">I", len(payload)) + kind + payload + struct.pack(">I", zlib.crc32(kind + payload) & 0xFFFFFFFF)

    with open(file_path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 16, 0, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(rows.tobytes(), 6)) + chunk(b"IEND", b""))


def read_png16(file_path):
    """16-bit greyscale PNG with filter type 0 rows (what write_png16 produces) -> uint16 [H,W] (tests, tools)."""
    raw = open(file_path, "rb").read()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w, h = 8, b"", 0, 0
    while pos < len(raw):
        n, kind = struct.unpack(">I", raw[pos:pos + 4])[0], raw[pos + 4:pos + 8]
        body = raw[pos + 8:pos + 8 + n]
        if kind == b"IHDR":
            w, h, depth, color = struct.unpack(">IIBB", body[:10])
            assert depth == 16 and color == 0
        elif kind == b"IDAT":
            idat += body
        pos += 12 + n
    rows = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h, 1 + 2 * w)
    assert not rows[:, 0].any(), "only filter type 0 rows are supported"
    return rows[:, 1:].copy().view(">u2").astype(np.uint16).reshape(h, w)


def write_depth_map(file_path, image):
    """preprocess.py:253-256 (without the optional inverse-depth visualisation)."""
    write_png16(file_path, depth_map_to_uint16(image))


def write_confidence_map(file_path, image):
    """preprocess.py:261-270."""
    write_png16(file_path, confidence_map_to_uint16(image))


def write_reference_image(image, file_path):
    """[H,W,3] image -> JPEG (preprocess.py:208-212: the channels are swapped before scipy.misc.imsave, which scales
    a float image to the full 8-bit range; cv2.imwrite takes the swapped order directly)."""
    import cv2
    img = np.asarray(image)
    if img.dtype != np.uint8:
        lo, hi = float(img.min()), float(img.max())
        img = ((img - lo) * (255.0 / (hi - lo) if hi > lo else 0.0)).astype(np.uint8)      # scipy.misc.bytescale
    if not cv2.imwrite(file_path, img):
        raise IOError("could not write " + file_path)


# ------------------------------------------------------------------------------------------------ after the path
def filter_depth_by_probability(depth_map, prob_map, prob_threshold):
    """depthfusion.py:188: depth_map[prob_map < prob_threshold] = 0 (on a copy)."""
    out = np.array(depth_map, copy=True)
    out[np.asarray(prob_map) < prob_threshold] = 0
    return out


def probability_filter(dense_folder, prob_threshold):
    """depthfusion.py:172-191: for every <prefix>.jpg in depths_mvsnet/, <prefix>_init.pfm + <prefix>_prob.pfm ->
    <prefix>_prob_filtered.pfm.  A folder without any <prefix>.jpg is an error here (upstream silently fuses nothing)."""
    depth_folder = os.path.join(dense_folder, "depths_mvsnet")
    images = sorted(glob.glob(os.path.join(depth_folder, "*.jpg")))
    if not images:
        raise FileNotFoundError("probability_filter: no <index>.jpg in " + depth_folder +
                                " (write_output_slice writes it when given the reference image)")
    for image_path in images:
        prefix = os.path.splitext(os.path.basename(image_path))[0]
        with open(os.path.join(depth_folder, prefix + "_init.pfm"), "rb") as f:
            depth_map = load_pfm(f)
        with open(os.path.join(depth_folder, prefix + "_prob.pfm"), "rb") as f:
            prob_map = load_pfm(f)
        write_pfm(os.path.join(depth_folder, prefix + "_prob_filtered.pfm"),
                  filter_depth_by_probability(depth_map, prob_map, prob_threshold))


def write_output_slice(output_dir, out_depth_map, out_prob_map, out_ref_cam, out_index, out_ref_image=None):
    """Everything predictlib.py:105-159 writes for one reference view: <index>_init.pfm, <index>_prob.pfm, <index>.txt,
    <index>_depth.png, <index>_prob.png and -- when the reference image is given -- <index>.jpg, which is how
    depthfusion.py finds the view (it globs depths_mvsnet/*.jpg)."""
    os.makedirs(output_dir, exist_ok=True)
    depth = np.squeeze(np.asarray(out_depth_map, dtype=np.float32))
    prob = np.squeeze(np.asarray(out_prob_map, dtype=np.float32))

    def path(suffix):
        return os.path.join(output_dir, "{}{}".format(out_index, suffix))

    write_pfm(path("_init.pfm"), depth)
    write_pfm(path("_prob.pfm"), prob)
    write_depth_map(path("_depth.png"), depth)
    write_confidence_map(path("_prob.png"), prob)
    write_cam(path(".txt"), np.squeeze(np.asarray(out_ref_cam)))
    if out_ref_image is not None:
        write_reference_image(np.squeeze(np.asarray(out_ref_image)), path(".jpg"))
