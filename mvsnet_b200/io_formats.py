"""On-disk formats at the two edges of the hot path, with the reference's names and byte
layout, so that the path can be fed from a dense folder and can feed the unchanged depthfusion.py:

  load_cam / write_cam      camera text files      mvsnet/preprocess.py:116-155, 273-292
  load_pfm / write_pfm      PFM depth / prob maps  mvsnet/preprocess.py:294-356
  write_depth_map / write_confidence_map  16-bit PNG arrays  preprocess.py:253-270 (array conversion only)
  probability_filter        depthfusion.py:172-191 (the array rule and the folder walk)
  write_output_slice        predictlib.py:105-159 (the files depthfusion.py reads: *_init.pfm, *_prob.pfm, *.txt)

Pure host code (NumPy): nothing here touches the GPU.  The reference goes through tf.gfile (`file_io.FileIO`) and
cv2.flip; plain files and np.flipud are byte-identical for local paths.
"""
from __future__ import annotations

import glob
import os
import re
import sys

import numpy as np


def load_cam(file, interval_scale=1, max_d=None):
    """Camera text file -> [2,4,4] float64 (cam[0] = extrinsic, cam[1][:3,:3] = K, cam[1][3] = depth_min, interval,
    depth_num, depth_max).  preprocess.py:116-155; the 29-word form needs max_d (the reference reads FLAGS.max_d)."""
    cam = np.zeros((2, 4, 4))
    words = file.read().split()
    for i in range(4):
        for j in range(4):
            cam[0][i][j] = words[4 * i + j + 1]
    for i in range(3):
        for j in range(3):
            cam[1][i][j] = words[3 * i + j + 18]
    if len(words) == 29:
        if max_d is None:
            raise ValueError("load_cam: a 29-word camera file needs max_d (FLAGS.max_d in the reference)")
        cam[1][3][0] = words[27]
        cam[1][3][1] = float(words[28]) * interval_scale
        cam[1][3][2] = max_d
        cam[1][3][3] = cam[1][3][0] + cam[1][3][1] * cam[1][3][2]
    elif len(words) == 30:
        cam[1][3][0] = words[27]
        cam[1][3][1] = float(words[28]) * interval_scale
        cam[1][3][2] = words[29]
        cam[1][3][3] = cam[1][3][0] + cam[1][3][1] * cam[1][3][2]
    elif len(words) == 31:
        cam[1][3][0] = words[27]
        cam[1][3][1] = float(words[28]) * interval_scale
        cam[1][3][2] = words[29]
        cam[1][3][3] = words[30]
    return cam


def load_cam_from_path(path, interval_scale=1.0, max_d=None):
    with open(path) as f:
        return load_cam(f, interval_scale, max_d)


def write_cam(file, cam):
    """preprocess.py:273-292: 'extrinsic' block, 'intrinsic' block, then depth_min interval depth_num depth_max."""
    with open(file, "w") as f:
        f.write("extrinsic\n")
        for i in range(4):
            for j in range(4):
                f.write(str(cam[0][i][j]) + " ")
            f.write("\n")
        f.write("\n")
        f.write("intrinsic\n")
        for i in range(3):
            for j in range(3):
                f.write(str(cam[1][i][j]) + " ")
            f.write("\n")
        f.write("\n" + str(cam[1][3][0]) + " " + str(cam[1][3][1]) + " " + str(cam[1][3][2]) + " " + str(cam[1][3][3])
                + "\n")


def load_pfm(file):
    """PFM -> float32 array, rows top to bottom (preprocess.py:294-325).  `file` is opened in binary mode."""
    header = file.readline().decode("latin-1").rstrip()
    if header == "PF":
        color = True
    elif header == "Pf":
        color = False
    else:
        raise Exception("Not a PFM file.")
    dim_match = re.match(r"^(\d+)\s(\d+)\s$", file.readline().decode("latin-1"))
    if not dim_match:
        raise Exception("Malformed PFM header.")
    width, height = map(int, dim_match.groups())
    scale = float(file.readline().decode("latin-1").rstrip())
    data_type = "<f" if scale < 0 else ">f"           # negative scale = little-endian
    data = np.frombuffer(file.read(), data_type)
    data = np.reshape(data, (height, width, 3) if color else (height, width))
    return np.flipud(data).astype(np.float32)


def write_pfm(file, image, scale=1):
    """float32 [H,W], [H,W,1] or [H,W,3] -> PFM, rows bottom to top, scale sign = byte order (preprocess.py:327-356)."""
    if image.dtype.name != "float32":
        raise Exception("Image dtype must be float32.")
    image = np.flipud(image)
    if len(image.shape) == 3 and image.shape[2] == 3:
        color = True
    elif len(image.shape) == 2 or (len(image.shape) == 3 and image.shape[2] == 1):
        color = False
    else:
        raise Exception("Image must have H x W x 3, H x W x 1 or H x W dimensions.")
    endian = image.dtype.byteorder
    if endian == "<" or (endian == "=" and sys.byteorder == "little"):
        scale = -scale
    with open(file, "wb") as f:
        f.write(b"PF\n" if color else b"Pf\n")
        f.write(("%d %d\n" % (image.shape[1], image.shape[0])).encode())
        f.write(("%f\n" % scale).encode())
        f.write(np.ascontiguousarray(image).tobytes())


def depth_map_to_uint16(image):
    """preprocess.py:255: clip to [0, 65535] and truncate."""
    return np.clip(image, 0, 65535).astype(np.uint16)


def confidence_map_to_uint16(image):
    """preprocess.py:267-269: probabilities [0,1] -> [0, 65535]."""
    return np.clip(np.asarray(image, dtype=np.float32) * 65535, 0, 65535).astype(np.uint16)


def filter_depth_by_probability(depth_map, prob_map, prob_threshold):
    """depthfusion.py:188: depth_map[prob_map < prob_threshold] = 0 (on a copy)."""
    out = np.array(depth_map, copy=True)
    out[np.asarray(prob_map) < prob_threshold] = 0
    return out


def probability_filter(dense_folder, prob_threshold):
    """depthfusion.py:172-191: for every <prefix>.jpg in depths_mvsnet/, <prefix>_init.pfm + <prefix>_prob.pfm ->
    <prefix>_prob_filtered.pfm."""
    depth_folder = os.path.join(dense_folder, "depths_mvsnet")
    for image_path in glob.glob(os.path.join(depth_folder, "*.jpg")):
        prefix = os.path.splitext(os.path.basename(image_path))[0]
        with open(os.path.join(depth_folder, prefix + "_init.pfm"), "rb") as f:
            depth_map = load_pfm(f)
        with open(os.path.join(depth_folder, prefix + "_prob.pfm"), "rb") as f:
            prob_map = load_pfm(f)
        write_pfm(os.path.join(depth_folder, prefix + "_prob_filtered.pfm"),
                  filter_depth_by_probability(depth_map, prob_map, prob_threshold))


def write_output_slice(output_dir, out_depth_map, out_prob_map, out_ref_cam, out_index):
    """The files depthfusion.py consumes for one reference view (predictlib.py:105-159 without the image / PNG
    side outputs): <index>_init.pfm, <index>_prob.pfm, <index>.txt."""
    os.makedirs(output_dir, exist_ok=True)
    depth = np.squeeze(np.asarray(out_depth_map, dtype=np.float32))
    prob = np.squeeze(np.asarray(out_prob_map, dtype=np.float32))
    write_pfm(os.path.join(output_dir, "{}_init.pfm".format(out_index)), depth)
    write_pfm(os.path.join(output_dir, "{}_prob.pfm".format(out_index)), prob)
    write_cam(os.path.join(output_dir, "{}.txt".format(out_index)), np.squeeze(np.asarray(out_ref_cam)))
