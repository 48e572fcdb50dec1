"""Drop-in for mvsnet/homography_warping.py: same names, argument order and tensor conventions,
CUDA torch tensors instead of TF tensors, arithmetic in libmvsnet_b200.so (no TF, no CPU path)."""
from __future__ import annotations

import torch

from . import ops


def _scalar(v, b=0) -> float:
    if isinstance(v, torch.Tensor):
        return float(v.reshape(-1)[b].item())
    try:
        return float(v[b])
    except (TypeError, IndexError):
        return float(v)


def get_homographies(left_cam, right_cam, depth_num, depth_start, depth_interval, batch_index=0):
    """homography_warping.py:10-58.  cams [B,2,4,4] -> [B,D,3,3]; `batch_index` is unused upstream too."""
    B = left_cam.shape[0]
    outs = []
    for b in range(B):
        cams = torch.stack([left_cam[b], right_cam[b]], dim=0)
        outs.append(ops.homographies(cams, int(depth_num), _scalar(depth_start, b), _scalar(depth_interval, b))[0])
    return torch.stack(outs, dim=0)


def get_homographies_inv_depth(left_cam, right_cam, depth_num, depth_start, depth_end):
    """homography_warping.py:60-106 (batch size 1 only, as upstream :94)."""
    if left_cam.shape[0] != 1:
        raise ValueError("get_homographies_inv_depth supports batch size 1 only (homography_warping.py:94)")
    cams = torch.stack([left_cam[0], right_cam[0]], dim=0)
    H = ops.homographies(cams, int(depth_num), _scalar(depth_start), _scalar(depth_end), inverse_depth=True)
    return H[0][None]


def get_pixel_grids(height, width):
    """homography_warping.py:108-117 -> [3*H*W]."""
    return ops.pixel_grids(int(height), int(width))


def interpolate(image, x, y):
    """homography_warping.py:131-174 -> [B*H*W, C]."""
    return ops.interpolate(image, x, y)


def homography_warping(input_image, homography):
    """homography_warping.py:176-210 (legacy clamp sampler)."""
    return ops.warp(input_image, homography.reshape(-1, 3, 3), sampler="legacy")


def tf_transform_homography(input_image, homography):
    """homography_warping.py:211-253 (tf.contrib.image.transform, BILINEAR, zero fill)."""
    return ops.warp(input_image, homography.reshape(-1, 3, 3), sampler="transform")
