"""TEST INFRASTRUCTURE -- CPU restatement of the refinement glue after the hot path:
mvsnet/model.py:753-811 `depth_refine` and the 'original' tower RefineNetConv (cnn_wrapper/mvsnetworks.py:178-193).
Only tests/ may import this.  PARITY UNPINNED like the rest of oracle/: TensorFlow 1.12 cannot run here, so
`tf.image.resize_bilinear` (align_corners=False, the TF 1.x kernel without half-pixel centres) and
`tf.layers.conv2d(use_bias=True)` are restated from their documented semantics.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def resize_bilinear(x, out_h, out_w):
    """tf.image.resize_bilinear (TF 1.x, align_corners=False): x [N,H,W,C] -> [N,out_h,out_w,C].
    in = out * (in_size / out_size); lower = floor(in); upper = min(lower + 1, in_size - 1); lerp = in - lower;
    top = tl + (tr - tl) * xl; bottom = bl + (br - bl) * xl; out = top + (bottom - top) * yl."""
    x = np.asarray(x, dtype=F32)
    n, h, w, c = x.shape
    sy, sx = F32(h) / F32(out_h), F32(w) / F32(out_w)
    fy = (np.arange(out_h, dtype=F32) * sy).astype(F32)
    fx = (np.arange(out_w, dtype=F32) * sx).astype(F32)
    y0, x0 = np.floor(fy).astype(np.int64), np.floor(fx).astype(np.int64)
    y1, x1 = np.minimum(y0 + 1, h - 1), np.minimum(x0 + 1, w - 1)
    ly = (fy - y0.astype(F32)).astype(F32)[None, :, None, None]
    lx = (fx - x0.astype(F32)).astype(F32)[None, None, :, None]
    tl, tr = x[:, y0][:, :, x0], x[:, y0][:, :, x1]
    bl, br = x[:, y1][:, :, x0], x[:, y1][:, :, x1]
    top = (tl + ((tr - tl).astype(F32) * lx).astype(F32)).astype(F32)
    bot = (bl + ((br - bl).astype(F32) * lx).astype(F32)).astype(F32)
    return (top + ((bot - top).astype(F32) * ly).astype(F32)).astype(F32)


def conv2d_bias(x, kernel, bias, relu):
    """tf.layers.conv2d(3x3, SAME, stride 1, use_bias=True [, activation=relu]) (network.py:171-206): x [N,H,W,Cin],
    kernel [3,3,Cin,Cout]."""
    import torch
    import torch.nn.functional as TF
    xt = torch.from_numpy(np.ascontiguousarray(x, dtype=F32)).permute(0, 3, 1, 2)
    wt = torch.from_numpy(np.ascontiguousarray(kernel, dtype=F32)).permute(3, 2, 0, 1)
    y = TF.conv2d(xt, wt, None if bias is None else torch.from_numpy(np.asarray(bias, dtype=F32)), padding=1)
    if relu:
        y = torch.relu(y)
    return y.permute(0, 2, 3, 1).contiguous().numpy()


def refine_net_conv(color_image, depth_image, weights):
    """RefineNetConv (mvsnetworks.py:178-193): concat(color_image, depth_image) -> three biased 3x3 convs with ReLU ->
    a biased 3x3 conv to one channel without ReLU."""
    x = np.concatenate([color_image, depth_image], axis=3).astype(F32)
    for i in range(4):
        x = conv2d_bias(x, weights[f"refine_conv{i}/kernel"], weights[f"refine_conv{i}/bias"], relu=i < 3)
    return x


def depth_refine(init_depth_map, image, prob_map, depth_num, depth_start, depth_interval, weights,
                 upsample_depth=False, refine_with_confidence=False, stereo_image=None, residual_refinement=True):
    """model.py:753-811 with network_type 'original'.  init_depth_map / prob_map [B,Hd,Wd,1], image [B,H,W,3] ->
    (refined_depth_map, residual_depth_map)."""
    init = np.asarray(init_depth_map, dtype=F32)
    image = np.asarray(image, dtype=F32)
    depth_start, depth_interval = F32(depth_start), F32(depth_interval)
    depth_end = F32(depth_start + F32(F32(depth_num) - F32(1.0)) * depth_interval)      # :759
    depth_scale = F32(depth_end - depth_start)
    norm = ((init - depth_start) / depth_scale).astype(F32)                              # :763-764
    if upsample_depth:
        h, w = image.shape[1:3]
        norm, init = resize_bilinear(norm, h, w), resize_bilinear(init, h, w)            # :768-769
        if refine_with_confidence:
            prob_map = resize_bilinear(prob_map, h, w)
    else:
        h, w = init.shape[1:3]
        image = resize_bilinear(image, h, w)                                             # :775
        if stereo_image is not None:
            stereo_image = resize_bilinear(stereo_image, h, w)
    data = norm
    if refine_with_confidence:
        data = np.concatenate([data, np.asarray(prob_map, dtype=F32)], axis=3)           # :782-783
    if stereo_image is not None:
        data = np.concatenate([data, stereo_image], axis=3)
    residual_norm = refine_net_conv(image, data, weights)
    residual = (residual_norm * depth_scale).astype(F32)                                 # :803
    refined = (residual + init).astype(F32) if residual_refinement else residual        # :805-808
    return refined, residual
