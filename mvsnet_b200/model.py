"""Drop-in for the 3DCNN part of mvsnet/model.py: inference (:257), inference_mem (:374),
get_probability_map (:20), get_probability_map_slice (:45) with the reference's argument order.

`images` is either the reference's [B,N,H,W,3] (then every view goes through the UNetDS2GN tower of this package,
model.py:392-406, once its variables are registered with `mvsnetworks.set_unet_variables`, or through a caller's own
extractor registered with `set_feature_extractor(fn)`, fn: images [B,H,W,3] -> features [B,H/4,W/4,C]), or the
feature towers themselves, [B,N,Hf,Wf,C] (any last dimension other than 3).
"""
from __future__ import annotations

import types

import torch

from . import ops
from .cnn_wrapper import mvsnetworks
from .engine import HotPath, RegnetWeights, regnet_base_filter

# stand-in for tf.app.flags.FLAGS; the reference reads these inside the path (model.py:28,275,381-382,431)
# precision: arithmetic of the hot path ("bf16" = tensor cores, "fp32" = parity mode); tower_precision: arithmetic of the
# image feature towers when images are given ("fp32" = CUDA-core parity mode, the default: the regularizer's input then
# matches the reference to 1e-5; "bf16" = tensor cores, 4x faster, features within ~2 % rms of the fp32 ones)
FLAGS = types.SimpleNamespace(view_num=None, batch_size=1, height=None, width=None, reuse_vars=False,
                              precision="bf16", tower_precision="fp32")

_feature_extractor = None
_engines = {}


def set_feature_extractor(fn) -> None:
    global _feature_extractor
    _feature_extractor = fn


def _sc(v, b):
    return float(v.reshape(-1)[b].item()) if isinstance(v, torch.Tensor) else float(v[b] if hasattr(v, "__len__") else v)


def get_probability_map_slice(cv, depth_map, depth_start, depth_interval, inverse_depth=False, num_buckets=4):
    """model.py:45-144.  cv [1,D,H,W], depth_map [1,H,W,1] -> [1,H,W,1]."""
    d, h, w = cv.shape[-3:]
    p = ops.probability_map(cv.reshape(d, h, w), depth_map.reshape(h, w), _sc(depth_start, 0),
                            _sc(depth_interval, 0), inverse_depth, num_buckets)
    return p.reshape(1, h, w, 1)


def get_probability_map(cv_batch, depth_map_batch, depth_start_batch, depth_interval_batch, inverse_depth=False,
                        num_buckets=4):
    """model.py:20-39: slice by slice over the batch."""
    outs = []
    for i in range(cv_batch.shape[0]):
        outs.append(get_probability_map_slice(cv_batch[i:i + 1], depth_map_batch[i:i + 1], _sc(depth_start_batch, i),
                                              _sc(depth_interval_batch, i), inverse_depth, num_buckets))
    return torch.cat(outs, dim=0)


def _towers(images, network_mode="normal"):
    if images.shape[-1] != 3:
        return images                                    # already feature towers [B,N,Hf,Wf,C]
    n = images.shape[1]
    if _feature_extractor is not None:
        return torch.stack([_feature_extractor(images[:, v]) for v in range(n)], dim=1)
    if not mvsnetworks._UNET_VARIABLES:
        raise RuntimeError("images given but neither UNetDS2GN variables (mvsnetworks.set_unet_variables) nor a "
                           "feature extractor (model.set_feature_extractor) is registered")
    # one tower per view with shared variables (model.py:392-406); the views of a batch are independent, so they
    # go through the tower together
    b, _, h, w, _ = images.shape
    tower = mvsnetworks.UNetDS2GN({"data": images.reshape(b * n, h, w, 3)}, mode=network_mode, reuse=True,
                                  precision=FLAGS.tower_precision)
    f = tower.get_output()
    return f.reshape(b, n, f.shape[1], f.shape[2], f.shape[3])


def _run(images, cams, depth_num, depth_start, depth_interval, network_mode, inverse_depth, order):
    if not isinstance(depth_num, int):
        raise TypeError("depth_num must be a Python int (model.py:427 iterates range(depth_num))")
    feats = _towers(images, network_mode).to(torch.float32)
    B, N, hf, wf, c = feats.shape
    if FLAGS.view_num is not None and FLAGS.view_num != N:
        raise ValueError(f"FLAGS.view_num={FLAGS.view_num} but {N} views given")
    weights = mvsnetworks.get_variables()
    if not weights:
        raise RuntimeError("RegNetUS0 variables not set: call mvsnetworks.set_variables(weights)")
    # the checkpoint version, not id(weights): set_variables() refills the same dict in place
    key = (N, depth_num, hf, wf, c, network_mode, bool(inverse_depth), order, FLAGS.precision,
           mvsnetworks.variables_version(), feats.device.index)
    eng = _engines.get(key)
    if eng is None:
        w = RegnetWeights(weights, feats.device)
        if w.base_filter != regnet_base_filter(network_mode):
            raise ValueError(f"variables are for base_filter {w.base_filter}, network_mode {network_mode!r} "
                             f"needs {regnet_base_filter(network_mode)}")
        _engines.clear()
        eng = _engines[key] = HotPath(N, depth_num, hf, wf, w, channels=c, precision=FLAGS.precision, order=order,
                                      inverse_depth=inverse_depth, device=feats.device)
    depths, probs = [], []
    for b in range(B):
        d = torch.empty((hf, wf), dtype=torch.float32, device=feats.device)
        p = torch.empty((hf, wf), dtype=torch.float32, device=feats.device)
        eng.infer(feats[b].contiguous(), cams[b].to(torch.float32).contiguous(), _sc(depth_start, b),
                  _sc(depth_interval, b), d, p)
        depths.append(d)
        probs.append(p)
    return torch.stack(depths)[..., None], torch.stack(probs)[..., None]


def inference(images, cams, depth_num, depth_start, depth_interval, network_mode, is_master_gpu=True, trainable=True,
              inverse_depth=False):
    """model.py:257-372 (training-graph variant: variance as mean2 - mean^2, :330-332)."""
    return _run(images, cams, depth_num, depth_start, depth_interval, network_mode, inverse_depth, "train")


def inference_mem(images, cams, depth_num, depth_start, depth_interval, network_mode, is_master_gpu=True,
                  training=True, trainable=True, inverse_depth=False):
    """model.py:374-502 (inference variant: Q/N - S^2/(N*N), :458-461)."""
    if not training:
        raise NotImplementedError("inference_mem(training=False) is never exercised by the reference "
                                  "(predictlib.py:83-84 leaves the default)")
    return _run(images, cams, depth_num, depth_start, depth_interval, network_mode, inverse_depth, "mem")


def _resize_bilinear(x, out_h, out_w, subtract=0.0, multiply=1.0):
    """tf.image.resize_bilinear (TF 1.x, align_corners=False) of x [B,H,W,C], then (y - subtract) * multiply."""
    from . import _lib as L
    x = x.to(torch.float32).contiguous()
    b, h, w, c = x.shape
    y = torch.empty((b, out_h, out_w, c), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        L.require_cuda(x)
        L.check(L.load().mvsb200_resize_bilinear(L.ptr(x), b, h, w, c, L.ptr(y), out_h, out_w, float(subtract),
                                                 float(multiply), L.stream_ptr()), "resize_bilinear")
    return y


def depth_refine(init_depth_map, image, prob_map, depth_num, depth_start, depth_interval, network_mode, network_type,
                 is_master_gpu=True, training=True, trainable=True, upsample_depth=False, refine_with_confidence=False,
                 stereo_image=None, residual_refinement=True):
    """model.py:753-811: refine the depth map with the reference image.  init_depth_map, prob_map [B,Hd,Wd,1]; image
    [B,H,W,3] -> (refined_depth_map, residual_depth_map).  The depth map is normalised to [0,1] over the sweep (:763),
    depth or image are resized with tf.image.resize_bilinear (:766-777), the tower sees concat(image, depth[, prob][,
    stereo image]) and predicts a residual in normalised units, which is scaled back and added (:803-809).
    network_type 'original' = RefineNetConv; 'unet' (RefineUNetConv) is outside this package's scope."""
    import numpy as np

    from . import _lib as L
    if network_type == "unet":
        raise NotImplementedError("depth_refine: network_type 'unet' (RefineUNetConv) is not built; use 'original'")
    if network_type != "original":
        raise NotImplementedError                                                      # model.py:801
    init = init_depth_map.to(torch.float32).contiguous()
    b, hd, wd, _ = init.shape
    if b != 1 and (isinstance(depth_start, torch.Tensor) and depth_start.numel() > 1):
        raise NotImplementedError("depth_refine: one depth range per call (batch_size 1, as predictlib.py:86-91 runs it)")
    ds, di = np.float32(_sc(depth_start, 0)), np.float32(_sc(depth_interval, 0))
    depth_end = np.float32(ds + np.float32(np.float32(depth_num) - np.float32(1.0)) * di)
    depth_scale = np.float32(depth_end - ds)
    inv_scale_note = float(depth_scale)
    if upsample_depth:
        h, w = int(image.shape[1]), int(image.shape[2])
        # (x - start) / scale is a division upstream; resize first, then the same division (tf.div before the resize is
        # the same values up to rounding: both are linear) -- kept in the reference's order: normalise, then resize
        norm = _resize_bilinear((init - float(ds)) / inv_scale_note, h, w)
        init = _resize_bilinear(init, h, w)
        if refine_with_confidence:
            prob_map = _resize_bilinear(prob_map, h, w)
    else:
        norm = (init - float(ds)) / inv_scale_note
        image = _resize_bilinear(image, hd, wd)
        if stereo_image is not None:
            stereo_image = _resize_bilinear(stereo_image, hd, wd)
    data = norm
    if refine_with_confidence:
        data = torch.cat([data, prob_map.to(torch.float32)], dim=3)
    if stereo_image is not None:
        data = torch.cat([data, stereo_image.to(torch.float32)], dim=3)
    tower = mvsnetworks.RefineNetConv({"color_image": image.to(torch.float32), "depth_image": data},
                                      trainable=trainable, training=training, mode=network_mode,
                                      reuse=not is_master_gpu or FLAGS.reuse_vars)
    residual_norm = tower.get_output().contiguous()
    residual = torch.empty_like(residual_norm)
    refined = torch.empty_like(residual_norm)
    with torch.cuda.device(residual_norm.device):
        L.check(L.load().mvsb200_scale_add(L.ptr(residual_norm), float(depth_scale),
                                           L.ptr(init.contiguous()) if residual_refinement else None,
                                           residual_norm.numel(), L.ptr(residual), L.ptr(refined), L.stream_ptr()),
                "scale_add")
    return refined, residual
