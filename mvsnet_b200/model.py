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
FLAGS = types.SimpleNamespace(view_num=None, batch_size=1, height=None, width=None, reuse_vars=False,
                              precision="bf16")

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
    tower = mvsnetworks.UNetDS2GN({"data": images.reshape(b * n, h, w, 3)}, mode=network_mode, reuse=True)
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
