"""tf.py_func shim: the hot path as ONE op inside the reference's TensorFlow 1.12 graph, so that
mvsnet/predictlib.py:79-102 can keep building TF tensors around it.

UNTESTED SOURCE.  TensorFlow 1.12 / Python 2.7 cannot be installed in the image this package is developed and tested in
(SURVEY.md 8c), so nothing in tests/ exercises this file; it is written against the documented `tf.py_func` contract
(numpy arrays in, numpy arrays out, executed in the session's Python process) and against the entry points the tests do
cover (`HotPath.infer_host`, `mvsb200_infer_host` in include/mvsnet_b200.h).  A maintainer wiring it in should run the
reference's own `scripts/test_models.sh` once.

Usage inside predictlib.get_depth_and_prob_map (predictlib.py:79-102), replacing the call of model.inference_mem:

    from mvsnet_b200 import tf_binding
    # `towers`: the reference's own feature towers, [B, N, Hf, Wf, 32] (model.py:392-406 builds them per view; stack them)
    depth_map, prob_map = tf_binding.inference_mem_from_towers(
        towers, scaled_cams, FLAGS.max_d, depth_start, depth_interval, regnet_variables=regnet_vars)

`regnet_vars` is a dict of numpy arrays read from the checkpoint ('3dconv0_1/kernel', '3dconv0_1/bn/gamma', ...), e.g.
with tf.train.load_checkpoint(path).get_tensor(name).
"""
from __future__ import annotations

import numpy as np

_ENGINES = {}


def _engine(n_views, depth_num, hf, wf, channels, regnet_variables, precision, order):
    import torch

    from .engine import HotPath
    key = (n_views, depth_num, hf, wf, channels, precision, order, id(regnet_variables))
    eng = _ENGINES.get(key)
    if eng is None:
        _ENGINES.clear()
        eng = _ENGINES[key] = HotPath(n_views, depth_num, hf, wf, regnet_variables, channels=channels,
                                      precision=precision, order=order, device=torch.device("cuda", torch.cuda.current_device()))
    return eng


def _run_numpy(towers, cams, depth_start, depth_interval, depth_num, regnet_variables, precision, order):
    """numpy in / numpy out: what tf.py_func calls.  towers [B,N,Hf,Wf,C], cams [B,N,2,4,4], depth_start / interval [B]."""
    import torch
    b, n, hf, wf, c = towers.shape
    eng = _engine(n, int(depth_num), hf, wf, c, regnet_variables, precision, order)
    depth = np.empty((b, hf, wf, 1), dtype=np.float32)
    prob = np.empty((b, hf, wf, 1), dtype=np.float32)
    for i in range(b):
        f = torch.from_numpy(np.ascontiguousarray(towers[i], dtype=np.float32))
        k = torch.from_numpy(np.ascontiguousarray(cams[i], dtype=np.float32))
        d, p = torch.empty((hf, wf)), torch.empty((hf, wf))
        eng.infer_host(f, k, float(np.ravel(depth_start)[i]), float(np.ravel(depth_interval)[i]), d, p)
        depth[i, :, :, 0], prob[i, :, :, 0] = d.numpy(), p.numpy()
    return depth, prob


def inference_mem_from_towers(towers, cams, depth_num, depth_start, depth_interval, regnet_variables, precision="bf16",
                              order="mem"):
    """TF tensors in, TF tensors out (model.py:374-502 after the feature towers): (estimated_depth_map [B,Hf,Wf,1],
    prob_map [B,Hf,Wf,1]).  No gradient is registered: this is the inference op (training=True only selects batch-stat
    BN upstream, which is what the library computes)."""
    import tensorflow as tf
    if not isinstance(depth_num, int):
        raise TypeError("depth_num must be a Python int (model.py:427 iterates range(depth_num))")

    def fn(t, k, s, i):
        return _run_numpy(t, k, s, i, depth_num, regnet_variables, precision, order)

    depth, prob = tf.py_func(fn, [towers, cams, depth_start, depth_interval], [tf.float32, tf.float32], stateful=False,
                             name="mvsb200_hot_path")
    shape = towers.get_shape().as_list()
    for t in (depth, prob):
        t.set_shape([shape[0], shape[2], shape[3], 1])
    return depth, prob
