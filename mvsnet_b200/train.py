"""Training step of the hot path (BASELINE config 4): what mvsnet/train.py:300-349 `get_loss` + train.py:429
`opt.compute_gradients(loss)` compute for the 3DCNN regularisation -- `inference` (model.py:257-372), the regression
loss (loss.py:190-220, loss_type 'original') and the gradients of every RegNetUS0 variable and of the feature maps the
path receives (the feature tower's own backward stays with its owner).  fp32 on the GPU; there is no CPU path.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib as L
from . import ops
from .engine import RegnetWeights, _on_own_device


class TrainStep:
    """feats [N,Hf,Wf,32] + cams [N,2,4,4] + ground-truth depth [Hf,Wf] (0 = invalid pixel) ->
    dict(loss, less_one_accuracy, less_three_accuracy, depth_map, grads {variable name: tensor}, dfeats)."""

    def __init__(self, n_views, depth_num, hf, wf, weights, channels=32, order="train", bn_eps=1e-5, device="cuda"):
        self.lib = L.load()
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.n_views, self.depth_num, self.hf, self.wf, self.channels = n_views, depth_num, hf, wf, channels
        self.order = ops._ORDER[order]
        self.bn_eps = float(bn_eps)
        self.weights = weights if isinstance(weights, RegnetWeights) else RegnetWeights(weights, self.device)
        self.base_filter = self.weights.base_filter
        nbytes = self.lib.mvsb200_train_workspace_bytes(n_views, depth_num, hf, wf, channels, self.base_filter)
        if nbytes == 0:
            raise L.MVSB200Error("train_workspace_bytes rejected the shape")
        self.workspace = torch.empty((nbytes,), dtype=torch.uint8, device=self.device)
        self.grads = {k: torch.zeros_like(v) for k, v in self.weights.tensors.items()}
        self._g = L.RegnetGrads()
        for i, name in enumerate(L.REGNET_LAYER_NAMES):
            self._g.kernel[i] = self.grads[name + "/kernel"].data_ptr()
            if name != "3dconv6_2":
                self._g.gamma[i] = self.grads[name + "/bn/gamma"].data_ptr()
                self._g.beta[i] = self.grads[name + "/bn/beta"].data_ptr()
        self.dfeats = torch.empty((n_views, hf, wf, channels), dtype=torch.float32, device=self.device)
        self.depth_map = torch.empty((hf, wf), dtype=torch.float32, device=self.device)
        self.metrics = torch.empty((3,), dtype=torch.float32, device=self.device)

    @_on_own_device
    def step(self, feats: torch.Tensor, cams: torch.Tensor, gt_depth: torch.Tensor, depth_start: float,
             depth_interval: float) -> dict:
        L.require_cuda(feats, cams, gt_depth)
        if tuple(feats.shape) != (self.n_views, self.hf, self.wf, self.channels) or feats.dtype != torch.float32:
            raise ValueError(f"feats must be fp32 {(self.n_views, self.hf, self.wf, self.channels)}")
        if tuple(gt_depth.shape) != (self.hf, self.wf) or gt_depth.dtype != torch.float32:
            raise ValueError(f"gt_depth must be fp32 {(self.hf, self.wf)}")
        rc = self.lib.mvsb200_train_step(
            L.ptr(feats.contiguous()), L.ptr(cams.to(torch.float32).contiguous()), L.ptr(gt_depth.contiguous()),
            self.n_views, self.depth_num, self.hf, self.wf, self.channels, float(depth_start), float(depth_interval),
            self.order, ctypes.byref(self.weights.params), self.base_filter, self.bn_eps, ctypes.byref(self._g),
            L.ptr(self.dfeats), L.ptr(self.depth_map), L.ptr(self.metrics), L.ptr(self.workspace),
            self.workspace.numel(), L.stream_ptr())
        L.check(rc, "train_step")
        return dict(metrics=self.metrics, depth_map=self.depth_map, grads=self.grads, dfeats=self.dfeats)


def get_loss_and_grads(feats, cams, depth_image, depth_start, depth_interval, depth_num, weights, order="train"):
    """One-call form (train.py:300-349 + :429 for batch size 1): returns (loss, less_one_accuracy, less_three_accuracy,
    grads, dfeats) with the three scalars as Python floats."""
    n, hf, wf, c = feats.shape
    ts = TrainStep(n, int(depth_num), hf, wf, weights, channels=c, order=order, device=feats.device)
    out = ts.step(feats, cams, depth_image, float(depth_start), float(depth_interval))
    m = out["metrics"].cpu().tolist()
    return m[0], m[1], m[2], out["grads"], out["dfeats"]
