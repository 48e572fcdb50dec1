"""Image feature tower UNetDS2GN on the GPU: images [N,H,W,3] -> features [N,H/4,W/4,32].

The step before the hot path (mvsnet/model.py:392-406: one UNetDS2GN per view, shared variables).  Two implementations
in libmvsnet_b200.so, no fallback: precision "fp32" = CUDA-core parity mode (csrc/feature2d.cu), precision "bf16" =
tensor cores (csrc/feature2d_tc.cu: bf16 operands, fp32 accumulation, group normalisation folded into the consumer).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib as L
from .engine import _on_own_device


class UnetWeights:
    """UNetDS2GN variables on the device, keyed by their TF names ('2dconv0_1/kernel', '.../gn/gamma', ...)."""

    def __init__(self, weights: dict, device="cuda"):
        self.tensors = {}
        self.params = L.UnetParams()
        for i, (name, op, k, _s, _m, _src, gn, _relu) in enumerate(L.UNET_LAYER_TABLE):
            kern = self._put(weights, name + "/kernel", device)
            if kern.dim() != 4 or tuple(kern.shape[:2]) != (k, k):
                raise ValueError(f"{name}/kernel must be [{k},{k},*,*], got {tuple(kern.shape)}")
            self.params.kernel[i] = kern.data_ptr()
            if gn:
                self.params.gamma[i] = self._put(weights, name + "/gn/gamma", device).data_ptr()
                self.params.beta[i] = self._put(weights, name + "/gn/beta", device).data_ptr()
        self.base_filter = int(self.tensors["2dconv0_1/kernel"].shape[3])
        self.in_channels = int(self.tensors["2dconv0_1/kernel"].shape[2])
        if self.in_channels != 3:
            raise ValueError("UNetDS2GN reads 3-channel images")

    def _put(self, weights, key, device):
        if key not in weights:
            raise KeyError(f"missing UNetDS2GN variable '{key}'")
        v = weights[key]
        t = torch.as_tensor(np.asarray(v.detach().cpu()) if isinstance(v, torch.Tensor) else np.asarray(v),
                            dtype=torch.float32).contiguous().to(device)
        self.tensors[key] = t
        return t


class FeatureTower:
    """images [N,H,W,3] fp32 (centred) -> features [N,H/4,W/4,4*base_filter] fp32, all views in one call."""

    def __init__(self, weights, epsilon=1e-5, device="cuda", precision="fp32"):
        if precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        self.precision = precision
        self.lib = L.load()
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.weights = weights if isinstance(weights, UnetWeights) else UnetWeights(weights, self.device)
        self.epsilon = float(epsilon)
        self._ws = None
        self._shape = None

    def _workspace(self, n, h, w):
        if self._shape != (n, h, w):
            size_fn = self.lib.mvsb200_unet_tc_workspace_bytes if self.precision == "bf16" else self.lib.mvsb200_unet_workspace_bytes
            nbytes = size_fn(n, h, w, self.weights.base_filter)
            if nbytes == 0:
                raise L.MVSB200Error(f"unet_workspace_bytes rejected the shape: {L.last_error()}")
            self._ws = torch.empty((nbytes,), dtype=torch.uint8, device=self.device)
            self._shape = (n, h, w)
        return self._ws

    def __call__(self, images: torch.Tensor) -> torch.Tensor:
        return self.forward(images)

    @_on_own_device
    def forward(self, images: torch.Tensor) -> torch.Tensor:
        L.require_cuda(images)
        if images.dim() != 4 or images.shape[-1] != 3 or images.dtype != torch.float32:
            raise ValueError("images must be fp32 [N,H,W,3]")
        n, h, w, _ = images.shape
        ws = self._workspace(n, h, w)
        feats = torch.empty((n, h // 4, w // 4, 4 * self.weights.base_filter), dtype=torch.float32, device=images.device)
        fwd = self.lib.mvsb200_unet_tc_forward if self.precision == "bf16" else self.lib.mvsb200_unet_forward
        rc = fwd(L.ptr(images.contiguous()), ctypes.byref(self.weights.params), n, h, w, self.weights.base_filter,
                 self.epsilon, L.ptr(feats), L.ptr(ws), ws.numel(), L.stream_ptr())
        L.check(rc, "unet_forward")
        return feats

    def layer_raw(self, layer: int):
        """bf16 mode, after forward(): (raw output [N,C/8,Ho,Wo,8] bf16, statistics [N,C/8,2] fp64) of a layer, views of
        the workspace (tests)."""
        if self.precision != "bf16":
            raise ValueError("layer_raw is the bf16 tower's inspection call; use layer_output in fp32 mode")
        n, h, w = self._shape
        off, soff, dims = ctypes.c_size_t(), ctypes.c_size_t(), (ctypes.c_int * 3)()
        L.check(self.lib.mvsb200_unet_tc_layer_raw(n, h, w, self.weights.base_filter, layer, ctypes.byref(off), dims,
                                                   ctypes.byref(soff)), "unet_tc_layer_raw")
        ho, wo, c = list(dims)
        raw = self._ws[off.value:off.value + n * ho * wo * c * 2].view(torch.bfloat16).view(n, c // 8, ho, wo, 8)
        stats = self._ws[soff.value:soff.value + n * (c // 8) * 16].view(torch.float64).view(n, c // 8, 2)
        return raw, stats

    def layer_output(self, layer: int) -> torch.Tensor:
        """After forward() in fp32 mode: the (normalised) output [N,Ho,Wo,C] of a layer, a view of the workspace (tests)."""
        if self.precision != "fp32":
            raise ValueError("layer_output is the fp32 tower's inspection call; use layer_raw in bf16 mode")
        n, h, w = self._shape
        off, dims = ctypes.c_size_t(), (ctypes.c_int * 3)()
        L.check(self.lib.mvsb200_unet_layer_output(n, h, w, self.weights.base_filter, layer, ctypes.byref(off), dims),
                "unet_layer_output")
        ho, wo, c = list(dims)
        nbytes = n * ho * wo * c * 4
        return self._ws[off.value:off.value + nbytes].view(torch.float32).view(n, ho, wo, c)


def conv2d_layer(xa, kernel, stride=1, transposed=False, xb=None, with_stats=True):
    """One tf.layers.conv2d / conv2d_transpose (SAME, no bias) of concat(xa, xb): returns (y raw, stats [N,C/8,2])."""
    lib = L.load()
    L.require_cuda(xa, kernel)
    n, h, w, ca = xa.shape
    cb = 0 if xb is None else int(xb.shape[-1])
    k = int(kernel.shape[0])
    cout = int(kernel.shape[2] if transposed else kernel.shape[3])
    ho, wo = (2 * h, 2 * w) if transposed else (-(-h // stride), -(-w // stride))
    y = torch.empty((n, ho, wo, cout), dtype=torch.float32, device=xa.device)
    stats = torch.zeros((n, max(cout // 8, 1), 2), dtype=torch.float64, device=xa.device) if with_stats else None
    rc = lib.mvsb200_conv2d_layer(L.ptr(xa.contiguous()), ca, L.ptr(xb.contiguous()) if xb is not None else None, cb,
                                  L.ptr(kernel.contiguous()), n, h, w, cout, k, stride, int(transposed), L.ptr(y),
                                  L.ptr(stats) if with_stats else None, L.stream_ptr())
    L.check(rc, "conv2d_layer")
    return y, stats


def group_norm_(y, stats, gamma, beta, eps=1e-5, relu=True):
    """In place: group normalisation of y [N,H,W,C] from the (sum, sum of squares) the conv accumulated."""
    lib = L.load()
    n, h, w, c = y.shape
    rc = lib.mvsb200_group_norm(L.ptr(y), L.ptr(stats), L.ptr(gamma.contiguous()), L.ptr(beta.contiguous()), n, h * w, c,
                                float(eps), int(relu), L.stream_ptr())
    L.check(rc, "group_norm")
    return y
