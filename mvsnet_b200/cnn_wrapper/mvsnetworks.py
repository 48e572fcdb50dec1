"""Drop-in for mvsnet/cnn_wrapper/mvsnetworks.py: RegNetUS0 (:122-158) and the image tower UNetDS2GN (:53-115)."""
from __future__ import annotations

import ctypes

import torch

from .. import _lib as L
from .. import ops
from ..engine import NETWORK_MODE_DIVISOR, RegnetWeights, regnet_base_filter

_VARIABLES = {}      # the "checkpoint": TF variable name -> array, shared like a TF variable scope
_VARIABLES_VERSION = [0]   # bumped by every set_variables(): device copies made from an older checkpoint are stale


def set_variables(weights: dict) -> None:
    """Register RegNetUS0 variables (what tf.train.Saver.restore does upstream, predictlib.py:69-76)."""
    _VARIABLES.clear()
    _VARIABLES.update(weights)
    _VARIABLES.pop("__device__", None)
    _VARIABLES_VERSION[0] += 1


def variables_version() -> int:
    return _VARIABLES_VERSION[0]


def get_variables() -> dict:
    return _VARIABLES


class RegNetUS0:
    """RegNetUS0({'data': cost_volume}, trainable=, training=, mode=, reuse=).get_output() -> [B,D,Hf,Wf,1].

    Batch-norm always uses batch statistics, exactly as upstream (network.py:54,64: training defaults
    to True and inference never overrides it); `training=False` is therefore rejected rather than
    silently computing something the reference never ran.
    """

    precision = "bf16"

    def __init__(self, inputs, trainable=True, training=True, mode="normal", reuse=False, epsilon=1e-5, **kwargs):
        if mode not in NETWORK_MODE_DIVISOR:
            raise ValueError(f"unknown network mode {mode!r}")
        if not training:
            raise NotImplementedError("RegNetUS0(training=False) is never exercised by the reference")
        self.inputs = inputs
        self.base_filter = regnet_base_filter(mode)
        self.epsilon = float(epsilon)
        self.layers = dict(inputs)
        self._output = None

    def _forward_one(self, lib, w, prec, c):
        if self.precision == "bf16" and c.dtype != torch.bfloat16:
            c = c.to(torch.bfloat16)
        d, hf, wf, ch = c.shape
        nbytes = lib.mvsb200_regnet_workspace_bytes(d, hf, wf, ch, w.base_filter, prec)
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=c.device)
        out = torch.empty((d, hf, wf), dtype=torch.float32, device=c.device)
        L.check(lib.mvsb200_regnet_forward(L.ptr(c), ops._DTYPE[c.dtype], ctypes.byref(w.params), d, hf, wf, ch,
                                           w.base_filter, self.epsilon, prec, L.ptr(out), L.ptr(ws), nbytes,
                                           L.stream_ptr()), "regnet_forward")
        return out

    def get_output(self):
        if self._output is None:
            cost = self.layers["data"]
            if cost.dim() != 5:
                raise ValueError("Improper input rank for layer: 3dconv1_0")      # network.py:212-215
            if not _VARIABLES:
                raise RuntimeError("RegNetUS0 variables not set: call mvsnetworks.set_variables(weights)")
            w = RegnetWeights(_VARIABLES, cost.device)
            if w.base_filter != self.base_filter:
                raise ValueError(f"variables are for base_filter {w.base_filter}, mode needs {self.base_filter}")
            lib = L.load()
            outs = []
            prec = ops._PRECISION[self.precision]
            with torch.cuda.device(cost.device):
                for b in range(cost.shape[0]):
                    outs.append(self._forward_one(lib, w, prec, cost[b].contiguous()))
            self._output = torch.stack(outs, dim=0)[..., None]
        return self._output


_UNET_VARIABLES = {}


def set_unet_variables(weights: dict) -> None:
    """Register UNetDS2GN variables ('2dconv0_1/kernel', '2dconv0_1/gn/gamma', ...), shared by every tower like the
    reference's reuse=True variable scope (model.py:392-406)."""
    _UNET_VARIABLES.clear()
    _UNET_VARIABLES.update(weights)
    _UNET_CACHE.clear()


_UNET_CACHE = {}


class UNetDS2GN:
    """UNetDS2GN({'data': image}, trainable=, training=, mode=, reuse=).get_output() -> [B,H/4,W/4,32].

    `image` is [B,H,W,3] fp32 (centred).  Group normalisation has no batch dependence, so a batch of B images is B
    independent towers (upstream builds one tower per view and batch_size 1)."""

    def __init__(self, inputs, trainable=True, training=True, mode="normal", reuse=False, epsilon=1e-5, precision="fp32",
                 **kwargs):
        if mode not in NETWORK_MODE_DIVISOR:
            raise ValueError(f"unknown network mode {mode!r}")
        if precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        self.precision = precision          # "bf16": the tensor-core tower (csrc/feature2d_tc.cu)
        if regnet_base_filter(mode) % 8:
            raise NotImplementedError(f"UNetDS2GN mode {mode!r}: groups of fewer than 8 channels are not built")
        self.base_divisor = NETWORK_MODE_DIVISOR[mode]
        self.base_filter = regnet_base_filter(mode)
        self.epsilon = float(epsilon)
        self.layers = dict(inputs)
        self._output = None

    def get_output(self):
        if self._output is None:
            from ..features import FeatureTower
            image = self.layers["data"]
            if image.dim() != 4:
                raise ValueError("Improper input rank for layer: 2dconv1_0")
            if not _UNET_VARIABLES:
                raise RuntimeError("UNetDS2GN variables not set: call mvsnetworks.set_unet_variables(weights)")
            key = (str(image.device), self.epsilon, self.precision)
            if key not in _UNET_CACHE:
                _UNET_CACHE[key] = FeatureTower(_UNET_VARIABLES, self.epsilon, image.device, precision=self.precision)
            tower = _UNET_CACHE[key]
            if tower.weights.base_filter != self.base_filter:
                raise ValueError(f"variables are for base_filter {tower.weights.base_filter}, mode needs {self.base_filter}")
            self._output = tower(image.to(torch.float32))
        return self._output


_REFINE_VARIABLES = {}


def set_refine_variables(weights: dict) -> None:
    """Register the variables of the refinement tower: 'refine_conv{0..3}/kernel' [3,3,Cin,Cout] and '/bias' [Cout]."""
    _REFINE_VARIABLES.clear()
    _REFINE_VARIABLES.update(weights)


class RefineNetConv:
    """RefineNetConv({'color_image': image, 'depth_image': data}, ...).get_output() -> [B,H,W,1]: the refinement tower of
    network_type 'original' (mvsnetworks.py:178-193): concat, three biased 3x3 convolutions with ReLU, a biased 3x3
    convolution to one channel."""

    def __init__(self, inputs, trainable=True, training=True, mode="normal", reuse=False, **kwargs):
        if mode not in NETWORK_MODE_DIVISOR:
            raise ValueError(f"unknown network mode {mode!r}")
        self.base_filter = max(1, int(32 / NETWORK_MODE_DIVISOR[mode]))
        self.layers = dict(inputs)
        self._output = None

    def get_output(self):
        if self._output is None:
            color, depth = self.layers["color_image"], self.layers["depth_image"]
            if color.dim() != 4 or depth.dim() != 4:
                raise ValueError("Improper input rank for layer: refine_conv0")
            if not _REFINE_VARIABLES:
                raise RuntimeError("refinement variables not set: call mvsnetworks.set_refine_variables(weights)")
            lib = L.load()
            with torch.cuda.device(color.device):
                L.require_cuda(color, depth)
                n, h, w, _ = color.shape
                xa, xb = color.to(torch.float32).contiguous(), depth.to(torch.float32).contiguous()
                for i in range(4):
                    k = torch.as_tensor(_REFINE_VARIABLES[f"refine_conv{i}/kernel"], dtype=torch.float32).to(color.device).contiguous()
                    b = torch.as_tensor(_REFINE_VARIABLES[f"refine_conv{i}/bias"], dtype=torch.float32).to(color.device).contiguous()
                    ca, cb = xa.shape[3], 0 if xb is None else xb.shape[3]
                    if tuple(k.shape[:3]) != (3, 3, ca + cb):
                        raise ValueError(f"refine_conv{i}/kernel must be [3,3,{ca + cb},*], got {tuple(k.shape)}")
                    if i < 3 and k.shape[3] != self.base_filter:
                        raise ValueError(f"variables are for {k.shape[3]} filters, mode needs {self.base_filter}")
                    y = torch.empty((n, h, w, k.shape[3]), dtype=torch.float32, device=color.device)
                    L.check(lib.mvsb200_conv2d_bias(L.ptr(xa), ca, L.ptr(xb) if xb is not None else None, cb, L.ptr(k),
                                                    L.ptr(b), n, h, w, int(k.shape[3]), int(i < 3), L.ptr(y),
                                                    L.stream_ptr()), "conv2d_bias")
                    xa, xb = y, None
            self._output = xa
        return self._output
