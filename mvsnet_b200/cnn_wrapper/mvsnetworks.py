"""Drop-in for the 3-D branch of mvsnet/cnn_wrapper/mvsnetworks.py: RegNetUS0 (:122-158)."""
from __future__ import annotations

import ctypes

import torch

from .. import _lib as L
from .. import ops
from ..engine import NETWORK_MODE_DIVISOR, RegnetWeights, regnet_base_filter

_VARIABLES = {}      # the "checkpoint": TF variable name -> array, shared like a TF variable scope


def set_variables(weights: dict) -> None:
    """Register RegNetUS0 variables (what tf.train.Saver.restore does upstream, predictlib.py:69-76)."""
    _VARIABLES.clear()
    _VARIABLES.update(weights)
    _VARIABLES.pop("__device__", None)


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
            for b in range(cost.shape[0]):
                c = cost[b].contiguous()
                if self.precision == "bf16" and c.dtype != torch.bfloat16:
                    c = c.to(torch.bfloat16)
                d, hf, wf, ch = c.shape
                nbytes = lib.mvsb200_regnet_workspace_bytes(d, hf, wf, ch, w.base_filter, prec)
                ws = torch.empty((nbytes,), dtype=torch.uint8, device=c.device)
                out = torch.empty((d, hf, wf), dtype=torch.float32, device=c.device)
                L.check(lib.mvsb200_regnet_forward(L.ptr(c), ops._DTYPE[c.dtype], ctypes.byref(w.params), d, hf, wf,
                                                   ch, w.base_filter, self.epsilon, prec, L.ptr(out), L.ptr(ws),
                                                   nbytes, L.stream_ptr()), "regnet_forward")
                outs.append(out)
            self._output = torch.stack(outs, dim=0)[..., None]
        return self._output
