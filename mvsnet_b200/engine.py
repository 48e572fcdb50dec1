"""Device-resident state for the hot path: weights, workspace, staging; whole-path calls.

Mirrors what the reference keeps inside its TF session: variables restored once
(predictlib.py:69-76) and one `sess.run` per reference view (inference.py:105-112).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib as L
from . import ops

NETWORK_MODE_DIVISOR = {  # cnn_wrapper/network.py:75-85
    # 'semilite' is written `4/3` upstream (network.py:77) and the reference runs on Python 2.7 (xrange, iteritems),
    # where that is INTEGER division = 1: semilite checkpoints have base_filter 8, like 'normal'
    "normal": 1.0, "semilite": 1.0, "lite": 2.0, "ultralite": 4.0, "fat": 0.5, "ultrafat": 0.25,
}


def regnet_base_filter(network_mode: str = "normal") -> int:
    """mvsnetworks.py:126-127: max(1, int(8 / base_divisor))."""
    return max(1, int(8 / NETWORK_MODE_DIVISOR[network_mode]))


def _on_own_device(method):
    """Run a HotPath / FeatureTower method with the object's GPU as the current device: the library launches on the
    current device and the current stream, whatever device its pointer arguments live on."""
    import functools

    @functools.wraps(method)
    def wrapper(self, *args, **kwargs):
        with torch.cuda.device(self.device):
            return method(self, *args, **kwargs)
    return wrapper


class RegnetWeights:
    """RegNetUS0 variables on the device, keyed by their TF names ('3dconv0_1/kernel', '.../bn/gamma', ...)."""

    def __init__(self, weights: dict, device="cuda"):
        self.tensors = {}
        self.params = L.RegnetParams()
        for i, name in enumerate(L.REGNET_LAYER_NAMES):
            k = self._put(weights, name + "/kernel", device)
            if k.dim() != 5 or tuple(k.shape[:3]) != (3, 3, 3):
                raise ValueError(f"{name}/kernel must be [3,3,3,*,*], got {tuple(k.shape)}")
            self.params.kernel[i] = k.data_ptr()
            if name != "3dconv6_2":
                g = self._put(weights, name + "/bn/gamma", device)
                b = self._put(weights, name + "/bn/beta", device)
                self.params.gamma[i] = g.data_ptr()
                self.params.beta[i] = b.data_ptr()
        k01 = self.tensors["3dconv0_1/kernel"]
        self.in_channels = int(k01.shape[3])
        self.base_filter = int(k01.shape[4])

    def _put(self, weights, key, device):
        if key not in weights:
            raise KeyError(f"missing RegNetUS0 variable '{key}'")
        v = weights[key]
        t = torch.as_tensor(np.asarray(v.detach().cpu()) if isinstance(v, torch.Tensor) else np.asarray(v),
                            dtype=torch.float32).contiguous().to(device)
        self.tensors[key] = t
        return t


class HotPath:
    """feats [N,Hf,Wf,C] + cams [N,2,4,4] -> depth map + probability map, on one GPU."""

    def __init__(self, n_views, depth_num, hf, wf, weights, channels=32, precision="bf16", order="mem",
                 sampler="transform", inverse_depth=False, bn_eps=1e-5, device="cuda"):
        self.lib = L.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise L.MVSB200Error("mvsnet_b200 needs a CUDA device: there is no CPU path")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.n_views, self.depth_num, self.hf, self.wf, self.channels = n_views, depth_num, hf, wf, channels
        self.precision = ops._PRECISION[precision]
        self.precision_name = precision
        self.order = ops._ORDER[order]
        self.sampler = ops._SAMPLER[sampler]
        self.inverse_depth = int(bool(inverse_depth))
        self.bn_eps = float(bn_eps)
        self.weights = weights if isinstance(weights, RegnetWeights) else RegnetWeights(weights, self.device)
        if self.weights.in_channels != channels:
            raise ValueError(f"feature channels {channels} != 3dconv0_1 input channels {self.weights.in_channels}")
        self.base_filter = self.weights.base_filter
        nbytes = self.lib.mvsb200_infer_workspace_bytes(n_views, depth_num, hf, wf, channels, self.base_filter,
                                                        self.precision)
        if nbytes == 0:
            raise L.MVSB200Error("infer_workspace_bytes rejected the shape")
        self.workspace = torch.empty((nbytes,), dtype=torch.uint8, device=self.device)
        sbytes = self.lib.mvsb200_infer_host_staging_bytes(n_views, hf, wf, channels)
        self.staging = torch.empty((sbytes,), dtype=torch.uint8, device=self.device)
        self.depth_map = torch.empty((hf, wf), dtype=torch.float32, device=self.device)
        self.prob_map = torch.empty((hf, wf), dtype=torch.float32, device=self.device)
        self.h2d_bytes = n_views * hf * wf * channels * 4 + n_views * 32 * 4
        self.d2h_bytes = 2 * hf * wf * 4

    # -- whole path, device buffers --------------------------------------------------------------
    @_on_own_device
    def infer(self, feats: torch.Tensor, cams: torch.Tensor, depth_start: float, depth_interval: float,
              depth_map: torch.Tensor | None = None, prob_map: torch.Tensor | None = None):
        L.require_cuda(feats, cams)
        if tuple(feats.shape) != (self.n_views, self.hf, self.wf, self.channels) or feats.dtype != torch.float32:
            raise ValueError(f"feats must be fp32 {(self.n_views, self.hf, self.wf, self.channels)}")
        if tuple(cams.shape) != (self.n_views, 2, 4, 4) or cams.dtype != torch.float32:
            raise ValueError("cams must be fp32 [N,2,4,4]")
        depth_map = self.depth_map if depth_map is None else depth_map
        prob_map = self.prob_map if prob_map is None else prob_map
        rc = self.lib.mvsb200_infer(
            L.ptr(feats.contiguous()), L.ptr(cams.contiguous()), self.n_views, self.depth_num, self.hf, self.wf,
            self.channels, float(depth_start), float(depth_interval), self.inverse_depth, self.order, self.sampler,
            ctypes.byref(self.weights.params), self.base_filter, self.bn_eps, self.precision, L.ptr(depth_map),
            L.ptr(prob_map), L.ptr(self.workspace), self.workspace.numel(), L.stream_ptr())
        L.check(rc, "infer")
        return depth_map, prob_map

    # -- whole path, host buffers (feed/fetch boundary of sess.run, inference.py:111) -----------------
    @_on_own_device
    def infer_host(self, feats_host: torch.Tensor, cams_host: torch.Tensor, depth_start: float,
                   depth_interval: float, depth_out: torch.Tensor, prob_out: torch.Tensor):
        for t in (feats_host, cams_host, depth_out, prob_out):
            if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("infer_host takes contiguous fp32 HOST tensors")
        rc = self.lib.mvsb200_infer_host(
            L.ptr(feats_host), L.ptr(cams_host), self.n_views, self.depth_num, self.hf, self.wf, self.channels,
            float(depth_start), float(depth_interval), self.inverse_depth, self.order, self.sampler,
            ctypes.byref(self.weights.params), self.base_filter, self.bn_eps, self.precision, L.ptr(depth_out),
            L.ptr(prob_out), L.ptr(self.staging), L.ptr(self.workspace), self.workspace.numel(), L.stream_ptr())
        L.check(rc, "infer_host")
        return depth_out, prob_out

    @_on_own_device
    def infer_host_async(self, feats_host: torch.Tensor, cams_host: torch.Tensor, depth_start: float,
                         depth_interval: float, depth_out: torch.Tensor, prob_out: torch.Tensor):
        """infer_host without the final synchronisation: everything is enqueued on the current stream and the
        (pinned) host outputs are valid after that stream has been synchronised.  Two engines on two streams
        overlap the feed of one reference view with the kernels of the other."""
        for t in (feats_host, cams_host, depth_out, prob_out):
            if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or not t.is_pinned():
                raise ValueError("infer_host_async takes contiguous fp32 PINNED host tensors")
        rc = self.lib.mvsb200_infer_host_async(
            L.ptr(feats_host), L.ptr(cams_host), self.n_views, self.depth_num, self.hf, self.wf, self.channels,
            float(depth_start), float(depth_interval), self.inverse_depth, self.order, self.sampler,
            ctypes.byref(self.weights.params), self.base_filter, self.bn_eps, self.precision, L.ptr(depth_out),
            L.ptr(prob_out), L.ptr(self.staging), L.ptr(self.workspace), self.workspace.numel(), L.stream_ptr())
        L.check(rc, "infer_host_async")
        return depth_out, prob_out

    @_on_own_device
    def infer_host_pipelined(self, feats_host: torch.Tensor, cams_host: torch.Tensor, depth_start: float,
                             depth_interval: float, depth_out: torch.Tensor, prob_out: torch.Tensor,
                             compute_stream: torch.cuda.Stream, copy_stream: torch.cuda.Stream):
        """infer_host_async with the copies on `copy_stream` and the kernels on `compute_stream` (chained by events).
        Two engines, each with its own copy stream, alternating over ONE compute stream keep the kernels back to back
        while the feed of the next reference view runs beside them; the outputs are valid once copy_stream has been
        synchronised."""
        for t in (feats_host, cams_host, depth_out, prob_out):
            if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or not t.is_pinned():
                raise ValueError("infer_host_pipelined takes contiguous fp32 PINNED host tensors")
        rc = self.lib.mvsb200_infer_host_pipelined(
            L.ptr(feats_host), L.ptr(cams_host), self.n_views, self.depth_num, self.hf, self.wf, self.channels,
            float(depth_start), float(depth_interval), self.inverse_depth, self.order, self.sampler,
            ctypes.byref(self.weights.params), self.base_filter, self.bn_eps, self.precision, L.ptr(depth_out),
            L.ptr(prob_out), L.ptr(self.staging), L.ptr(self.workspace), self.workspace.numel(),
            ctypes.c_void_p(compute_stream.cuda_stream), ctypes.c_void_p(copy_stream.cuda_stream))
        L.check(rc, "infer_host_pipelined")
        return depth_out, prob_out

    @_on_own_device
    def set_stage_events(self, events) -> None:
        """events: five torch.cuda.Event(enable_timing=True) (or None) recorded at the stage boundaries of infer()."""
        if events is None:
            L.check(self.lib.mvsb200_infer_set_stage_events(None), "set_stage_events")
            return
        for e in events:
            e.record()            # torch creates the underlying cudaEvent lazily on first record
        arr = (ctypes.c_void_p * 5)(*[e.cuda_event for e in events])
        L.check(self.lib.mvsb200_infer_set_stage_events(arr), "set_stage_events")

    # -- stages (tests, profiling) ---------------------------------------------------------------------
    @_on_own_device
    def cost_volume_planar(self, feats: torch.Tensor, cams: torch.Tensor, depth_start: float, depth_interval: float):
        """Run the whole path once and return views of the cost volume inside the workspace in the regularizer's two
        layouts: CP8 [D, C/8, Hf, Wf, 8] and PS8 [D, C/8, 4, Hf/2, Wf/2, 8] (bf16 mode, even Hf / Wf)."""
        self.infer(feats, cams, depth_start, depth_interval)
        cp8_off, ps8_off = ctypes.c_size_t(), ctypes.c_size_t()
        L.check(self.lib.mvsb200_infer_cost_offsets(self.n_views, self.depth_num, self.hf, self.wf, self.channels,
                                                    self.base_filter, self.precision, ctypes.byref(cp8_off),
                                                    ctypes.byref(ps8_off)), "infer_cost_offsets")
        n = self.depth_num * self.hf * self.wf * self.channels * 2
        cp8 = self.workspace[cp8_off.value:cp8_off.value + n].view(torch.bfloat16)
        ps8 = self.workspace[ps8_off.value:ps8_off.value + n].view(torch.bfloat16)
        return cp8, ps8

    def filtered_volume(self) -> torch.Tensor:
        """The filtered cost volume [D,Hf,Wf] fp32 the last infer() left in the workspace (a view, not a copy)."""
        off = ctypes.c_size_t()
        L.check(self.lib.mvsb200_infer_filtered_offset(self.n_views, self.depth_num, self.hf, self.wf, self.channels,
                                                       self.base_filter, self.precision, ctypes.byref(off)),
                "infer_filtered_offset")
        n = self.depth_num * self.hf * self.wf * 4
        return self.workspace[off.value:off.value + n].view(torch.float32).view(self.depth_num, self.hf, self.wf)

    @_on_own_device
    def regnet(self, cost: torch.Tensor) -> torch.Tensor:
        d, hf, wf, c = cost.shape
        nbytes = self.lib.mvsb200_regnet_workspace_bytes(d, hf, wf, c, self.base_filter, self.precision)
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=cost.device)
        out = torch.empty((d, hf, wf), dtype=torch.float32, device=cost.device)
        rc = self.lib.mvsb200_regnet_forward(L.ptr(cost.contiguous()), ops._DTYPE[cost.dtype],
                                             ctypes.byref(self.weights.params), d, hf, wf, c, self.base_filter,
                                             self.bn_eps, self.precision, L.ptr(out), L.ptr(ws), nbytes,
                                             L.stream_ptr())
        L.check(rc, "regnet_forward")
        self._last_regnet_ws = ws
        return out

    def regnet_layer_raw(self, layer: int, d: int, hf: int, wf: int):
        """After regnet(): (raw output, scale, shift) of one layer as tensors viewing the workspace."""
        ws = self._last_regnet_ws
        sc, sh = ctypes.c_void_p(), ctypes.c_void_p()
        raw = self.lib.mvsb200_regnet_layer_raw(L.ptr(ws), d, hf, wf, self.channels, self.base_filter, self.precision,
                                                layer, ctypes.byref(sc), ctypes.byref(sh))
        return raw, sc.value, sh.value
