"""Tensor-level wrappers over the C ABI (one function per entry point of include/mvsnet_b200.h).

Inputs and outputs are CUDA torch tensors (torch owns the memory; all arithmetic happens in
libmvsnet_b200.so).  Shapes follow the reference: channels-last, fp32.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib as L

_ORDER = {"mem": L.ORDER_MEM, "train": L.ORDER_TRAIN}
_SAMPLER = {"transform": L.SAMPLER_TRANSFORM, "legacy": L.SAMPLER_LEGACY}
_PRECISION = {"fp32": L.PRECISION_FP32, "bf16": L.PRECISION_BF16}
_DTYPE = {torch.float32: L.F32, torch.bfloat16: L.BF16}


def _f32c(t: torch.Tensor) -> torch.Tensor:
    L.require_cuda(t)
    return t.to(torch.float32).contiguous()


def depth_end_f32(depth_num: int, depth_start: float, depth_interval: float) -> float:
    """model.py:378-379 in fp32: start + (float(D) - 1) * interval."""
    f = np.float32
    return float(f(depth_start) + f(f(depth_num) - f(1.0)) * f(depth_interval))


def device_info():
    lib = L.load()
    sm, major, minor = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    L.check(lib.mvsb200_device_info(ctypes.byref(sm), ctypes.byref(major), ctypes.byref(minor)), "device_info")
    return sm.value, major.value, minor.value


def launch_count() -> int:
    return int(L.load().mvsb200_launch_count())


def homographies(cams: torch.Tensor, depth_num: int, depth_start: float, depth_step: float,
                 inverse_depth: bool = False, want_transforms: bool = False):
    """cams [N,2,4,4] -> H [(N-1),D,3,3] (and T [(N-1),D,8])."""
    lib = L.load()
    cams = _f32c(cams)
    n = cams.shape[0]
    H = torch.empty((n - 1, depth_num, 3, 3), device=cams.device, dtype=torch.float32)
    T = torch.empty((n - 1, depth_num, 8), device=cams.device, dtype=torch.float32) if want_transforms else None
    L.check(lib.mvsb200_homographies(L.ptr(cams), n, int(depth_num), float(depth_start), float(depth_step),
                                     int(bool(inverse_depth)), L.ptr(H), L.ptr(T), L.stream_ptr()), "homographies")
    return (H, T) if want_transforms else H


def transform_coefs(H: torch.Tensor) -> torch.Tensor:
    lib = L.load()
    H = _f32c(H).reshape(-1, 9)
    T = torch.empty((H.shape[0], 8), device=H.device, dtype=torch.float32)
    L.check(lib.mvsb200_transform_coefs(L.ptr(H), H.shape[0], L.ptr(T), L.stream_ptr()), "transform_coefs")
    return T


def warp(image: torch.Tensor, H: torch.Tensor, sampler: str = "transform") -> torch.Tensor:
    """image [B or 1,H,W,C], H [B,3,3] -> [B,H,W,C]."""
    lib = L.load()
    image = _f32c(image)
    H = _f32c(H).reshape(-1, 9)
    ic, h, w, c = image.shape
    out = torch.empty((H.shape[0], h, w, c), device=image.device, dtype=torch.float32)
    L.check(lib.mvsb200_warp(L.ptr(image), ic, L.ptr(H), H.shape[0], h, w, c, _SAMPLER[sampler], L.ptr(out),
                             L.stream_ptr()), "warp")
    return out


def interpolate(image: torch.Tensor, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    lib = L.load()
    image, x, y = _f32c(image), _f32c(x).reshape(-1), _f32c(y).reshape(-1)
    b, h, w, c = image.shape
    if x.numel() != b * h * w or y.numel() != b * h * w:
        raise ValueError("interpolate: x and y must hold B*H*W coordinates")
    out = torch.empty((b * h * w, c), device=image.device, dtype=torch.float32)
    L.check(lib.mvsb200_interpolate(L.ptr(image), L.ptr(x), L.ptr(y), b, h, w, c, L.ptr(out), L.stream_ptr()),
            "interpolate")
    return out


def pixel_grids(height: int, width: int, device="cuda") -> torch.Tensor:
    lib = L.load()
    out = torch.empty((3 * height * width,), device=device, dtype=torch.float32)
    L.check(lib.mvsb200_pixel_grids(int(height), int(width), L.ptr(out), L.stream_ptr()), "pixel_grids")
    return out


def sample_coords(H: torch.Tensor, height: int, width: int, sampler: str = "transform") -> torch.Tensor:
    lib = L.load()
    H = _f32c(H).reshape(-1, 9)
    out = torch.empty((H.shape[0], height, width, 2), device=H.device, dtype=torch.float32)
    L.check(lib.mvsb200_sample_coords(L.ptr(H), H.shape[0], height, width, _SAMPLER[sampler], L.ptr(out),
                                      L.stream_ptr()), "sample_coords")
    return out


def cost_volume(feats: torch.Tensor, H: torch.Tensor, order: str = "mem", sampler: str = "transform",
                out_dtype=torch.float32, variant: int = 0, out: torch.Tensor | None = None) -> torch.Tensor:
    """feats [N,Hf,Wf,C], H [(N-1),D,3,3] -> [D,Hf,Wf,C]."""
    lib = L.load()
    feats = _f32c(feats)
    H = _f32c(H)
    n, hf, wf, c = feats.shape
    d = H.shape[1]
    if H.shape[0] != n - 1:
        raise ValueError("cost_volume: need one homography stack per source view")
    if out is None:
        out = torch.empty((d, hf, wf, c), device=feats.device, dtype=out_dtype)
    L.check(lib.mvsb200_cost_volume(L.ptr(feats), L.ptr(H), n, d, hf, wf, c, _ORDER[order], _SAMPLER[sampler],
                                    _DTYPE[out.dtype], L.ptr(out), int(variant), L.stream_ptr()), "cost_volume")
    return out


def conv3d_layer(x, kernel_tf, stride=1, transposed=False, precision="fp32", x_affine=None, skip=None,
                 skip_affine=None, out_dtype=None, want_stats=True):
    """One regularizer layer; returns (y_raw [Do,Ho,Wo,Cout], stats [2*Cout] float64 or None)."""
    lib = L.load()
    L.require_cuda(x, kernel_tf)
    x = x.contiguous()
    kernel_tf = _f32c(kernel_tf)
    d, h, w, cin = x.shape
    cout = kernel_tf.shape[3] if transposed else kernel_tf.shape[4]
    if transposed:
        od, oh, ow = 2 * d, 2 * h, 2 * w
    else:
        od, oh, ow = -(-d // stride), -(-h // stride), -(-w // stride)
    if out_dtype is None:
        out_dtype = torch.bfloat16 if precision == "bf16" else torch.float32
    y = torch.empty((od, oh, ow, cout), device=x.device, dtype=out_dtype)
    stats = torch.zeros((2 * cout,), device=x.device, dtype=torch.float64) if want_stats else None
    xs, xb = (None, None) if x_affine is None else (_f32c(x_affine[0]), _f32c(x_affine[1]))
    ss, sb = (None, None) if skip_affine is None else (_f32c(skip_affine[0]), _f32c(skip_affine[1]))
    if skip is not None:
        skip = skip.contiguous()
        if skip.dtype != x.dtype or skip.shape != x.shape:
            raise ValueError("conv3d_layer: skip must match x in dtype and shape")
    L.check(lib.mvsb200_conv3d_layer(L.ptr(x), _DTYPE[x.dtype], L.ptr(xs), L.ptr(xb), L.ptr(skip), L.ptr(ss),
                                     L.ptr(sb), L.ptr(kernel_tf), d, h, w, cin, cout, int(stride),
                                     int(bool(transposed)), _PRECISION[precision], L.ptr(y), _DTYPE[y.dtype],
                                     L.ptr(stats), L.stream_ptr()), "conv3d_layer")
    return y, stats


def bn_finalize(stats, gamma, beta, count, eps=1e-5):
    lib = L.load()
    gamma, beta = _f32c(gamma), _f32c(beta)
    c = gamma.numel()
    scale = torch.empty((c,), device=gamma.device, dtype=torch.float32)
    shift = torch.empty((c,), device=gamma.device, dtype=torch.float32)
    L.check(lib.mvsb200_bn_finalize(L.ptr(stats), L.ptr(gamma), L.ptr(beta), c, float(count), float(eps),
                                    L.ptr(scale), L.ptr(shift), L.stream_ptr()), "bn_finalize")
    return scale, shift


def depth_regress(filtered: torch.Tensor, depth_start: float, depth_interval: float, inverse_depth=False,
                  num_buckets: int = 4, want_prob_volume: bool = False):
    """filtered [D,Hf,Wf] -> depth [Hf,Wf], prob [Hf,Wf] (and P [D,Hf,Wf])."""
    lib = L.load()
    filtered = _f32c(filtered)
    d, hf, wf = filtered.shape
    depth = torch.empty((hf, wf), device=filtered.device, dtype=torch.float32)
    prob = torch.empty((hf, wf), device=filtered.device, dtype=torch.float32)
    pv = torch.empty_like(filtered) if want_prob_volume else None
    L.check(lib.mvsb200_depth_regress(L.ptr(filtered), d, hf, wf, float(depth_start), float(depth_interval),
                                      int(bool(inverse_depth)), int(num_buckets), L.ptr(depth), L.ptr(prob),
                                      L.ptr(pv), L.stream_ptr()), "depth_regress")
    return (depth, prob, pv) if want_prob_volume else (depth, prob)


def probability_map(prob_volume: torch.Tensor, depth_map: torch.Tensor, depth_start: float, depth_interval: float,
                    inverse_depth=False, num_buckets: int = 4) -> torch.Tensor:
    lib = L.load()
    prob_volume = _f32c(prob_volume)
    depth_map = _f32c(depth_map)
    d, h, w = prob_volume.shape
    out = torch.empty((h, w), device=prob_volume.device, dtype=torch.float32)
    L.check(lib.mvsb200_probability_map(L.ptr(prob_volume), L.ptr(depth_map), d, h, w, float(depth_start),
                                        float(depth_interval), int(bool(inverse_depth)), int(num_buckets),
                                        L.ptr(out), L.stream_ptr()), "probability_map")
    return out


def umma_probe(a_image: torch.Tensor, b_image: torch.Tensor, n: int, kblocks: int, a_kblock_stride: int,
               a_start: int, a_lbo: int, a_sbo: int, b_kblock_stride: int, b_lbo: int, b_sbo: int) -> torch.Tensor:
    lib = L.load()
    L.require_cuda(a_image, b_image)
    out = torch.zeros((128, n), device=a_image.device, dtype=torch.float32)
    L.check(lib.mvsb200_umma_probe(L.ptr(a_image), a_image.numel() * a_image.element_size(), L.ptr(b_image),
                                   b_image.numel() * b_image.element_size(), n, kblocks, a_kblock_stride, a_start,
                                   a_lbo, a_sbo, b_kblock_stride, b_lbo, b_sbo, L.ptr(out), L.stream_ptr()),
            "umma_probe")
    return out
