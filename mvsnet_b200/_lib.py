"""ctypes binding of libmvsnet_b200.so (include/mvsnet_b200.h).

There is no fallback: if the library has not been built, or a call fails, this
raises.  PyTorch is used only to own device memory and streams.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_size_t, c_uint64, c_void_p

LIB_PATH = os.environ.get("MVSB200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libmvsnet_b200.so")   # MVSB200_LIB: development override (A/B builds)

OK = 0
ORDER_MEM, ORDER_TRAIN = 0, 1
SAMPLER_TRANSFORM, SAMPLER_LEGACY = 0, 1
F32, BF16 = 0, 1
PRECISION_FP32, PRECISION_BF16 = 0, 1
REGNET_LAYERS = 11
REGNET_LAYER_NAMES = ["3dconv1_0", "3dconv2_0", "3dconv3_0", "3dconv0_1", "3dconv1_1", "3dconv2_1",
                      "3dconv3_1", "3dconv4_0", "3dconv5_0", "3dconv6_0", "3dconv6_2"]


class MVSB200Error(RuntimeError):
    """A call into libmvsnet_b200.so returned a non-zero status."""


UNET_LAYERS = 32
# name, op, kernel, stride, filters / base_filter, sources (-1 = the images), group norm, relu: mvsnetworks.py:58-115
UNET_LAYER_TABLE = [
    ("2dconv1_0", "conv", 3, 2, 2, (-1,), True, True), ("2dconv2_0", "conv", 3, 2, 4, (0,), True, True),
    ("2dconv3_0", "conv", 3, 2, 8, (1,), True, True), ("2dconv4_0", "conv", 3, 2, 16, (2,), True, True),
    ("2dconv0_1", "conv", 3, 1, 1, (-1,), True, True), ("2dconv0_2", "conv", 3, 1, 1, (4,), True, True),
    ("2dconv1_1", "conv", 3, 1, 2, (0,), True, True), ("2dconv1_2", "conv", 3, 1, 2, (6,), True, True),
    ("2dconv2_1", "conv", 3, 1, 4, (1,), True, True), ("2dconv2_2", "conv", 3, 1, 4, (8,), True, True),
    ("2dconv3_1", "conv", 3, 1, 8, (2,), True, True), ("2dconv3_2", "conv", 3, 1, 8, (10,), True, True),
    ("2dconv4_1", "conv", 3, 1, 16, (3,), True, True), ("2dconv4_2", "conv", 3, 1, 16, (12,), True, True),
    ("2dconv5_0", "deconv", 3, 2, 8, (13,), True, False), ("2dconv5_1", "conv", 3, 1, 8, (14, 11), True, True),
    ("2dconv5_2", "conv", 3, 1, 8, (15,), True, True), ("2dconv6_0", "deconv", 3, 2, 4, (16,), True, False),
    ("2dconv6_1", "conv", 3, 1, 4, (17, 9), True, True), ("2dconv6_2", "conv", 3, 1, 4, (18,), True, True),
    ("2dconv7_0", "deconv", 3, 2, 2, (19,), True, False), ("2dconv7_1", "conv", 3, 1, 2, (20, 7), True, True),
    ("2dconv7_2", "conv", 3, 1, 2, (21,), True, True), ("2dconv8_0", "deconv", 3, 2, 1, (22,), True, False),
    ("2dconv8_1", "conv", 3, 1, 1, (23, 5), True, True), ("2dconv8_2", "conv", 3, 1, 1, (24,), True, True),
    ("conv9_0", "conv", 5, 2, 2, (25,), True, True), ("conv9_1", "conv", 3, 1, 2, (26,), True, True),
    ("conv9_2", "conv", 3, 1, 2, (27,), True, True), ("conv10_0", "conv", 5, 2, 4, (28,), True, True),
    ("conv10_1", "conv", 3, 1, 4, (29,), True, True), ("conv10_2", "conv", 3, 1, 4, (30,), False, False),
]
UNET_LAYER_NAMES = [t[0] for t in UNET_LAYER_TABLE]


class UnetParams(ctypes.Structure):
    _fields_ = [("kernel", c_void_p * UNET_LAYERS), ("gamma", c_void_p * UNET_LAYERS), ("beta", c_void_p * UNET_LAYERS)]


class RegnetParams(ctypes.Structure):
    _fields_ = [("kernel", c_void_p * REGNET_LAYERS), ("gamma", c_void_p * REGNET_LAYERS),
                ("beta", c_void_p * REGNET_LAYERS)]


class RegnetGrads(ctypes.Structure):
    _fields_ = [("kernel", c_void_p * REGNET_LAYERS), ("gamma", c_void_p * REGNET_LAYERS),
                ("beta", c_void_p * REGNET_LAYERS)]


# name -> (restype, argtypes); every symbol include/mvsnet_b200.h declares
_P = c_void_p
SIGNATURES = {
    "mvsb200_last_error": (c_char_p, []),
    "mvsb200_version": (c_int, []),
    "mvsb200_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "mvsb200_homographies": (c_int, [_P, c_int, c_int, c_float, c_float, c_int, _P, _P, _P]),
    "mvsb200_transform_coefs": (c_int, [_P, c_int, _P, _P]),
    "mvsb200_warp": (c_int, [_P, c_int, _P, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "mvsb200_interpolate": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, _P, _P]),
    "mvsb200_pixel_grids": (c_int, [c_int, c_int, _P, _P]),
    "mvsb200_sample_coords": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P]),
    "mvsb200_cost_volume": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_int, _P]),
    "mvsb200_conv3d_layer": (c_int, [_P, c_int, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int,
                                     c_int, c_int, _P, c_int, _P, _P]),
    "mvsb200_bn_finalize": (c_int, [_P, _P, _P, c_int, c_double, c_float, _P, _P, _P]),
    "mvsb200_regnet_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "mvsb200_regnet_forward": (c_int, [_P, c_int, POINTER(RegnetParams), c_int, c_int, c_int, c_int, c_int, c_float,
                                       c_int, _P, _P, c_size_t, _P]),
    "mvsb200_regnet_layer_raw": (c_void_p, [_P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_void_p),
                                            POINTER(c_void_p)]),
    "mvsb200_conv2d_layer": (c_int, [_P, c_int, _P, c_int, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P]),
    "mvsb200_group_norm": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_float, c_int, _P]),
    "mvsb200_unet_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "mvsb200_unet_forward": (c_int, [_P, POINTER(UnetParams), c_int, c_int, c_int, c_int, c_float, _P, _P, c_size_t, _P]),
    "mvsb200_unet_layer_output": (c_int, [c_int, c_int, c_int, c_int, c_int, POINTER(c_size_t), POINTER(c_int)]),
    "mvsb200_unet_tc_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "mvsb200_unet_tc_forward": (c_int, [_P, POINTER(UnetParams), c_int, c_int, c_int, c_int, c_float, _P, _P, c_size_t, _P]),
    "mvsb200_unet_tc_plan": (c_int, [c_int, c_int, c_int, c_int, c_int, POINTER(c_int)]),
    "mvsb200_unet_tc_layer_raw": (c_int, [c_int, c_int, c_int, c_int, c_int, POINTER(c_size_t), POINTER(c_int),
                                          POINTER(c_size_t)]),
    "mvsb200_depth_regress": (c_int, [_P, c_int, c_int, c_int, c_float, c_float, c_int, c_int, _P, _P, _P, _P]),
    "mvsb200_probability_map": (c_int, [_P, _P, c_int, c_int, c_int, c_float, c_float, c_int, c_int, _P, _P]),
    "mvsb200_infer_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "mvsb200_infer": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_int, c_int, c_int,
                              POINTER(RegnetParams), c_int, c_float, c_int, _P, _P, _P, c_size_t, _P]),
    "mvsb200_infer_set_stage_events": (c_int, [POINTER(c_void_p)]),
    "mvsb200_conv3d_plan": (c_int, [c_int] * 10 + [_P, ctypes.c_char_p, c_int]),
    "mvsb200_slab_workspace_bytes": (c_size_t, [c_int] * 7),
    "mvsb200_slab_begin": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_int, c_int,
                                   _P, c_int, _P, c_size_t, _P]),
    "mvsb200_slab_layer": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_int, c_float, _P, _P]),
    "mvsb200_slab_layer_p2p": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_int, c_float, _P, _P, _P,
                                       ctypes.c_uint, _P]),
    "mvsb200_slab_p2p_error": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "mvsb200_slab_p2p_abort": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "mvsb200_ipc_alloc": (c_int, [c_size_t, _P]),
    "mvsb200_ipc_free": (c_int, [_P]),
    "mvsb200_ipc_export": (c_int, [_P, _P]),
    "mvsb200_ipc_open": (c_int, [_P, _P]),
    "mvsb200_ipc_close": (c_int, [_P]),
    "mvsb200_regress_partial": (c_int, [_P, c_int, c_int, c_int, c_int, c_float, c_float, c_int, _P, _P]),
    "mvsb200_regress_combine": (c_int, [_P, c_int, _P, c_int, c_int, c_int, c_int, c_float, c_float, c_int, c_int, _P, _P,
                                        _P]),
    "mvsb200_slab_regions": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "mvsb200_infer_cost_offsets": (c_int, [c_int] * 7 + [_P, _P]),
    "mvsb200_infer_filtered_offset": (c_int, [c_int] * 7 + [_P]),
    "mvsb200_infer_host_staging_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "mvsb200_infer_host": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_int, c_int, c_int,
                                   POINTER(RegnetParams), c_int, c_float, c_int, _P, _P, _P, _P, c_size_t, _P]),
    "mvsb200_infer_host_async": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_int, c_int, c_int,
                                   POINTER(RegnetParams), c_int, c_float, c_int, _P, _P, _P, _P, c_size_t, _P]),
    "mvsb200_infer_host_pipelined": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_int, c_int, c_int,
                                   POINTER(RegnetParams), c_int, c_float, c_int, _P, _P, _P, _P, c_size_t, _P, _P]),
    "mvsb200_umma_probe": (c_int, [_P, c_int, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                   c_int, _P, _P]),
    "mvsb200_train_workspace_bytes": (c_size_t, [c_int] * 6),
    "mvsb200_train_step": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_int,
                                   POINTER(RegnetParams), c_int, c_float, POINTER(RegnetGrads), _P, _P, _P, _P, c_size_t, _P]),
    "mvsb200_resize_bilinear": (c_int, [_P, c_int, c_int, c_int, c_int, _P, c_int, c_int, c_float, c_float, _P]),
    "mvsb200_scale_add": (c_int, [_P, c_float, _P, c_size_t, _P, _P, _P]),
    "mvsb200_conv2d_bias": (c_int, [_P, c_int, _P, c_int, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "mvsb200_launch_count": (c_uint64, []),
    "mvsb200_set_tuning": (c_int, [c_char_p, c_char_p]),
    "mvsb200_cost_volume_window_stats": (c_int, [POINTER(c_uint64), c_int]),
}

_lib = None


def load():
    """Load the shared library (once) and bind every declared symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m mvsnet_b200.build` "
            "(nvcc, sm_100a).  mvsnet_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def set_tuning(name: str, value=None) -> None:
    """Development / tuning switch of the library (MVSB200_<name> in the environment is only its initial value)."""
    v = None if value is None else str(int(value) if isinstance(value, bool) else value).encode()
    check(load().mvsb200_set_tuning(name.encode(), v), f"set_tuning({name})")


def last_error() -> str:
    return load().mvsb200_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = "") -> None:
    if rc != OK:
        raise MVSB200Error(f"{what or 'mvsb200 call'} failed ({rc}): {last_error()}")


def ptr(t):
    """Device (or host) address of a tensor / None."""
    if t is None:
        return None
    return c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    """Every tensor must live on the CURRENT CUDA device: the library launches on the current device and on its
    current stream, so a tensor of another GPU would be dereferenced by the wrong one (wrap the call in
    `torch.cuda.device(t.device)`, as engine.HotPath does)."""
    import torch
    cur = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise MVSB200Error("mvsnet_b200 needs CUDA tensors: there is no CPU path")
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise MVSB200Error(f"tensor on cuda:{t.device.index} but the current device is cuda:{cur}: "
                               "wrap the call in torch.cuda.device(tensor.device)")
