"""ctypes access to the plain-C restatement (oracle/mvs_oracle.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None
F = ctypes.POINTER(ctypes.c_float)


def load(build=True):
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "liboracle_c.so")
        if build and not os.path.exists(path):
            subprocess.run(["make", "-C", HERE, "-s"], check=True)
        _lib = ctypes.CDLL(path)
        _lib.oracle_num_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(F)


def homographies(cams, depth_num, depth_start, depth_step, inverse=False):
    cams = np.ascontiguousarray(cams, dtype=np.float32)
    n = cams.shape[0]
    H = np.empty((n - 1, depth_num, 3, 3), dtype=np.float32)
    load().oracle_homographies(_p(cams), n, int(depth_num), ctypes.c_float(depth_start), ctypes.c_float(depth_step),
                               int(bool(inverse)), _p(H))
    return H


def transform_coefs(H):
    H = np.ascontiguousarray(H, dtype=np.float32).reshape(-1, 9)
    T = np.empty((H.shape[0], 8), dtype=np.float32)
    for i in range(H.shape[0]):
        load().oracle_transform_coefs(_p(H[i]), _p(T[i]))
    return T


def transform_warp(img, t):
    img = np.ascontiguousarray(img, dtype=np.float32)
    t = np.ascontiguousarray(t, dtype=np.float32)
    h, w, c = img.shape
    out = np.empty_like(img)
    coords = np.empty((h, w, 2), dtype=np.float32)
    load().oracle_transform_warp(_p(img), h, w, c, _p(t), _p(out), _p(coords))
    return out, coords


def cost_volume(feats, H, order="mem"):
    feats = np.ascontiguousarray(feats, dtype=np.float32)
    H = np.ascontiguousarray(H, dtype=np.float32)
    n, hf, wf, c = feats.shape
    d = H.shape[1]
    out = np.empty((d, hf, wf, c), dtype=np.float32)
    load().oracle_cost_volume(_p(feats), _p(H), n, d, hf, wf, c, 0 if order == "mem" else 1, _p(out))
    return out


def depth_regress(F_, depth_start, depth_interval, num_buckets=4):
    F_ = np.ascontiguousarray(F_, dtype=np.float32)
    d, hf, wf = F_.shape
    depth = np.empty((hf, wf), dtype=np.float32)
    prob = np.empty((hf, wf), dtype=np.float32)
    P = np.empty_like(F_)
    load().oracle_depth_regress(_p(F_), d, hf * wf, ctypes.c_float(depth_start), ctypes.c_float(depth_interval),
                                int(num_buckets), _p(depth), _p(prob), _p(P))
    return depth, prob, P
