"""The C-ABI library loads and exports every symbol include/mvsnet_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "mvsnet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mvsb200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from mvsnet_b200 import _lib
    from mvsnet_b200.build import build_library
    build_library()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert set(names) == set(_lib.SIGNATURES), "ctypes SIGNATURES out of sync with the header"


def test_no_compute_errors_are_reported_without_gpu():
    from mvsnet_b200 import _lib
    lib = _lib.load()
    assert lib.mvsb200_version() >= 100
    # argument validation happens before any CUDA call
    rc = lib.mvsb200_homographies(None, 2, 4, 1.0, 1.0, 0, None, None, None)
    assert rc == -1 and b"cams" in lib.mvsb200_last_error()
    rc = lib.mvsb200_depth_regress(None, 4, 4, 4, 1.0, 1.0, 0, 3, None, None, None, None)
    assert rc == -1
    assert lib.mvsb200_regnet_workspace_bytes(16, 24, 32, 32, 8, 1) > 0
    assert lib.mvsb200_infer_workspace_bytes(1, 16, 24, 32, 32, 8, 1) == 0


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "mvsnet_b200")
    for dp, _dn, fn in os.walk(pkg):
        for f in fn:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_python_mirror_signatures_match_reference():
    import inspect
    from mvsnet_b200 import homography_warping as hw, model
    assert list(inspect.signature(hw.get_homographies).parameters) == [
        "left_cam", "right_cam", "depth_num", "depth_start", "depth_interval", "batch_index"]
    assert list(inspect.signature(hw.get_homographies_inv_depth).parameters) == [
        "left_cam", "right_cam", "depth_num", "depth_start", "depth_end"]
    assert list(inspect.signature(hw.tf_transform_homography).parameters) == ["input_image", "homography"]
    assert list(inspect.signature(hw.homography_warping).parameters) == ["input_image", "homography"]
    assert list(inspect.signature(hw.interpolate).parameters) == ["image", "x", "y"]
    assert list(inspect.signature(model.inference_mem).parameters) == [
        "images", "cams", "depth_num", "depth_start", "depth_interval", "network_mode", "is_master_gpu", "training",
        "trainable", "inverse_depth"]
    assert list(inspect.signature(model.inference).parameters) == [
        "images", "cams", "depth_num", "depth_start", "depth_interval", "network_mode", "is_master_gpu", "trainable",
        "inverse_depth"]
    assert list(inspect.signature(model.get_probability_map_slice).parameters) == [
        "cv", "depth_map", "depth_start", "depth_interval", "inverse_depth", "num_buckets"]
