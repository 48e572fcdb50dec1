"""Host logic of the bf16 / tcgen05 regularizer: the launch planner (tile, z-/x-fold, ring depth, shared and tensor
memory budgets) for every RegNetUS0 layer shape of the BASELINE configs.  No device work (runs without a GPU)."""
import ctypes

import pytest

from mvsnet_b200 import _lib as L
from mvsnet_b200 import synthetic

LEVEL = {"3dconv1_0": 0, "3dconv2_0": 1, "3dconv3_0": 2, "3dconv0_1": 0, "3dconv1_1": 1, "3dconv2_1": 2,
         "3dconv3_1": 3, "3dconv4_0": 3, "3dconv5_0": 2, "3dconv6_0": 1, "3dconv6_2": 0}
SKIP = {"3dconv5_0", "3dconv6_0", "3dconv6_2"}
SMEM_MAX = 227 * 1024
NAMES = ["launches", "tx", "ty", "px", "ry", "mb", "mma_n", "zf", "xfold", "ring", "smem", "tmem"]


def plan(d, h, w, cin, cout, stride, transposed, has_skip, transform):
    lib = L.load()
    nums = (ctypes.c_int * 12)()
    text = ctypes.create_string_buffer(2048)
    rc = lib.mvsb200_conv3d_plan(d, h, w, cin, cout, stride, int(transposed), int(has_skip), int(transform), 148, nums,
                                 text, 2048)
    L.check(rc, "conv3d_plan")
    return dict(zip(NAMES, list(nums))), text.value.decode()


@pytest.mark.parametrize("config", ["cfg1", "cfg2", "cfg5"])
def test_plans_fit_the_sm(config):
    cfg = synthetic.CONFIGS[config]
    D, hf, wf = cfg["depth_num"], cfg["height"] // 4, cfg["width"] // 4
    for name, (cin, cout, op, stride) in synthetic.regnet_channels(32, 8).items():
        lv = LEVEL[name]
        p, text = plan(D >> lv, hf >> lv, wf >> lv, cin, cout, stride, op == "deconv", name in SKIP, name != "3dconv0_1"
                       and name != "3dconv1_0")
        assert p["launches"] == (cout + 31) // 32, (name, text)
        assert 0 < p["smem"] <= SMEM_MAX, (name, text)
        assert p["tmem"] in (32, 64, 128, 256, 512), (name, text)
        assert 1 <= p["mb"] <= 4 and p["mma_n"] % 16 == 0 and 16 <= p["mma_n"] <= 256, (name, text)
        assert p["px"] * 8 <= 256 and p["ry"] <= 256, (name, text)          # TMA box limits
        assert p["ring"] >= 3, (name, text)
        if op == "conv" and stride == 1:
            assert p["xfold"] in (0, 1) and p["zf"] in (1, 2, 4)
            if p["xfold"]:
                assert p["px"] in (8, 16, 32) and (p["ty"] * p["px"]) % 128 == 0, (name, text)
        else:
            assert p["xfold"] == 0 and p["zf"] == 1, (name, text)


def test_headline_layer_folds():
    """3dconv0_1 at config 2 (60 % of the FLOPs, Cout = 8): both folds on, N = 96."""
    p, text = plan(192, 216, 288, 32, 8, 1, False, False, False)
    assert p["xfold"] == 1 and p["zf"] == 4 and p["mma_n"] == 96, text


def test_plan_rejects_unsupported_channels():
    with pytest.raises(L.MVSB200Error, match="Cin"):
        plan(8, 16, 16, 24, 8, 1, False, False, False)
