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


@pytest.mark.parametrize("config", ["cfg1", "cfg2", "cfg5"])
def test_ops_skip_only_zero_column_blocks(config):
    """Every MMA of a step keeps N a multiple of 16 inside the launch's N and starts on a 16-column boundary, the first
    op of a step is full width (it overwrites the accumulators); layers whose B images have all-zero column blocks
    (z-folded convs, merged transposed-conv classes) issue fewer columns than ops x N, and MVSB200_TC_TRIM=0 turns the
    skipping off."""
    import re
    cfg = synthetic.CONFIGS[config]
    D, hf, wf = cfg["depth_num"], cfg["height"] // 4, cfg["width"] // 4
    seen_less = 0
    for name, (cin, cout, op, stride) in synthetic.regnet_channels(32, 8).items():
        lv = LEVEL[name]
        args = (D >> lv, hf >> lv, wf >> lv, cin, cout, stride, op == "deconv", name in SKIP,
                name != "3dconv0_1" and name != "3dconv1_0")
        p, text = plan(*args)
        for line in text.strip().splitlines():
            m = re.search(r"ops (\d+) .* cols (\d+)/(\d+) ok (\d)", line)
            assert m, line
            nops, cols, full, ok = map(int, m.groups())
            assert ok == 1, (name, line)
            assert full == nops * p["mma_n"] and 16 * nops <= cols <= full, (name, line)
            zfold = int(re.search(r"zfold (\d+)", line).group(1))
            merged = op == "deconv" and cout <= 16
            if zfold > 1 and p["mma_n"] >= 32 or merged:
                assert cols < full, (name, line)
                seen_less += 1
            else:
                assert cols == full, (name, line)
    assert seen_less >= 2
    L.set_tuning("TC_TRIM", 0)
    try:
        p, text = plan(D >> 1, hf >> 1, wf >> 1, 16, 8, 2, True, True, True)       # 3dconv6_0
        m = re.search(r"cols (\d+)/(\d+) ok (\d)", text)
        assert m and m.group(1) == m.group(2) and m.group(3) == "1", text
    finally:
        L.set_tuning("TC_TRIM", None)


def test_plan_rejects_unsupported_channels():
    with pytest.raises(L.MVSB200Error, match="Cin"):
        plan(8, 16, 16, 24, 8, 1, False, False, False)


# ---- tensor-core image tower (csrc/feature2d_tc.cu): host-side plans of the 32 layers ------------------------------
TOWER = ["kind", "nch", "cs", "slices", "n", "mb", "ops", "smem", "tmem", "tiles", "wbytes", "obuf"]


@pytest.mark.parametrize("config", ["cfg1", "cfg2", "cfg5"])
def test_tower_plans_fit_the_sm(config):
    import oracle.feature_oracle as FO
    cfg = synthetic.CONFIGS[config]
    lib = L.load()
    specs = FO.unet_layer_specs(8)
    assert len(specs) == 32
    for i, (name, op, k, stride, cin, cout, srcs, gn, relu) in enumerate(specs):
        nums = (ctypes.c_int * 12)()
        L.check(lib.mvsb200_unet_tc_plan(cfg["n_views"], cfg["height"], cfg["width"], 8, i, nums), "unet_tc_plan")
        p = dict(zip(TOWER, list(nums)))
        kind = 4 if op == "deconv" else (1 if stride == 1 else (2 if k == 3 else 3))
        assert p["kind"] == kind, (name, p)
        assert p["nch"] == max(cin, 8) // 8 and p["cs"] * p["slices"] == cout, (name, p)      # the 3-channel images fill one chunk
        assert p["n"] % 16 == 0 and p["n"] >= p["cs"] and p["mb"] in (1, 2), (name, p)
        assert 0 < p["smem"] <= 220 * 1024 and p["wbytes"] <= 160 * 1024, (name, p)
        assert p["tmem"] in (32, 64, 128, 256, 512) and p["tmem"] >= p["mb"] * (4 if kind == 4 else 1) * p["n"], (name, p)
        assert p["ops"] <= 72 and p["wbytes"] == p["ops"] * 2 * p["n"] * 16, (name, p)
        # one MMA per (tap, channel pair) -- or per pair of taps when the input is a single chunk
        taps = {1: 9, 2: 9, 3: 25, 4: 9}[kind]
        if p["nch"] == 1:
            assert p["ops"] == ({1: 5, 2: 5, 3: 13}[kind] if kind != 4 else 2 + 1 + 1 + 1), (name, p)
        else:
            assert p["ops"] == taps * p["nch"] // 2, (name, p)


def test_tower_plan_rejects_bad_shapes():
    lib = L.load()
    nums = (ctypes.c_int * 12)()
    assert lib.mvsb200_unet_tc_plan(5, 100, 160, 8, 0, nums) != 0          # H not a multiple of 16
    assert lib.mvsb200_unet_tc_plan(5, 96, 160, 4, 0, nums) != 0           # groups of 8 channels need base_filter % 8 == 0
    assert lib.mvsb200_unet_tc_workspace_bytes(5, 100, 160, 8) == 0
    assert lib.mvsb200_unet_tc_workspace_bytes(5, 96, 160, 8) > 0
