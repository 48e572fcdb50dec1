"""CPU checks of the feature-tower restatement (oracle/feature_oracle.py): known answers for the
SAME padding, the transposed convolution as the adjoint of the stride-2 convolution, group statistics over groups of
8 channels, the layer table (mvsnetworks.py:58-115) and its agreement with the product's table.  No GPU."""
import numpy as np
import pytest

import oracle.feature_oracle as FO
from mvsnet_b200 import _lib as L
from mvsnet_b200 import synthetic

F32 = np.float32


def test_layer_table_matches_the_product_table():
    names = {n: i for i, n in enumerate(L.UNET_LAYER_NAMES)}
    assert len(FO.UNET_LAYERS) == L.UNET_LAYERS == 32
    for (name, op, k, s, mult, srcs, gn, relu), prod in zip(FO.UNET_LAYERS, L.UNET_LAYER_TABLE):
        assert prod[:5] == (name, op, k, s, mult)
        assert tuple(-1 if x == "data" else names[x] for x in srcs) == prod[5]
        assert (gn, relu) == prod[6:]


def test_channel_bookkeeping():
    specs = {s[0]: s for s in FO.unet_layer_specs(8)}
    assert specs["2dconv1_0"][4:6] == (3, 16) and specs["2dconv0_1"][4:6] == (3, 8)
    assert specs["2dconv5_1"][4] == 128 and specs["2dconv6_1"][4] == 64      # concat(deconv, skip)
    assert specs["2dconv7_1"][4] == 32 and specs["2dconv8_1"][4] == 16
    assert specs["conv9_0"][2:4] == (5, 2) and specs["conv10_0"][2:4] == (5, 2)
    assert specs["conv10_2"][5] == 32 and not specs["conv10_2"][7] and not specs["conv10_2"][8]
    # deconv_gn: normalised but no ReLU (network.py:356: relu=False and the U-Net never overrides it)
    for d in ("2dconv5_0", "2dconv6_0", "2dconv7_0", "2dconv8_0"):
        assert specs[d][1] == "deconv" and specs[d][7] and not specs[d][8]
    assert [(n, ci, co) for n, _o, _k, _s, ci, co, _g in synthetic.unet_channels(8)] == \
        [(s[0], s[4], s[5]) for s in FO.unet_layer_specs(8)]


@pytest.mark.parametrize("size,k,s,pads", [(8, 3, 1, (1, 1)), (8, 3, 2, (0, 1)), (8, 5, 2, (1, 2)), (7, 3, 2, (1, 1)),
                                           (7, 5, 2, (2, 2))])
def test_same_padding(size, k, s, pads):
    assert FO.tf_same_pads(size, k, s) == pads


def test_conv5x5_stride2_known_answer():
    """y[o] = sum_k x[2 o + k - 1] w[k] on an even extent (one zero before, two after)."""
    rng = np.random.RandomState(0)
    x = rng.normal(size=(1, 6, 8, 2)).astype(F32)
    w = rng.normal(size=(5, 5, 2, 3)).astype(F32)
    y = FO.conv2d_same(x, w, 2)
    assert y.shape == (1, 3, 4, 3)
    ref = np.zeros_like(y, dtype=np.float64)
    for oy in range(3):
        for ox in range(4):
            for kh in range(5):
                for kw in range(5):
                    iy, ix = 2 * oy + kh - 1, 2 * ox + kw - 1
                    if 0 <= iy < 6 and 0 <= ix < 8:
                        ref[0, oy, ox] += x[0, iy, ix].astype(np.float64) @ w[kh, kw].astype(np.float64)
    np.testing.assert_allclose(y, ref, rtol=1e-5, atol=1e-5)


def test_transposed_conv_is_the_adjoint_of_the_stride2_conv():
    """tf.layers.conv2d_transpose(SAME, stride 2) is the gradient of conv2d(SAME, stride 2) w.r.t. its input."""
    rng = np.random.RandomState(1)
    u = rng.normal(size=(1, 8, 12, 4)).astype(F32)          # fine grid
    v = rng.normal(size=(1, 4, 6, 5)).astype(F32)           # coarse grid
    w = rng.normal(size=(3, 3, 4, 5)).astype(F32)           # conv kernel [k,k,Cin=4,Cout=5] = deconv kernel [k,k,Cout,Cin]
    lhs = float((FO.conv2d_same(u, w, 2).astype(np.float64) * v).sum())
    rhs = float((u.astype(np.float64) * FO.conv2d_transpose_same(v, w)).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs))
    assert FO.conv2d_transpose_same(v, w).shape == (1, 8, 12, 4)


def test_group_norm_groups_of_eight_channels():
    rng = np.random.RandomState(2)
    x = (rng.normal(size=(2, 5, 7, 16)) * 3.0 + 1.5).astype(F32)
    ones, zeros = np.ones(16, F32), np.zeros(16, F32)
    y = FO.group_norm(x, ones, zeros, relu=False)
    g = y.reshape(2, 35, 2, 8)
    np.testing.assert_allclose(g.mean(axis=(1, 3)), 0.0, atol=1e-5)
    np.testing.assert_allclose(g.var(axis=(1, 3)), 1.0, atol=1e-3)
    # the two groups of a sample, and the two samples, do not see each other
    x2 = x.copy()
    x2[0, :, :, 8:] *= 10.0
    y2 = FO.group_norm(x2, ones, zeros, relu=False)
    np.testing.assert_array_equal(y2[0, :, :, :8], y[0, :, :, :8])
    np.testing.assert_array_equal(y2[1], y[1])
    # gamma / beta per channel, then ReLU
    gam, bet = rng.uniform(0.5, 1.5, 16).astype(F32), rng.normal(size=16).astype(F32)
    np.testing.assert_allclose(FO.group_norm(x, gam, bet, relu=True), np.maximum(y * gam + bet, 0.0), rtol=1e-6, atol=1e-6)
    # fewer than 8 channels: one group (network.py:246-247, G = max(1, C / 8))
    y4 = FO.group_norm(x[..., :4], ones[:4], zeros[:4], relu=False)
    np.testing.assert_allclose(y4.reshape(2, -1).mean(axis=1), 0.0, atol=1e-5)


def test_tower_shapes_and_deconv_without_relu():
    w = synthetic.make_unet_weights(8)
    im = synthetic.make_images(2, 32, 48)
    f, outs = FO.unet_ds2gn(im, w, return_layers=True)
    assert f.shape == (2, 8, 12, 32) and f.dtype == F32 and np.isfinite(f).all()
    assert outs["2dconv4_2"].shape == (2, 2, 3, 128) and outs["2dconv8_2"].shape == (2, 32, 48, 8)
    assert outs["2dconv5_0"].min() < 0.0 and outs["2dconv0_1"].min() >= 0.0
    with pytest.raises(ValueError):
        FO.unet_ds2gn(im[:, :30], w)


def test_c_abi_tower_plan_is_host_only_and_matches_the_layer_table():
    """mvsb200_unet_workspace_bytes / _layer_output do no device work: extents and channel counts of every layer as the
    library lays them out equal the restatement's, offsets are disjoint and inside the workspace, bad shapes give 0."""
    import ctypes
    lib = L.load()
    n, h, w = 3, 64, 96
    total = lib.mvsb200_unet_workspace_bytes(n, h, w, 8)
    assert total > 0
    im = np.zeros((1, h, w, 3), F32)
    shapes = {"data": (h, w)}
    ends = []
    for i, (name, op, k, s, cin, cout, srcs, gn, relu) in enumerate(FO.unet_layer_specs(8)):
        ih, iw = shapes[srcs[0]]
        shapes[name] = (2 * ih, 2 * iw) if op == "deconv" else (-(-ih // s), -(-iw // s))
        off, dims = ctypes.c_size_t(), (ctypes.c_int * 3)()
        L.check(lib.mvsb200_unet_layer_output(n, h, w, 8, i, ctypes.byref(off), dims), "unet_layer_output")
        assert tuple(dims) == (*shapes[name], cout), name
        ends.append((off.value, off.value + n * dims[0] * dims[1] * dims[2] * 4))
    ends.sort()
    assert all(a[1] <= b[0] for a, b in zip(ends, ends[1:])) and ends[-1][1] <= total
    assert shapes["conv10_2"] == (h // 4, w // 4)
    for bad in [(n, 40, 96, 8), (n, 64, 100, 8), (0, 64, 96, 8), (n, 64, 96, 4)]:
        assert lib.mvsb200_unet_workspace_bytes(*bad) == 0
    assert b"multiples of 16" in lib.mvsb200_last_error() or b"base_filter" in lib.mvsb200_last_error()


def test_golden_tower_fixture_matches_the_restatement(golden_tower):
    """tests/golden/tiny_tower.npz (made by tests/golden/make_golden_tower.py) freezes the restatement's output."""
    g = golden_tower
    w = synthetic.make_unet_weights(8)
    np.testing.assert_allclose(np.array([np.float64(np.abs(w[k]).sum()) for k in sorted(w)]), g["weight_digest"], rtol=1e-12)
    np.testing.assert_array_equal(synthetic.make_images(2, 32, 48), g["images"])
    feats, outs = FO.unet_ds2gn(g["images"], w, return_layers=True)
    # torch-CPU convolutions may pick different kernels on different hosts: a few ulps, not bit equality
    np.testing.assert_allclose(feats, g["feats"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(outs["2dconv5_0"], g["l2dconv5_0"], rtol=0, atol=2e-5)
    names = [s[0] for s in FO.unet_layer_specs(8)]
    digest = np.array([[outs[n].mean(), np.abs(outs[n]).mean(), outs[n].max()] for n in names], dtype=np.float64)
    np.testing.assert_allclose(digest, g["layer_digest"], rtol=1e-4, atol=1e-5)


def test_bf16_operand_model_stays_close_to_fp32():
    """round_fn models a tensor-core tower (bf16 operands, fp32 accumulation, fp32 normalisation): on the seeded
    problem the features move by about 1 % of their range -- the tolerance budget of the round-2 implementation."""
    import torch

    def bf16(a):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=F32)).to(torch.bfloat16).to(torch.float32).numpy()

    w = synthetic.make_unet_weights(8)
    im = synthetic.make_images(1, 32, 48)
    f32 = FO.unet_ds2gn(im, w)
    f16 = FO.unet_ds2gn(im, w, round_fn=bf16)
    err = np.abs(f16 - f32)
    scale = float(np.abs(f32).max())
    assert float(err.max()) <= 0.08 * scale and float(err.mean()) <= 0.01 * scale, (float(err.max()), float(err.mean()), scale)
