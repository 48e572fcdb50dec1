"""GPU parity of the regularizer: single layers and the whole RegNetUS0 in fp32 parity mode (CUDA-core
direct conv) and bf16 product mode (tcgen05), against the oracle (torch-CPU fp32 conv3d)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
from conftest import to_dev  # noqa: E402


@pytest.fixture(scope="module")
def O():
    import oracle
    return oracle


@pytest.fixture(scope="module")
def ops():
    from mvsnet_b200 import ops
    return ops


def bf16_round(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(torch.bfloat16).float().numpy()


LAYER_CASES = [
    # (D, H, W, Cin, Cout, stride, transposed)
    (8, 16, 24, 32, 8, 1, False),      # 3dconv0_1
    (8, 16, 24, 32, 16, 2, False),     # 3dconv1_0
    (8, 8, 16, 16, 16, 1, False),      # 3dconv1_1
    (8, 8, 16, 16, 32, 2, False),      # 3dconv2_0
    (4, 8, 8, 32, 32, 1, False),       # 3dconv2_1
    (4, 8, 8, 32, 64, 2, False),       # 3dconv3_0
    (3, 5, 6, 64, 64, 1, False),       # 3dconv3_1 (odd extents)
    (3, 5, 6, 64, 32, 2, True),        # 3dconv4_0
    (4, 8, 8, 32, 16, 2, True),        # 3dconv5_0
    (8, 8, 16, 16, 8, 2, True),        # 3dconv6_0
    (8, 16, 24, 8, 1, 1, False),       # 3dconv6_2
    (7, 9, 11, 32, 16, 2, False),      # stride 2 on odd extents: TF SAME pads (1,1)
]


def _layer_ref(O, x, w, stride, transposed):
    return O.conv3d_transpose_same(x, w) if transposed else O.conv3d_same(x, w, stride)


@pytest.mark.parametrize("case", LAYER_CASES)
def test_fp32_layer_vs_oracle(ops, O, case):
    D, H, W, cin, cout, stride, tr = case
    rng = np.random.RandomState(21)
    x = rng.randn(D, H, W, cin).astype(np.float32)
    w = (rng.randn(*((3, 3, 3, cout, cin) if tr else (3, 3, 3, cin, cout))) * 0.1).astype(np.float32)
    y, stats = ops.conv3d_layer(to_dev(x), to_dev(w), stride, tr, "fp32")
    ref = _layer_ref(O, x, w, stride, tr)
    assert tuple(y.shape) == ref.shape
    err = np.abs(y.cpu().numpy() - ref).max()
    assert err <= 2e-4 * max(1.0, np.abs(ref).max()), err
    st = stats.cpu().numpy()
    r64 = ref.reshape(-1, cout).astype(np.float64)
    np.testing.assert_allclose(st[:cout], r64.sum(0), rtol=1e-4, atol=1e-2)
    np.testing.assert_allclose(st[cout:], (r64 ** 2).sum(0), rtol=1e-4, atol=1e-2)


def test_fp32_layer_with_affine_relu_and_skip(ops, O):
    rng = np.random.RandomState(22)
    D, H, W, cin, cout = 4, 8, 8, 16, 8
    x = rng.randn(D, H, W, cin).astype(np.float32)
    sk = rng.randn(D, H, W, cin).astype(np.float32)
    xs, xb, ss, sb = (rng.uniform(0.5, 1.5, cin).astype(np.float32) for _ in range(4))
    w = (rng.randn(3, 3, 3, cout, cin) * 0.1).astype(np.float32)
    inp = np.maximum(x * xs + xb, 0) + np.maximum(sk * ss + sb, 0)
    ref = O.conv3d_transpose_same(inp.astype(np.float32), w)
    y, _ = ops.conv3d_layer(to_dev(x), to_dev(w), 2, True, "fp32", x_affine=(to_dev(xs), to_dev(xb)),
                            skip=to_dev(sk), skip_affine=(to_dev(ss), to_dev(sb)))
    assert np.abs(y.cpu().numpy() - ref).max() <= 3e-4 * max(1.0, np.abs(ref).max())


def test_bn_finalize(ops, O):
    rng = np.random.RandomState(23)
    x = (rng.randn(4, 6, 8, 5) * 2 + 0.5).astype(np.float32)
    g = rng.uniform(0.5, 1.5, 5).astype(np.float32)
    b = rng.randn(5).astype(np.float32)
    x64 = x.reshape(-1, 5).astype(np.float64)
    stats = to_dev(np.concatenate([x64.sum(0), (x64 ** 2).sum(0)]))
    scale, shift = ops.bn_finalize(stats, to_dev(g), to_dev(b), x64.shape[0])
    y = np.maximum(x * scale.cpu().numpy() + shift.cpu().numpy(), 0)
    np.testing.assert_allclose(y, O.batch_norm_train(x, g, b), rtol=1e-5, atol=1e-5)


def _regnet(precision, p, cost):
    from mvsnet_b200.engine import HotPath
    eng = HotPath(p["n_views"], cost.shape[0], cost.shape[1], cost.shape[2], p["weights"], precision=precision)
    c = to_dev(cost)
    if precision == "bf16":
        c = c.to(torch.bfloat16)
    return eng.regnet(c).cpu().numpy()


@pytest.mark.parametrize("which", ["tiny", "small"])
def test_regnet_fp32_vs_oracle(O, tiny_problem, small_problem, golden_tiny, which):
    p = tiny_problem if which == "tiny" else small_problem
    H = np.stack([O.get_homographies(p["cams"][0:1], p["cams"][v:v + 1], p["depth_num"], p["depth_start"],
                                     p["depth_interval"])[0] for v in range(1, p["n_views"])])
    cost = O.cost_volume(p["feats"], H)
    ref = O.regnet_us0(cost, p["weights"])
    out = _regnet("fp32", p, cost)
    scale = np.abs(ref).max()
    assert np.abs(out - ref).max() <= 2e-3 * scale, (np.abs(out - ref).max(), scale)
    if which == "tiny":
        assert np.abs(out - golden_tiny["filtered"]).max() <= 2e-3 * scale


def test_regnet_rejects_bad_extents(tiny_problem):
    from mvsnet_b200._lib import MVSB200Error
    from mvsnet_b200.engine import HotPath
    p = tiny_problem
    eng = HotPath(3, 16, 24, 32, p["weights"], precision="fp32")
    with pytest.raises(MVSB200Error, match="multiples of 8"):
        eng.regnet(torch.zeros((12, 24, 32, 32), device="cuda"))


# ---- bf16 / tcgen05 product mode ----------------------------------------------------------------------------
@pytest.mark.parametrize("case", LAYER_CASES)
def test_bf16_layer_vs_oracle(ops, O, case):
    D, H, W, cin, cout, stride, tr = case
    rng = np.random.RandomState(31)
    x = bf16_round(rng.randn(D, H, W, cin))
    w = (rng.randn(*((3, 3, 3, cout, cin) if tr else (3, 3, 3, cin, cout))) * 0.1).astype(np.float32)
    y, stats = ops.conv3d_layer(to_dev(x).to(torch.bfloat16), to_dev(w), stride, tr, "bf16", out_dtype=torch.float32)
    ref = _layer_ref(O, x, bf16_round(w), stride, tr)             # same bf16 operands, fp32 accumulation
    assert tuple(y.shape) == ref.shape
    err = np.abs(y.cpu().numpy() - ref).max()
    assert err <= 1e-3 * max(1.0, np.abs(ref).max()), f"{case}: max abs err {err}"
    st = stats.cpu().numpy()
    r64 = ref.reshape(-1, cout).astype(np.float64)
    np.testing.assert_allclose(st[:cout], r64.sum(0), rtol=1e-3, atol=0.5)
    np.testing.assert_allclose(st[cout:], (r64 ** 2).sum(0), rtol=1e-3, atol=0.5)


def test_bf16_layer_with_affine_relu_and_skip(ops, O):
    rng = np.random.RandomState(32)
    D, H, W, cin, cout = 4, 8, 8, 16, 8
    x = bf16_round(rng.randn(D, H, W, cin))
    sk = bf16_round(rng.randn(D, H, W, cin))
    xs, xb, ss, sb = (rng.uniform(0.5, 1.5, cin).astype(np.float32) for _ in range(4))
    w = (rng.randn(3, 3, 3, cout, cin) * 0.1).astype(np.float32)
    inp = bf16_round(np.maximum(x * xs + xb, 0) + np.maximum(sk * ss + sb, 0))
    ref = O.conv3d_transpose_same(inp, bf16_round(w))
    y, _ = ops.conv3d_layer(to_dev(x).to(torch.bfloat16), to_dev(w), 2, True, "bf16",
                            x_affine=(to_dev(xs), to_dev(xb)), skip=to_dev(sk).to(torch.bfloat16),
                            skip_affine=(to_dev(ss), to_dev(sb)), out_dtype=torch.float32)
    assert np.abs(y.cpu().numpy() - ref).max() <= 2e-2 * max(1.0, np.abs(ref).max())


def test_regnet_bf16_vs_oracle(O, small_problem):
    """Whole bf16 RegNetUS0 against (a) the oracle run with bf16-rounded conv operands (round_fn: the arithmetic this
    mode implements -- what is left is the bf16 storage of the raw layer outputs and summation order) and (b) the plain
    fp32 oracle (the distance the bf16 operands themselves cost)."""
    p = small_problem
    H = np.stack([O.get_homographies(p["cams"][0:1], p["cams"][v:v + 1], p["depth_num"], p["depth_start"],
                                     p["depth_interval"])[0] for v in range(1, p["n_views"])])
    cost = bf16_round(O.cost_volume(p["feats"], H))
    out = _regnet("bf16", p, cost)
    ref_model = O.regnet_us0(cost, p["weights"], round_fn=bf16_round)
    ref = O.regnet_us0(cost, p["weights"])
    rel_model = np.abs(out - ref_model).max() / np.abs(ref_model).max()
    rms_model = np.sqrt(np.mean((out - ref_model) ** 2)) / np.sqrt(np.mean(ref_model ** 2))
    rel = np.abs(out - ref).max() / np.abs(ref).max()
    print(f"bf16 RegNetUS0 vs bf16-operand oracle: max {rel_model:.4f} rms {rms_model:.5f} of range; vs fp32 oracle: max {rel:.4f}")
    # measured on B200: 0.0064 / 0.0034 / 0.0058
    assert rel_model <= 0.015, rel_model
    assert rms_model <= 0.006, rms_model
    assert rel <= 0.02, rel
    assert np.corrcoef(out.ravel(), ref.ravel())[0, 1] >= 0.999


@pytest.mark.parametrize("mode,base_filter", [("lite", 4), ("ultralite", 2)])
def test_regnet_bf16_narrow_network_modes(O, small_problem, mode, base_filter):
    """network_mode lite / ultralite (network.py:75-85; train.py:80 defaults to lite): 4 / 2 base filters.  The tensor-core
    path pads every channel count to whole 8-channel cells (zero weights, zero scale / shift); against the oracle run
    with bf16-rounded operands, and through the reference-named entry point."""
    from mvsnet_b200 import model, synthetic
    from mvsnet_b200.cnn_wrapper import mvsnetworks
    from mvsnet_b200.engine import HotPath, regnet_base_filter
    assert regnet_base_filter(mode) == base_filter and regnet_base_filter("semilite") == 8
    p = small_problem
    weights = synthetic.make_regnet_weights(32, base_filter, seed=7)
    H = np.stack([O.get_homographies(p["cams"][0:1], p["cams"][v:v + 1], p["depth_num"], p["depth_start"],
                                     p["depth_interval"])[0] for v in range(1, p["n_views"])])
    cost = bf16_round(O.cost_volume(p["feats"], H))
    ref = O.regnet_us0(cost, weights, round_fn=bf16_round)
    eng = HotPath(p["n_views"], p["depth_num"], p["hf"], p["wf"], weights, precision="bf16")
    out = eng.regnet(to_dev(cost).to(torch.bfloat16)).cpu().numpy()
    rel = np.abs(out - ref).max() / np.abs(ref).max()
    print(f"{mode}: bf16 RegNetUS0 (base filter {base_filter}) vs bf16-operand oracle: max {rel:.4f} of range")
    assert rel <= 0.02, rel
    # whole path by its reference name
    rd, _ = O.inference_from_features(p["feats"], p["cams"], p["depth_num"], p["depth_start"], p["depth_interval"], weights)
    mvsnetworks.set_variables(weights)
    model.FLAGS.view_num, model.FLAGS.precision = p["n_views"], "bf16"
    depth, _ = model.inference_mem(to_dev(p["feats"])[None], to_dev(p["cams"])[None], p["depth_num"],
                                   torch.tensor([p["depth_start"]]), torch.tensor([p["depth_interval"]]), mode)
    frac = float(np.mean(np.abs(depth[0, :, :, 0].cpu().numpy() - rd) <= 0.1 * p["depth_interval"]))
    assert frac >= 0.95, frac
    with pytest.raises(ValueError):            # the checkpoint's width must match the mode
        model.inference_mem(to_dev(p["feats"])[None], to_dev(p["cams"])[None], p["depth_num"],
                            torch.tensor([p["depth_start"]]), torch.tensor([p["depth_interval"]]), "normal")


@pytest.mark.parametrize("xfold", [0, 1])
@pytest.mark.parametrize("zf", [1, 2, 4])
@pytest.mark.parametrize("case", [(8, 16, 24, 32, 8, 1, False), (7, 9, 11, 16, 16, 1, False), (10, 16, 24, 8, 1, 1, False),
                                  (18, 20, 40, 32, 8, 1, False), (8, 16, 24, 32, 16, 2, False),
                                  (7, 9, 11, 32, 16, 2, False)])
def test_bf16_zfold_variants(ops, O, tuning, case, zf, xfold):
    """z-fold (several output planes per MMA N, master B images) and x-fold (kw taps in N, shuffled epilogue)
    against the oracle, incl. ragged D and tiles wider than the volume."""
    D, H, W, cin, cout, stride, tr = case
    if stride == 2 and zf != 1:
        pytest.skip("z-fold applies to stride-1 convs only")
    if zf * cout > 32 or (xfold and stride != 1):
        pytest.skip("fold does not apply")
    tuning("TC_ZF", zf)
    tuning("TC_XFOLD", xfold)
    rng = np.random.RandomState(41)
    x = bf16_round(rng.randn(D, H, W, cin))
    w = (rng.randn(3, 3, 3, cin, cout) * 0.1).astype(np.float32)
    y, stats = ops.conv3d_layer(to_dev(x).to(torch.bfloat16), to_dev(w), stride, tr, "bf16", out_dtype=torch.float32)
    ref = _layer_ref(O, x, bf16_round(w), stride, tr)
    assert tuple(y.shape) == ref.shape
    err = np.abs(y.cpu().numpy() - ref).max()
    assert err <= 1e-3 * max(1.0, np.abs(ref).max()), f"{case} zf={zf} xfold={xfold}: max abs err {err}"
    st = stats.cpu().numpy()
    r64 = ref.reshape(-1, cout).astype(np.float64)
    np.testing.assert_allclose(st[:cout], r64.sum(0), rtol=1e-3, atol=0.5)
    np.testing.assert_allclose(st[cout:], (r64 ** 2).sum(0), rtol=1e-3, atol=0.5)


@pytest.mark.parametrize("case", [(8, 16, 24, 32, 8, 1, False), (7, 9, 11, 16, 16, 2, False), (6, 10, 12, 32, 16, 2, True)])
def test_bf16_layer_bf16_output_layouts(ops, O, case):
    """bf16 output goes through the chunk-planar layout and back to NDHWC (stand-alone entry)."""
    D, H, W, cin, cout, stride, tr = case
    rng = np.random.RandomState(52)
    x = bf16_round(rng.randn(D, H, W, cin))
    w = (rng.randn(*((3, 3, 3, cout, cin) if tr else (3, 3, 3, cin, cout))) * 0.1).astype(np.float32)
    y, _ = ops.conv3d_layer(to_dev(x).to(torch.bfloat16), to_dev(w), stride, tr, "bf16")
    assert y.dtype == torch.bfloat16
    ref = _layer_ref(O, x, bf16_round(w), stride, tr)
    assert tuple(y.shape) == ref.shape
    err = np.abs(y.float().cpu().numpy() - ref).max()
    assert err <= 1e-2 * max(1.0, np.abs(ref).max()), f"{case}: max abs err {err}"


@pytest.mark.parametrize("base_filter", [8, 4])
def test_regnet_bf16_first_layers_in_one_launch(O, small_problem, tuning, base_filter):
    """3dconv0_1 and 3dconv1_0 as ONE launch over the cost volume (the stride-2 layer rides on the odd positions of
    the stride-1 layer's MMAs; default) against the two separate launches (tuning TC_FUSE01=0): same raw layer outputs up
    to the bf16 rounding of differently ordered fp32 sums, same batch-norm scale / shift, same filtered volume."""
    import ctypes
    from mvsnet_b200 import synthetic
    from mvsnet_b200.engine import HotPath
    p = small_problem
    weights = p["weights"] if base_filter == 8 else synthetic.make_regnet_weights(32, base_filter, seed=7)
    D, hf, wf = p["depth_num"], p["hf"], p["wf"]
    rng = np.random.RandomState(5)
    cost = to_dev(np.abs(rng.randn(D, hf, wf, 32)).astype(np.float32)).to(torch.bfloat16)
    eng = HotPath(p["n_views"], D, hf, wf, weights, precision="bf16")
    got = {}
    for fuse in (0, 1):
        tuning("TC_FUSE01", fuse)
        out = eng.regnet(cost).clone()
        ws = eng._last_regnet_ws
        layers = {}
        for name, layer, lvl, c in (("3dconv1_0", 0, 1, max(2 * base_filter, 8)), ("3dconv0_1", 3, 0, 8)):
            raw, sc, sh = eng.regnet_layer_raw(layer, D, hf, wf)
            off = raw - ws.data_ptr()
            n = (D >> lvl) * (hf >> lvl) * (wf >> lvl) * c
            t = ws[off:off + 2 * n].view(torch.bfloat16).float().clone()
            soff, hoff = sc - ws.data_ptr(), sh - ws.data_ptr()
            layers[name] = (t, ws[soff:soff + 4 * c].view(torch.float32).clone(), ws[hoff:hoff + 4 * c].view(torch.float32).clone())
        got[fuse] = (out, layers)
    for name in ("3dconv1_0", "3dconv0_1"):
        a, b = got[0][1][name], got[1][1][name]
        err = (a[0] - b[0]).abs()
        assert float(a[0].abs().max()) > 0.1
        assert float((err <= 2.0 ** -7 * a[0].abs() + 1e-3).float().mean()) == 1.0, (name, float(err.max()))
        assert float((err == 0).float().mean()) >= 0.98, name
        ct = base_filter * (2 if name == "3dconv1_0" else 1)        # the layer's own channels (the rest is padding)
        assert torch.allclose(a[1][:ct], b[1][:ct], rtol=2e-3, atol=1e-5), name
        assert torch.allclose(a[2][:ct], b[2][:ct], rtol=2e-3, atol=2e-4), name
    # the filtered volumes differ by what the one-ulp flips above grow into through nine more bf16 layers
    scale = float(got[0][0].abs().max())
    diff = got[0][0] - got[1][0]
    assert float(diff.abs().max()) <= 2e-2 * scale and float(diff.pow(2).mean().sqrt()) <= 2e-3 * scale
