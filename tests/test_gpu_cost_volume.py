"""GPU parity of the fused warp+variance kernel (all variants) against the oracle."""
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


def _homs(O, p, D=None):
    D = D or p["depth_num"]
    return np.stack([O.get_homographies(p["cams"][0:1], p["cams"][v:v + 1], D, p["depth_start"],
                                        p["depth_interval"])[0] for v in range(1, p["cams"].shape[0])])


@pytest.mark.parametrize("variant", [1, 2, 3])
@pytest.mark.parametrize("order", ["mem", "train"])
def test_cost_volume_variants_vs_oracle(ops, O, small_problem, variant, order):
    p = small_problem
    H = _homs(O, p)
    ref = O.cost_volume(p["feats"], H, order=order)
    out = ops.cost_volume(to_dev(p["feats"]), to_dev(H), order=order, variant=variant).cpu().numpy()
    err = np.abs(out - ref).max()
    assert err <= 2e-5, f"variant {variant} {order}: max abs err {err}"


def test_cost_volume_golden(ops, golden_tiny):
    g = golden_tiny
    for order, key in (("mem", "cost_mem_sub"), ("train", "cost_train_sub")):
        out = ops.cost_volume(to_dev(g["feats"]), to_dev(g["homographies"]), order=order).cpu().numpy()
        assert np.abs(out[::2, ::2, ::2, :] - g[key]).max() <= 2e-5


@pytest.mark.parametrize("variant", [1, 2, 3])
def test_ragged_extents(ops, O, small_problem, variant):
    """Wf not a multiple of the tile, Hf odd, D not a multiple of the 8-plane chunk."""
    p = small_problem
    feats = np.ascontiguousarray(p["feats"][:4, :37, :45, :])
    cams = p["cams"][:4]
    D = 13
    H = np.stack([O.get_homographies(cams[0:1], cams[v:v + 1], D, p["depth_start"], p["depth_interval"])[0]
                  for v in range(1, 4)])
    ref = O.cost_volume(feats, H)
    out = ops.cost_volume(to_dev(feats), to_dev(H), variant=variant).cpu().numpy()
    assert out.shape == ref.shape
    assert np.abs(out - ref).max() <= 2e-5


def test_bf16_output_is_rounded_fp32(ops, O, small_problem):
    p = small_problem
    H = to_dev(_homs(O, p))
    f = to_dev(p["feats"])
    for variant in (1, 2):
        a = ops.cost_volume(f, H, out_dtype=torch.float32, variant=variant)
        b = ops.cost_volume(f, H, out_dtype=torch.bfloat16, variant=variant)
        assert b.dtype == torch.bfloat16 and torch.equal(b, a.to(torch.bfloat16))


@pytest.mark.parametrize("channels", [16, 6])
def test_other_channel_counts_and_legacy_sampler(ops, O, tiny_problem, channels):
    p = tiny_problem
    feats = np.ascontiguousarray(p["feats"][..., :channels])
    H = _homs(O, p)
    for sampler in ("transform", "legacy"):
        ref = O.cost_volume(feats, H, sampler=sampler)
        out = ops.cost_volume(to_dev(feats), to_dev(H), sampler=sampler).cpu().numpy()
        assert np.abs(out - ref).max() <= 2e-5, (channels, sampler)


def test_eight_views_and_iid_features(ops, O):
    from mvsnet_b200 import synthetic
    cams = synthetic.make_cameras(8, 96, 128, 16, 8.0)
    feats = synthetic.make_features(cams, 24, 32, 32, iid=True)
    H = np.stack([O.get_homographies(cams[0:1], cams[v:v + 1], 16, 425.0, 20.0)[0] for v in range(1, 8)])
    ref = O.cost_volume(feats, H)
    for variant in (1, 2):
        out = ops.cost_volume(to_dev(feats), to_dev(H), variant=variant).cpu().numpy()
        assert np.abs(out - ref).max() <= 5e-5


def test_full_size_properties(ops):
    """Config-2 size: identical views give (numerically) zero variance; variants agree with each other."""
    from mvsnet_b200 import synthetic
    cfg = synthetic.CONFIGS["cfg2"]
    cams = synthetic.make_cameras(5, cfg["height"], cfg["width"], cfg["depth_num"], cfg["interval_scale"])
    hf, wf = cfg["height"] // 4, cfg["width"] // 4
    g = torch.Generator(device="cuda").manual_seed(0)
    feats = torch.randn((5, hf, wf, 32), device="cuda", generator=g)
    H = ops.homographies(to_dev(cams), 192, 425.0, 2.65)
    a = ops.cost_volume(feats, H, variant=2)
    b = ops.cost_volume(feats, H, variant=3)
    c = ops.cost_volume(feats, H, variant=1)
    assert (a - b).abs().max().item() <= 1e-5 and (a - c).abs().max().item() <= 1e-5
    assert torch.isfinite(a).all()
    eye = torch.eye(3, device="cuda").expand(4, 192, 3, 3).contiguous()
    same = feats[0:1].expand(5, hf, wf, 32).contiguous()
    z = ops.cost_volume(same, eye, variant=2)
    assert z.abs().max().item() <= 1e-5
    assert (z >= -1e-5).all()


def test_errors(ops, tiny_problem):
    from mvsnet_b200._lib import MVSB200Error
    f = to_dev(tiny_problem["feats"])
    H = torch.zeros((2, 4, 3, 3), device="cuda")
    with pytest.raises(MVSB200Error):
        ops.cost_volume(f, H, variant=9)
    with pytest.raises(ValueError):
        ops.cost_volume(f, H[:1])
    with pytest.raises(MVSB200Error):
        ops.cost_volume(f[..., :16].contiguous(), H, variant=3)      # fast path is C=32 only
