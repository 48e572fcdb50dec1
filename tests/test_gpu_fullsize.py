"""BASELINE config-2 size (5 views 1152x864, D=192: 11.9 M voxels), where the oracle is too slow: size-independent
properties of the path -- exact scaling of a layer with power-of-two weights, translation consistency of the cost
volume against a small crop the oracle can do, agreement of the two cost-volume layouts, bounds and the
double-count rule of the regression, determinism of the whole path."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from mvsnet_b200 import synthetic  # noqa: E402

CFG = synthetic.CONFIGS["cfg2"]
N, D = CFG["n_views"], CFG["depth_num"]
HF, WF = CFG["height"] // 4, CFG["width"] // 4


@pytest.fixture(scope="module")
def big():
    cams = synthetic.make_cameras(N, CFG["height"], CFG["width"], D, CFG["interval_scale"])
    g = torch.Generator(device="cuda").manual_seed(7)
    feats = torch.randn((N, HF, WF, 32), device="cuda", generator=g)
    return dict(cams=torch.from_numpy(cams).cuda(), cams_np=cams, feats=feats, ds=float(cams[0, 1, 3, 0]),
                di=float(cams[0, 1, 3, 1]))


def test_cost_volume_crop_against_oracle(big):
    """Full-size fused warp+variance: a 6-plane x 8x8-pixel crop equals the oracle run on the same inputs."""
    import oracle as O
    from mvsnet_b200 import ops
    H = ops.homographies(big["cams"], D, big["ds"], big["di"])
    cost = ops.cost_volume(big["feats"], H)                     # fp32, NDHWC, reference op order
    assert tuple(cost.shape) == (D, HF, WF, 32)
    planes = [0, 1, 95, 96, 190, 191]
    Hn = H.cpu().numpy()[:, planes]
    ref = O.cost_volume(big["feats"].cpu().numpy(), Hn)         # [6, HF, WF, 32]
    for (y0, x0) in ((0, 0), (100, 140), (HF - 8, WF - 8)):
        got = cost[planes][:, y0:y0 + 8, x0:x0 + 8].cpu().numpy()
        assert np.abs(got - ref[:, y0:y0 + 8, x0:x0 + 8]).max() <= 2e-5


def test_planar_cost_volume_matches_ndhwc(big, tuning):
    """Product mode writes the volume chunk-planar (and, when 3dconv1_0 runs as a launch of its own, parity-split) with
    fp16 taps: both copies hold the same cells and agree with the fp32-tap bf16 volume within the bf16 + fp16-blend
    tolerance."""
    from mvsnet_b200 import ops
    from mvsnet_b200.engine import HotPath
    tuning("TC_FUSE01", 0)
    eng = HotPath(N, D, HF, WF, synthetic.make_regnet_weights(), precision="bf16")
    H = ops.homographies(big["cams"], D, big["ds"], big["di"])
    ref = ops.cost_volume(big["feats"], H, out_dtype=torch.bfloat16).float()
    cp8, ps8 = eng.cost_volume_planar(big["feats"], big["cams"], big["ds"], big["di"])
    a = cp8.view(D, 4, HF, WF, 8).permute(0, 2, 3, 1, 4).reshape(D, HF, WF, 32).float()
    b = ps8.view(D, 4, 2, 2, HF // 2, WF // 2, 8)               # [z, chunk, y parity, x parity, ys, xs, 8]
    b = b.permute(0, 4, 2, 5, 3, 1, 6).reshape(D, HF, WF, 32).float()
    assert torch.equal(a, b)
    err = (a - ref).abs()
    scale = ref.abs().max()
    assert float(err.max()) <= 2e-2 * float(scale) and float(err.mean()) <= 2e-3 * float(ref.abs().mean() + 1e-6)


def test_layer_scales_exactly_with_power_of_two_weights(big):
    """3dconv0_1 at full size (z- and x-fold, N = 96): doubling the weights doubles every output bit for bit, and the
    batch statistics scale by 2 and 4 (fp32 accumulation of bf16 products is exact under a power-of-two scale)."""
    from mvsnet_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn((D, HF, WF, 32), device="cuda", generator=g).to(torch.bfloat16)
    w = torch.from_numpy(synthetic.make_regnet_weights()["3dconv0_1/kernel"]).cuda()
    y1, s1 = ops.conv3d_layer(x, w, 1, False, "bf16", out_dtype=torch.float32)
    y2, s2 = ops.conv3d_layer(x, 2.0 * w, 1, False, "bf16", out_dtype=torch.float32)
    assert torch.equal(2.0 * y1, y2)
    np.testing.assert_allclose(s2[:8].cpu().numpy(), 2.0 * s1[:8].cpu().numpy(), rtol=1e-9)
    np.testing.assert_allclose(s2[8:].cpu().numpy(), 4.0 * s1[8:].cpu().numpy(), rtol=1e-9)
    # and the statistics are the sums of what was stored
    np.testing.assert_allclose(s1[:8].cpu().numpy(), y1.double().sum(dim=(0, 1, 2)).cpu().numpy(), rtol=1e-6, atol=1e-2)


def test_whole_path_is_deterministic_and_bounded(big):
    from mvsnet_b200.engine import HotPath
    eng = HotPath(N, D, HF, WF, synthetic.make_regnet_weights(), precision="bf16")
    d1, p1 = eng.infer(big["feats"], big["cams"], big["ds"], big["di"])
    d1, p1 = d1.clone(), p1.clone()
    d2, p2 = eng.infer(big["feats"], big["cams"], big["ds"], big["di"])
    # batch statistics are accumulated with fp64 atomics in a fixed set of partial sums: the order of the adds can
    # differ by an ulp of a double, far below one bf16 step of any activation
    assert float((d1 - d2).abs().max()) <= 1e-3 * big["di"]
    end = big["ds"] + (D - 1) * big["di"]
    assert float(d1.min()) >= big["ds"] - 1e-3 and float(d1.max()) <= end + 1e-3       # soft-argmin is a convex mix
    assert float(p1.min()) >= 0.0 and float(p1.max()) <= 2.0 + 1e-5                      # 4 buckets, doubles counted twice
    assert torch.isfinite(d1).all() and torch.isfinite(p1).all()


def test_fused_soft_argmin_matches_regression_kernel_at_full_size(big, tuning):
    """At this size every CTA of 3dconv6_2 walks the whole depth range, so the soft-argmin is folded into its
    epilogue; the stand-alone regression kernel on the same filtered volume must agree (see test_gpu_e2e)."""
    from mvsnet_b200.engine import HotPath
    eng = HotPath(N, D, HF, WF, synthetic.make_regnet_weights(), precision="bf16")
    d1, p1 = [t.clone() for t in eng.infer(big["feats"], big["cams"], big["ds"], big["di"])]
    tuning("NO_FUSED_REGRESS", 1)
    d2, p2 = [t.clone() for t in eng.infer(big["feats"], big["cams"], big["ds"], big["di"])]
    assert float((d1 - d2).abs().max()) <= 2e-3 * big["di"]
    assert float(((p1 - p2).abs() <= 1e-4).float().mean()) >= 0.999
