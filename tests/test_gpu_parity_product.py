"""Parity of what bench.py actually runs -- bf16 product mode at BASELINE configs 1 and 2 -- against the CPU oracle.

* whole path: depth within 0.1 depth-interval of the oracle on >= 95 % of pixels (bf16 product mode; the achieved
  share is printed) and >= 99.9 % in fp32 parity mode at config 1 (north_star gates), model.py:374-502;
* product-mode cost volume (shared-memory window kernel, fp16 taps, planar bf16 output) read back from the workspace
  against oracle.cost_volume (model.py:423-463), with a tolerance derived below;
* the soft-argmin fused into 3dconv6_2's epilogue + the probability gather against oracle.depth_regress
  (model.py:472-498, :45-144) on the very volume the GPU regressed.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from conftest import to_dev  # noqa: E402


def _frac_within(depth, ref, interval, tol=0.1):
    return float(np.mean(np.abs(depth - ref) <= tol * interval))


def _engine(p, precision):
    from mvsnet_b200.engine import HotPath
    return HotPath(p["n_views"], p["depth_num"], p["hf"], p["wf"], p["weights"], precision=precision)


def _cost_tolerance(ref_cost, fmax):
    """Per-voxel bound on |product-mode cost - fp32 reference cost|.

    Product mode reads the source views as fp16 (relative rounding 2^-12 per tap), blends the four taps in packed
    fp16 (three more roundings of 2^-12 relative to the partial sums, each <= max|tap|) and stores bf16.  So a warped
    value is off by at most e = 4 * 2^-12 * fmax (doubled below for slack: 2^-9 * fmax), and with
    cost = mean((w_v - m)^2): |d cost| <= (2/N) * sum_v |w_v - m| * e + e^2 <= 2 * e * sqrt(cost) + e^2
    (Cauchy-Schwarz), plus the bf16 rounding of the stored value (2^-9 relative, 2^-8 with slack) and the fp32
    cancellation noise of Q/N - S^2/N^2 itself (~1e-6 * fmax^2)."""
    e = 2.0 ** -9 * fmax
    c = np.maximum(ref_cost, 0.0)
    return 2.0 ** -8 * np.abs(ref_cost) + 2.0 * e * np.sqrt(c) + e * e + 2e-6 * fmax * fmax


def _planar_to_ndhwc(cp8, D, hf, wf):
    return cp8.view(D, 4, hf, wf, 8).permute(0, 2, 3, 1, 4).reshape(D, hf, wf, 32).float()


# ---------------------------------------------------------------------------------------------- whole path
def test_cfg1_fp32_and_bf16_depth_vs_oracle(oracle_cfg1):
    p, r = oracle_cfg1
    feats, cams = to_dev(p["feats"]), to_dev(p["cams"])
    for precision, need in (("fp32", 0.999), ("bf16", 0.95)):
        eng = _engine(p, precision)
        d, pm = eng.infer(feats, cams, p["depth_start"], p["depth_interval"])
        frac = _frac_within(d.cpu().numpy(), r["depth"], p["depth_interval"])
        pfrac = float(np.mean(np.abs(pm.cpu().numpy() - r["prob"]) <= 0.05))
        print(f"cfg1 {precision}: {100 * frac:.3f}% of pixels within 0.1 interval of the oracle; "
              f"probability map within 0.05 on {100 * pfrac:.2f}%")
        assert frac >= need, (precision, frac)
        del eng


def test_cfg2_bf16_depth_vs_oracle(oracle_cfg2):
    """The benchmarked configuration and instantiation (5 views 1152x864, D = 192, bf16 product mode)."""
    p, r = oracle_cfg2
    eng = _engine(p, "bf16")
    d, pm = eng.infer(to_dev(p["feats"]), to_dev(p["cams"]), p["depth_start"], p["depth_interval"])
    frac = _frac_within(d.cpu().numpy(), r["depth"], p["depth_interval"])
    frac01 = _frac_within(d.cpu().numpy(), r["depth"], p["depth_interval"], 0.01)
    pfrac = float(np.mean(np.abs(pm.cpu().numpy() - r["prob"]) <= 0.05))
    print(f"cfg2 bf16: {100 * frac:.3f}% of pixels within 0.1 interval of the oracle ({100 * frac01:.2f}% within 0.01); "
          f"probability map within 0.05 on {100 * pfrac:.2f}%")
    assert frac >= 0.95, frac
    assert pfrac >= 0.90, pfrac


# ---------------------------------------------------------------------------------------------- cost volume
@pytest.mark.parametrize("blend32", [0, 1, 2])
def test_product_cost_volume_vs_oracle_small(small_problem, tuning, blend32):
    import oracle as O
    p = small_problem
    tuning("CV_FP32_BLEND", blend32)
    eng = _engine(p, "bf16")
    cp8, ps8 = eng.cost_volume_planar(to_dev(p["feats"]), to_dev(p["cams"]), p["depth_start"], p["depth_interval"])
    D, hf, wf = p["depth_num"], p["hf"], p["wf"]
    got = _planar_to_ndhwc(cp8, D, hf, wf).cpu().numpy()
    H = np.stack([O.get_homographies(p["cams"][0:1], p["cams"][v:v + 1], D, p["depth_start"], p["depth_interval"])[0]
                  for v in range(1, p["n_views"])])
    ref = O.cost_volume(p["feats"], H)
    tol = _cost_tolerance(ref, float(np.abs(p["feats"]).max()))
    ratio = np.abs(got - ref) / tol
    print(f"small, blend32={blend32}: max |err| / tolerance = {ratio.max():.3f}, mean = {ratio.mean():.4f}")
    assert ratio.max() <= 1.0
    # the parity-split copy (only written when 3dconv1_0 runs as a launch of its own) holds the same cells
    tuning("TC_FUSE01", 0)
    cp8, ps8 = eng.cost_volume_planar(to_dev(p["feats"]), to_dev(p["cams"]), p["depth_start"], p["depth_interval"])
    b = ps8.view(D, 4, 2, 2, hf // 2, wf // 2, 8).permute(0, 4, 2, 5, 3, 1, 6).reshape(D, hf, wf, 32).float()
    assert torch.equal(b.cpu(), torch.from_numpy(got))


def test_product_cost_volume_vs_oracle_cfg2_planes(oracle_cfg2):
    """Config 2, the shipped kernel: 12 planes across the sweep, every pixel, against the oracle."""
    import oracle as O
    p, r = oracle_cfg2
    eng = _engine(p, "bf16")
    cp8, _ = eng.cost_volume_planar(to_dev(p["feats"]), to_dev(p["cams"]), p["depth_start"], p["depth_interval"])
    D, hf, wf = p["depth_num"], p["hf"], p["wf"]
    planes = [0, 1, 7, 8, 63, 64, 95, 96, 127, 128, 190, 191]
    got = cp8.view(D, 4, hf, wf, 8)[planes].permute(0, 2, 3, 1, 4).reshape(len(planes), hf, wf, 32).float().cpu().numpy()
    ref = O.cost_volume(p["feats"], r["homographies"][:, planes])
    tol = _cost_tolerance(ref, float(np.abs(p["feats"]).max()))
    ratio = np.abs(got - ref) / tol
    print(f"cfg2 planes: max |err| / tolerance = {ratio.max():.3f}, mean = {ratio.mean():.4f}")
    assert ratio.max() <= 1.0


def test_window_kernel_agrees_with_gather_kernel(small_problem, tuning):
    """The round-1 gather kernel (tuning CV_KERNEL=1) and the window kernel compute the same fp16-tap blend: the volumes
    agree to the last bf16 bit almost everywhere (the weights are rounded at different points)."""
    p = small_problem
    feats, cams = to_dev(p["feats"]), to_dev(p["cams"])
    eng = _engine(p, "bf16")
    a = eng.cost_volume_planar(feats, cams, p["depth_start"], p["depth_interval"])[0].float().clone()
    tuning("CV_KERNEL", 1)
    b = eng.cost_volume_planar(feats, cams, p["depth_start"], p["depth_interval"])[0].float().clone()
    err = (a - b).abs()
    assert float(err.max()) <= 0.02 * float(b.abs().max())
    assert float((err <= 2.0 ** -7 * b.abs() + 1e-4).float().mean()) >= 0.99


def test_window_kernel_falls_back_outside_the_window(tiny_problem, tuning):
    """Geometry the shared-memory window cannot hold (a source view rotated by 90 degrees about the optical axis: a row of
    reference pixels maps to a column of the source) still gives the oracle's volume: those voxels read global memory."""
    import oracle as O
    from mvsnet_b200 import _lib
    p = tiny_problem
    cams = p["cams"].copy()
    Rz = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1]], dtype=np.float32)
    cams[1, 0, :3, :3] = Rz @ cams[1, 0, :3, :3]
    cams[1, 0, :3, 3] = Rz @ cams[1, 0, :3, 3]
    tuning("CV_STATS", 1)
    lib = _lib.load()
    import ctypes
    n = ctypes.c_uint64()
    eng = _engine(p, "bf16")
    lib.mvsb200_cost_volume_window_stats(ctypes.byref(n), 1)
    cp8, _ = eng.cost_volume_planar(to_dev(p["feats"]), to_dev(cams), p["depth_start"], p["depth_interval"])
    lib.mvsb200_cost_volume_window_stats(ctypes.byref(n), 1)
    D, hf, wf = p["depth_num"], p["hf"], p["wf"]
    got = _planar_to_ndhwc(cp8, D, hf, wf).cpu().numpy()
    H = np.stack([O.get_homographies(cams[0:1], cams[v:v + 1], D, p["depth_start"], p["depth_interval"])[0]
                  for v in range(1, p["n_views"])])
    ref = O.cost_volume(p["feats"], H)
    tol = _cost_tolerance(ref, float(np.abs(p["feats"]).max()))
    assert (np.abs(got - ref) / tol).max() <= 1.0
    print(f"rotated view: {n.value} (voxel, view) pairs served from global memory")
    assert n.value > 0


# ---------------------------------------------------------------------------------------------- regression
@pytest.mark.parametrize("cfg", ["small", "cfg1"])
def test_fused_soft_argmin_vs_oracle(cfg, small_problem, oracle_cfg1):
    """bf16 mode regresses inside the epilogue of 3dconv6_2 (fast exp, running rescale); the oracle's softmax /
    soft-argmin / 4-bucket probability on the same filtered volume must agree."""
    import oracle as O
    p = small_problem if cfg == "small" else oracle_cfg1[0]
    eng = _engine(p, "bf16")
    d, pm = eng.infer(to_dev(p["feats"]), to_dev(p["cams"]), p["depth_start"], p["depth_interval"])
    F = eng.filtered_volume().cpu().numpy()
    rd, rp, _ = O.depth_regress(F, p["depth_start"], p["depth_interval"])
    derr = np.abs(d.cpu().numpy() - rd) / p["depth_interval"]
    print(f"{cfg}: fused soft-argmin vs oracle on the same volume: max {derr.max():.2e} intervals")
    assert derr.max() <= 2e-3
    # the bucket gather flips only where the index sits within rounding of an integer (model.py:113-120)
    assert float(np.mean(np.abs(pm.cpu().numpy() - rp) <= 1e-4)) >= 0.995
