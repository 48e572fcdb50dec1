"""Whole hot path on the GPU against the oracle: depth within 0.1 depth-interval on >= 99.9 % of pixels in
fp32 parity mode and >= 95 % in bf16 product mode (north_star gates), through the device and the host-buffer
C-ABI entry points and the reference-named Python API."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
from conftest import to_dev  # noqa: E402


@pytest.fixture(scope="module")
def oracle_small(small_problem):
    import oracle as O
    p = small_problem
    depth, prob, allr = O.inference_from_features(p["feats"], p["cams"], p["depth_num"], p["depth_start"],
                                                  p["depth_interval"], p["weights"], return_all=True)
    return depth, prob, allr


def _frac_within(depth, ref, interval, tol=0.1):
    return float(np.mean(np.abs(depth - ref) <= tol * interval))


def test_fp32_path_vs_oracle(small_problem, oracle_small):
    from mvsnet_b200.engine import HotPath
    p = small_problem
    rd, rp, _ = oracle_small
    eng = HotPath(p["n_views"], p["depth_num"], p["hf"], p["wf"], p["weights"], precision="fp32")
    d, pm = eng.infer(to_dev(p["feats"]), to_dev(p["cams"]), p["depth_start"], p["depth_interval"])
    frac = _frac_within(d.cpu().numpy(), rd, p["depth_interval"])
    assert frac >= 0.999, frac
    assert np.mean(np.abs(pm.cpu().numpy() - rp) <= 1e-2) >= 0.99


def test_bf16_path_vs_oracle(small_problem, oracle_small):
    from mvsnet_b200.engine import HotPath
    p = small_problem
    rd, rp, _ = oracle_small
    eng = HotPath(p["n_views"], p["depth_num"], p["hf"], p["wf"], p["weights"], precision="bf16")
    d, pm = eng.infer(to_dev(p["feats"]), to_dev(p["cams"]), p["depth_start"], p["depth_interval"])
    frac = _frac_within(d.cpu().numpy(), rd, p["depth_interval"])
    assert frac >= 0.95, frac


def test_fused_soft_argmin_matches_regression_kernel(small_problem, tuning):
    """bf16 mode folds the soft-argmin (model.py:472-495) into the epilogue of 3dconv6_2; the tuning switch NO_FUSED_REGRESS
    runs the stand-alone regression kernel on the filtered volume instead.  Same volume, same answer: depth to a
    thousandth of an interval (fast exp, different summation order); the probability sum of the four planes around
    the estimate may pick different planes only where the index sits on an integer."""
    from mvsnet_b200.engine import HotPath
    p = small_problem
    eng = HotPath(p["n_views"], p["depth_num"], p["hf"], p["wf"], p["weights"], precision="bf16")
    feats, cams = to_dev(p["feats"]), to_dev(p["cams"])
    d1, p1 = [t.clone() for t in eng.infer(feats, cams, p["depth_start"], p["depth_interval"])]
    tuning("NO_FUSED_REGRESS", 1)
    d2, p2 = [t.clone() for t in eng.infer(feats, cams, p["depth_start"], p["depth_interval"])]
    assert float((d1 - d2).abs().max()) <= 1e-3 * p["depth_interval"]
    assert float(((p1 - p2).abs() <= 1e-4).float().mean()) >= 0.999
    assert torch.isfinite(d1).all() and torch.isfinite(p1).all()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_host_entry_point_matches_device(small_problem, precision):
    from mvsnet_b200.engine import HotPath
    p = small_problem
    eng = HotPath(p["n_views"], p["depth_num"], p["hf"], p["wf"], p["weights"], precision=precision)
    d, pm = eng.infer(to_dev(p["feats"]), to_dev(p["cams"]), p["depth_start"], p["depth_interval"])
    d, pm = d.clone(), pm.clone()
    fh = torch.from_numpy(p["feats"]).pin_memory()
    ch = torch.from_numpy(p["cams"]).pin_memory()
    dh = torch.empty((p["hf"], p["wf"])).pin_memory()
    ph = torch.empty((p["hf"], p["wf"])).pin_memory()
    eng.infer_host(fh, ch, p["depth_start"], p["depth_interval"], dh, ph)
    # batch-statistic BN sums use atomics: allow last-bit noise
    assert np.allclose(dh.numpy(), d.cpu().numpy(), rtol=1e-4, atol=1e-2)
    assert np.allclose(ph.numpy(), pm.cpu().numpy(), atol=1e-2)


def test_reference_named_inference(small_problem, oracle_small):
    from mvsnet_b200 import model
    from mvsnet_b200.cnn_wrapper import mvsnetworks
    p = small_problem
    rd, _, allr = oracle_small
    mvsnetworks.set_variables(p["weights"])
    model.FLAGS.view_num = p["n_views"]
    model.FLAGS.precision = "fp32"
    feats = to_dev(p["feats"])[None]
    cams = to_dev(p["cams"])[None]
    depth, prob = model.inference_mem(feats, cams, p["depth_num"], torch.tensor([p["depth_start"]]),
                                      torch.tensor([p["depth_interval"]]), "normal")
    assert depth.shape == (1, p["hf"], p["wf"], 1) and prob.shape == depth.shape
    assert _frac_within(depth[0, :, :, 0].cpu().numpy(), rd, p["depth_interval"]) >= 0.999
    depth2, _ = model.inference(feats, cams, p["depth_num"], torch.tensor([p["depth_start"]]),
                                torch.tensor([p["depth_interval"]]), "normal")
    assert _frac_within(depth2[0, :, :, 0].cpu().numpy(), rd, p["depth_interval"]) >= 0.99
    with pytest.raises(TypeError):
        model.inference_mem(feats, cams, torch.tensor(32), torch.tensor([425.0]), torch.tensor([10.0]), "normal")
    with pytest.raises(RuntimeError):
        model.inference_mem(torch.zeros((1, 5, 8, 8, 3), device="cuda"), cams, 32, torch.tensor([425.0]),
                            torch.tensor([10.0]), "normal")
    # RegNetUS0 by its reference name
    mvsnetworks.RegNetUS0.precision = "fp32"
    out = mvsnetworks.RegNetUS0({"data": to_dev(allr["cost"])[None]}, trainable=True, training=True,
                                mode="normal", reuse=False).get_output()
    assert out.shape == (1,) + allr["filtered"].shape + (1,)
    assert np.abs(out[0, ..., 0].cpu().numpy() - allr["filtered"]).max() <= 2e-3 * np.abs(allr["filtered"]).max()
    mvsnetworks.RegNetUS0.precision = "bf16"
    model.FLAGS.precision = "bf16"


def test_infer_host_async_two_streams(tiny_problem):
    """Two engines on two streams (own staging + workspace) fed from pinned host buffers: both results match
    the synchronous call bit for bit (no state is shared between calls in flight)."""
    from mvsnet_b200.engine import HotPath
    p = tiny_problem
    fh = torch.from_numpy(p["feats"]).pin_memory()
    ch = torch.from_numpy(p["cams"]).pin_memory()
    # a second, different problem: permute the source views
    perm = [0] + list(range(p["n_views"] - 1, 0, -1))
    fh2 = fh[perm].contiguous().pin_memory()
    ch2 = ch[perm].contiguous().pin_memory()
    engs = [HotPath(p["n_views"], p["depth_num"], p["hf"], p["wf"], p["weights"], precision="bf16") for _ in range(2)]
    ref = []
    for f, c in ((fh, ch), (fh2, ch2)):
        d = torch.empty((p["hf"], p["wf"])).pin_memory()
        q = torch.empty((p["hf"], p["wf"])).pin_memory()
        engs[0].infer_host(f, c, p["depth_start"], p["depth_interval"], d, q)
        ref.append((d.clone(), q.clone()))
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = [(torch.empty((p["hf"], p["wf"])).pin_memory(), torch.empty((p["hf"], p["wf"])).pin_memory()) for _ in range(2)]
    for rep in range(3):
        for k, (f, c) in enumerate(((fh, ch), (fh2, ch2))):
            with torch.cuda.stream(streams[k]):
                engs[k].infer_host_async(f, c, p["depth_start"], p["depth_interval"], outs[k][0], outs[k][1])
    torch.cuda.synchronize()
    for k in range(2):
        assert torch.equal(outs[k][0], ref[k][0])
        assert torch.equal(outs[k][1], ref[k][1])


def test_infer_host_pipelined_one_compute_stream(tiny_problem):
    """Feed / kernels / fetch on separate streams chained by events: same maps as the synchronous host entry point,
    for several views in flight over one compute stream and one copy stream per staging buffer."""
    from mvsnet_b200.engine import HotPath
    p = tiny_problem
    engs = [HotPath(p["n_views"], p["depth_num"], p["hf"], p["wf"], p["weights"], precision="bf16") for _ in range(2)]
    fh = [torch.from_numpy(p["feats"] * s).pin_memory() for s in (1.0, 0.5, 2.0)]
    ch = torch.from_numpy(p["cams"]).pin_memory()
    ref = []
    for f in fh:
        d, q = torch.empty((p["hf"], p["wf"])).pin_memory(), torch.empty((p["hf"], p["wf"])).pin_memory()
        engs[0].infer_host(f, ch, p["depth_start"], p["depth_interval"], d, q)
        ref.append((d.clone(), q.clone()))
    compute, copies = torch.cuda.Stream(), [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = [(torch.empty((p["hf"], p["wf"])).pin_memory(), torch.empty((p["hf"], p["wf"])).pin_memory()) for _ in fh]
    for i, f in enumerate(fh):
        engs[i % 2].infer_host_pipelined(f, ch, p["depth_start"], p["depth_interval"], outs[i][0], outs[i][1], compute,
                                         copies[i % 2])
    torch.cuda.synchronize()
    for (d, q), (rd, rq) in zip(outs, ref):
        assert np.allclose(d.numpy(), rd.numpy(), rtol=1e-4, atol=1e-2)
        assert np.allclose(q.numpy(), rq.numpy(), atol=1e-2)
