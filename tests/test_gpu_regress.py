"""GPU parity of softmax / soft-argmin / probability map (kernel 4) against the oracle."""
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


def _check(ops, O, F, ds, di, inverse=False, buckets=4):
    depth, prob, P = ops.depth_regress(to_dev(F), ds, di, inverse, buckets, want_prob_volume=True)
    depth, prob, P = depth.cpu().numpy(), prob.cpu().numpy(), P.cpu().numpy()
    rd, rp, rP = O.depth_regress(F, ds, di, inverse, buckets)
    assert np.abs(P - rP).max() <= 1e-6
    np.testing.assert_allclose(depth, rd, rtol=2e-6)
    # the bucket indices are discontinuous in depth: evaluate the oracle's map at the kernel's own depth
    rp_at = O.get_probability_map_slice(P, depth, ds, di, inverse, buckets)
    np.testing.assert_allclose(prob, rp_at, rtol=1e-6, atol=1e-7)
    frac = np.mean(np.abs(prob - rp) <= 1e-5)
    assert frac >= 0.999, frac
    # stand-alone probability map entry point
    pm = ops.probability_map(to_dev(P), to_dev(depth), ds, di, inverse, buckets).cpu().numpy()
    np.testing.assert_array_equal(pm, rp_at)
    return depth, prob


@pytest.mark.parametrize("shape", [(16, 24, 32), (192, 16, 64), (33, 7, 13), (8, 1, 1)])
def test_random_volumes(ops, O, shape):
    rng = np.random.RandomState(11)
    F = (rng.randn(*shape) * 3).astype(np.float32)
    _check(ops, O, F, 425.0, 2.65)
    _check(ops, O, F, 425.0, 2.65, buckets=2)
    _check(ops, O, F, 425.0, 2.65, inverse=True)


def test_golden(ops, golden_tiny):
    g = golden_tiny
    depth, prob = ops.depth_regress(to_dev(g["filtered"]), float(g["depth_start"]), float(g["depth_interval"]))
    np.testing.assert_allclose(depth.cpu().numpy(), g["depth"], rtol=2e-6)
    assert np.mean(np.abs(prob.cpu().numpy() - g["prob"]) <= 1e-5) >= 0.999


def test_one_hot_double_count(ops, O):                           # Appendix B.7
    D, Lv = 8, 30.0
    for k in (0, 3, 7):
        F = np.full((D, 4, 4), Lv, dtype=np.float32)
        F[k] = -Lv
        depth, prob = ops.depth_regress(to_dev(F), 100.0, 2.0)
        rd, rp, _ = O.depth_regress(F, 100.0, 2.0)
        np.testing.assert_array_equal(depth.cpu().numpy(), rd)
        np.testing.assert_allclose(prob.cpu().numpy(), rp, rtol=1e-6)
        assert prob.max().item() > 1.0


def test_large_depth_falls_back_to_global_path(ops, O):
    rng = np.random.RandomState(12)
    F = rng.randn(1024, 4, 16).astype(np.float32)                 # D*64*4 B > 200 KB of shared memory
    _check(ops, O, F, 100.0, 0.5)


def test_full_size_properties(ops):
    D, hf, wf = 192, 216, 288
    g = torch.Generator(device="cuda").manual_seed(1)
    F = torch.randn((D, hf, wf), device="cuda", generator=g) * 4
    depth, prob, P = ops.depth_regress(F, 425.0, 2.65, want_prob_volume=True)
    assert torch.allclose(P.sum(0), torch.ones_like(depth), atol=1e-5)
    assert (depth >= 425.0).all() and (depth <= 425.0 + 191 * 2.65 + 1e-2).all()
    assert (prob > 0).all() and (prob <= 2.0).all()
    shifted, _ = ops.depth_regress(F + 7.5, 425.0, 2.65)           # softmax shift invariance
    assert torch.allclose(shifted, depth, rtol=1e-5)


def test_reference_named_probability_map(ops, O):
    from mvsnet_b200 import model
    rng = np.random.RandomState(13)
    F = rng.randn(2, 12, 6, 8).astype(np.float32)
    Ps, ds = [], []
    for b in range(2):
        d, _, P = O.depth_regress(F[b], 400.0 + b, 3.0)
        Ps.append(P)
        ds.append(d)
    cv = to_dev(np.stack(Ps))
    dm = to_dev(np.stack(ds)[..., None])
    out = model.get_probability_map(cv, dm, torch.tensor([400.0, 401.0]), torch.tensor([3.0, 3.0]))
    ref = O.get_probability_map(np.stack(Ps), np.stack(ds)[..., None], [400.0, 401.0], [3.0, 3.0])
    assert out.shape == (2, 6, 8, 1)
    np.testing.assert_array_equal(out.cpu().numpy(), ref)


@pytest.mark.parametrize("slabs", [1, 2, 4])
def test_slab_regression_matches_whole_volume(ops, slabs):
    """D-slab mode: per-slab soft-argmin partials + combine + summed probability shares equal the whole-volume
    kernel up to the association of the sums (all slabs emulated on one GPU)."""
    import ctypes
    from mvsnet_b200 import _lib as L
    lib = L.load()
    D, hf, wf = 32, 24, 40
    ds, di = 425.0, 2.5
    g = torch.Generator(device="cuda").manual_seed(3)
    F = torch.randn((D, hf, wf), device="cuda", generator=g) * 3.0
    F[5, 3, 7] = -40.0                                   # a sharp minimum: exact-integer index, double-counted buckets
    F[D - 1, 0, 0] = -40.0                               # and one at the clipped end
    depth_ref, prob_ref = ops.depth_regress(F, ds, di)
    npix, dl = hf * wf, D // slabs
    partials = torch.empty((slabs, 3, npix), device="cuda")
    for r in range(slabs):
        L.check(lib.mvsb200_regress_partial(L.ptr(F[r * dl:(r + 1) * dl].contiguous()), dl, r * dl, D, npix, ds, di, 0,
                                            L.ptr(partials[r]), L.stream_ptr()), "regress_partial")
    prob = torch.zeros((hf, wf), device="cuda")
    for r in range(slabs):
        depth = torch.empty((hf, wf), device="cuda")
        share = torch.empty((hf, wf), device="cuda")
        L.check(lib.mvsb200_regress_combine(L.ptr(partials), slabs, L.ptr(F[r * dl:(r + 1) * dl].contiguous()), dl, r * dl,
                                            D, npix, ds, di, 0, 4, L.ptr(depth), L.ptr(share), L.stream_ptr()),
                "regress_combine")
        prob += share
        assert float((depth - depth_ref).abs().max()) <= 1e-5 * float(depth_ref.abs().max())
    assert float((prob - prob_ref).abs().max()) <= 2e-5
    assert abs(float(prob[3, 7]) - float(prob_ref[3, 7])) <= 2e-5 and float(prob_ref[3, 7]) > 1.5    # counted twice
