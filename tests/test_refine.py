"""Refinement glue after the path (model.py:753-811): known answers of the CPU restatement
(oracle/refine_oracle.py) and, on the GPU, mvsnet_b200.model.depth_refine against it."""
import numpy as np
import pytest

from oracle import refine_oracle as R


def make_refine_weights(in_channels=4, filters=32, seed=5):
    rng = np.random.RandomState(seed)
    w, cin = {}, in_channels
    for i, cout in enumerate((filters, filters, filters, 1)):
        lim = np.sqrt(6.0 / (9 * cin + 9 * cout))
        w[f"refine_conv{i}/kernel"] = rng.uniform(-lim, lim, (3, 3, cin, cout)).astype(np.float32)
        w[f"refine_conv{i}/bias"] = rng.normal(0, 0.05, (cout,)).astype(np.float32)
        cin = cout
    return w


def test_resize_bilinear_known_answers():
    x = np.arange(12, dtype=np.float32).reshape(1, 3, 4, 1)
    np.testing.assert_array_equal(R.resize_bilinear(x, 3, 4), x)                        # same size: identity
    up = R.resize_bilinear(x, 6, 8)[0, :, :, 0]
    # TF 1.x convention (no half-pixel centres): output (2i, 2j) is input (i, j); odd positions are midpoints; the last
    # row / column repeat the edge (upper index clamps)
    np.testing.assert_array_equal(up[::2, ::2], x[0, :, :, 0])
    np.testing.assert_allclose(up[0, 1], 0.5)
    np.testing.assert_allclose(up[1, 0], 2.0)
    np.testing.assert_array_equal(up[:, 7], up[:, 6])
    np.testing.assert_array_equal(up[5, :], up[4, :])
    down = R.resize_bilinear(x, 2, 2)[0, :, :, 0]                                       # scale 1.5 / 2: samples (0,0),(0,2),(1.5,*)
    np.testing.assert_allclose(down, [[0.0, 2.0], [6.0, 8.0]])


def test_depth_refine_identity_tower():
    """A tower with zero weights and zero bias predicts no residual: refined = (resized) initial depth map."""
    rng = np.random.RandomState(1)
    w = {k: np.zeros_like(v) for k, v in make_refine_weights().items()}
    depth = (500.0 + 100.0 * rng.rand(1, 6, 8, 1)).astype(np.float32)
    image = rng.randn(1, 24, 32, 3).astype(np.float32)
    refined, residual = R.depth_refine(depth, image, None, 64, 425.0, 2.65, w)
    np.testing.assert_array_equal(refined, depth)
    assert not residual.any()
    refined_up, _ = R.depth_refine(depth, image, None, 64, 425.0, 2.65, w, upsample_depth=True)
    np.testing.assert_array_equal(refined_up, R.resize_bilinear(depth, 24, 32))


# ---------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("shape", [((1, 5, 7, 3), (11, 13)), ((2, 12, 16, 1), (6, 8)), ((1, 8, 8, 2), (8, 8))])
def test_gpu_resize_bilinear_bit_exact(shape):
    torch = pytest.importorskip("torch")
    from mvsnet_b200 import model
    (n, h, w, c), (oh, ow) = shape
    x = np.random.RandomState(3).randn(n, h, w, c).astype(np.float32)
    y = model._resize_bilinear(torch.from_numpy(x).cuda(), oh, ow).cpu().numpy()
    np.testing.assert_array_equal(y, R.resize_bilinear(x, oh, ow))


@pytest.mark.gpu
@pytest.mark.parametrize("upsample,confidence", [(False, False), (True, True), (False, True)])
def test_gpu_depth_refine_vs_oracle(upsample, confidence):
    torch = pytest.importorskip("torch")
    from mvsnet_b200 import model
    from mvsnet_b200.cnn_wrapper import mvsnetworks
    rng = np.random.RandomState(7)
    w = make_refine_weights(in_channels=4 + int(confidence))
    depth = (500.0 + 100.0 * rng.rand(1, 12, 16, 1)).astype(np.float32)
    prob = rng.rand(1, 12, 16, 1).astype(np.float32)
    image = rng.randn(1, 48, 64, 3).astype(np.float32)
    ref_refined, ref_residual = R.depth_refine(depth, image, prob, 192, 425.0, 2.65, w, upsample_depth=upsample,
                                               refine_with_confidence=confidence)
    mvsnetworks.set_refine_variables(w)
    refined, residual = model.depth_refine(torch.from_numpy(depth).cuda(), torch.from_numpy(image).cuda(),
                                           torch.from_numpy(prob).cuda(), 192, torch.tensor([425.0]), torch.tensor([2.65]),
                                           "normal", "original", upsample_depth=upsample, refine_with_confidence=confidence)
    assert refined.shape == ref_refined.shape
    scale = 191 * 2.65
    assert np.abs(residual.cpu().numpy() - ref_residual).max() <= 1e-4 * scale      # fp32 convs, different summation order
    assert np.abs(refined.cpu().numpy() - ref_refined).max() <= 1e-4 * scale
    with pytest.raises(NotImplementedError):
        model.depth_refine(torch.from_numpy(depth).cuda(), torch.from_numpy(image).cuda(), None, 192, 425.0, 2.65,
                           "normal", "unet")
