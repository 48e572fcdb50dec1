"""Oracle self-checks: the known-answer tests of SURVEY.md Appendix B (none exist upstream) and the
frozen golden fixture.  CPU only."""
import numpy as np
import pytest

import oracle as O
from mvsnet_b200 import synthetic

F32 = np.float32


def _cam(K, R, t, d0=425.0, di=2.5, D=8):
    cam = np.zeros((1, 2, 4, 4), dtype=F32)
    cam[0, 0, :3, :3] = R
    cam[0, 0, :3, 3] = t
    cam[0, 0, 3, 3] = 1
    cam[0, 1, :3, :3] = K
    cam[0, 1, 3] = (d0, di, D, d0 + (D - 1) * di)
    return cam


K0 = np.array([[100.0, 0, 16], [0, 100.0, 12], [0, 0, 1]], dtype=F32)


def test_inverse_lu_matches_numpy():
    rng = np.random.RandomState(0)
    for _ in range(20):
        A = (rng.randn(3, 3) + 3 * np.eye(3)).astype(F32)
        np.testing.assert_allclose(O.inv3x3_lu(A), np.linalg.inv(A.astype(np.float64)), rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(O.inv3x3_lu(K0) @ K0, np.eye(3), atol=1e-6)


def test_identity_cameras_give_identity_homography():            # Appendix B.1
    cam = _cam(K0, np.eye(3), np.zeros(3))
    H = O.get_homographies(cam, cam, 8, 425.0, 2.5)
    assert H.shape == (1, 8, 3, 3)
    np.testing.assert_allclose(H[0], np.broadcast_to(np.eye(3), (8, 3, 3)), atol=2e-6)
    T = O.transform_coefs(H[0])
    np.testing.assert_allclose(T, np.broadcast_to(np.array([1, 0, 0, 0, 1, 0, 0, 0], F32), (8, 8)), atol=2e-6)
    img = np.random.RandomState(1).randn(1, 24, 32, 4).astype(F32)
    np.testing.assert_array_equal(O.tf_transform_homography(img, np.eye(3, dtype=F32)[None]), img)
    feats = np.repeat(img, 3, axis=0)
    cost = O.cost_volume(feats, np.broadcast_to(np.eye(3, dtype=F32), (2, 4, 3, 3)))
    np.testing.assert_allclose(cost, 0.0, atol=1e-5)


def test_fronto_parallel_shift_and_half_pixel_convention():      # Appendix B.2
    f, b, d = 100.0, 5.0, 500.0                                  # shift f*b/d = 1.0 px exactly
    left = _cam(K0, np.eye(3), np.zeros(3))
    right = _cam(K0, np.eye(3), np.array([-b, 0, 0]))            # camera centre at +b along x
    H = O.get_homographies(left, right, 1, d, 1.0)[0, 0]
    T = O.transform_coefs(H)[0]
    ix, iy = O.sample_coords(T, 24, 32)
    xs = np.arange(32, dtype=F32)[None, :].repeat(24, 0)
    np.testing.assert_allclose(ix, xs - 1.0, atol=1e-4)
    np.testing.assert_allclose(iy, np.arange(24, dtype=F32)[:, None].repeat(32, 1), atol=1e-4)
    img = np.random.RandomState(2).randn(1, 24, 32, 3).astype(F32)
    w = O.tf_transform_homography(img, H[None])[0]
    np.testing.assert_allclose(w[:, 2:], img[0][:, 1:-1], atol=2e-4)


def test_zero_fill_vs_clamp_border():                            # Appendix B.3
    img = np.ones((1, 8, 8, 1), dtype=F32)
    # pixel-coordinate shift of -0.5 px in x: ix = x - 0.5
    Himg = np.array([[1, 0, -0.5], [0, 1, 0], [0, 0, 1]], dtype=F32)
    w = O.tf_transform_homography(img, Himg[None])[0, :, :, 0]
    np.testing.assert_allclose(w[:, 0], 0.5)                     # linear ramp into the zero fill
    np.testing.assert_allclose(w[:, 1:], 1.0)
    wl = O.homography_warping(img, Himg[None])[0, :, :, 0]
    np.testing.assert_array_equal(wl[:, 0], 0.0)                 # both corners clamp to column 0 -> exactly 0
    np.testing.assert_allclose(wl[:-1, 1:], 1.0)
    # legacy quirk: y exactly on the last row -> y1 clamps onto y0, both y-weights are 0 -> output 0
    np.testing.assert_array_equal(wl[-1, :], 0.0)


def test_variance_orders_agree():                                # Appendix B.4
    p = synthetic.make_problem("tiny")
    H = np.stack([O.get_homographies(p["cams"][0:1], p["cams"][v:v + 1], 4, 425.0, 20.0)[0] for v in (1, 2)])
    a = O.cost_volume(p["feats"], H, order="mem")
    b = O.cost_volume(p["feats"], H, order="train")
    assert np.abs(a - b).max() <= 4 * np.spacing(np.abs(a).max() + 4.0)


def test_same_padding_and_deconv_alignment():                    # Appendix B.5 / A.5
    assert O.tf_same_pads(8, 3, 1) == (1, 1)
    assert O.tf_same_pads(8, 3, 2) == (0, 1)
    assert O.tf_same_pads(7, 3, 2) == (1, 1)
    w = np.zeros((3, 3, 3, 1, 1), dtype=F32)
    w[0, 0, :, 0, 0] = (1, 10, 100)                              # taps along W only (kd=kh=0 reads d=h=0)
    x = np.zeros((2, 2, 8, 1), dtype=F32)
    x[0, 0, 4, 0] = 1
    y = O.conv3d_same(x, w, 2)                                   # out[o] = sum_k in[2o+k] w[k], pad (0,1)
    np.testing.assert_array_equal(y[0, 0, :, 0], [0, 100, 1, 0])
    x[:] = 0
    x[0, 0, 5, 0] = 1
    np.testing.assert_array_equal(O.conv3d_same(x, w, 2)[0, 0, :, 0], [0, 0, 10, 0])
    wt = np.zeros((3, 3, 3, 1, 1), dtype=F32)
    wt[0, 0, :, 0, 0] = (1, 10, 100)
    xi = np.zeros((1, 1, 4, 1), dtype=F32)
    xi[0, 0, 1, 0] = 1
    yt = O.conv3d_transpose_same(xi, wt)
    assert yt.shape == (2, 2, 8, 1)
    np.testing.assert_array_equal(yt[0, 0, :, 0], [0, 0, 1, 10, 100, 0, 0, 0])     # lands on 2i, 2i+1, 2i+2
    xi[:] = 0
    xi[0, 0, 3, 0] = 1
    np.testing.assert_array_equal(O.conv3d_transpose_same(xi, wt)[0, 0, :, 0], [0, 0, 0, 0, 0, 0, 1, 10])


def test_batch_norm_moments():                                   # Appendix B.6
    rng = np.random.RandomState(3)
    x = (rng.randn(4, 6, 8, 5) * 3 + 1).astype(F32)
    g = rng.uniform(0.5, 1.5, 5).astype(F32)
    b = rng.randn(5).astype(F32)
    y = O.batch_norm_train(x, g, b, relu=False)
    var = x.reshape(-1, 5).astype(np.float64).var(axis=0)
    np.testing.assert_allclose(y.reshape(-1, 5).mean(0), b, atol=2e-5)
    np.testing.assert_allclose(y.reshape(-1, 5).var(0), g.astype(np.float64) ** 2 * var / (var + 1e-5), rtol=1e-4)


def test_regression_one_hot_and_double_count():                  # Appendix B.7
    D, L = 8, 30.0
    for k in (0, 3, 7):
        F = np.full((D, 2, 2), L, dtype=F32)
        F[k] = -L
        depth, prob, P = O.depth_regress(F, 100.0, 2.0)
        np.testing.assert_allclose(depth, 100.0 + 2.0 * k, rtol=1e-6)
        idx = (depth - F32(100.0)) / F32(2.0)
        assert np.all(idx == k)                                  # exactly integer -> l0 == r0 (double count)
        if k == 0:
            expect = 3 * P[0] + P[1]
        elif k == D - 1:
            expect = P[k - 1] + 3 * P[k]
        else:
            expect = P[k - 1] + 2 * P[k] + P[k + 1]
        np.testing.assert_allclose(prob, expect, rtol=1e-6)
        assert prob.max() > 1.0
    # non-integer index: four distinct buckets
    F = np.zeros((D, 1, 1), dtype=F32)
    depth, prob, P = O.depth_regress(F, 100.0, 2.0)              # uniform -> depth = 107, idx = 3.5
    np.testing.assert_allclose(prob, 4.0 / D, rtol=1e-6)


def test_inverse_depth_planes_and_prob_map():
    d = O.inv_depth_planes(16, 400.0, 900.0)
    assert d[0] == F32(1.0) / (F32(1.0) / F32(400.0)) and d[-1] < 901 and np.all(np.diff(d) > 0)
    rng = np.random.RandomState(4)
    F = rng.randn(16, 3, 4).astype(F32)
    depth, prob, P = O.depth_regress(F, 400.0, 10.0, inverse_depth=True)
    assert depth.min() >= 400 and depth.max() <= 550 and np.all(prob > 0) and np.all(prob <= 2.0)


def test_plane_scene_cost_minimum():                             # Appendix B.8 (cost volume part)
    p = synthetic.make_problem("tiny")
    H = np.stack([O.get_homographies(p["cams"][0:1], p["cams"][v:v + 1], p["depth_num"], p["depth_start"],
                                     p["depth_interval"])[0] for v in (1, 2)])
    cost = O.cost_volume(p["feats"], H)
    am = cost.mean(axis=3).argmin(axis=0)
    planes = O.plane_depths(p["depth_num"], p["depth_start"], p["depth_interval"])
    assert abs(np.median(planes[am[4:-4, 4:-4]]) - 680.0) <= p["depth_interval"]


def test_regnet_specs_and_flops():
    specs = O.regnet_layer_specs(32, 8)
    assert [s[0] for s in specs] == synthetic.REGNET_LAYER_ORDER
    ch = synthetic.regnet_channels(32, 8)
    for name, op, cin, cout, stride in specs:
        assert (cin, cout, op, stride) == ch[name], name


def test_golden_fixture_matches_oracle(golden_tiny):
    g = golden_tiny
    p = synthetic.make_problem("tiny")
    np.testing.assert_array_equal(p["feats"], g["feats"])        # generator is deterministic
    np.testing.assert_array_equal(p["cams"], g["cams"])
    D, ds, di = int(g["depth_num"]), float(g["depth_start"]), float(g["depth_interval"])
    H = np.stack([O.get_homographies(g["cams"][0:1], g["cams"][v:v + 1], D, ds, di)[0] for v in (1, 2)])
    np.testing.assert_array_equal(H, g["homographies"])
    np.testing.assert_array_equal(O.transform_coefs(H.reshape(-1, 3, 3)).reshape(2, D, 8), g["transforms"])
    cost = O.cost_volume(g["feats"], H)
    np.testing.assert_array_equal(cost[::2, ::2, ::2, :], g["cost_mem_sub"])
    filtered = O.regnet_us0(cost, p["weights"])
    np.testing.assert_allclose(filtered, g["filtered"], rtol=1e-4, atol=1e-4)   # torch conv may reorder sums
    depth, prob, _ = O.depth_regress(g["filtered"], ds, di)
    np.testing.assert_array_equal(depth, g["depth"])
    np.testing.assert_array_equal(prob, g["prob"])


def test_c_restatement_matches_numpy_oracle(golden_tiny):
    """oracle/mvs_oracle.c (independent plain-C restatement) against the NumPy oracle and the golden fixture."""
    from oracle import c_oracle as C
    g = golden_tiny
    D, ds, di = int(g["depth_num"]), float(g["depth_start"]), float(g["depth_interval"])
    H = C.homographies(g["cams"], D, ds, di)
    np.testing.assert_array_equal(H, g["homographies"])
    np.testing.assert_array_equal(C.transform_coefs(H).reshape(2, D, 8), g["transforms"])
    de = F32(ds) + F32(D - 1) * F32(di)
    np.testing.assert_array_equal(C.homographies(g["cams"], D, ds, float(de), inverse=True), g["homographies_inv"])
    w, coords = C.transform_warp(g["feats"][1], g["transforms"][0, 5])
    np.testing.assert_array_equal(w, g["warped_v0_d5"])
    _, c0 = C.transform_warp(g["feats"][1], g["transforms"][0, 0])
    np.testing.assert_array_equal(c0, g["coords"][0])
    cost = C.cost_volume(g["feats"], H, "mem")
    np.testing.assert_array_equal(cost[::2, ::2, ::2, :], g["cost_mem_sub"])
    cost_t = C.cost_volume(g["feats"], H, "train")
    np.testing.assert_array_equal(cost_t[::2, ::2, ::2, :], g["cost_train_sub"])
    depth, prob, P = C.depth_regress(g["filtered"], ds, di)
    np.testing.assert_allclose(depth, g["depth"], rtol=2e-6)               # libm expf vs numpy exp
    assert np.mean(np.abs(prob - g["prob"]) <= 1e-5) >= 0.999
