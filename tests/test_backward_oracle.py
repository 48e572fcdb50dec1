"""The CPU oracle of the training step (oracle/backward_oracle.py; BASELINE config 4): its forward agrees with the fp32
forward oracle, its gradients agree with central finite differences of its own loss (fp64), and the two warp-gradient
flavours differ only in the gradient that flows through the warp."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

import oracle as O  # noqa: E402
from oracle import backward_oracle as B  # noqa: E402
from mvsnet_b200 import synthetic  # noqa: E402


def micro_problem(seed=0, channels=8, base_filter=2, size=16, depth_num=16):
    cams = synthetic.make_cameras(3, 4 * size, 4 * size, depth_num, interval_scale=8.0, seed=1234 + seed)
    feats = synthetic.make_features(cams, size, size, channels, seed=5678 + seed)
    weights = synthetic.make_regnet_weights(channels, base_filter, seed=42)
    ds, di = float(cams[0, 1, 3, 0]), float(cams[0, 1, 3, 1])
    rng = np.random.RandomState(9)
    gt = (ds + di * rng.uniform(2, depth_num - 3, size=(size, size))).astype(np.float32)
    gt[rng.rand(size, size) < 0.2] = 0.0                      # invalid pixels of the ground truth (loss.py:21)
    return dict(feats=feats, cams=cams, weights=weights, depth_num=depth_num, depth_start=ds, depth_interval=di, gt=gt)


def test_forward_agrees_with_forward_oracle():
    p = micro_problem()
    rd, _ = O.inference_from_features(p["feats"], p["cams"], p["depth_num"], p["depth_start"], p["depth_interval"],
                                      p["weights"], order="train")
    f = torch.tensor(p["feats"], dtype=torch.float64)
    w = {k: torch.tensor(v, dtype=torch.float64) for k, v in p["weights"].items()}
    d, prob, _ = B.forward(f, p["cams"], p["depth_num"], p["depth_start"], p["depth_interval"], w)
    assert np.abs(d.numpy() - rd).max() <= 1e-3 * p["depth_interval"]
    np.testing.assert_allclose(prob.sum(dim=0).numpy(), 1.0, atol=1e-9)


def test_loss_known_answers():
    gt = torch.tensor([[500.0, 0.0], [600.0, 700.0]], dtype=torch.float64)
    est = torch.tensor([[510.0, 123.0], [600.0, 690.0]], dtype=torch.float64)
    loss, l1, l3 = B.regression_loss(est, gt, 425.0, 425.0 + 191.0 * 5.0)                 # interval 5
    np.testing.assert_allclose(float(loss), (10 + 0 + 10) / 5.0 / 3.0, rtol=1e-6)          # masked MAE / interval
    np.testing.assert_allclose(float(l1), 1.0 / 3.0, rtol=1e-5)                            # |err| <= 1 interval
    np.testing.assert_allclose(float(l3), 1.0, rtol=1e-5)


@pytest.mark.parametrize("order", ["train", "mem"])
def test_gradients_match_finite_differences(order):
    p = micro_problem()
    res = B.loss_and_grads(p["feats"], p["cams"], p["gt"], p["depth_num"], p["depth_start"], p["depth_interval"],
                           p["weights"], order=order)
    depth_end = p["depth_start"] + (p["depth_num"] - 1) * p["depth_interval"]

    def loss_of(feats, weights):
        f = torch.tensor(feats, dtype=torch.float64)
        w = {k: torch.tensor(v, dtype=torch.float64) for k, v in weights.items()}
        d, _, _ = B.forward(f, p["cams"], p["depth_num"], p["depth_start"], p["depth_interval"], w, order)
        return float(B.regression_loss(d, torch.tensor(p["gt"], dtype=torch.float64), p["depth_start"], depth_end)[0])

    rng = np.random.RandomState(3)
    h = 1e-5
    for name in ("3dconv0_1/kernel", "3dconv3_1/kernel", "3dconv5_0/kernel", "3dconv6_2/kernel", "3dconv1_0/bn/gamma",
                 "3dconv4_0/bn/beta"):
        base = p["weights"][name].astype(np.float64)
        idx = tuple(rng.randint(0, s) for s in base.shape)
        num = 0.0
        for sgn in (1.0, -1.0):
            w = dict(p["weights"])
            pert = base.copy()
            pert[idx] += sgn * h
            w[name] = pert
            num += sgn * loss_of(p["feats"], w)
        num /= 2 * h
        assert abs(num - res["grads"][name][idx]) <= 1e-5 + 2e-4 * abs(num), (name, num, res["grads"][name][idx])
    for _ in range(3):
        idx = (rng.randint(0, 3), rng.randint(2, 14), rng.randint(2, 14), rng.randint(0, 8))
        num = 0.0
        for sgn in (1.0, -1.0):
            f = p["feats"].astype(np.float64).copy()
            f[idx] += sgn * h
            num += sgn * loss_of(f, p["weights"])
        num /= 2 * h
        assert abs(num - res["dfeats"][idx]) <= 1e-6 + 2e-4 * abs(num), (idx, num, res["dfeats"][idx])


def test_warp_gradient_flavours():
    """tf_compat (gradient resampled with the inverse transform) and exact_adjoint (scatter) agree on everything that
    does not flow through the warp -- the loss, the weight gradients, the reference view's feature gradient -- and
    differ on the source views' feature gradients."""
    p = micro_problem()
    a = B.loss_and_grads(p["feats"], p["cams"], p["gt"], p["depth_num"], p["depth_start"], p["depth_interval"], p["weights"])
    b = B.loss_and_grads(p["feats"], p["cams"], p["gt"], p["depth_num"], p["depth_start"], p["depth_interval"], p["weights"],
                         flavour="tf_compat")
    assert a["loss"] == b["loss"]
    for k in a["grads"]:
        np.testing.assert_allclose(a["grads"][k], b["grads"][k], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(a["dfeats"][0], b["dfeats"][0], rtol=1e-9, atol=1e-12)
    assert np.abs(a["dfeats"][1:] - b["dfeats"][1:]).max() > 0
    # identity transform: both flavours are the identity map, forward and backward
    img = torch.arange(4 * 5 * 2, dtype=torch.float64).reshape(4, 5, 2).requires_grad_(True)
    ident = np.array([1, 0, 0, 0, 1, 0, 0, 0], dtype=np.float32)
    for flavour in ("exact_adjoint", "tf_compat"):
        out = B.warp(img, ident, flavour)
        g, = torch.autograd.grad((out * out).sum(), img)
        np.testing.assert_allclose(out.detach().numpy(), img.detach().numpy())
        np.testing.assert_allclose(g.numpy(), 2 * img.detach().numpy())
