"""Training step on the GPU (csrc/backward.cu; BASELINE config 4) against the fp64 autograd oracle
(oracle/backward_oracle.py, warp flavour 'exact_adjoint'): loss, accuracies, depth map, the gradient of every RegNetUS0
variable and of the feature maps, <= 1e-4 relative (of each tensor's largest entry; measured 2e-6 .. 5e-6)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from conftest import to_dev  # noqa: E402


def _problem(seed, size, depth_num, base_filter=8):
    from mvsnet_b200 import synthetic
    cams = synthetic.make_cameras(3, 4 * size[0], 4 * size[1], depth_num, interval_scale=8.0, seed=1234 + seed)
    feats = synthetic.make_features(cams, size[0], size[1], 32, seed=5678 + seed)
    weights = synthetic.make_regnet_weights(32, base_filter, seed=42)
    ds, di = float(cams[0, 1, 3, 0]), float(cams[0, 1, 3, 1])
    rng = np.random.RandomState(9)
    gt = (ds + di * rng.uniform(2, depth_num - 3, size=size)).astype(np.float32)
    gt[rng.rand(*size) < 0.2] = 0.0
    return dict(feats=feats, cams=cams, weights=weights, depth_num=depth_num, ds=ds, di=di, gt=gt, hf=size[0], wf=size[1])


@pytest.mark.parametrize("order,size,depth_num,base_filter", [("train", (16, 24), 16, 8), ("mem", (16, 16), 24, 8),
                                                              ("train", (24, 16), 16, 4)])
def test_train_step_vs_oracle(order, size, depth_num, base_filter):
    from oracle import backward_oracle as B
    from mvsnet_b200.train import TrainStep
    p = _problem(1, size, depth_num, base_filter)
    ref = B.loss_and_grads(p["feats"], p["cams"], p["gt"], depth_num, p["ds"], p["di"], p["weights"], order=order)
    ts = TrainStep(3, depth_num, p["hf"], p["wf"], p["weights"], order=order)
    out = ts.step(to_dev(p["feats"]), to_dev(p["cams"]), to_dev(p["gt"]), p["ds"], p["di"])
    loss, l1, l3 = out["metrics"].cpu().tolist()
    assert abs(loss - ref["loss"]) <= 1e-4 * max(1.0, abs(ref["loss"])), (loss, ref["loss"])
    assert abs(l1 - ref["less_one"]) <= 0.01 and abs(l3 - ref["less_three"]) <= 0.01
    assert np.abs(out["depth_map"].cpu().numpy() - ref["depth"]).max() <= 1e-3 * p["di"]
    worst = 0.0
    for name, g_ref in ref["grads"].items():
        g = out["grads"][name].cpu().numpy().astype(np.float64)
        scale = np.abs(g_ref).max()
        err = np.abs(g - g_ref).max() / (scale + 1e-30)
        worst = max(worst, err)
        assert err <= 1e-4, (name, err, scale)
    d_ref = ref["dfeats"]
    err = np.abs(out["dfeats"].cpu().numpy() - d_ref).max() / np.abs(d_ref).max()
    print(f"{order} {size} D={depth_num} b={base_filter}: worst variable-gradient error {worst:.2e}, feature-gradient error {err:.2e} "
          f"(relative to the largest entry); loss {loss:.5f} vs {ref['loss']:.5f}")
    assert err <= 1e-4, err


def test_train_step_rejects_bad_arguments():
    from mvsnet_b200._lib import MVSB200Error
    from mvsnet_b200.train import TrainStep
    p = _problem(0, (16, 16), 16)
    with pytest.raises(MVSB200Error):
        TrainStep(3, 12, 16, 16, p["weights"]).step(to_dev(p["feats"]), to_dev(p["cams"]), to_dev(p["gt"]), p["ds"], p["di"])
