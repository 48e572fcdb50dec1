"""TEST INFRASTRUCTURE -- CPU oracle of the training step of the hot path (BASELINE config 4):
forward of mvsnet/model.py:257-372 (`inference`, the training graph: variance as mean2 - mean^2), the loss of
mvsnet/loss.py:15-29,190-220 (`mvsnet_regression_loss`, loss_type 'original') and the gradients
`opt.compute_gradients(loss)` asks for (train.py:429) with respect to every RegNetUS0 variable and to the feature
maps the path receives.  Only tests/ and bench.py's cpu_baseline may import this.

The forward is the restatement of oracle/mvs_oracle.py written in torch (fp64 by default, so that the GPU's fp32
gradients are compared with something better than themselves); torch autograd supplies the derivatives of
convolutions, batch normalisation with BATCH statistics (the mean and variance are differentiated through, as TF
does with training=True), ReLU, softmax and the soft-argmin.  The warp has two gradient flavours:

  exact_adjoint  the adjoint of the bilinear zero-fill gather (a scatter of the incoming gradient with the same four
                 weights) -- what BASELINE config 4 calls "bilinear-warp backward scatter"; autograd derives it.
  tf_compat      what TF 1.12 registers for ImageProjectiveTransform: the gradient image resampled with the INVERSE
                 transform (`_image_projective_transform_grad`), no gradient to the transform.  Implemented as a
                 custom autograd function.

PARITY UNPINNED like the rest of oracle/ (TensorFlow cannot run here); pinned instead by finite differences of its own
forward (tests/test_backward_oracle.py) and by agreement of that forward with oracle/mvs_oracle.py.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as Fn

from . import mvs_oracle as O

REGNET = [  # name, op, source, skip, stride  (mvsnetworks.py:131-158)
    ("3dconv1_0", "conv", "data", None, 2), ("3dconv2_0", "conv", "3dconv1_0", None, 2),
    ("3dconv3_0", "conv", "3dconv2_0", None, 2), ("3dconv0_1", "conv", "data", None, 1),
    ("3dconv1_1", "conv", "3dconv1_0", None, 1), ("3dconv2_1", "conv", "3dconv2_0", None, 1),
    ("3dconv3_1", "conv", "3dconv3_0", None, 1), ("3dconv4_0", "deconv", "3dconv3_1", None, 2),
    ("3dconv5_0", "deconv", "3dconv4_0", "3dconv2_1", 2), ("3dconv6_0", "deconv", "3dconv5_0", "3dconv1_1", 2),
    ("3dconv6_2", "conv", "3dconv6_0", "3dconv0_1", 1),
]


# ------------------------------------------------------------------------------------------------ warp
def _footprints(coefs, height, width, dtype):
    """Sample positions of every output pixel (in fp32 exactly as the forward oracle computes them) ->
    flat tap indices [4, H*W] (clamped), validity masks and weights as torch tensors of `dtype`."""
    ix, iy = O.sample_coords(np.asarray(coefs, dtype=np.float32).reshape(1, 8), height, width)
    ix, iy = ix.reshape(-1).astype(np.float32), iy.reshape(-1).astype(np.float32)
    finite = np.isfinite(ix) & np.isfinite(iy)
    ix = np.where(finite, ix, -10.0).astype(np.float32)
    iy = np.where(finite, iy, -10.0).astype(np.float32)
    xf, yf = np.floor(ix), np.floor(iy)
    wx1, wy1 = (ix - xf).astype(np.float32), (iy - yf).astype(np.float32)
    wx0, wy0 = ((xf + 1) - ix).astype(np.float32), ((yf + 1) - iy).astype(np.float32)
    taps = []
    for dy, wy in ((0, wy0), (1, wy1)):
        for dx, wx in ((0, wx0), (1, wx1)):
            xx, yy = xf + dx, yf + dy
            valid = finite & (xx >= 0) & (xx < width) & (yy >= 0) & (yy < height)
            idx = (np.clip(yy, 0, height - 1) * width + np.clip(xx, 0, width - 1)).astype(np.int64)
            taps.append((torch.from_numpy(idx), torch.from_numpy((wy * wx * valid).astype(np.float64)).to(dtype)))
    return taps


def warp_gather(image, coefs):
    """tf.contrib.image.transform(BILINEAR) with zero fill: image [H,W,C] torch -> [H,W,C]; differentiable with respect
    to the image (the exact adjoint falls out of autograd).  The association differs from the forward oracle's
    wy*(wx*a + wx*b) by rounding only."""
    h, w, c = image.shape
    flat = image.reshape(h * w, c)
    out = 0
    for idx, wgt in _footprints(coefs, h, w, image.dtype):
        out = out + wgt[:, None] * flat[idx]
    return out.reshape(h, w, c)


class _WarpTfCompat(torch.autograd.Function):
    """Forward = warp_gather; backward = the gradient image resampled with the inverse transform (TF 1.12
    `_image_projective_transform_grad`: transforms -> 3x3 matrix, inverted, flattened again with the last entry
    normalised to 1)."""

    @staticmethod
    def forward(ctx, image, coefs):
        ctx.coefs = np.asarray(coefs, dtype=np.float64)
        with torch.no_grad():
            return warp_gather(image, coefs)

    @staticmethod
    def backward(ctx, grad):
        m = np.concatenate([ctx.coefs.reshape(8), [1.0]]).reshape(3, 3)
        inv = np.linalg.inv(m)
        inv = (inv / inv[2, 2]).reshape(9)[:8]
        with torch.no_grad():
            return warp_gather(grad.contiguous(), inv.astype(np.float32)), None


def warp(image, coefs, flavour="exact_adjoint"):
    if flavour == "exact_adjoint":
        return warp_gather(image, coefs)
    if flavour == "tf_compat":
        return _WarpTfCompat.apply(image, coefs)
    raise ValueError(flavour)


# ------------------------------------------------------------------------------------------------ forward
def _conv(x, w, stride):
    """x [D,H,W,Cin], w [3,3,3,Cin,Cout] (TF SAME, no bias) -> [Do,Ho,Wo,Cout]."""
    xt = x.permute(3, 0, 1, 2)[None]
    pads = []
    for dim in (3, 2, 1):
        pads.extend(O.tf_same_pads(xt.shape[dim + 1], 3, stride))
    y = Fn.conv3d(Fn.pad(xt, pads), w.permute(4, 3, 0, 1, 2), stride=stride)
    return y[0].permute(1, 2, 3, 0)


def _deconv(x, w):
    """x [D,H,W,Cin], w [3,3,3,Cout,Cin] -> [2D,2H,2W,Cout] (conv_transpose3d cropped to 2*in)."""
    d, h, wd = x.shape[:3]
    y = Fn.conv_transpose3d(x.permute(3, 0, 1, 2)[None], w.permute(4, 3, 0, 1, 2), stride=2)
    return y[0, :, :2 * d, :2 * h, :2 * wd].permute(1, 2, 3, 0)


def _bn_relu(x, gamma, beta, eps):
    """tf.layers.batch_normalization(training=True) + relu: batch moments over all voxels, biased variance (A.6)."""
    mean = x.mean(dim=(0, 1, 2))
    var = ((x - mean) ** 2).mean(dim=(0, 1, 2))
    inv = gamma / torch.sqrt(var + eps)
    return torch.relu(x * inv + (beta - mean * inv))


def regnet(cost, weights, eps=1e-5):
    acts = {"data": cost}
    for name, op, src, skip, stride in REGNET:
        x = acts[src] if skip is None else acts[src] + acts[skip]
        y = _conv(x, weights[name + "/kernel"], stride) if op == "conv" else _deconv(x, weights[name + "/kernel"])
        if name != "3dconv6_2":
            y = _bn_relu(y, weights[name + "/bn/gamma"], weights[name + "/bn/beta"], eps)
        acts[name] = y
    return acts["3dconv6_2"][..., 0]


def cost_volume(feats, coefs, order="train", flavour="exact_adjoint"):
    """feats [N,Hf,Wf,C] torch, coefs [N-1,D,8] numpy -> [D,Hf,Wf,C] (model.py:315-334 / :423-463)."""
    n = feats.shape[0]
    planes = []
    for d in range(coefs.shape[1]):
        s, q = feats[0], feats[0] * feats[0]
        for v in range(n - 1):
            w = warp(feats[v + 1], coefs[v, d], flavour)
            s, q = s + w, q + w * w
        if order == "train":
            mean = s / n
            planes.append(q / n - mean * mean)
        else:
            planes.append(q / n - (s * s) / (n * n))
    return torch.stack(planes)


def regression_loss(depth, gt_depth, depth_start, depth_end):
    """mvsnet_regression_loss(loss_type='original', grad_loss=False) (loss.py:15-29, 190-220): masked mean absolute
    error in units of (depth_end - depth_start) / 191, plus the two accuracy figures."""
    interval = (depth_end - depth_start) / 191.0
    mask = (gt_depth != 0).to(depth.dtype)
    denom = mask.sum().abs() + 1e-6
    err = (mask * (gt_depth - depth)).abs()
    loss = (err.sum() / interval) / denom
    rel = (gt_depth - depth).abs() / interval
    less_one = (mask * (rel <= 1.0).to(depth.dtype)).sum() / denom
    less_three = (mask * (rel <= 3.0).to(depth.dtype)).sum() / denom
    return loss, less_one, less_three


def forward(feats, cams, depth_num, depth_start, depth_interval, weights, order="train", flavour="exact_adjoint", eps=1e-5):
    """torch feats / weights -> (depth [Hf,Wf], prob volume [D,Hf,Wf], filtered [D,Hf,Wf]); geometry from the fp32
    forward oracle (homographies and transform coefficients carry no gradient, as upstream: the cameras are data)."""
    cams = np.asarray(cams, dtype=np.float32)
    n = feats.shape[0]
    homs = np.stack([O.get_homographies(cams[0:1], cams[v:v + 1], depth_num, depth_start, depth_interval)[0]
                     for v in range(1, n)])
    coefs = np.asarray(O.transform_coefs(homs.reshape(-1, 3, 3)), dtype=np.float32).reshape(n - 1, depth_num, 8)
    cost = cost_volume(feats, coefs, order, flavour)
    filtered = regnet(cost, weights, eps)
    prob = torch.softmax(-filtered, dim=0)                                              # model.py:345
    samples = torch.from_numpy(np.asarray(O.depth_samples(depth_num, depth_start, depth_interval), dtype=np.float64)
                               ).to(feats.dtype)
    depth = (samples[:, None, None] * prob).sum(dim=0)                                   # model.py:358-366
    return depth, prob, filtered


def loss_and_grads(feats, cams, gt_depth, depth_num, depth_start, depth_interval, weights, order="train",
                   flavour="exact_adjoint", dtype=torch.float64, eps=1e-5):
    """numpy in, numpy out: dict(loss, less_one, less_three, depth, grads={variable name: array}, dfeats [N,Hf,Wf,C])."""
    f = torch.tensor(np.asarray(feats), dtype=dtype, requires_grad=True)
    w = {k: torch.tensor(np.asarray(v), dtype=dtype, requires_grad=True) for k, v in weights.items()}
    depth_end = float(np.float32(depth_start) + np.float32(np.float32(depth_num) - np.float32(1.0)) * np.float32(depth_interval))
    depth, _prob, _filtered = forward(f, cams, depth_num, depth_start, depth_interval, w, order, flavour, eps)
    gt = torch.tensor(np.asarray(gt_depth), dtype=dtype)
    loss, l1, l3 = regression_loss(depth, gt, float(depth_start), depth_end)
    loss.backward()
    grads = {k: (v.grad.numpy().copy() if v.grad is not None else np.zeros(v.shape)) for k, v in w.items()}
    return dict(loss=float(loss.detach()), less_one=float(l1.detach()), less_three=float(l3.detach()), depth=depth.detach().numpy(),
                grads=grads, dfeats=f.grad.numpy().copy())
