"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's image feature tower.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may import this module; the product
path (mvsnet_b200/) never does.  PARITY UNPINNED: TensorFlow 1.12 cannot run here and the reference ships no golden
vectors for this path, so the restatement below is anchored on the reference's own source lines only.

What it restates (file:line are /root/reference/mvsnet/...):
  * UNetDS2GN, the 2-D U-Net with group normalisation every view's image goes through before the cost volume
    (cnn_wrapper/mvsnetworks.py:53-115; called per view with shared weights, model.py:392-406);
  * conv_gn / deconv_gn (cnn_wrapper/network.py:218-276, :349-409): tf.layers.conv2d / conv2d_transpose, SAME padding,
    no bias, then group normalisation over groups of 8 channels (channel_wise=True, group_channel=8 -> G = max(1, C/8)),
    biased variance from tf.nn.moments over (C/G, H, W), x = (x - mean) / sqrt(var + eps) with eps = 1e-5
    (network.py:55), per-channel gamma / beta, ReLU for conv_gn (default relu=True) and NO ReLU for deconv_gn
    (default relu=False, never overridden by the U-Net);
  * the last layer conv10_2 is a plain convolution: no bias, no normalisation, no ReLU (mvsnetworks.py:115).

Arithmetic: fp32 convolutions (torch CPU); the moments are accumulated in fp64 and rounded to fp32, like the 3-D
restatement in mvs_oracle.py (TF reduces in fp32 with an unspecified tree; the difference is ~1e-7 relative).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32

# (name, op, kernel, stride, filters / base_filter, sources, group norm, relu); mvsnetworks.py:58-115.
# A source is 'data' or a layer name; two sources = tf.concat(axis=-1) in that order (the 2dconcatN_0 layers).
UNET_LAYERS = [
    ("2dconv1_0", "conv", 3, 2, 2, ("data",), True, True),                    # :59
    ("2dconv2_0", "conv", 3, 2, 4, ("2dconv1_0",), True, True),               # :60
    ("2dconv3_0", "conv", 3, 2, 8, ("2dconv2_0",), True, True),               # :61
    ("2dconv4_0", "conv", 3, 2, 16, ("2dconv3_0",), True, True),              # :62
    ("2dconv0_1", "conv", 3, 1, 1, ("data",), True, True),                    # :65
    ("2dconv0_2", "conv", 3, 1, 1, ("2dconv0_1",), True, True),               # :66
    ("2dconv1_1", "conv", 3, 1, 2, ("2dconv1_0",), True, True),               # :69
    ("2dconv1_2", "conv", 3, 1, 2, ("2dconv1_1",), True, True),               # :70
    ("2dconv2_1", "conv", 3, 1, 4, ("2dconv2_0",), True, True),               # :73
    ("2dconv2_2", "conv", 3, 1, 4, ("2dconv2_1",), True, True),               # :74
    ("2dconv3_1", "conv", 3, 1, 8, ("2dconv3_0",), True, True),               # :77
    ("2dconv3_2", "conv", 3, 1, 8, ("2dconv3_1",), True, True),               # :78
    ("2dconv4_1", "conv", 3, 1, 16, ("2dconv4_0",), True, True),              # :81
    ("2dconv4_2", "conv", 3, 1, 16, ("2dconv4_1",), True, True),              # :82
    ("2dconv5_0", "deconv", 3, 2, 8, ("2dconv4_2",), True, False),            # :83
    ("2dconv5_1", "conv", 3, 1, 8, ("2dconv5_0", "2dconv3_2"), True, True),   # :85-87
    ("2dconv5_2", "conv", 3, 1, 8, ("2dconv5_1",), True, True),               # :88
    ("2dconv6_0", "deconv", 3, 2, 4, ("2dconv5_2",), True, False),            # :89
    ("2dconv6_1", "conv", 3, 1, 4, ("2dconv6_0", "2dconv2_2"), True, True),   # :91-93
    ("2dconv6_2", "conv", 3, 1, 4, ("2dconv6_1",), True, True),               # :94
    ("2dconv7_0", "deconv", 3, 2, 2, ("2dconv6_2",), True, False),            # :95
    ("2dconv7_1", "conv", 3, 1, 2, ("2dconv7_0", "2dconv1_2"), True, True),   # :97-99
    ("2dconv7_2", "conv", 3, 1, 2, ("2dconv7_1",), True, True),               # :100
    ("2dconv8_0", "deconv", 3, 2, 1, ("2dconv7_2",), True, False),            # :101
    ("2dconv8_1", "conv", 3, 1, 1, ("2dconv8_0", "2dconv0_2"), True, True),   # :103-105
    ("2dconv8_2", "conv", 3, 1, 1, ("2dconv8_1",), True, True),               # :107
    ("conv9_0", "conv", 5, 2, 2, ("2dconv8_2",), True, True),                 # :108
    ("conv9_1", "conv", 3, 1, 2, ("conv9_0",), True, True),                   # :109
    ("conv9_2", "conv", 3, 1, 2, ("conv9_1",), True, True),                   # :110
    ("conv10_0", "conv", 5, 2, 4, ("conv9_2",), True, True),                  # :111
    ("conv10_1", "conv", 3, 1, 4, ("conv10_0",), True, True),                 # :112
    ("conv10_2", "conv", 3, 1, 4, ("conv10_1",), False, False),               # :113-115 (biased=False, relu=False)
]


def unet_layer_specs(base_filter=8, in_channels=3):
    """[(name, op, k, stride, Cin, Cout, sources, gn, relu)] in execution order."""
    ch = {"data": in_channels}
    specs = []
    for name, op, k, stride, mult, srcs, gn, relu in UNET_LAYERS:
        cin = sum(ch[s] for s in srcs)
        cout = base_filter * mult
        specs.append((name, op, k, stride, cin, cout, srcs, gn, relu))
        ch[name] = cout
    return specs


def _f(x):
    return np.asarray(x, dtype=F32)


def tf_same_pads(size, k, s):
    """TF SAME: out = ceil(in / s); total = max((out - 1) s + k - in, 0); before = total // 2."""
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


def conv2d_same(x, w, stride):
    """tf.layers.conv2d, SAME, no bias (network.py:172-215).  x [N,H,W,Cin], w [k,k,Cin,Cout] -> [N,ceil(H/s),ceil(W/s),Cout]."""
    import torch
    import torch.nn.functional as Fn
    k = w.shape[0]
    xt = torch.from_numpy(np.ascontiguousarray(_f(x))).permute(0, 3, 1, 2)
    wt = torch.from_numpy(np.ascontiguousarray(_f(w))).permute(3, 2, 0, 1).contiguous()
    pw, ph = tf_same_pads(xt.shape[3], k, stride), tf_same_pads(xt.shape[2], k, stride)
    xt = Fn.pad(xt, (pw[0], pw[1], ph[0], ph[1]))
    y = Fn.conv2d(xt, wt, stride=stride, padding=0)
    return y.permute(0, 2, 3, 1).contiguous().numpy()


def conv2d_transpose_same(x, w):
    """tf.layers.conv2d_transpose, SAME, stride 2, no bias (network.py:295-348).  x [N,H,W,Cin], w [k,k,Cout,Cin].

    The transposed convolution is the adjoint of the SAME stride-2 convolution from the 2H x 2W output grid to the
    H x W input grid, whose padding is (0 before, 1 after) for k = 3: out[2i + k] += x[i] * w[k], cropped to 2 * in.
    """
    import torch
    import torch.nn.functional as Fn
    assert w.shape[0] == 3
    xt = torch.from_numpy(np.ascontiguousarray(_f(x))).permute(0, 3, 1, 2)
    wt = torch.from_numpy(np.ascontiguousarray(_f(w))).permute(3, 2, 0, 1).contiguous()      # [Cin, Cout, k, k]
    y = Fn.conv_transpose2d(xt, wt, stride=2, padding=0)
    H, W = x.shape[1:3]
    return y[:, :, :2 * H, :2 * W].permute(0, 2, 3, 1).contiguous().numpy()


def group_norm(x, gamma, beta, eps=1e-5, relu=True, group_channel=8):
    """network.py:237-276: groups of `group_channel` consecutive channels, statistics per (sample, group)."""
    x = _f(x)
    n, h, w, c = x.shape
    g = max(1, c // group_channel)                                  # :246-247 (Python 2 integer division)
    xg = x.reshape(n, h * w, g, c // g).astype(np.float64)
    mean = xg.mean(axis=(1, 3), keepdims=True)
    var = ((xg - mean) ** 2).mean(axis=(1, 3), keepdims=True)       # tf.nn.moments: biased
    mean, var = mean.astype(F32), var.astype(F32)
    xn = (x.reshape(n, h * w, g, c // g) - mean) / np.sqrt(var + F32(eps))       # :253
    y = (xn.reshape(n, h, w, c) * _f(gamma) + _f(beta)).astype(F32)   # :269
    if relu:
        y = np.maximum(y, F32(0.0))
    return y


def unet_ds2gn(images, weights, base_filter=8, eps=1e-5, return_layers=False, round_fn=None):
    """UNetDS2GN forward (mvsnetworks.py:53-115).  images [N,H,W,3] (centred, mvs_data_generation/utils.py:33-38)
    -> features [N,H/4,W/4,4*base_filter].  H and W must be multiples of 16 (the concats need equal extents).

    weights: TF variable names '<layer>/kernel' ([k,k,Cin,Cout]; deconv [k,k,Cout,Cin]), '<layer>/gn/gamma', '<layer>/gn/beta'.
    round_fn (optional) is applied to every convolution input and kernel -- used by tests to model an implementation
    with bf16 operands and fp32 accumulation; the reference itself is plain fp32.
    """
    images = _f(images)
    if images.shape[1] % 16 or images.shape[2] % 16:
        raise ValueError("UNetDS2GN needs H and W to be multiples of 16")
    outs = {"data": images}
    for name, op, k, stride, cin, cout, srcs, gn, relu in unet_layer_specs(base_filter, images.shape[-1]):
        x = outs[srcs[0]] if len(srcs) == 1 else np.concatenate([outs[s] for s in srcs], axis=-1)
        kern = weights[name + "/kernel"]
        if round_fn is not None:
            x, kern = round_fn(x), round_fn(kern)
        y = conv2d_same(x, kern, stride) if op == "conv" else conv2d_transpose_same(x, kern)
        if gn:
            y = group_norm(y, weights[name + "/gn/gamma"], weights[name + "/gn/beta"], eps, relu)
        outs[name] = y
    return (outs["conv10_2"], outs) if return_layers else outs["conv10_2"]
