"""NumPy / PyTorch-CPU fp32 restatement of the MVSNet cost-volume hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: the
reference has no golden vectors and TensorFlow 1.12 cannot run here, so every
function below follows the reference *source text* (cited as file:line under
/root/reference) plus the TF-1.12 kernel semantics of SURVEY.md Appendix A.

Conventions fixed by this oracle (and mirrored bit-for-bit by the CUDA code):
  * all geometry is IEEE fp32, one rounding per operation, NO fused
    multiply-add, sums of three products associated left to right
    ``(p0 + p1) + p2``;
  * ``tf.matrix_inverse`` is restated as a partial-pivot LU (Doolittle) solve
    against the identity, the algorithm family TF's Eigen kernel uses;
  * batch-norm moments are accumulated in fp64 and rounded to fp32 once (TF's
    fp32 reduction tree is unspecified; fp64 is the value every tree
    approximates);
  * non-finite sample coordinates read as "outside the image" (value 0).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32

__all__ = [
    "inv3x3_lu", "matmul3", "tf_linspace", "plane_depths", "inv_depth_planes", "bilinear_fill0",
    "get_homographies", "get_homographies_inv_depth", "transform_coefs",
    "sample_coords", "projective_transform_bilinear", "tf_transform_homography",
    "get_pixel_grids", "interpolate", "homography_warping", "legacy_coords",
    "cost_volume", "tf_same_pads", "conv3d_same", "conv3d_transpose_same",
    "batch_norm_train", "regnet_us0", "regnet_layer_specs", "softmax_neg",
    "depth_samples", "soft_argmin", "get_probability_map_slice",
    "get_probability_map", "depth_regress", "inference_from_features",
    "REGNET_LAYERS",
]


def _f(x):
    return np.asarray(x, dtype=F32)


# --------------------------------------------------------------------------
# small dense algebra, explicit op order
# --------------------------------------------------------------------------
def matmul3(A, B):
    """Batched [...,3,3] @ [...,3,n] with C_ij = (a_i0 b_0j + a_i1 b_1j) + a_i2 b_2j.

    Restates the ``tf.matmul`` calls of homography_warping.py:39-56 with a
    fixed association and one fp32 rounding per multiply/add.
    """
    A = _f(A)
    B = _f(B)
    p0 = A[..., :, 0:1] * B[..., 0:1, :]
    p1 = A[..., :, 1:2] * B[..., 1:2, :]
    p2 = A[..., :, 2:3] * B[..., 2:3, :]
    return (p0 + p1) + p2


def inv3x3_lu(A):
    """fp32 3x3 inverse: LU with partial pivoting, then 3 triangular solves.

    Restates ``tf.matrix_inverse`` (homography_warping.py:33,81).  Scalar fp32
    arithmetic; elimination ``u_ij = u_ij - l_ik*u_kj``; forward substitution
    ``y_i = (b_i - l_i0 y_0) - l_i1 y_1``; back substitution
    ``x_i = ((y_i - u_i,i+1 x_i+1) - u_i,i+2 x_i+2) / u_ii``.
    """
    A = _f(A)
    if A.ndim > 2:
        return np.stack([inv3x3_lu(a) for a in A], axis=0)
    lu = [[F32(A[i, j]) for j in range(3)] for i in range(3)]
    perm = [0, 1, 2]
    for k in range(3):
        p = k
        best = abs(lu[k][k])
        for i in range(k + 1, 3):
            if abs(lu[i][k]) > best:
                best = abs(lu[i][k])
                p = i
        if p != k:
            lu[k], lu[p] = lu[p], lu[k]
            perm[k], perm[p] = perm[p], perm[k]
        for i in range(k + 1, 3):
            lu[i][k] = F32(lu[i][k] / lu[k][k])
            for j in range(k + 1, 3):
                lu[i][j] = F32(lu[i][j] - F32(lu[i][k] * lu[k][j]))
    inv = np.zeros((3, 3), dtype=F32)
    for c in range(3):
        b = [F32(1.0) if perm[i] == c else F32(0.0) for i in range(3)]
        y0 = b[0]
        y1 = F32(b[1] - F32(lu[1][0] * y0))
        y2 = F32(F32(b[2] - F32(lu[2][0] * y0)) - F32(lu[2][1] * y1))
        x2 = F32(y2 / lu[2][2])
        x1 = F32(F32(y1 - F32(lu[1][2] * x2)) / lu[1][1])
        x0 = F32(F32(F32(y0 - F32(lu[0][1] * x1)) - F32(lu[0][2] * x2)) / lu[0][0])
        inv[0, c], inv[1, c], inv[2, c] = x0, x1, x2
    return inv


def tf_linspace(start, stop, num):
    """TF-1.12 LinSpace: step=(stop-start)/(num-1); out[i]=start+step*i (Appendix A.2)."""
    start = F32(start)
    stop = F32(stop)
    if num == 1:
        return np.array([start], dtype=F32)
    step = F32(F32(stop - start) / F32(num - 1))
    i = np.arange(num, dtype=F32)
    return (start + step * i).astype(F32)


def plane_depths(depth_num, depth_start, depth_interval):
    """depth[i] = float(i)*interval + start (homography_warping.py:28-30)."""
    i = np.arange(int(depth_num), dtype=F32)
    return (i * F32(depth_interval) + F32(depth_start)).astype(F32)


# --------------------------------------------------------------------------
# a1 / a2: plane-sweep homographies
# --------------------------------------------------------------------------
def _homographies_from_depths(left_cam, right_cam, depth):
    """Shared body of homography_warping.py:33-56 and :81-104.  depth: [B,D]."""
    left_cam = _f(left_cam)
    right_cam = _f(right_cam)
    R_left = left_cam[:, 0, :3, :3]
    R_right = right_cam[:, 0, :3, :3]
    t_left = left_cam[:, 0, :3, 3:4]
    t_right = right_cam[:, 0, :3, 3:4]
    K_left = left_cam[:, 1, :3, :3]
    K_right = right_cam[:, 1, :3, :3]

    K_left_inv = inv3x3_lu(K_left)                                  # :33
    R_left_trans = np.swapaxes(R_left, 1, 2)                        # :34
    R_right_trans = np.swapaxes(R_right, 1, 2)                      # :35
    fronto_direction = R_left[:, 2:3, :]                            # :37  [B,1,3]
    c_left = -matmul3(R_left_trans, t_left)                         # :39
    c_right = -matmul3(R_right_trans, t_right)                      # :40
    c_relative = (c_right - c_left).astype(F32)                     # :41  [B,3,1]
    temp_vec = (c_relative * fronto_direction).astype(F32)          # :45  inner dim 1
    depth_mat = _f(depth)[:, :, None, None]                         # :46
    eye = np.eye(3, dtype=F32)[None, None]
    middle_mat0 = (eye - temp_vec[:, None] / depth_mat).astype(F32)  # :50
    middle_mat1 = matmul3(R_left_trans, K_left_inv)[:, None]        # :51
    middle_mat2 = matmul3(middle_mat0, middle_mat1)                 # :52
    homographies = matmul3(K_right[:, None],
                           matmul3(R_right[:, None], middle_mat2))  # :54-56
    return homographies.astype(F32)


def get_homographies(left_cam, right_cam, depth_num, depth_start, depth_interval,
                     batch_index=0):
    """homography_warping.py:10-58.  cams [B,2,4,4] -> [B,D,3,3] (ref -> source image coords)."""
    depth_start = np.atleast_1d(_f(depth_start))
    depth_interval = np.atleast_1d(_f(depth_interval))
    depth = np.stack([plane_depths(depth_num, s, i)
                      for s, i in zip(depth_start, depth_interval)], axis=0)
    return _homographies_from_depths(left_cam, right_cam, depth)


def inv_depth_planes(depth_num, depth_start, depth_end):
    """d = 1 / linspace(1/start, 1/end, D) (homography_warping.py:74-77)."""
    inv_start = F32(F32(1.0) / F32(depth_start))
    inv_end = F32(F32(1.0) / F32(depth_end))
    inv_depth = tf_linspace(inv_start, inv_end, int(depth_num))
    return (F32(1.0) / inv_depth).astype(F32)


def get_homographies_inv_depth(left_cam, right_cam, depth_num, depth_start, depth_end):
    """homography_warping.py:60-106 (batch size 1 only, :94)."""
    ds = np.atleast_1d(_f(depth_start))
    de = np.atleast_1d(_f(depth_end))
    depth = inv_depth_planes(depth_num, ds[0], de[0])[None]
    return _homographies_from_depths(left_cam, right_cam, depth)


# --------------------------------------------------------------------------
# a3: tf_transform_homography = coefficient conversion + contrib transform
# --------------------------------------------------------------------------
def transform_coefs(homography):
    """homography_warping.py:216-250: image-coord H -> 8 pixel-coord coefficients."""
    h = _f(homography).reshape(-1, 9)
    a0, a1, a2, b0, b1, b2, c0, c1, c2 = [h[:, i] for i in range(9)]
    two = F32(2)
    four = F32(4)
    a_0 = a0 - c0 / two
    a_1 = a1 - c1 / two
    a_2 = (a0 + a1) / two + a2 - (c0 + c1) / four - c2 / two
    b_0 = b0 - c0 / two
    b_1 = b1 - c1 / two
    b_2 = (b0 + b1) / two + b2 - (c0 + c1) / four - c2 / two
    c_0 = c0
    c_1 = c1
    c_2 = c2 + (c0 + c1) / two
    lin = np.stack([a_0, a_1, a_2, b_0, b_1, b_2, c_0, c_1], axis=1).astype(F32)
    return (lin / c_2[:, None]).astype(F32)


def sample_coords(coefs, height, width):
    """Source sample position of every output pixel (TF ImageProjectiveTransform, Appendix A.3).

    proj = (t6*x + t7*y) + 1;  ix = ((t0*x + t1*y) + t2)/proj;  iy likewise.
    """
    t = _f(coefs).reshape(8)
    xs = np.arange(width, dtype=F32)[None, :]
    ys = np.arange(height, dtype=F32)[:, None]
    with np.errstate(all="ignore"):
        proj = (t[6] * xs + t[7] * ys) + F32(1.0)
        ix = ((t[0] * xs + t[1] * ys) + t[2]) / proj
        iy = ((t[3] * xs + t[4] * ys) + t[5]) / proj
    return ix.astype(F32), iy.astype(F32)


def _read_fill0(image, yy, xx):
    H, W = image.shape[:2]
    with np.errstate(all="ignore"):
        valid = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
    yi = np.where(valid, yy, 0).astype(np.int64)
    xi = np.where(valid, xx, 0).astype(np.int64)
    return image[yi, xi] * valid[..., None].astype(F32)


def bilinear_fill0(image, ix, iy):
    """Bilinear read with zero fill, TF contrib image_ops.h order (Appendix A.3)."""
    image = _f(image)
    with np.errstate(all="ignore"):
        finite = np.isfinite(ix) & np.isfinite(iy)
        ixs = np.where(finite, ix, F32(-10.0)).astype(F32)
        iys = np.where(finite, iy, F32(-10.0)).astype(F32)
        xf = np.floor(ixs)
        yf = np.floor(iys)
        xc = xf + F32(1.0)
        yc = yf + F32(1.0)
        wxl = (xc - ixs)[..., None]
        wxr = (ixs - xf)[..., None]
        v_yf = wxl * _read_fill0(image, yf, xf) + wxr * _read_fill0(image, yf, xc)
        v_yc = wxl * _read_fill0(image, yc, xf) + wxr * _read_fill0(image, yc, xc)
        out = (yc - iys)[..., None] * v_yf + (iys - yf)[..., None] * v_yc
    return np.where(finite[..., None], out, F32(0.0)).astype(F32)


def projective_transform_bilinear(image, coefs):
    """tf.contrib.image.transform(image[H,W,C], coefs[8], 'BILINEAR') (homography_warping.py:251)."""
    image = _f(image)
    ix, iy = sample_coords(coefs, image.shape[0], image.shape[1])
    return bilinear_fill0(image, ix, iy)


def tf_transform_homography(input_image, homography):
    """homography_warping.py:211-253.  input [B,H,W,C], homography [B,3,3]."""
    input_image = _f(input_image)
    coefs = transform_coefs(homography)
    return np.stack([projective_transform_bilinear(input_image[b], coefs[b])
                     for b in range(input_image.shape[0])], axis=0)


# --------------------------------------------------------------------------
# a4: legacy homography_warping / interpolate (clamp sampler)
# --------------------------------------------------------------------------
def get_pixel_grids(height, width):
    """homography_warping.py:108-117: [3*H*W] = concat(x, y, 1) at pixel centres."""
    x_lin = tf_linspace(0.5, F32(width) - F32(0.5), width)
    y_lin = tf_linspace(0.5, F32(height) - F32(0.5), height)
    xg, yg = np.meshgrid(x_lin, y_lin)
    xg = xg.reshape(-1)
    yg = yg.reshape(-1)
    return np.concatenate([xg, yg, np.ones_like(xg)], axis=0).astype(F32)


def legacy_coords(homography, height, width):
    """Warped (x, y) image coordinates of homography_warping.py:190-203 for one 3x3."""
    h = _f(homography).reshape(3, 3)
    grid = get_pixel_grids(height, width).reshape(3, -1)
    gx, gy, g1 = grid[0], grid[1], grid[2]
    with np.errstate(all="ignore"):
        ax = (h[0, 0] * gx + h[0, 1] * gy) + h[0, 2] * g1
        ay = (h[1, 0] * gx + h[1, 1] * gy) + h[1, 2] * g1
        dv = (h[2, 0] * gx + h[2, 1] * gy) + h[2, 2] * g1
        dv = dv + (dv == 0).astype(F32) * F32(1e-7)                  # :197-198
        xw = (ax / dv).astype(F32)
        yw = (ay / dv).astype(F32)
    return xw, yw


def _floor_to_int(v):
    """floor -> int32 with the out-of-range / NaN cases pinned (NaN -> 0, clamp to +-2^30)."""
    with np.errstate(all="ignore"):
        fl = np.floor(v)
    fl = np.where(np.isnan(fl), F32(0.0), fl)
    fl = np.clip(fl, F32(-1073741824.0), F32(1073741824.0))
    return fl.astype(np.int64)


def interpolate(image, x, y):
    """homography_warping.py:131-174.  image [B,H,W,C]; x,y flat [B*H*W] image coords."""
    image = _f(image)
    B, H, W, C = image.shape
    x = (_f(x) - F32(0.5)).astype(F32)                               # :138
    y = (_f(y) - F32(0.5)).astype(F32)
    x0 = _floor_to_int(x)
    x1 = x0 + 1
    y0 = _floor_to_int(y)
    y1 = y0 + 1
    x0 = np.clip(x0, 0, W - 1)                                       # :146-149
    x1 = np.clip(x1, 0, W - 1)
    y0 = np.clip(y0, 0, H - 1)
    y1 = np.clip(y1, 0, H - 1)
    b = np.repeat(np.arange(B), H * W)
    pa = image[b, y0, x0]
    pb = image[b, y0, x1]
    pc = image[b, y1, x0]
    pd = image[b, y1, x1]
    x0f, x1f, y0f, y1f = (v.astype(F32) for v in (x0, x1, y0, y1))
    with np.errstate(all="ignore"):
        area_a = ((y1f - y) * (x1f - x))[:, None]                    # :166-169
        area_b = ((y1f - y) * (x - x0f))[:, None]
        area_c = ((y - y0f) * (x1f - x))[:, None]
        area_d = ((y - y0f) * (x - x0f))[:, None]
        out = ((area_a * pa + area_b * pb) + area_c * pc) + area_d * pd  # add_n :170-173
    return out.astype(F32)


def homography_warping(input_image, homography):
    """homography_warping.py:176-210 (dead code upstream; sampler mode 'legacy' here)."""
    input_image = _f(input_image)
    B, H, W, C = input_image.shape
    homography = _f(homography).reshape(B, 3, 3)
    xs, ys = [], []
    for b in range(B):
        xw, yw = legacy_coords(homography[b], H, W)
        xs.append(xw)
        ys.append(yw)
    out = interpolate(input_image, np.concatenate(xs), np.concatenate(ys))
    return out.reshape(B, H, W, C)


# --------------------------------------------------------------------------
# a5 / a6: N-view variance cost volume
# --------------------------------------------------------------------------
def cost_volume(feats, homographies, order="mem", sampler="transform", return_warped=False):
    """model.py:423-463 (order='mem', inference_mem) / model.py:315-334 (order='train').

    feats [N,Hf,Wf,C] (view 0 = reference), homographies [N-1,D,3,3] -> [D,Hf,Wf,C].
    """
    feats = _f(feats)
    N = feats.shape[0]
    D = homographies.shape[1]
    ref = feats[0]
    ref2 = (ref * ref).astype(F32)
    n_f = F32(N)
    nn_f = F32(N * N)
    out = np.empty((D,) + ref.shape, dtype=F32)
    warped_all = np.empty((N - 1, D) + ref.shape, dtype=F32) if return_warped else None
    for d in range(D):
        S = ref
        Q = ref2
        for v in range(N - 1):
            if sampler == "transform":
                w = tf_transform_homography(feats[v + 1][None], homographies[v, d][None])[0]
            elif sampler == "legacy":
                w = homography_warping(feats[v + 1][None], homographies[v, d][None])[0]
            else:
                raise ValueError(sampler)
            if return_warped:
                warped_all[v, d] = w
            S = (S + w).astype(F32)
            Q = (Q + w * w).astype(F32)
        if order == "mem":                                           # model.py:458-461
            A = ((S * S) / nn_f).astype(F32)
            cost = (Q / n_f - A).astype(F32)
        elif order == "train":                                       # model.py:330-332
            mean = (S / n_f).astype(F32)
            mean2 = (Q / n_f).astype(F32)
            cost = (mean2 - mean * mean).astype(F32)
        else:
            raise ValueError(order)
        out[d] = cost
    if return_warped:
        return out, warped_all
    return out


# --------------------------------------------------------------------------
# a7: RegNetUS0 (3-D U-Net regularizer)
# --------------------------------------------------------------------------
# (name, op, input layer, skip-add layer or None, Cout multiplier of base_filter, stride, has_bn_relu)
REGNET_LAYERS = [
    ("3dconv1_0", "conv", "data", 2, 2),        # mvsnetworks.py:131
    ("3dconv2_0", "conv", "3dconv1_0", 4, 2),   # :132
    ("3dconv3_0", "conv", "3dconv2_0", 8, 2),   # :133
    ("3dconv0_1", "conv", "data", 1, 1),        # :135-136
    ("3dconv1_1", "conv", "3dconv1_0", 2, 1),   # :138-139
    ("3dconv2_1", "conv", "3dconv2_0", 4, 1),   # :141-142
    ("3dconv3_1", "conv", "3dconv3_0", 8, 1),   # :144-145
    ("3dconv4_0", "deconv", "3dconv3_1", 4, 2),  # :146
    ("3dconv5_0", "deconv", "3dconv4_1", 2, 2),  # :150   input = add(4_0, 2_1) :148-149
    ("3dconv6_0", "deconv", "3dconv5_1", 1, 2),  # :154   input = add(5_0, 1_1) :152-153
    ("3dconv6_2", "conv", "3dconv6_1", 0, 1),   # :158   input = add(6_0, 0_1) :156-157; Cout=1, no BN
]


def regnet_layer_specs(in_channels=32, base_filter=8):
    """[(name, op, Cin, Cout, stride)] in execution order (mvsnetworks.py:125-158)."""
    ch = {"data": in_channels}
    specs = []
    for name, op, src, mult, stride in REGNET_LAYERS:
        cout = base_filter * mult if mult else 1
        src_ch = {"3dconv4_1": ch.get("3dconv4_0"), "3dconv5_1": ch.get("3dconv5_0"),
                  "3dconv6_1": ch.get("3dconv6_0")}.get(src, ch.get(src))
        specs.append((name, op, src_ch, cout, stride))
        ch[name] = cout
    return specs


def tf_same_pads(size, k, s):
    """TF SAME: out=ceil(in/s); total=max((out-1)s+k-in,0); before=total//2 (Appendix A.5)."""
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


def conv3d_same(x, w, stride):
    """tf.layers.conv3d SAME, no bias (network.py:210).  x [D,H,W,Cin], w [3,3,3,Cin,Cout]."""
    import torch
    import torch.nn.functional as Fn
    xt = torch.from_numpy(np.ascontiguousarray(_f(x))).permute(3, 0, 1, 2)[None]
    wt = torch.from_numpy(np.ascontiguousarray(_f(w))).permute(4, 3, 0, 1, 2).contiguous()
    pads = []
    for dim in (3, 2, 1):  # F.pad wants last dim first: W, H, D
        pads.extend(tf_same_pads(xt.shape[dim + 1], 3, stride))
    xt = Fn.pad(xt, pads)
    y = Fn.conv3d(xt, wt, stride=stride, padding=0)
    return y[0].permute(1, 2, 3, 0).contiguous().numpy()


def conv3d_transpose_same(x, w):
    """tf.layers.conv3d_transpose SAME stride 2, no bias (network.py:327).

    x [D,H,W,Cin], w [3,3,3,Cout,Cin];  out[2i+k] += x[i]*w[k], cropped to 2*in (Appendix A.5).
    """
    import torch
    import torch.nn.functional as Fn
    xt = torch.from_numpy(np.ascontiguousarray(_f(x))).permute(3, 0, 1, 2)[None]
    wt = torch.from_numpy(np.ascontiguousarray(_f(w))).permute(4, 3, 0, 1, 2).contiguous()
    y = Fn.conv_transpose3d(xt, wt, stride=2, padding=0)
    D, H, W = x.shape[:3]
    y = y[:, :, :2 * D, :2 * H, :2 * W]
    return y[0].permute(1, 2, 3, 0).contiguous().numpy()


def batch_norm_train(x, gamma, beta, eps=1e-5, relu=True):
    """tf.layers.batch_normalization(training=True) + relu (network.py:493-509, Appendix A.6)."""
    x = _f(x)
    x64 = x.astype(np.float64).reshape(-1, x.shape[-1])
    mean64 = x64.mean(axis=0)
    var64 = ((x64 - mean64) ** 2).mean(axis=0)
    mean = mean64.astype(F32)
    var = var64.astype(F32)
    inv = (F32(1.0) / np.sqrt(var + F32(eps))).astype(F32) * _f(gamma)
    shift = (_f(beta) - mean * inv).astype(F32)
    y = (x * inv + shift).astype(F32)
    if relu:
        y = np.maximum(y, F32(0.0))
    return y


def regnet_us0(cost_volume_, weights, eps=1e-5, return_layers=False, round_fn=None):
    """RegNetUS0 forward (mvsnetworks.py:122-158).  cost [D,H,W,Cin] -> [D,H,W].

    weights: dict with TF variable names '<layer>/kernel', '<layer>/bn/gamma', '<layer>/bn/beta'.
    round_fn (optional) is applied to every conv input and kernel -- used by tests to
    model a bf16-operand implementation; the reference itself is plain fp32.
    """
    rf = (lambda a: a) if round_fn is None else round_fn
    layers = {"data": _f(cost_volume_)}

    def conv_bn(name, src, stride):
        y = conv3d_same(rf(layers[src]), rf(weights[name + "/kernel"]), stride)
        layers[name + "/raw"] = y
        layers[name] = batch_norm_train(y, weights[name + "/bn/gamma"], weights[name + "/bn/beta"], eps)

    def deconv_bn(name, src):
        y = conv3d_transpose_same(rf(layers[src]), rf(weights[name + "/kernel"]))
        layers[name + "/raw"] = y
        layers[name] = batch_norm_train(y, weights[name + "/bn/gamma"], weights[name + "/bn/beta"], eps)

    conv_bn("3dconv1_0", "data", 2)
    conv_bn("3dconv2_0", "3dconv1_0", 2)
    conv_bn("3dconv3_0", "3dconv2_0", 2)
    conv_bn("3dconv0_1", "data", 1)
    conv_bn("3dconv1_1", "3dconv1_0", 1)
    conv_bn("3dconv2_1", "3dconv2_0", 1)
    conv_bn("3dconv3_1", "3dconv3_0", 1)
    deconv_bn("3dconv4_0", "3dconv3_1")
    layers["3dconv4_1"] = (layers["3dconv4_0"] + layers["3dconv2_1"]).astype(F32)   # add_n :148-149
    deconv_bn("3dconv5_0", "3dconv4_1")
    layers["3dconv5_1"] = (layers["3dconv5_0"] + layers["3dconv1_1"]).astype(F32)
    deconv_bn("3dconv6_0", "3dconv5_1")
    layers["3dconv6_1"] = (layers["3dconv6_0"] + layers["3dconv0_1"]).astype(F32)
    out = conv3d_same(rf(layers["3dconv6_1"]), rf(weights["3dconv6_2/kernel"]), 1)   # :158
    layers["3dconv6_2"] = out
    out = out[..., 0]                                                               # model.py:468-469
    if return_layers:
        return out, layers
    return out


# --------------------------------------------------------------------------
# a8 / a9: softmax, soft-argmin, probability map
# --------------------------------------------------------------------------
def softmax_neg(filtered):
    """P = softmax(-F) along axis 0 (model.py:474): exp(x-max)/sum exp(x-max)."""
    x = -_f(filtered)
    m = x.max(axis=0, keepdims=True)
    e = np.exp((x - m).astype(F32)).astype(F32)
    return (e / e.sum(axis=0, keepdims=True, dtype=F32)).astype(F32)


def depth_samples(depth_num, depth_start, depth_interval, inverse_depth=False):
    """model.py:378-379,480-490: the D depth hypotheses used by the soft-argmin."""
    depth_start = F32(depth_start)
    depth_interval = F32(depth_interval)
    depth_end = F32(depth_start + F32(F32(depth_num) - F32(1.0)) * depth_interval)
    if inverse_depth:
        return inv_depth_planes(depth_num, depth_start, depth_end)
    return tf_linspace(depth_start, depth_end, int(depth_num))


def soft_argmin(prob_volume, depth_num, depth_start, depth_interval, inverse_depth=False):
    """depth = sum_i samples_i * P_i (model.py:487-494)."""
    samples = depth_samples(depth_num, depth_start, depth_interval, inverse_depth)
    return (samples[:, None, None] * _f(prob_volume)).sum(axis=0, dtype=F32).astype(F32)


def get_probability_map_slice(cv, depth_map, depth_start, depth_interval,
                              inverse_depth=False, num_buckets=4):
    """model.py:45-144.  cv [D,H,W] probability volume, depth_map [H,W] -> prob [H,W]."""
    cv = _f(cv)
    depth_map = _f(depth_map)
    D = cv.shape[0]
    depth_start = F32(depth_start)
    depth_interval = F32(depth_interval)
    if inverse_depth:                                                # :83-107
        depth_end = F32(depth_start + F32(F32(D) - F32(1.0)) * depth_interval)
        inv_start = F32(F32(1.0) / depth_start)
        inv_end = F32(F32(1.0) / depth_end)
        inv_interval = F32(F32(inv_start - inv_end) / F32(F32(D) - F32(1.0)))
        with np.errstate(all="ignore"):
            inv_data = (F32(1.0) / depth_map).astype(F32)
            inv_data = ((inv_data - inv_end) / inv_interval).astype(F32)
        l0 = D - _floor_to_int(np.ceil(inv_data)) - 1
        l0 = np.clip(l0, 0, D - 1)
        r0 = D - _floor_to_int(inv_data) - 1
        r0 = np.clip(r0, 0, D - 1)
    else:                                                            # :108-120
        with np.errstate(all="ignore"):
            idx = ((depth_map - depth_start) / depth_interval).astype(F32)
        l0 = np.clip(_floor_to_int(idx), 0, D - 1)
        r0 = np.clip(_floor_to_int(np.ceil(idx)), 0, D - 1)
    l1 = np.clip(l0 - 1, 0, D - 1)
    r1 = np.clip(r0 + 1, 0, D - 1)
    yy, xx = np.meshgrid(np.arange(cv.shape[1]), np.arange(cv.shape[2]), indexing="ij")
    prob = (cv[l0, yy, xx] + cv[r0, yy, xx]).astype(F32)             # :128-130
    if num_buckets == 4:                                             # :132-140
        prob = (prob + (cv[l1, yy, xx] + cv[r1, yy, xx]).astype(F32)).astype(F32)
    return prob


def get_probability_map(cv_batch, depth_map_batch, depth_start_batch, depth_interval_batch,
                        inverse_depth=False, num_buckets=4):
    """model.py:20-39.  cv [B,D,H,W], depth [B,H,W,1] -> [B,H,W,1]."""
    outs = []
    for i in range(cv_batch.shape[0]):
        outs.append(get_probability_map_slice(
            cv_batch[i], np.asarray(depth_map_batch[i]).reshape(cv_batch.shape[2], cv_batch.shape[3]),
            np.atleast_1d(depth_start_batch)[i], np.atleast_1d(depth_interval_batch)[i],
            inverse_depth, num_buckets))
    return np.stack(outs, axis=0)[..., None]


def depth_regress(filtered, depth_start, depth_interval, inverse_depth=False, num_buckets=4):
    """model.py:472-498: filtered cost [D,H,W] -> (depth [H,W], prob [H,W], P [D,H,W])."""
    P = softmax_neg(filtered)
    depth = soft_argmin(P, filtered.shape[0], depth_start, depth_interval, inverse_depth)
    prob = get_probability_map_slice(P, depth, depth_start, depth_interval, inverse_depth, num_buckets)
    return depth, prob, P


# --------------------------------------------------------------------------
# whole hot path (model.py:407-502 after the feature towers)
# --------------------------------------------------------------------------
def inference_from_features(feats, cams, depth_num, depth_start, depth_interval, weights,
                            order="mem", sampler="transform", inverse_depth=False,
                            round_fn=None, return_all=False):
    """feats [N,Hf,Wf,C], cams [N,2,4,4] -> (depth [Hf,Wf], prob [Hf,Wf])."""
    feats = _f(feats)
    cams = _f(cams)
    N = feats.shape[0]
    depth_start = F32(depth_start)
    depth_interval = F32(depth_interval)
    depth_end = F32(depth_start + F32(F32(depth_num) - F32(1.0)) * depth_interval)   # model.py:378-379
    homs = []
    for v in range(1, N):                                                            # model.py:410-420
        if inverse_depth:
            h = get_homographies_inv_depth(cams[0:1], cams[v:v + 1], depth_num, depth_start, depth_end)
        else:
            h = get_homographies(cams[0:1], cams[v:v + 1], depth_num, depth_start, depth_interval)
        homs.append(h[0])
    homs = np.stack(homs, axis=0)
    cost = cost_volume(feats, homs, order=order, sampler=sampler)
    filtered = regnet_us0(cost, weights, round_fn=round_fn)
    depth, prob, P = depth_regress(filtered, depth_start, depth_interval, inverse_depth)
    if return_all:
        return depth, prob, dict(homographies=homs, cost=cost, filtered=filtered, prob_volume=P)
    return depth, prob
