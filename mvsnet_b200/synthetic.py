"""Deterministic synthetic inputs for the hot path.

NumPy only; used by tests, bench.py and smoke().  No dataset or checkpoint is
needed: cameras follow the DTU-like geometry the reference is run on
(README.md:125), features are a textured fronto-parallel plane rendered
consistently into every view (so the cost volume has a true minimum), weights
are Glorot-uniform in TF variable layout (mvsnetworks.py:125-158).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32

# name, op, Cin multiplier source, Cout multiplier of base_filter (0 -> 1 channel), stride
REGNET_LAYER_ORDER = [
    "3dconv1_0", "3dconv2_0", "3dconv3_0", "3dconv0_1", "3dconv1_1", "3dconv2_1",
    "3dconv3_1", "3dconv4_0", "3dconv5_0", "3dconv6_0", "3dconv6_2",
]

CONFIGS = {
    # BASELINE.json configs -> (n_views, H, W, D, interval_scale)
    "cfg1": dict(n_views=3, height=512, width=640, depth_num=128, interval_scale=1.06),
    "cfg2": dict(n_views=5, height=864, width=1152, depth_num=192, interval_scale=1.06),
    "cfg5": dict(n_views=5, height=1184, width=1600, depth_num=256, interval_scale=0.8),
    # small shapes for CPU-speed parity tests (D, Hf, Wf multiples of 8)
    "tiny": dict(n_views=3, height=96, width=128, depth_num=16, interval_scale=8.0),
    "small": dict(n_views=5, height=192, width=256, depth_num=32, interval_scale=4.0),
}


def regnet_channels(in_channels=32, base_filter=8):
    """(Cin, Cout, op, stride) per layer of RegNetUS0 in REGNET_LAYER_ORDER (mvsnetworks.py:125-158)."""
    b = base_filter
    return {
        "3dconv1_0": (in_channels, 2 * b, "conv", 2),
        "3dconv2_0": (2 * b, 4 * b, "conv", 2),
        "3dconv3_0": (4 * b, 8 * b, "conv", 2),
        "3dconv0_1": (in_channels, b, "conv", 1),
        "3dconv1_1": (2 * b, 2 * b, "conv", 1),
        "3dconv2_1": (4 * b, 4 * b, "conv", 1),
        "3dconv3_1": (8 * b, 8 * b, "conv", 1),
        "3dconv4_0": (8 * b, 4 * b, "deconv", 2),
        "3dconv5_0": (4 * b, 2 * b, "deconv", 2),
        "3dconv6_0": (2 * b, b, "deconv", 2),
        "3dconv6_2": (b, 1, "conv", 1),
    }


def make_regnet_weights(in_channels=32, base_filter=8, seed=42):
    """TF-layout weights: conv '<l>/kernel' [3,3,3,Cin,Cout], deconv [3,3,3,Cout,Cin]; BN gamma/beta."""
    rng = np.random.RandomState(seed)
    w = {}
    for name, (cin, cout, op, _s) in regnet_channels(in_channels, base_filter).items():
        limit = np.sqrt(6.0 / (27 * cin + 27 * cout))       # glorot_uniform, tf.layers default
        shape = (3, 3, 3, cin, cout) if op == "conv" else (3, 3, 3, cout, cin)
        w[name + "/kernel"] = rng.uniform(-limit, limit, size=shape).astype(F32)
        if name != "3dconv6_2":
            w[name + "/bn/gamma"] = rng.uniform(0.5, 1.5, size=(cout,)).astype(F32)
            w[name + "/bn/beta"] = rng.normal(0.0, 0.1, size=(cout,)).astype(F32)
    return w


def unet_channels(base_filter=8, in_channels=3):
    """[(name, op, k, stride, Cin, Cout, has_gn)] of UNetDS2GN in the reference's build order (mvsnetworks.py:58-115)."""
    from ._lib import UNET_LAYER_TABLE
    out, ch = [], []
    for name, op, k, stride, mult, srcs, gn, _relu in UNET_LAYER_TABLE:
        cin = sum(in_channels if s < 0 else ch[s] for s in srcs)
        ch.append(base_filter * mult)
        out.append((name, op, k, stride, cin, ch[-1], gn))
    return out


def make_unet_weights(base_filter=8, seed=43):
    """TF-layout weights of the feature tower: conv '<l>/kernel' [k,k,Cin,Cout], deconv [k,k,Cout,Cin]; '<l>/gn/gamma',
    '<l>/gn/beta' (network.py:256-266).  Glorot-uniform kernels (tf.layers default), gamma ~ U(0.5,1.5), beta ~ N(0,0.1)."""
    rng = np.random.RandomState(seed)
    w = {}
    for name, op, k, _stride, cin, cout, gn in unet_channels(base_filter):
        limit = np.sqrt(6.0 / (k * k * cin + k * k * cout))
        shape = (k, k, cin, cout) if op == "conv" else (k, k, cout, cin)
        w[name + "/kernel"] = rng.uniform(-limit, limit, size=shape).astype(F32)
        if gn:
            w[name + "/gn/gamma"] = rng.uniform(0.5, 1.5, size=(cout,)).astype(F32)
            w[name + "/gn/beta"] = rng.normal(0.0, 0.1, size=(cout,)).astype(F32)
    return w


def make_images(n_views, height, width, seed=91):
    """Centred images [N,H,W,3] (zero mean, unit variance per image: mvs_data_generation/utils.py:33-38): a smooth
    random texture plus noise, a different crop per view."""
    rng = np.random.RandomState(seed)
    base = rng.normal(size=(height + 64, width + 64, 3))
    for _ in range(2):                                     # cheap blur: neighbour averages along both axes
        base = (base + np.roll(base, 1, 0) + np.roll(base, -1, 0)) / 3.0
        base = (base + np.roll(base, 1, 1) + np.roll(base, -1, 1)) / 3.0
    imgs = []
    for v in range(n_views):
        oy, ox = rng.randint(0, 64, size=2)
        im = base[oy:oy + height, ox:ox + width] + 0.05 * rng.normal(size=(height, width, 3))
        imgs.append((im - im.mean()) / np.sqrt(im.var() + 1e-8))
    return np.stack(imgs).astype(F32)


def make_scene_images(cams, height, width, seed=97, noise=0.02):
    """Centred images [N,H,W,3] of ONE scene: the textured plane make_features renders, seen through every camera of
    `cams` at full resolution (cams carry the intrinsics of the H/4 x W/4 feature maps, scale_camera in
    mvs_data_generation/utils.py:64-73, so K is scaled back by 4).  Unlike make_images the views are photo-consistent:
    the cost volume built from their feature towers has a real minimum at the plane."""
    full = np.array(cams, dtype=F32, copy=True)
    full[:, 1, :2, :3] *= 4.0
    im = make_features(full, height, width, channels=3, seed=seed, noise=noise)
    out = []
    for v in range(im.shape[0]):
        out.append((im[v] - im[v].mean()) / np.sqrt(im[v].var() + 1e-8))
    return np.stack(out).astype(F32)


def _look_at(cam_pos, target, roll_deg):
    """World->camera rotation for a camera at cam_pos looking at target (z forward, y down)."""
    z = target - cam_pos
    z = z / np.linalg.norm(z)
    up = np.array([0.0, 1.0, 0.0])
    x = np.cross(up, z)
    x = x / np.linalg.norm(x)
    y = np.cross(z, x)
    R = np.stack([x, y, z], axis=0)
    a = np.deg2rad(roll_deg)
    Rz = np.array([[np.cos(a), -np.sin(a), 0.0], [np.sin(a), np.cos(a), 0.0], [0.0, 0.0, 1.0]])
    return Rz @ R


def make_cameras(n_views, height, width, depth_num, interval_scale=1.06, seed=1234,
                 sample_scale=0.25, depth_start=425.0, plane_depth=680.0):
    """cams [N,2,4,4] fp32 in the reference layout (mvs_cluster.py:103-111), K scaled to feature res."""
    rng = np.random.RandomState(seed)
    f = 2892.3 * (width / 1600.0)
    K = np.array([[f, 0.0, width / 2.0], [0.0, f, height / 2.0], [0.0, 0.0, 1.0]])
    K[:2, :] *= sample_scale                                   # scale_camera, mvs_data_generation/utils.py:64-73
    depth_interval = 2.5 * interval_scale
    cams = np.zeros((n_views, 2, 4, 4), dtype=np.float64)
    target = np.array([0.0, 0.0, plane_depth])
    for v in range(n_views):
        if v == 0:
            R = np.eye(3)
            pos = np.zeros(3)
        else:
            ang = 2.0 * np.pi * (v - 1) / max(n_views - 1, 1) + rng.uniform(-0.3, 0.3)
            rad = rng.uniform(80.0, 160.0)
            pos = np.array([rad * np.cos(ang), rad * np.sin(ang), rng.uniform(-10.0, 10.0)])
            R = _look_at(pos, target, rng.uniform(-2.0, 2.0))
        t = -R @ pos
        cams[v, 0, :3, :3] = R
        cams[v, 0, :3, 3] = t
        cams[v, 0, 3, 3] = 1.0
        cams[v, 1, :3, :3] = K
        cams[v, 1, 3, 0] = depth_start
        cams[v, 1, 3, 1] = depth_interval
        cams[v, 1, 3, 2] = depth_num
        cams[v, 1, 3, 3] = depth_start + (depth_num - 1) * depth_interval
    return cams.astype(F32)


def _smooth_texture(h, w, c, rng, sigma=2.0):
    from scipy.ndimage import gaussian_filter
    tex = rng.standard_normal((h, w, c)).astype(F32)
    tex = gaussian_filter(tex, sigma=(sigma, sigma, 0.0), mode="wrap")
    tex /= tex.std() + 1e-12
    return tex.astype(F32)


def make_features(cams, hf, wf, channels=32, seed=5678, plane_depth=680.0, noise=0.05,
                  iid=False):
    """feats [N,Hf,Wf,C] fp32: a textured plane z=plane_depth seen from every camera (+ noise)."""
    rng = np.random.RandomState(seed)
    n = cams.shape[0]
    if iid:
        return rng.standard_normal((n, hf, wf, channels)).astype(F32)
    K0 = cams[0, 1, :3, :3].astype(np.float64)
    mm_per_texel = plane_depth / K0[0, 0]
    margin = int(np.ceil(220.0 / mm_per_texel)) + 8
    th, tw = hf + 2 * margin, wf + 2 * margin
    tex = _smooth_texture(th, tw, channels, rng)
    feats = np.empty((n, hf, wf, channels), dtype=F32)
    xs, ys = np.meshgrid(np.arange(wf) + 0.5, np.arange(hf) + 0.5)
    pix = np.stack([xs, ys, np.ones_like(xs)], axis=-1).reshape(-1, 3).T     # [3, HW]
    for v in range(n):
        K = cams[v, 1, :3, :3].astype(np.float64)
        R = cams[v, 0, :3, :3].astype(np.float64)
        t = cams[v, 0, :3, 3].astype(np.float64)
        ray = R.T @ (np.linalg.inv(K) @ pix)                 # world-space ray directions
        c = -R.T @ t
        s = (plane_depth - c[2]) / ray[2]
        X = c[0] + s * ray[0]
        Y = c[1] + s * ray[1]
        # texture coordinates: reference pixel grid at plane depth, shifted by margin
        u = X / mm_per_texel + K0[0, 2] + margin - 0.5
        w_ = Y / mm_per_texel + K0[1, 2] + margin - 0.5
        u = np.clip(u, 0.0, tw - 1.001)
        w_ = np.clip(w_, 0.0, th - 1.001)
        u0 = np.floor(u).astype(np.int64)
        v0 = np.floor(w_).astype(np.int64)
        fu = (u - u0).astype(F32)[:, None]
        fv = (w_ - v0).astype(F32)[:, None]
        val = ((1 - fv) * ((1 - fu) * tex[v0, u0] + fu * tex[v0, u0 + 1])
               + fv * ((1 - fu) * tex[v0 + 1, u0] + fu * tex[v0 + 1, u0 + 1]))
        feats[v] = val.reshape(hf, wf, channels)
    feats += rng.normal(0.0, noise, size=feats.shape).astype(F32)
    return feats.astype(F32)


def make_problem(name="cfg2", seed=0, channels=32, base_filter=8, iid=False):
    """One hot-path problem: dict(feats, cams, depth_num, depth_start, depth_interval, weights)."""
    cfg = CONFIGS[name]
    hf, wf = cfg["height"] // 4, cfg["width"] // 4
    cams = make_cameras(cfg["n_views"], cfg["height"], cfg["width"], cfg["depth_num"],
                        cfg["interval_scale"], seed=1234 + seed)
    feats = make_features(cams, hf, wf, channels, seed=5678 + seed, iid=iid)
    return dict(
        name=name, feats=feats, cams=cams, depth_num=cfg["depth_num"],
        depth_start=float(cams[0, 1, 3, 0]), depth_interval=float(cams[0, 1, 3, 1]),
        weights=make_regnet_weights(channels, base_filter, seed=42),
        hf=hf, wf=wf, n_views=cfg["n_views"],
    )
