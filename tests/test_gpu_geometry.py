"""GPU parity: homographies / transform coefficients / sample coordinates bit-exact in fp32, warped
features <= 1e-5 abs (north_star gates), legacy sampler bit-exact, against the oracle and the golden fixture."""
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
    sm, major, minor = ops.device_info()
    assert major == 10
    return ops


def _assert_bits(a, b, what):
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
    assert same.all(), f"{what}: {np.count_nonzero(~same)} of {same.size} words differ; max abs diff " \
                       f"{np.nanmax(np.abs(a - b))}"


@pytest.mark.parametrize("name", ["tiny", "small", "cfg2"])
def test_homographies_and_coefs_bit_exact(ops, O, name):
    from mvsnet_b200 import synthetic
    cfg = synthetic.CONFIGS[name]
    cams = synthetic.make_cameras(cfg["n_views"], cfg["height"], cfg["width"], cfg["depth_num"], cfg["interval_scale"])
    D, ds, di = cfg["depth_num"], float(cams[0, 1, 3, 0]), float(cams[0, 1, 3, 1])
    H, T = ops.homographies(to_dev(cams), D, ds, di, want_transforms=True)
    Href = np.stack([O.get_homographies(cams[0:1], cams[v:v + 1], D, ds, di)[0] for v in range(1, cams.shape[0])])
    _assert_bits(H.cpu().numpy(), Href, "homographies")
    Tref = O.transform_coefs(Href.reshape(-1, 3, 3)).reshape(Href.shape[0], D, 8)
    _assert_bits(T.cpu().numpy(), Tref, "transform coefficients")
    _assert_bits(ops.transform_coefs(H).cpu().numpy().reshape(Tref.shape), Tref, "transform_coefs entry point")
    # inverse-depth planes
    de = ops.depth_end_f32(D, ds, di)
    Hi = ops.homographies(to_dev(cams), D, ds, de, inverse_depth=True)
    Hiref = np.stack([O.get_homographies_inv_depth(cams[0:1], cams[v:v + 1], D, ds, de)[0]
                      for v in range(1, cams.shape[0])])
    _assert_bits(Hi.cpu().numpy(), Hiref, "inverse-depth homographies")


def test_golden_homographies(ops, golden_tiny):
    g = golden_tiny
    H, T = ops.homographies(to_dev(g["cams"]), int(g["depth_num"]), float(g["depth_start"]),
                            float(g["depth_interval"]), want_transforms=True)
    _assert_bits(H.cpu().numpy(), g["homographies"], "golden homographies")
    _assert_bits(T.cpu().numpy(), g["transforms"], "golden transforms")
    de = ops.depth_end_f32(int(g["depth_num"]), float(g["depth_start"]), float(g["depth_interval"]))
    Hi = ops.homographies(to_dev(g["cams"]), int(g["depth_num"]), float(g["depth_start"]), de, inverse_depth=True)
    _assert_bits(Hi.cpu().numpy(), g["homographies_inv"], "golden inverse-depth homographies")


def test_random_cameras_bit_exact(ops, O):
    """Appendix B.9: random well-conditioned cameras (general K with skew, general R, t)."""
    rng = np.random.RandomState(7)
    for trial in range(6):
        n = 3
        cams = np.zeros((n, 2, 4, 4), dtype=np.float32)
        for v in range(n):
            q, _ = np.linalg.qr(rng.randn(3, 3))
            if np.linalg.det(q) < 0:
                q[:, 0] *= -1
            R = np.eye(3) if v == 0 else (np.eye(3) * 0.97 + 0.03 * q)
            u, _, vt = np.linalg.svd(R)
            R = u @ vt
            cams[v, 0, :3, :3] = R
            cams[v, 0, :3, 3] = rng.uniform(-150, 150, 3) * (v > 0)
            cams[v, 0, 3, 3] = 1
            f = rng.uniform(300, 700)
            cams[v, 1, :3, :3] = [[f, rng.uniform(-2, 2), rng.uniform(100, 200)], [0, f * rng.uniform(0.9, 1.1),
                                  rng.uniform(80, 160)], [0, 0, 1]]
        D, ds, di = 24, float(rng.uniform(300, 600)), float(rng.uniform(1, 5))
        H = ops.homographies(to_dev(cams), D, ds, di)
        Href = np.stack([O.get_homographies(cams[0:1], cams[v:v + 1], D, ds, di)[0] for v in range(1, n)])
        _assert_bits(H.cpu().numpy(), Href, f"random cameras trial {trial}")
        hh, ww = 40, 56
        for sampler in ("transform", "legacy"):
            c = ops.sample_coords(H[0, :4], hh, ww, sampler).cpu().numpy()
            for d in range(4):
                if sampler == "transform":
                    ix, iy = O.sample_coords(O.transform_coefs(Href[0, d])[0], hh, ww)
                else:
                    ix, iy = (a.reshape(hh, ww) for a in O.legacy_coords(Href[0, d], hh, ww))
                _assert_bits(c[d, :, :, 0], ix, f"{sampler} x coords")
                _assert_bits(c[d, :, :, 1], iy, f"{sampler} y coords")


def test_sample_coords_and_warp_vs_golden(ops, golden_tiny):
    g = golden_tiny
    D = int(g["depth_num"])
    H = to_dev(g["homographies"])
    hf, wf = g["feats"].shape[1:3]
    sel = torch.stack([H[0, 0], H[0, D - 1], H[1, 0], H[1, D - 1]])
    c = ops.sample_coords(sel, hf, wf, "transform").cpu().numpy()
    _assert_bits(c, g["coords"], "golden sample coordinates")
    w = ops.warp(to_dev(g["feats"][1][None]), H[0, 5][None], "transform").cpu().numpy()[0]
    assert np.abs(w - g["warped_v0_d5"]).max() <= 1e-5
    wl = ops.warp(to_dev(g["feats"][2][None]), H[1, 9][None], "legacy").cpu().numpy()[0]
    _assert_bits(wl, g["warped_legacy_v1_d9"], "golden legacy warp")


@pytest.mark.parametrize("channels", [32, 6, 1])
def test_warp_all_planes_vs_oracle(ops, O, small_problem, channels):
    p = small_problem
    feats = p["feats"][:, :, :, :channels]
    D = p["depth_num"]
    Href = np.stack([O.get_homographies(p["cams"][0:1], p["cams"][v:v + 1], D, p["depth_start"],
                                        p["depth_interval"])[0] for v in range(1, 3)])
    for v in range(2):
        out = ops.warp(to_dev(feats[v + 1][None]), to_dev(Href[v]), "transform").cpu().numpy()
        ref = np.stack([O.tf_transform_homography(feats[v + 1][None], Href[v, d][None])[0] for d in range(D)])
        assert np.abs(out - ref).max() <= 1e-5, np.abs(out - ref).max()
        outl = ops.warp(to_dev(feats[v + 1][None]), to_dev(Href[v, ::8]), "legacy").cpu().numpy()
        refl = np.stack([O.homography_warping(feats[v + 1][None], Href[v, d][None])[0] for d in range(0, D, 8)])
        _assert_bits(outl, refl, "legacy warp")


def test_border_kats(ops, O):
    """Appendix B.3 on the GPU: zero-fill ramp vs clamp zeros; identity; far-outside and degenerate H."""
    img = np.ones((1, 8, 8, 4), dtype=np.float32)
    Hs = np.array([[[1, 0, -0.5], [0, 1, 0], [0, 0, 1]],
                   [[1, 0, 0], [0, 1, 0], [0, 0, 1]],
                   [[1, 0, 500.0], [0, 1, -300.0], [0, 0, 1]],
                   [[1, 0, 0], [0, 1, 0], [0.5, 0.25, -2.0]],         # projective, denominator crosses zero
                   [[0, 0, 0], [0, 0, 0], [0, 0, 0]]], dtype=np.float32)
    for sampler, fn in (("transform", O.tf_transform_homography), ("legacy", O.homography_warping)):
        out = ops.warp(to_dev(img), to_dev(Hs), sampler).cpu().numpy()
        for i in range(Hs.shape[0]):
            with np.errstate(all="ignore"):
                ref = fn(img, Hs[i][None])[0]
            ok = np.isclose(out[i], ref, atol=1e-5) | (np.isnan(out[i]) & np.isnan(ref))
            assert ok.all(), (sampler, i, out[i][..., 0], ref[..., 0])


def test_interpolate_and_pixel_grids(ops, O):
    rng = np.random.RandomState(5)
    img = rng.randn(2, 12, 16, 8).astype(np.float32)
    x = rng.uniform(-3, 19, 2 * 12 * 16).astype(np.float32)
    y = rng.uniform(-3, 15, 2 * 12 * 16).astype(np.float32)
    out = ops.interpolate(to_dev(img), to_dev(x), to_dev(y)).cpu().numpy()
    _assert_bits(out, O.interpolate(img, x, y), "interpolate")
    _assert_bits(ops.pixel_grids(12, 16).cpu().numpy(), O.get_pixel_grids(12, 16), "pixel grids")


def test_reference_named_api(ops, O, tiny_problem):
    from mvsnet_b200 import homography_warping as hw
    p = tiny_problem
    cams = to_dev(p["cams"])
    H = hw.get_homographies(cams[0:1], cams[1:2], p["depth_num"], torch.tensor([p["depth_start"]]),
                            torch.tensor([p["depth_interval"]]))
    Href = O.get_homographies(p["cams"][0:1], p["cams"][1:2], p["depth_num"], p["depth_start"], p["depth_interval"])
    _assert_bits(H.cpu().numpy(), Href, "hw.get_homographies")
    f = to_dev(p["feats"][1][None])
    w = hw.tf_transform_homography(f, H[:, 3])
    assert np.abs(w.cpu().numpy() - O.tf_transform_homography(p["feats"][1][None], Href[:, 3])).max() <= 1e-5
    wl = hw.homography_warping(f, H[:, 3])
    _assert_bits(wl.cpu().numpy(), O.homography_warping(p["feats"][1][None], Href[:, 3]), "hw.homography_warping")
    with pytest.raises(Exception):
        hw.tf_transform_homography(f.cpu(), H[:, 3])             # no CPU path
