"""On-disk formats at the edges of the path: byte-level known answers derived from the reference
source (preprocess.py:116-155, 273-356; depthfusion.py:172-191) and round trips."""
import io
import os
import struct

import numpy as np
import pytest

from mvsnet_b200 import io_formats as F

CAM_TXT = """extrinsic
0.970263 0.00747983 0.241939 -191.02
-0.0147429 0.999493 0.0282234 3.28832
-0.241605 -0.030951 0.969881 22.5401
0.0 0.0 0.0 1.0

intrinsic
2892.33 0 823.205
0 2883.18 619.071
0 0 1

425 2.5
"""


def test_load_cam_word_counts():
    cam = F.load_cam(io.StringIO(CAM_TXT), interval_scale=1.06, max_d=192)          # 29 words: depth_min interval
    assert cam.shape == (2, 4, 4) and cam.dtype == np.float64
    assert cam[0][0][3] == -191.02 and cam[1][1][1] == 2883.18 and cam[1][0][2] == 823.205
    assert cam[1][3][0] == 425 and cam[1][3][1] == 2.5 * 1.06 and cam[1][3][2] == 192
    assert cam[1][3][3] == 425 + 2.5 * 1.06 * 192                                     # preprocess.py:139
    cam30 = F.load_cam(io.StringIO(CAM_TXT.rstrip() + " 128\n"), interval_scale=2.0)
    assert cam30[1][3][2] == 128 and cam30[1][3][3] == 425 + 5.0 * 128
    cam31 = F.load_cam(io.StringIO(CAM_TXT.rstrip() + " 128 999.5\n"))
    assert cam31[1][3][3] == 999.5
    with pytest.raises(ValueError):
        F.load_cam(io.StringIO(CAM_TXT))                                              # FLAGS.max_d has no stand-in


def test_write_cam_text_and_round_trip(tmp_path):
    cam = F.load_cam(io.StringIO(CAM_TXT), max_d=192)
    p = tmp_path / "00000000.txt"
    F.write_cam(str(p), cam)
    text = p.read_text()
    lines = text.split("\n")
    assert lines[0] == "extrinsic" and lines[6] == "intrinsic" and lines[5] == "" and lines[10] == ""
    assert lines[1] == "0.970263 0.00747983 0.241939 -191.02 "                        # str(float64) + trailing blank
    assert lines[11] == "425.0 2.5 192.0 905.0" and text.endswith("\n")
    back = F.load_cam(io.StringIO(text))                                               # 31 words now
    np.testing.assert_array_equal(back, cam)


def test_write_pfm_bytes(tmp_path):
    img = np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]], dtype=np.float32)
    p = tmp_path / "a.pfm"
    F.write_pfm(str(p), img)
    raw = p.read_bytes()
    head = b"Pf\n3 2\n-1.000000\n"                                                    # grey, width height, little-endian
    assert raw[:len(head)] == head
    assert raw[len(head):] == struct.pack("<6f", 4.0, 5.0, 6.0, 1.0, 2.0, 3.0)        # bottom row first
    with open(p, "rb") as f:
        back = F.load_pfm(f)
    np.testing.assert_array_equal(back, img)
    with pytest.raises(Exception, match="float32"):
        F.write_pfm(str(p), img.astype(np.float64))
    with pytest.raises(Exception, match="dimensions"):
        F.write_pfm(str(p), np.zeros((2, 2, 2), np.float32))


def test_pfm_colour_big_endian_and_bad_headers(tmp_path):
    rgb = np.arange(2 * 2 * 3, dtype=np.float32).reshape(2, 2, 3)
    p = tmp_path / "c.pfm"
    F.write_pfm(str(p), rgb)
    assert p.read_bytes().startswith(b"PF\n2 2\n")
    with open(p, "rb") as f:
        np.testing.assert_array_equal(F.load_pfm(f), rgb)
    be = b"Pf\n2 1\n1.0\n" + struct.pack(">2f", 7.5, -2.0)                             # positive scale = big-endian
    np.testing.assert_array_equal(F.load_pfm(io.BytesIO(be)), np.array([[7.5, -2.0]], np.float32))
    with pytest.raises(Exception, match="Not a PFM"):
        F.load_pfm(io.BytesIO(b"P5\n2 1\n1.0\n"))
    with pytest.raises(Exception, match="Malformed"):
        F.load_pfm(io.BytesIO(b"Pf\n2x1\n1.0\n"))


def test_png_conversions_and_probability_filter(tmp_path):
    d = np.array([[-3.0, 0.4, 700.9], [65534.7, 65536.0, 1e9]], np.float32)
    np.testing.assert_array_equal(F.depth_map_to_uint16(d), np.array([[0, 0, 700], [65534, 65535, 65535]], np.uint16))
    np.testing.assert_array_equal(F.confidence_map_to_uint16(np.array([0.0, 0.5, 1.0, 1.5], np.float32)),
                                  np.array([0, 32767, 65535, 65535], np.uint16))
    depth = np.array([[500.0, 600.0], [700.0, 800.0]], np.float32)
    prob = np.array([[0.9, 0.29], [0.3, 0.0]], np.float32)
    np.testing.assert_array_equal(F.filter_depth_by_probability(depth, prob, 0.3),
                                  np.array([[500.0, 0.0], [700.0, 0.0]], np.float32))    # strict <, depthfusion.py:188
    folder = tmp_path / "dense"
    cam = F.load_cam(io.StringIO(CAM_TXT), max_d=192)
    out_dir = folder / "depths_mvsnet"
    F.write_output_slice(str(out_dir), depth[None, :, :, None], prob[None, :, :, None], cam[None], 7)
    assert sorted(os.listdir(out_dir)) == ["7.txt", "7_depth.png", "7_init.pfm", "7_prob.pfm", "7_prob.png"]
    with pytest.raises(FileNotFoundError):              # depthfusion.py finds the views by <index>.jpg: none written yet
        F.probability_filter(str(folder), 0.3)
    image = np.linspace(-1.0, 1.0, 2 * 2 * 3, dtype=np.float32).reshape(2, 2, 3)         # a centred image
    F.write_output_slice(str(out_dir), depth, prob, cam, 7, out_ref_image=image)
    assert (out_dir / "7.jpg").read_bytes()[:3] == b"\xff\xd8\xff"                       # a JPEG
    np.testing.assert_array_equal(F.read_png16(str(out_dir / "7_depth.png")), np.array([[500, 600], [700, 800]], np.uint16))
    np.testing.assert_array_equal(F.read_png16(str(out_dir / "7_prob.png")), F.confidence_map_to_uint16(prob))
    F.probability_filter(str(folder), 0.3)
    with open(out_dir / "7_prob_filtered.pfm", "rb") as f:
        np.testing.assert_array_equal(F.load_pfm(f), np.array([[500.0, 0.0], [700.0, 0.0]], np.float32))


def test_png16_is_a_standard_png(tmp_path):
    img = (np.arange(5 * 7, dtype=np.uint32).reshape(5, 7) * 1999 % 65536).astype(np.uint16)
    p = str(tmp_path / "x.png")
    F.write_png16(p, img)
    np.testing.assert_array_equal(F.read_png16(p), img)
    cv2 = pytest.importorskip("cv2")
    back = cv2.imread(p, cv2.IMREAD_UNCHANGED)                                          # an independent decoder
    assert back.dtype == np.uint16
    np.testing.assert_array_equal(back, img)
    with pytest.raises(ValueError):
        F.write_png16(p, img.astype(np.uint8))


def _camera_json(tx):
    pose = {"{},{}".format(i, j): float(i == j) for i in range(4) for j in range(4)}
    pose["0,3"], pose["1,3"], pose["2,3"] = tx, 0.25, -0.5                              # metres
    return {"pose": {"matrix": pose}, "intrinsics": {"fx": 700.0, "fy": 710.0, "px": 320.0, "py": 240.0}}


def test_cluster_and_covisibility(tmp_path):
    """mvs_cluster.py:91-140, cluster_generator.py:139-156: camera JSON -> cam [2,4,4] (translation in mm, depth row),
    view lists padded with the reference, empty clusters skipped."""
    import json
    session = tmp_path / "session"
    (session / "cameras").mkdir(parents=True)
    for i in range(4):
        (session / "cameras" / f"{i}.json").write_text(json.dumps(_camera_json(0.1 * i)))
    covis = {"0": {"views": [1, 2, 3], "min_depth": 400.0, "max_depth": 900.0},
             "1": {"views": [0], "min_depth": 410.0, "max_depth": 910.0},
             "2": {"views": [], "min_depth": 420.0, "max_depth": 920.0},
             "3": {"views": [2, 1, 0], "min_depth": 430.0, "max_depth": 930.0}}
    (session / "covisibility.json").write_text(json.dumps(covis))
    clusters = F.load_covisibility(str(session), view_num=3, depth_num=128, interval_scale=1.06)
    assert [c.ref_index for c in clusters] == [0, 1, 3]                                  # the empty one is skipped
    assert clusters[0].indices == [0, 1, 2] and clusters[1].indices == [1, 0, 1] and clusters[2].indices == [3, 2, 1]
    assert [c.ref_index for c in F.load_covisibility(str(session), 3, include_empty=True)] == [0, 1, 2, 3]
    assert len(F.load_covisibility(str(session), 3, max_clusters=2)) == 2
    cams = clusters[1].cameras()
    assert cams.shape == (3, 2, 4, 4) and cams.dtype == np.float32
    np.testing.assert_allclose(cams[0, 0, :3, 3], [100.0, 250.0, -500.0])                # metres -> millimetres
    np.testing.assert_array_equal(cams[0], cams[2])                                      # padded with the reference view
    np.testing.assert_allclose(cams[0, 1, :3, :3], [[700.0, 0, 320.0], [0, 710.0, 240.0], [0, 0, 1.0]])
    interval = (910.0 - 410.0) / 127 * 1.06
    np.testing.assert_allclose(cams[0, 1, 3], [410.0, interval, 128, 910.0], rtol=1e-6)
    assert clusters[0].image_path(2).endswith(os.path.join("images", "2.jpg"))
