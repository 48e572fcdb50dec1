"""Regenerates tests/golden/tiny_hotpath.npz from the oracle (run from the repo root).

The reference ships no golden vectors and TensorFlow 1.12 cannot run here (PARITY UNPINNED), so the
fixture pins the ORACLE: it freezes the oracle's outputs on the seeded 'tiny' problem so that later
oracle edits, the C restatement and the CUDA kernels are all compared against the same bytes.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle as O  # noqa: E402
from mvsnet_b200 import synthetic  # noqa: E402


def main():
    p = synthetic.make_problem("tiny")
    feats, cams = p["feats"], p["cams"]
    D, ds, di = p["depth_num"], p["depth_start"], p["depth_interval"]
    depth, prob, allr = O.inference_from_features(feats, cams, D, ds, di, p["weights"], return_all=True)
    H = allr["homographies"]
    T = O.transform_coefs(H.reshape(-1, 3, 3)).reshape(H.shape[0], D, 8)
    hf, wf = feats.shape[1:3]
    coords = np.stack([np.stack(O.sample_coords(T[v, d], hf, wf), axis=-1) for v in range(2) for d in (0, D - 1)])
    warped = O.tf_transform_homography(feats[1][None], H[0, 5][None])[0]
    warped_legacy = O.homography_warping(feats[2][None], H[1, 9][None])[0]
    cost_train = O.cost_volume(feats, H, order="train")
    H_inv = np.stack([O.get_homographies_inv_depth(cams[0:1], cams[v:v + 1], D, ds,
                                                   np.float32(ds) + np.float32(D - 1) * np.float32(di))[0]
                      for v in range(1, feats.shape[0])])
    out = dict(
        feats=feats, cams=cams, depth_num=np.int32(D), depth_start=np.float32(ds), depth_interval=np.float32(di),
        homographies=H, transforms=T, homographies_inv=H_inv, coords=coords, warped_v0_d5=warped,
        warped_legacy_v1_d9=warped_legacy, cost_mem_sub=allr["cost"][::2, ::2, ::2, :],
        cost_train_sub=cost_train[::2, ::2, ::2, :], filtered=allr["filtered"], depth=depth, prob=prob)
    path = os.path.join(ROOT, "tests", "golden", "tiny_hotpath.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
