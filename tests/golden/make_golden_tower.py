"""Regenerates tests/golden/tiny_tower.npz from the feature-tower restatement (run from the repo root).

Like tiny_hotpath.npz this pins the ORACLE (the reference ships no vectors; PARITY UNPINNED): the images, the
features and a digest of every layer's output on a seeded problem, so that later edits of the restatement and the
CUDA tower are compared against the same bytes.  The weights are regenerated from their seed by the tests
(synthetic.make_unet_weights(8)); a digest of them is stored to catch a change of the generator.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle.feature_oracle as FO  # noqa: E402
from mvsnet_b200 import synthetic  # noqa: E402


def weight_digest(w):
    return np.array([np.float64(np.abs(w[k]).sum()) for k in sorted(w)])


def main():
    w = synthetic.make_unet_weights(8)
    im = synthetic.make_images(2, 32, 48)
    feats, outs = FO.unet_ds2gn(im, w, return_layers=True)
    names = [s[0] for s in FO.unet_layer_specs(8)]
    digest = np.array([[np.float64(outs[n].mean()), np.float64(np.abs(outs[n]).mean()), np.float64(outs[n].max())] for n in names])
    path = os.path.join(ROOT, "tests", "golden", "tiny_tower.npz")
    np.savez_compressed(path, images=im, feats=feats, layer_digest=digest, weight_digest=weight_digest(w),
                        l2dconv5_0=outs["2dconv5_0"], l2dconv8_2=outs["2dconv8_2"][:, ::4, ::4, :])
    print(path, os.path.getsize(path) / 1e3, "kB")


if __name__ == "__main__":
    main()
