"""Time the image feature tower (UNetDS2GN, fp32 parity mode) at a config's image size: all views of one reference
view in one call.  Prints one JSON line (ms per call, images/s, GFLOP/s of the convolutions)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvsnet_b200 import synthetic  # noqa: E402
from mvsnet_b200.features import FeatureTower  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    cfg = synthetic.CONFIGS[a.config]
    n, h, w = cfg["n_views"], cfg["height"], cfg["width"]
    im = torch.randn((n, h, w, 3), device="cuda")
    tower = FeatureTower(synthetic.make_unet_weights(8))
    flop, hh, ww = 0.0, {-1: h}, {-1: w}
    from mvsnet_b200._lib import UNET_LAYER_TABLE
    ch = {-1: 3}
    for i, (name, op, k, s, mult, srcs, gn, relu) in enumerate(UNET_LAYER_TABLE):
        ih, iw = hh[srcs[0]], ww[srcs[0]]
        hh[i], ww[i] = (2 * ih, 2 * iw) if op == "deconv" else (-(-ih // s), -(-iw // s))
        ch[i] = 8 * mult
        cin = sum(ch[x] for x in srcs)
        px = ih * iw if op == "deconv" else hh[i] * ww[i]
        flop += 2.0 * n * px * k * k * cin * ch[i]
    for _ in range(2):
        tower(im)
    ts = []
    for _ in range(a.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tower(im)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    med = ts[len(ts) // 2]
    print(json.dumps({"config": a.config, "views": n, "image": [h, w], "ms": med, "min_ms": ts[0],
                      "images_per_s": n / med * 1e3, "conv_gflop": flop / 1e9, "tflops": flop / med / 1e9}))


if __name__ == "__main__":
    main()
