"""Feature tower timing at a named config (development aid): fp32 CUDA-core tower vs bf16 tensor-core tower."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvsnet_b200 import synthetic  # noqa: E402
from mvsnet_b200.features import FeatureTower  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--precisions", default="bf16,fp32")
    ap.add_argument("--size", default="", help="n,h,w instead of the config's (small sizes show the per-layer fixed cost)")
    ap.add_argument("--out", default="gpurun_out/tower_bench.json")
    a = ap.parse_args()
    cfg = synthetic.CONFIGS[a.config]
    n, h, w = cfg["n_views"], cfg["height"], cfg["width"]
    if a.size:
        n, h, w = (int(v) for v in a.size.split(","))
    im = torch.randn((n, h, w, 3), device="cuda")
    wts = synthetic.make_unet_weights(8)
    res = {"config": a.config, "views": n, "height": h, "width": w}
    for prec in a.precisions.split(","):
        tower = FeatureTower(wts, precision=prec)
        for _ in range(3):
            tower(im)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            tower(im)
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        res[prec + "_ms"] = float(np.median(ts))
        print(prec, res[prec + "_ms"], flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
