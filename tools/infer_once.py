"""One whole-path inference at a named config (ncu target)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvsnet_b200 import synthetic  # noqa: E402
from mvsnet_b200.engine import HotPath  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--iters", type=int, default=1)
    ap.add_argument("--fast-features", action="store_true", help="i.i.d. features instead of the rendered plane")
    a = ap.parse_args()
    cfg = synthetic.CONFIGS[a.config]
    n, D = cfg["n_views"], cfg["depth_num"]
    hf, wf = cfg["height"] // 4, cfg["width"] // 4
    cams = synthetic.make_cameras(n, cfg["height"], cfg["width"], D, cfg["interval_scale"])
    if a.fast_features:
        feats = torch.randn((n, hf, wf, 32), device="cuda")
    else:
        feats = torch.from_numpy(synthetic.make_features(cams, hf, wf, 32)).cuda()
    camsd = torch.from_numpy(cams).cuda()
    ds, di = float(cams[0, 1, 3, 0]), float(cams[0, 1, 3, 1])
    eng = HotPath(n, D, hf, wf, synthetic.make_regnet_weights(), precision="bf16")
    for _ in range(a.iters):
        d, p = eng.infer(feats, camsd, ds, di)
    torch.cuda.synchronize()
    print("depth checksum", float(d.sum()))


if __name__ == "__main__":
    main()
