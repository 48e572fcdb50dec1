"""Run one RegNetUS0 layer (bf16 / tcgen05) at a config's shapes: timing or ncu target."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvsnet_b200 import ops, synthetic  # noqa: E402

LV = {"3dconv1_0": 0, "3dconv2_0": 1, "3dconv3_0": 2, "3dconv0_1": 0, "3dconv1_1": 1, "3dconv2_1": 2,
      "3dconv3_1": 3, "3dconv4_0": 3, "3dconv5_0": 2, "3dconv6_0": 1, "3dconv6_2": 0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--layer", default="3dconv0_1")
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    cfg = synthetic.CONFIGS[a.config]
    D, hf, wf = cfg["depth_num"], cfg["height"] // 4, cfg["width"] // 4
    w = synthetic.make_regnet_weights()
    cin, cout, op, stride = synthetic.regnet_channels(32, 8)[a.layer]
    l = LV[a.layer]
    d, h, wd = D >> l, hf >> l, wf >> l
    x = torch.randn((d, h, wd, cin), device="cuda").to(torch.bfloat16)
    kern = torch.from_numpy(w[a.layer + "/kernel"]).cuda()
    # the two layers that read the cost volume have no normalisation on their input
    aff = (torch.ones(cin, device="cuda"), torch.zeros(cin, device="cuda")) if a.layer not in ("3dconv0_1", "3dconv1_0") else None
    skip = x if a.layer in ("3dconv5_0", "3dconv6_0", "3dconv6_2") else None
    od = torch.float32 if a.layer == "3dconv6_2" else torch.bfloat16
    ts = []
    for i in range(a.iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        ops.conv3d_layer(x, kern, stride, op == "deconv", "bf16", x_affine=aff, skip=skip,
                         skip_affine=aff if skip is not None else None, out_dtype=od)
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    print(a.layer, "dbg=%s" % os.environ.get("MVSB200_TC_DBG", "0"), "ms:", " ".join("%.3f" % t for t in ts), flush=True)


if __name__ == "__main__":
    main()
