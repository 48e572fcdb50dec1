"""Is the planner's cycle model ranking the plans of a layer right?  Runs the model's top-K plans (tuning TC_RANK) of
the RegNetUS0 layers at a config's shapes through the stand-alone layer entry point (its layout conversions are a
constant on top) and prints device time per rank and the plan text (development aid)."""
import argparse
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvsnet_b200 import _lib, ops, synthetic  # noqa: E402

LV = {"3dconv2_0": 1, "3dconv3_0": 2, "3dconv1_1": 1, "3dconv2_1": 2, "3dconv3_1": 3, "3dconv4_0": 3, "3dconv5_0": 2,
      "3dconv6_0": 1, "3dconv6_2": 0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--ranks", type=int, default=6)
    ap.add_argument("--layers", default=",".join(LV))
    a = ap.parse_args()
    cfg = synthetic.CONFIGS[a.config]
    D, hf, wf = cfg["depth_num"], cfg["height"] // 4, cfg["width"] // 4
    w = synthetic.make_regnet_weights()
    lib = _lib.load()
    for layer in a.layers.split(","):
        cin, cout, op, stride = synthetic.regnet_channels(32, 8)[layer]
        l = LV[layer]
        d, h, wd = D >> l, hf >> l, wf >> l
        x = torch.randn((d, h, wd, cin), device="cuda").to(torch.bfloat16)
        kern = torch.from_numpy(w[layer + "/kernel"]).cuda()
        aff = (torch.ones(cin, device="cuda"), torch.zeros(cin, device="cuda"))
        skip = x if layer in ("3dconv5_0", "3dconv6_0", "3dconv6_2") else None
        od = torch.float32 if layer == "3dconv6_2" else torch.bfloat16
        mode = 2 if op == "deconv" else (1 if stride == 2 else 0)
        _lib.set_tuning("TC_LAYER", f"{cin},{cout},{mode}")
        print(layer, flush=True)
        for rank in range(a.ranks):
            _lib.set_tuning("TC_RANK", rank)
            ts = []
            for i in range(8):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                ops.conv3d_layer(x, kern, stride, op == "deconv", "bf16", x_affine=aff, skip=skip,
                                 skip_affine=aff if skip is not None else None, out_dtype=od)
                e.record()
                torch.cuda.synchronize()
                ts.append(s.elapsed_time(e))
            print("   rank %d: %.3f ms (min of 8, with layout conversions)" % (rank, min(ts[2:])), flush=True)
        _lib.set_tuning("TC_RANK", None)
        _lib.set_tuning("TC_LAYER", None)


if __name__ == "__main__":
    main()
