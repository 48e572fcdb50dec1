"""Per-stage device timings at a named config (development aid; bench.py is the contract)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvsnet_b200 import ops, synthetic  # noqa: E402
from mvsnet_b200.engine import HotPath  # noqa: E402


def timeit(fn, iters=10, warm=3, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--regnet", default="", help="comma list of precisions to time the regularizer in")
    ap.add_argument("--layers", action="store_true", help="time every regularizer layer alone (bf16)")
    ap.add_argument("--skip-cv", action="store_true", help="skip the cost-volume variant sweep")
    ap.add_argument("--out", default="gpurun_out/stage_bench.json")
    a = ap.parse_args()
    cfg = synthetic.CONFIGS[a.config]
    n, D = cfg["n_views"], cfg["depth_num"]
    hf, wf = cfg["height"] // 4, cfg["width"] // 4
    cams = synthetic.make_cameras(n, cfg["height"], cfg["width"], D, cfg["interval_scale"])
    t0 = time.time()
    feats = torch.from_numpy(synthetic.make_features(cams, hf, wf, 32)).cuda()
    print("features built in %.1fs" % (time.time() - t0), flush=True)
    camsd = torch.from_numpy(cams).cuda()
    ds, di = float(cams[0, 1, 3, 0]), float(cams[0, 1, 3, 1])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    res = {"config": a.config, "V": D * hf * wf}
    H = ops.homographies(camsd, D, ds, di)
    res["homographies_ms"] = timeit(lambda: ops.homographies(camsd, D, ds, di))
    V = D * hf * wf
    for dt, name in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
        out = torch.empty((D, hf, wf, 32), device="cuda", dtype=dt)
        for variant in (() if a.skip_cv else (1, 2, 3)):
            med, mn = timeit(lambda: ops.cost_volume(feats, H, variant=variant, out=out), flush=flush)
            nbytes = n * hf * wf * 32 * 4 + V * 32 * out.element_size()
            res[f"cost_volume_{name}_v{variant}"] = dict(ms=med, min_ms=mn, gbs=nbytes / med / 1e6)
            print(name, variant, med, nbytes / med / 1e6, "GB/s", flush=True)
    F = torch.randn((D, hf, wf), device="cuda")
    med, mn = timeit(lambda: ops.depth_regress(F, ds, di), flush=flush)
    res["regress"] = dict(ms=med, min_ms=mn, gbs=(V * 4 + 2 * hf * wf * 4) / med / 1e6)
    print("regress", med, flush=True)
    w = synthetic.make_regnet_weights()
    for prec in [p for p in a.regnet.split(",") if p]:
        eng = HotPath(n, D, hf, wf, w, precision=prec)
        cost = ops.cost_volume(feats, H, out_dtype=torch.bfloat16 if prec == "bf16" else torch.float32)
        med, mn = timeit(lambda: eng.regnet(cost), iters=3, warm=1)
        res[f"regnet_{prec}"] = dict(ms=med, min_ms=mn, tflops=22896.0 * V / med / 1e9)
        print("regnet", prec, med, flush=True)
        med, mn = timeit(lambda: eng.infer(feats, camsd, ds, di), iters=3, warm=1)
        res[f"infer_{prec}"] = dict(ms=med, min_ms=mn)
        print("infer", prec, med, flush=True)
    if a.layers:
        # every RegNetUS0 layer alone at this config's shapes (bf16 / tcgen05)
        ch = synthetic.regnet_channels(32, 8)
        lv = {"3dconv1_0": 0, "3dconv2_0": 1, "3dconv3_0": 2, "3dconv0_1": 0, "3dconv1_1": 1, "3dconv2_1": 2,
              "3dconv3_1": 3, "3dconv4_0": 3, "3dconv5_0": 2, "3dconv6_0": 1, "3dconv6_2": 0}
        res["layers"] = {}
        for name, (cin, cout, op, stride) in ch.items():
            l = lv[name]
            d, h, wd = D >> l, hf >> l, wf >> l
            x = torch.randn((d, h, wd, cin), device="cuda").to(torch.bfloat16)
            kern = torch.from_numpy(w[name + "/kernel"]).cuda()
            aff = (torch.ones(cin, device="cuda"), torch.zeros(cin, device="cuda")) if name != "3dconv0_1" else None
            skip = x if name in ("3dconv5_0", "3dconv6_0", "3dconv6_2") else None
            od = torch.float32 if name == "3dconv6_2" else torch.bfloat16
            fn = lambda: ops.conv3d_layer(x, kern, stride, op == "deconv", "bf16", x_affine=aff, skip=skip,
                                          skip_affine=aff if skip is not None else None, out_dtype=od)
            med, mn = timeit(fn, iters=5, warm=2, flush=flush)
            vin, vout = d * h * wd, (d * h * wd * 8 if op == "deconv" else d * h * wd // (stride ** 3))
            flop = 2.0 * 27 * cin * cout * (vin if op == "deconv" else vout)
            nbytes = vin * cin * 2 * (2 if skip is not None else 1) + vout * cout * (4 if name == "3dconv6_2" else 2)
            res["layers"][name] = dict(ms=med, min_ms=mn, tflops=flop / med / 1e9, gbs=nbytes / med / 1e6)
            print(name, "%.3f ms  %.1f TFLOP/s  %.0f GB/s" % (med, flop / med / 1e9, nbytes / med / 1e6), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
