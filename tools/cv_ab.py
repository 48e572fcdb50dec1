"""Cost-volume stage under the tuning switches (development aid): device time of stage 1 of mvsb200_infer per variant,
and the window kernel's global-memory fallback count."""
import argparse
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvsnet_b200 import _lib, synthetic  # noqa: E402
from mvsnet_b200.engine import HotPath  # noqa: E402

VARIANTS = {
    "window_fp16_blend_fp32_sums": {},
    "window_fp16_blend_fp16_sums": {"CV_FP32_BLEND": 2},
    "window_fp32_blend": {"CV_FP32_BLEND": 1},
    "gather_fp16_taps": {"CV_KERNEL": 1},
    "gather_fp32_taps": {"CV_KERNEL": 1, "CV_FP32_TAPS": 1},
}


def stage_ms(eng, feats, cams, ds, di, iters=10):
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(iters)]
    for _ in range(3):
        eng.infer(feats, cams, ds, di)
    for e in evs:
        eng.set_stage_events(e)
        eng.infer(feats, cams, ds, di)
    torch.cuda.synchronize()
    eng.set_stage_events(None)
    return np.array([[e[j].elapsed_time(e[j + 1]) for j in range(4)] for e in evs])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--variants", default=",".join(VARIANTS))
    ap.add_argument("--out", default="gpurun_out/cv_ab.json")
    a = ap.parse_args()
    p = synthetic.make_problem(a.config)
    feats, cams = torch.from_numpy(p["feats"]).cuda(), torch.from_numpy(p["cams"]).cuda()
    eng = HotPath(p["n_views"], p["depth_num"], p["hf"], p["wf"], p["weights"], precision="bf16")
    res = {"config": a.config}
    lib = _lib.load()
    for name in a.variants.split(","):
        for k, v in VARIANTS[name].items():
            _lib.set_tuning(k, v)
        ms = stage_ms(eng, feats, cams, p["depth_start"], p["depth_interval"])
        d, _ = eng.infer(feats, cams, p["depth_start"], p["depth_interval"])
        res[name] = {"cost_volume_ms": float(np.median(ms[:, 1])), "min_ms": float(ms[:, 1].min()),
                     "regularizer_ms": float(np.median(ms[:, 2])), "depth_checksum": float(d.sum())}
        for k in VARIANTS[name]:
            _lib.set_tuning(k, None)
        print(name, res[name], flush=True)
    _lib.set_tuning("CV_STATS", 1)
    n = ctypes.c_uint64()
    lib.mvsb200_cost_volume_window_stats(ctypes.byref(n), 1)
    eng.infer(feats, cams, p["depth_start"], p["depth_interval"])
    lib.mvsb200_cost_volume_window_stats(ctypes.byref(n), 1)
    _lib.set_tuning("CV_STATS", None)
    pairs = p["depth_num"] * p["hf"] * p["wf"] * (p["n_views"] - 1)
    res["window_fallback_pairs"] = int(n.value)
    res["window_fallback_share"] = n.value / pairs
    print("fallback pairs", n.value, "of", pairs, flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
