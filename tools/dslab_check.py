"""torchrun target: D-slab mode against the single-GPU path on the same inputs, and its timing.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/dslab_check.py [--config small]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvsnet_b200 import synthetic  # noqa: E402
from mvsnet_b200.dslab import DSlabHotPath  # noqa: E402
from mvsnet_b200.engine import HotPath  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="small")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--iid", action="store_true", help="i.i.d. features (fast to build at the large configs)")
    ap.add_argument("--p2p", action="store_true", help="also run the peer-memory variant (exchange inside the kernels)")
    ap.add_argument("--abort-test", action="store_true",
                    help="with --p2p: rank 0 runs the layers alone (nobody publishes), a timer calls abort(): the waits give up")
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = synthetic.CONFIGS[a.config]
    n, D = cfg["n_views"], cfg["depth_num"]
    hf, wf = cfg["height"] // 4, cfg["width"] // 4
    cams = synthetic.make_cameras(n, cfg["height"], cfg["width"], D, cfg["interval_scale"])
    if a.iid:
        feats = torch.from_numpy(np.random.RandomState(5).randn(n, hf, wf, 32).astype(np.float32)).to(dev)
    else:
        feats = torch.from_numpy(synthetic.make_features(cams, hf, wf, 32)).to(dev)
    camsd = torch.from_numpy(cams).to(dev)
    ds, di = float(cams[0, 1, 3, 0]), float(cams[0, 1, 3, 1])
    w = synthetic.make_regnet_weights()
    single = HotPath(n, D, hf, wf, w, precision="bf16", device=dev)
    d1, p1 = single.infer(feats, camsd, ds, di)
    d1, p1 = d1.clone(), p1.clone()
    slab = DSlabHotPath(n, D, hf, wf, single.weights, device=dev)
    d2, p2 = slab.infer(feats, camsd, ds, di)
    torch.cuda.synchronize()
    err = (d1 - d2).abs()
    res = {"rank": rank, "world": world, "config": a.config, "max_abs_depth_diff_in_intervals": float(err.max()) / di,
           "frac_within_0.01_interval": float((err <= 0.01 * di).float().mean()),
           "max_abs_prob_diff": float((p1 - p2).abs().max())}

    def timeit(fn):
        for _ in range(3):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(a.iters):
            fn()
        e.record()
        torch.cuda.synchronize()
        t = torch.tensor([s.elapsed_time(e) / a.iters], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    res["single_gpu_ms"] = timeit(lambda: single.infer(feats, camsd, ds, di))
    res["dslab_ms"] = timeit(lambda: slab.infer(feats, camsd, ds, di))
    if a.p2p:
        fused = DSlabHotPath(n, D, hf, wf, single.weights, device=dev, p2p=True)
        d3, p3 = fused.infer(feats, camsd, ds, di)
        torch.cuda.synchronize()
        res["p2p_max_abs_depth_diff_in_intervals"] = float((d1 - d3).abs().max()) / di
        res["p2p_vs_nccl_max_abs_depth_diff_in_intervals"] = float((d2 - d3).abs().max()) / di
        res["dslab_p2p_ms"] = timeit(lambda: fused.infer(feats, camsd, ds, di))
        res["p2p_wait_timeouts"] = int(fused.lib.mvsb200_slab_p2p_error(n, D, world, hf, wf, 32, fused.base_filter,
                                                                        fused.ws.data_ptr(), None))
        if a.abort_test:
            # rank 0 runs the layers of one more inference ALONE: every consumer kernel waits for flags the other ranks
            # never raise.  A host timer aborts after 50 ms; without it each layer would sit out its ~2 s bound.
            import ctypes
            import threading
            import time
            from mvsnet_b200 import _lib as L
            from mvsnet_b200.dslab import SLAB_ORDER
            dist.barrier()
            if rank == 0:
                fused.seq += 1
                args = (n, D, rank, world, hf, wf, 32)
                t0 = time.perf_counter()
                timer = threading.Timer(0.05, fused.abort)
                timer.start()
                L.check(fused.lib.mvsb200_slab_begin(L.ptr(feats), L.ptr(camsd), *args, ds, di, 0, 0,
                                                     ctypes.byref(fused.weights.params), fused.base_filter, L.ptr(fused.ws),
                                                     fused.ws.numel(), L.stream_ptr()), "slab_begin")
                for layer in SLAB_ORDER:
                    L.check(fused.lib.mvsb200_slab_layer_p2p(layer, *args, ctypes.byref(fused.weights.params), fused.base_filter,
                                                             fused.bn_eps, L.ptr(fused.ws), L.ptr(fused.peers_dev),
                                                             fused.peers_host, fused.seq, L.stream_ptr()), "slab_layer_p2p")
                torch.cuda.synchronize()
                timer.join()
                res["abort_released_ms"] = 1e3 * (time.perf_counter() - t0)
                res["abort_reported"] = bool(fused.p2p_error())
            dist.barrier()
        dist.barrier()
        fused.close()
    if rank == 0:
        print(json.dumps(res), flush=True)
    # same gate as the bf16 path against the oracle: depth within 0.1 interval (here: of the single-GPU result)
    ok = res["max_abs_depth_diff_in_intervals"] <= 0.1 and res.get("p2p_max_abs_depth_diff_in_intervals", 0.0) <= 0.1 \
        and res.get("p2p_wait_timeouts", 0) == 0 and res.get("abort_reported", True) \
        and res.get("abort_released_ms", 0.0) < 1500.0
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
