"""D-slab mode on real GPUs (needs >= 2 visible devices, skipped otherwise): the depth map of a volume split over two
ranks (NCCL halo exchange + statistics all-reduce between layers) against the single-GPU bf16 path."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="D-slab mode needs two GPUs")
def test_two_slabs_match_one_gpu():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tools", "dslab_check.py"), "--config", "small",
           "--iters", "2", "--p2p", "--abort-test"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = next(l for l in out.stdout.splitlines() if l.startswith("{"))
    res = json.loads(line)
    assert res["world"] == 2 and res["max_abs_depth_diff_in_intervals"] <= 0.1
    # the peer-memory variant (exchange inside the kernels) computes exactly what the NCCL-driven one does
    assert res["p2p_vs_nccl_max_abs_depth_diff_in_intervals"] == 0.0 and res["p2p_wait_timeouts"] == 0
    # a rank that waits for peers that never publish is released by abort() (not after 11 x 2 s) and says so
    assert res["abort_reported"] and res["abort_released_ms"] < 1500.0


@pytest.mark.parametrize("slabs", [2, 4])
def test_slabs_on_one_gpu_vs_oracle(small_problem, slabs):
    """D-slab kernels with every slab on this GPU (LocalSlabHotPath: the library calls of the multi-GPU mode, the exchange
    done by device copies): depth within 0.1 interval of the ORACLE on >= 95 % of pixels (the bf16 gate), and equal to
    the one-volume bf16 path up to the summation order of the batch statistics."""
    import numpy as np
    import oracle as O
    from conftest import to_dev
    from mvsnet_b200.dslab import LocalSlabHotPath
    from mvsnet_b200.engine import HotPath
    p = small_problem
    rd, rp = O.inference_from_features(p["feats"], p["cams"], p["depth_num"], p["depth_start"], p["depth_interval"],
                                       p["weights"])
    feats, cams = to_dev(p["feats"]), to_dev(p["cams"])
    eng = LocalSlabHotPath(p["n_views"], p["depth_num"], p["hf"], p["wf"], p["weights"], slabs)
    d, pm = eng.infer(feats, cams, p["depth_start"], p["depth_interval"])
    frac = float(np.mean(np.abs(d.cpu().numpy() - rd) <= 0.1 * p["depth_interval"]))
    print(f"{slabs} slabs on one GPU: {100 * frac:.2f}% of pixels within 0.1 interval of the oracle")
    assert frac >= 0.95, frac
    one = HotPath(p["n_views"], p["depth_num"], p["hf"], p["wf"], p["weights"], precision="bf16")
    d1, p1 = one.infer(feats, cams, p["depth_start"], p["depth_interval"])
    assert float((d - d1).abs().max()) <= 0.1 * p["depth_interval"]
    assert float(((pm - p1).abs() <= 0.02).float().mean()) >= 0.99
    # the regression of the slabs against the oracle's on the same (assembled) filtered volume
    F = eng.filtered_volume().cpu().numpy()
    od, op_, _ = O.depth_regress(F, p["depth_start"], p["depth_interval"])
    assert np.abs(d.cpu().numpy() - od).max() <= 2e-3 * p["depth_interval"]
    assert float(np.mean(np.abs(pm.cpu().numpy() - op_) <= 1e-4)) >= 0.99
