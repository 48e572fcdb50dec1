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
           "--iters", "2", "--p2p"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = next(l for l in out.stdout.splitlines() if l.startswith("{"))
    res = json.loads(line)
    assert res["world"] == 2 and res["max_abs_depth_diff_in_intervals"] <= 0.1
    # the peer-memory variant (exchange inside the kernels) computes exactly what the NCCL-driven one does
    assert res["p2p_vs_nccl_max_abs_depth_diff_in_intervals"] == 0.0 and res["p2p_wait_timeouts"] == 0
