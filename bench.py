#!/usr/bin/env python
"""bench.py -- depth maps/s of the MVSNet cost-volume hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2]
    python bench.py --config cfg3 [--gpus N]                 BASELINE config 3: 49 reference views sharded over the ranks
    python bench.py --config cfg4 [--gpus N]                 BASELINE config 4: training step (forward + backward)
    python bench.py --config cfg5 --mode dslab [--gpus N]    BASELINE config 5: ONE volume, depth slabs over the ranks

A step = one pass of the hot path over one reference view (cluster): feats [5,216,288,32] + cams
-> depth map + probability map at 1152x864, D=192, N=5 (BASELINE.json configs[1]), bf16 regularizer.
Ranks (one process per GPU, torchrun) hold full replicas and process different clusters: no
data-path collective ("weak" scaling).  Prints ONE JSON line on rank 0.

  value     depth maps/s, inputs resident in HBM, device-timed (CUDA events, max over ranks)
  e2e       the same through the host-buffer C-ABI call (pinned host -> device copy of feats+cams and
            device -> host read of depth+prob inside the timed region; two views in flight on two streams)
  roofline  the dominant kernel (fused warp+variance) against the measured HBM copy bandwidth
  cpu_baseline / --impl reference   the CPU restatement of the reference (oracle/) on the host cores
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "depth_maps_per_s"
UNIT = "depth maps/s"
N_CLUSTERS = 4          # distinct input sets rotated through the timed loop (4 x 40 MB > nothing reused from L2)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def tensor_peak():
    """Dense bf16 tensor throughput for a kernel timed inside a long step: the sustained cuBLAS figure the driver measured."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        if "bf16_tflops_sustained" in d:
            return float(d["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    return 2250.0, "nominal (2.25 PFLOP/s dense bf16)"


class ClockSampler:
    """SM clock and throttle reasons sampled WHILE the timed region runs: NVML polled from a thread every few
    milliseconds (the timed region of 20 steps is ~60 ms, too short for `nvidia-smi -lms`), nvidia-smi as a fallback."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4), ("hw_power_brake_slowdown", 0x80))

    def __init__(self, index: int):
        self.index, self.rows, self.thread, self.stop_flag, self.nvml, self.max_mhz = index, [], None, False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.index]) if visible and visible.split(",")[self.index].isdigit() else self.index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None
        self.stop_flag = False
        self.thread = threading.Thread(target=self._poll if self.nvml else self._poll_smi, daemon=True)
        self.thread.start()

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                bits = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)) if hasattr(
                    n, "nvmlDeviceGetCurrentClocksEventReasons") else int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.rows.append((mhz, bits))
            except Exception:
                pass
            time.sleep(0.003)

    def _poll_smi(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.max_mhz = float(out[1])
                self.rows.append((float(out[0]), int(out[2].strip(), 16)))
            except Exception:
                time.sleep(0.05)

    def stop(self):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no clock samples"], "samples": 0}
        sm = [r[0] for r in self.rows]
        bits = 0
        for r in self.rows:
            bits |= r[1]
        reasons = [name for name, mask in self.REASONS if bits & mask]
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU reference arm: the oracle restatement on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_step(problem, sample_planes, threads):
    """One pass of the oracle over the first `sample_planes` depth planes of the workload; returns seconds."""
    from concurrent.futures import ThreadPoolExecutor

    import torch

    import oracle as O
    torch.set_num_threads(threads)
    feats, cams = problem["feats"], problem["cams"]
    ds, di = problem["depth_start"], problem["depth_interval"]
    n = feats.shape[0]
    t0 = time.perf_counter()
    H = np.stack([O.get_homographies(cams[0:1], cams[v:v + 1], sample_planes, ds, di)[0] for v in range(1, n)])

    def plane(d):
        return O.cost_volume(feats, H[:, d:d + 1])[0]

    with ThreadPoolExecutor(max_workers=threads) as ex:      # numpy releases the GIL in the heavy ops
        cost = np.stack(list(ex.map(plane, range(sample_planes))))
    filtered = O.regnet_us0(cost, problem["weights"])
    O.depth_regress(filtered, ds, di)
    return time.perf_counter() - t0


def cpu_sample_planes(problem, D, passes, budget_s, threads, forced=0):
    """Depth planes per CPU step: the whole sweep when `passes` of it fit the time budget (an 8-plane probe estimates
    the cost per plane), else the largest multiple of 8 that does (RegNetUS0 halves the depth three times)."""
    if forced:
        return min(D, max(8, (forced + 7) // 8 * 8)), None
    t8 = cpu_reference_step(problem, 8, threads)
    t8 = min(t8, cpu_reference_step(problem, 8, threads))
    per_plane = t8 / 8.0
    if per_plane * D * passes <= budget_s:
        return D, per_plane
    return max(8, int(budget_s / (per_plane * passes)) // 8 * 8), per_plane


def sample_text(sample, D, threads):
    what = (f"all {D} depth planes of the workload per step" if sample == D else
            f"{sample} of {D} depth planes of the same workload per step (value scaled by {sample}/{D}; per-voxel work "
            "is plane-independent)")
    return (f"{what}; CPU restatement of the reference (oracle/): numpy warp + variance over a {threads}-thread pool, "
            "torch-CPU fp32 conv3d / conv_transpose3d with all host threads (TensorFlow 1.12 cannot be installed here)")


def core_config(name, cfg):
    """The part of `config` both arms print identically."""
    hf, wf = cfg["height"] // 4, cfg["width"] // 4
    return {"workload": workload_name(name, cfg), "voxels_per_map": cfg["depth_num"] * hf * wf}


def run_reference(args, rank, world):
    if rank != 0:
        return
    from mvsnet_b200 import synthetic
    cfg_name = "cfg2" if args.config == "cfg3" else args.config
    cfg = synthetic.CONFIGS[cfg_name]
    D = cfg["depth_num"]
    threads = os.cpu_count() or 1
    problem = synthetic.make_problem(cfg_name)
    sample, _ = cpu_sample_planes(problem, D, args.steps + args.warmup, args.cpu_budget, threads, args.cpu_planes)
    for _ in range(args.warmup):
        cpu_reference_step(problem, sample, threads)
    ts = [cpu_reference_step(problem, sample, threads) for _ in range(args.steps)]
    frac = sample / D
    value = frac / float(np.mean(ts))          # whole depth maps per second
    V = D * (cfg["height"] // 4) * (cfg["width"] // 4)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(ts)) / frac, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": core_config(cfg_name, cfg),
        "detail": {"gvox_per_s": value * V / 1e9, "planes_per_step": sample, "measured_ms_per_step": 1e3 * float(np.mean(ts))},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample_text(sample, D, threads)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_name(name, cfg):
    return (f"{name}: MVSNet 3DCNN inference hot path, {cfg['n_views']} views {cfg['width']}x{cfg['height']}, "
            f"D={cfg['depth_num']}, interval_scale {cfg['interval_scale']} (feature maps "
            f"{cfg['width'] // 4}x{cfg['height'] // 4}x32 in, depth+prob map out)")


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def make_clusters(synthetic, cfg, count, first_seed, dev):
    """`count` distinct reference views (cluster = cams + feature maps), pinned on the host and resident on the device."""
    import torch
    n, D = cfg["n_views"], cfg["depth_num"]
    hf, wf = cfg["height"] // 4, cfg["width"] // 4
    feats_h, cams_h, feats_d, cams_d = [], [], [], []
    for c in range(count):
        cams = synthetic.make_cameras(n, cfg["height"], cfg["width"], D, cfg["interval_scale"], seed=1234 + first_seed + c)
        feats = synthetic.make_features(cams, hf, wf, 32, seed=5678 + first_seed + c)
        feats_h.append(torch.from_numpy(feats).pin_memory())
        cams_h.append(torch.from_numpy(cams).pin_memory())
        feats_d.append(feats_h[-1].to(dev))
        cams_d.append(cams_h[-1].to(dev))
    return feats_h, cams_h, feats_d, cams_d


def stage_times(eng, feats, cams, ds, di, iters):
    """Median device time (ms) of the four stages of mvsb200_infer over `iters` calls (stage-boundary events)."""
    import torch
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(iters)]
    for e in evs:
        eng.set_stage_events(e)
        eng.infer(feats, cams, ds, di)
    torch.cuda.synchronize()
    eng.set_stage_events(None)
    return np.median(np.array([[e[j].elapsed_time(e[j + 1]) for j in range(4)] for e in evs]), axis=0)


def cost_volume_variants(eng, feats, cams, ds, di):
    """The cost-volume stage under its arithmetic variants, side by side (device ms, config of this run): the shipped
    kernel reads fp16-rounded features, blends the difference to the reference pixel in packed fp16 and sums in fp32;
    north_star's arithmetic is fp32 features.  (The variants are timed inside the whole pass: with the gather kernel the
    two first regularizer layers still run as one launch.)"""
    from mvsnet_b200 import _lib
    variants = (("window_fp16_taps_fp16_blend_fp32_sums (shipped)", {}),
                ("window_fp16_taps_fp16_blend_fp16_sums", {"CV_FP32_BLEND": 2}),
                ("window_fp16_taps_fp32_blend", {"CV_FP32_BLEND": 1}),
                ("gather_fp16_taps_fp16_blend (round 1)", {"CV_KERNEL": 1}),
                ("gather_fp32_taps_fp32_blend (north_star arithmetic)", {"CV_KERNEL": 1, "CV_FP32_TAPS": 1}))
    out = {}
    for name, sw in variants:
        for k, v in sw.items():
            _lib.set_tuning(k, v)
        try:
            eng.infer(feats, cams, ds, di)
            out[name] = float(stage_times(eng, feats, cams, ds, di, 5)[1])
        finally:
            for k in sw:
                _lib.set_tuning(k, None)
    return out


def images_in(eng, synthetic, cfg, cams, ds, di, dev, iters=10):
    """The step BEFORE the path in front of it: images [N,H,W,3] resident in HBM -> UNetDS2GN towers ->
    hot path -> depth map, device-timed, with the tensor-core tower (bf16) and the CUDA-core parity tower (fp32).  Not the
    headline metric (BASELINE.json defines it from feature maps); random tower weights."""
    import torch
    from mvsnet_b200.features import FeatureTower
    n, h, w = cfg["n_views"], cfg["height"], cfg["width"]
    images = torch.randn((n, h, w, 3), device=dev)
    wts = synthetic.make_unet_weights(8)
    out = {}
    for prec in ("bf16", "fp32"):
        tower = FeatureTower(wts, precision=prec, device=dev)
        for _ in range(3):
            eng.infer(tower(images), cams, ds, di)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        t_tower, t_all = [], []
        for _ in range(iters):
            ev[0].record()
            f = tower(images)
            ev[1].record()
            eng.infer(f, cams, ds, di)
            ev[2].record()
            torch.cuda.synchronize()
            t_tower.append(ev[0].elapsed_time(ev[1]))
            t_all.append(ev[0].elapsed_time(ev[2]))
        out[f"tower_{prec}_ms"] = float(np.median(t_tower))
        out[f"depth_maps_per_s_{prec}_tower"] = 1e3 / float(np.median(t_all))
        del tower
    return out


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from mvsnet_b200 import ops, synthetic
    from mvsnet_b200.engine import HotPath

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = synthetic.CONFIGS[args.config]
    n, D = cfg["n_views"], cfg["depth_num"]
    hf, wf = cfg["height"] // 4, cfg["width"] // 4
    V = D * hf * wf
    weights = synthetic.make_regnet_weights()
    eng = HotPath(n, D, hf, wf, weights, precision="bf16", device=dev)
    # distinct clusters per rank (reference views are independent problems, inference.py:105-119)
    feats_h, cams_h, feats_d, cams_d = make_clusters(synthetic, cfg, N_CLUSTERS, rank * N_CLUSTERS, dev)
    ds, di = float(cams_h[0][0, 1, 3, 0]), float(cams_h[0][0, 1, 3, 1])
    depth_h = torch.empty((hf, wf)).pin_memory()
    prob_h = torch.empty((hf, wf)).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput -------------------------------------------------------------
    for i in range(args.warmup):
        eng.infer(feats_d[i % N_CLUSTERS], cams_d[i % N_CLUSTERS], ds, di)
    stage_events = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(args.steps)]
    for evs in stage_events:
        for e in evs:
            e.record()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = ops.launch_count()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for i in range(args.steps):
        eng.set_stage_events(stage_events[i])
        eng.infer(feats_d[i % N_CLUSTERS], cams_d[i % N_CLUSTERS], ds, di)
    t_end.record()
    barrier()
    clocks = sampler.stop()
    eng.set_stage_events(None)
    launches = ops.launch_count() - launches0
    ms = t_start.elapsed_time(t_end)
    stage_ms = np.array([[evs[j].elapsed_time(evs[j + 1]) for j in range(4)] for evs in stage_events]).mean(axis=0)

    # ---- end to end through the host-buffer C-ABI call ---------------------------------------------
    # Two reference views in flight (own staging + workspace + copy stream each) over ONE compute stream
    # (mvsb200_infer_host_pipelined): the pinned-host -> device feed of the next view and the fetch of the previous
    # one run beside the kernels of the current view, as a prefetching input pipeline would feed sess.run.  Every
    # step still copies its own inputs in and its own depth + probability maps out inside the timed region.
    engs = [eng, HotPath(n, D, hf, wf, eng.weights, precision="bf16", device=dev)]
    compute_stream = torch.cuda.Stream(device=dev)
    copy_streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
    outs = [(depth_h, prob_h), (torch.empty((hf, wf)).pin_memory(), torch.empty((hf, wf)).pin_memory())]

    def e2e_pass(count):
        for i in range(count):
            k = i % 2
            engs[k].infer_host_pipelined(feats_h[i % N_CLUSTERS], cams_h[i % N_CLUSTERS], ds, di, outs[k][0], outs[k][1],
                                         compute_stream, copy_streams[k])

    e2e_pass(min(args.warmup, 4))
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)]
    e0.record()
    for st in (compute_stream, *copy_streams):
        st.wait_event(e0)
    e2e_pass(args.steps)
    e1[0].record(copy_streams[0])
    e1[1].record(copy_streams[1])
    barrier()
    ms_e2e = max(e0.elapsed_time(e1[0]), e0.elapsed_time(e1[1]))
    checksum = float(outs[(args.steps - 1) % 2][0].sum())

    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    maps = args.steps * world
    value = maps / (ms / 1e3)
    e2e_value = maps / (ms_e2e / 1e3)
    peak, peak_src = measured_peaks()
    cv_bytes = n * hf * wf * 32 * 4 + V * 32 * 2            # feature reads once + bf16 volume write
    cv_ms = float(stage_ms[1])
    achieved = cv_bytes / (cv_ms * 1e-3) / 1e9
    # DRAM traffic of the dominant kernel: from the ncu capture of THIS kernel (tools/profile_round.sh regenerates the
    # file and names the kernel it measured); a capture of another kernel is not reported
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "cost_volume_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if "cost_volume_window_kernel" in tj.get("kernel", "") and tj.get("config", "cfg2") == args.config:
            traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
    rg_bytes = V * 4 + 2 * hf * wf * 4                       # filtered volume once + two maps
    rg_ms = float(stage_ms[3])
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": core_config(args.config, cfg),
        "detail": {"gvox_per_s": value * V / 1e9, "l2": f"inputs rotate over {N_CLUSTERS} clusters; every step streams "
                   ">2.5 GB of intermediates through HBM (L2 is 126 MB)", "parallelism": f"view-sharded x{world}",
                   "stage_ms": {"homographies": float(stage_ms[0]), "cost_volume": cv_ms,
                                "regularizer": float(stage_ms[2]), "regression": rg_ms},
                   "arithmetic": "source views and reference view read as fp16, 4-tap blend of (warped - reference) in "
                                 "packed fp16, sums of the differences and their squares and the variance in fp32, bf16 "
                                 "volume (chunk-planar copy only); regularizer bf16 operands / fp32 accumulation (tcgen05), "
                                 "3dconv0_1 + 3dconv1_0 in one launch; parity of exactly this mode vs the oracle: "
                                 "tests/test_gpu_parity_product.py"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": eng.h2d_bytes,
                "d2h_bytes_per_step": eng.d2h_bytes, "ms_per_step": ms_e2e / args.steps, "depth_checksum": checksum},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"kernel": "cost_volume_window_kernel (fused warp + variance, TMA-staged source windows)",
                     "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes": cv_bytes, "avg_launch_ms": cv_ms},
        # the second stage as a whole (12 launches of conv3d_tc_kernel): SURVEY 8(d) unfused compulsory bytes 230 B
        # per voxel and 22 896 FLOP per voxel, against the same measured HBM peak
        "roofline_regularizer": {"kernel": "conv3d_tc_kernel x12 (RegNetUS0, bf16 tcgen05)", "bound": "hbm",
                                 "achieved": 230.0 * V / (float(stage_ms[2]) * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                 "frac": 230.0 * V / (float(stage_ms[2]) * 1e-3) / 1e9 / peak,
                                 "algorithmic_bytes": 230 * V, "tflops": 22896.0 * V / (float(stage_ms[2]) * 1e-3) / 1e12,
                                 "tensor_peak_tflops": tensor_peak()[0], "tensor_peak_source": tensor_peak()[1],
                                 "tensor_frac": 22896.0 * V / (float(stage_ms[2]) * 1e-3) / 1e12 / tensor_peak()[0],
                                 "tensor_pipe_active_per_launch": "profiles/r2d_conv3d_ncu.json (ncu sm__pipe_tensor_cycles_active: "
                                                                  "41.3 % in the 3dconv0_1 + 3dconv1_0 launch, 3-19 % elsewhere)",
                                 "stage_ms": float(stage_ms[2])},
        # K4: in bf16 mode the soft-argmin runs inside 3dconv6_2's epilogue (the filtered volume is not re-read for it);
        # what is timed here is regress_combine_kernel (three partial maps in, four probability gathers, two maps out).
        # Against SURVEY 8(d)'s stand-alone figure (read the volume once) the stage therefore exceeds the roofline.
        "roofline_regression": {"kernel": "regress_combine_kernel (+ soft-argmin fused into 3dconv6_2)", "bound": "hbm",
                                "achieved": rg_bytes / (rg_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                "frac": rg_bytes / (rg_ms * 1e-3) / 1e9 / peak, "algorithmic_bytes": rg_bytes,
                                "stage_ms": rg_ms},
    }
    if world == 1:
        line["detail"]["cost_volume_variants_ms"] = cost_volume_variants(eng, feats_d[0], cams_d[0], ds, di)
        line["detail"]["from_images"] = images_in(eng, synthetic, cfg, cams_d[0], ds, di, dev)
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        problem = synthetic.make_problem(args.config)
        sample, _ = cpu_sample_planes(problem, D, 1, 30.0, threads, args.cpu_planes)
        t = cpu_reference_step(problem, sample, threads)
        line["cpu_baseline"] = {"value": (sample / D) / t, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "one pass over " + sample_text(sample, D, threads)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# BASELINE config 3: a DTU-scan-sized batch, 49 reference views x 5 views, sharded over the ranks
# ------------------------------------------------------------------------------------------------
def run_cfg3(args, rank, world, local_rank):
    """Every rank takes its round-robin share of the 49 reference views (mvsnet_b200.sharding.shard_views, the loop of
    inference.py:105-119 split over GPUs), runs them through the host-buffer entry point (feed + kernels + fetch per
    view) and the maps are gathered on rank 0 (sharding.gather_maps).  value = 49 / (time of the slowest rank +
    gather); 49 views over 8 ranks is 7 / 6 views per rank: the partition alone caps 8-GPU efficiency at 87.5 %."""
    import torch
    import torch.distributed as dist

    from mvsnet_b200 import ops, sharding, synthetic
    from mvsnet_b200.engine import HotPath

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = synthetic.CONFIGS["cfg2"]
    n, D = cfg["n_views"], cfg["depth_num"]
    hf, wf = cfg["height"] // 4, cfg["width"] // 4
    n_ref = 49
    mine = sharding.shard_views(n_ref, rank, world)
    weights = synthetic.make_regnet_weights()
    engs = [HotPath(n, D, hf, wf, weights, precision="bf16", device=dev)]
    engs.append(HotPath(n, D, hf, wf, engs[0].weights, precision="bf16", device=dev))
    distinct = min(len(mine), N_CLUSTERS)
    feats_h, cams_h, _, _ = make_clusters(synthetic, cfg, distinct, rank * N_CLUSTERS, dev)
    ds, di = float(cams_h[0][0, 1, 3, 0]), float(cams_h[0][0, 1, 3, 1])
    depth_out = [torch.empty((hf, wf)).pin_memory() for _ in mine]
    prob_out = [torch.empty((hf, wf)).pin_memory() for _ in mine]
    compute_stream = torch.cuda.Stream(device=dev)
    copy_streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]

    def one_pass():
        for i in range(len(mine)):
            k = i % 2
            engs[k].infer_host_pipelined(feats_h[i % distinct], cams_h[i % distinct], ds, di, depth_out[i], prob_out[i],
                                         compute_stream, copy_streams[k])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(1, min(args.warmup, 2))):
        one_pass()
    if world > 1:
        # the gather's first call builds NCCL's channels (> 1 s): part of the warm-up, like the first kernel launches
        torch.cuda.synchronize()
        sharding.gather_maps(mine, [d.to(dev) for d in depth_out], n_ref)
    barrier()
    launches0 = ops.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t0 = time.perf_counter()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)]
    e0.record()
    for st in (compute_stream, *copy_streams):
        st.wait_event(e0)
    one_pass()
    e1[0].record(copy_streams[0])
    e1[1].record(copy_streams[1])
    torch.cuda.synchronize()
    ms_dev = max(e0.elapsed_time(e1[0]), e0.elapsed_time(e1[1]))
    gathered = None
    if world > 1:
        gathered = sharding.gather_maps(mine, [d.to(dev) for d in depth_out], n_ref)
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    clocks = sampler.stop()
    launches = ops.launch_count() - launches0
    if world > 1:
        t = torch.tensor([ms_dev, wall_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, wall_ms = float(t[0]), float(t[1])
    if rank == 0:
        if gathered is not None:
            assert len(gathered) == n_ref and all(g is not None for g in gathered)
        V = D * hf * wf
        value = n_ref / (wall_ms / 1e3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": n_ref, "warmup": max(1, min(args.warmup, 2)),
            "ms_per_step": wall_ms / n_ref, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "cfg3: DTU-scan-sized batch, 49 reference views x 5 views 1152x864, D=192, sharded "
                                   f"round-robin over {world} GPU(s), maps gathered on rank 0", "voxels_per_map": V},
            "detail": {"gvox_per_s": value * V / 1e9, "views_per_rank": [len(sharding.shard_views(n_ref, r, world)) for r in range(world)],
                       "partition_ceiling": n_ref / (world * max(len(sharding.shard_views(n_ref, r, world)) for r in range(world))),
                       "device_ms_slowest_rank": ms_dev, "wall_ms_incl_gather": wall_ms,
                       "timing": "wall clock of the whole batch on the slowest rank (host-buffer entry point per view: feed, "
                                 "kernels, fetch; then the gather), after a barrier"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": engs[0].h2d_bytes, "d2h_bytes_per_step": engs[0].d2h_bytes},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# BASELINE config 5: ONE R-MVSNet-scale volume, depth slabs over the ranks (NVLink halo exchange)
# ------------------------------------------------------------------------------------------------
def run_dslab(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from mvsnet_b200 import ops, synthetic
    from mvsnet_b200.dslab import DSlabHotPath
    from mvsnet_b200.engine import HotPath

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = synthetic.CONFIGS[args.config]
    n, D = cfg["n_views"], cfg["depth_num"]
    hf, wf = cfg["height"] // 4, cfg["width"] // 4
    V = D * hf * wf
    cams = synthetic.make_cameras(n, cfg["height"], cfg["width"], D, cfg["interval_scale"])
    feats = torch.from_numpy(synthetic.make_features(cams, hf, wf, 32)).to(dev)
    camsd = torch.from_numpy(cams).to(dev)
    ds, di = float(cams[0, 1, 3, 0]), float(cams[0, 1, 3, 1])
    weights = synthetic.make_regnet_weights()
    if world == 1:
        eng = HotPath(n, D, hf, wf, weights, precision="bf16", device=dev)
        step = lambda: eng.infer(feats, camsd, ds, di)
        mode = "one GPU (no slabs)"
    else:
        eng = DSlabHotPath(n, D, hf, wf, weights, device=dev, p2p=not args.no_p2p)
        step = lambda: eng.infer(feats, camsd, ds, di)
        mode = f"{world} depth slabs of {D // world} planes; " + ("halo planes + statistics exchanged inside the kernels "
               "over NVLink peer memory" if not args.no_p2p else "halo planes + statistics exchanged by NCCL between layers")
    for _ in range(max(args.warmup, 3)):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        d, pm = step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    launches = ops.launch_count() - launches0
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    if rank == 0:
        value = args.steps / (ms / 1e3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(args.config, cfg) + " -- ONE volume split into depth slabs over the GPUs",
                       "voxels_per_map": V},
            "detail": {"gvox_per_s": value * V / 1e9, "mode": mode, "depth_checksum": float(d.sum()),
                       "l2": "every step streams > 6 GB of intermediates through HBM (L2 is 126 MB)"},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        if isinstance(eng, DSlabHotPath) and eng.p2p:
            dist.barrier()
            eng.close()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# BASELINE config 4: training step (forward + backward) of the path, 3 views 640x512, D=128
# ------------------------------------------------------------------------------------------------
def run_train(args, rank, world, local_rank):
    """A step = forward in the fp32 parity mode + mvsnet_regression_loss + the gradients of every RegNetUS0 variable and
    of the feature maps (mvsb200_train_step; train.py:314-315,429).  Ranks hold replicas and take different clusters
    (data parallel without a gradient exchange: the optimizer and its all-reduce are outside the path)."""
    import torch
    import torch.distributed as dist

    from mvsnet_b200 import ops, synthetic
    from mvsnet_b200.train import TrainStep

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = synthetic.CONFIGS["cfg1"]                    # config 4 has config 1's shape
    n, D = cfg["n_views"], cfg["depth_num"]
    hf, wf = cfg["height"] // 4, cfg["width"] // 4
    V = D * hf * wf
    ts = TrainStep(n, D, hf, wf, synthetic.make_regnet_weights(), order="train", device=dev)
    _, cams_h, feats_d, cams_d = make_clusters(synthetic, cfg, 2, rank * 2, dev)
    ds, di = float(cams_h[0][0, 1, 3, 0]), float(cams_h[0][0, 1, 3, 1])
    rng = np.random.RandomState(11 + rank)
    gt = (ds + di * rng.uniform(2, D - 3, size=(hf, wf))).astype(np.float32)
    gt[rng.rand(hf, wf) < 0.2] = 0.0
    gt_d = torch.from_numpy(gt).to(dev)
    for i in range(args.warmup):
        ts.step(feats_d[i % 2], cams_d[i % 2], gt_d, ds, di)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        out = ts.step(feats_d[i % 2], cams_d[i % 2], gt_d, ds, di)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    launches = ops.launch_count() - launches0
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    if rank == 0:
        value = args.steps * world / (ms / 1e3)
        loss = float(out["metrics"][0])
        line = {
            "metric": "train_steps_per_s", "value": value, "unit": "training steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg4: MVSNet training step fwd+bwd of the 3DCNN path, 3 views 640x512, D=128 (feature "
                                   "maps 160x128x32 in; loss + gradients of RegNetUS0 and of the feature maps out)",
                       "voxels_per_map": V},
            "detail": {"loss": loss, "arithmetic": "fp32 on CUDA cores (parity mode); bilinear-warp backward = exact adjoint scatter",
                       "gvox_per_s": value * V / 1e9},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", help="cfg1 | cfg2 (BASELINE metric config) | cfg3 (49-view batch) | cfg4 (training step) | cfg5")
    ap.add_argument("--mode", default="views", choices=["views", "dslab"],
                    help="views: every rank whole reference views; dslab: ONE volume, depth slabs over the ranks (cfg5)")
    ap.add_argument("--no-p2p", action="store_true", help="dslab: exchange by NCCL between layers instead of peer memory")
    ap.add_argument("--cpu-planes", type=int, default=0,
                    help="depth planes per CPU-reference step (0: the whole sweep if it fits --cpu-budget, else as many as fit)")
    ap.add_argument("--cpu-budget", type=float, default=240.0, help="seconds the CPU-reference arm may take in total")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1 and args.impl == "ours":
        # convenience: re-launch under torchrun so that `python bench.py --gpus N` works on its own
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"),
               os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.config == "cfg3":
        run_cfg3(args, rank, world, local_rank)
    elif args.config == "cfg4":
        run_train(args, rank, world, local_rank)
    elif args.mode == "dslab":
        run_dslab(args, rank, world, local_rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
