"""Turn the ncu outputs in gpurun_out/ into the per-round summaries under profiles/ (tracked)."""
import csv
import json
import os
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
if tag.startswith("-"):
    sys.exit("usage: summarize_profiles.py [tag]   (reads gpurun_out/, writes profiles/<tag>_*)")


def read_csv(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r]
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    return rows[start], rows[start + 1:]


# ---- launch list (gpu__time_duration per launch) -> per-kernel totals and shares ---------------------------
hdr, rows = read_csv(os.path.join(G, "launches.csv"))
ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
iu = hdr.index("Metric Unit")
per = OrderedDict()
total = 0.0
with open(os.path.join(P, f"{tag}_launches_raw.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["id", "kernel", "duration_us"])
    for r in rows:
        if r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        us = v / 1e3 if r[iu] in ("ns", "nsecond") else (v if r[iu] in ("us", "usecond") else v * 1e3)
        name = r[ik].split("(")[0].replace("void ", "").replace("mvsb200::", "")
        w.writerow([r[0], name, f"{us:.3f}"])
        d = per.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += us
        total += us
with open(os.path.join(P, f"{tag}_launches_by_kernel.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "launches", "total_us", "share_of_all_launches"])
    for k, (n, us) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        w.writerow([k, n, f"{us:.1f}", f"{us / total:.4f}"])

# ---- cost volume: full capture -> key counters + dram traffic for bench.py's roofline.traffic --------------
def raw_metrics(path):
    rows = list(csv.reader(open(path, errors="replace")))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    return rows[start], rows[start + 1], rows[start + 2:]


def tensor_keys(hdr):
    """Every tensor-pipe / TMEM counter the capture holds (the hmma sub-pipe counter does not see tcgen05's UTCHMMA)."""
    return [k for k in hdr if any(t in k for t in ("pipe_tensor", "tmem", "utc", "tcgen")) and ("pct_of_peak" in k or k.endswith(".sum"))]


KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__inst_executed.sum",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max"]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def summarise(src, dst, names=None):
    hdr, units, rows = raw_metrics(src)
    out = []
    for i, r in enumerate(rows):
        rec = OrderedDict(kernel=r[hdr.index("Kernel Name")].split("(")[0])
        if names:
            rec["layer"] = names[i] if i < len(names) else ""
        for k in KEYS + [t for t in tensor_keys(hdr) if t not in KEYS]:
            if k in hdr and r[hdr.index(k)] not in ("", "n/a"):
                rec[k] = r[hdr.index(k)] + " " + units[hdr.index(k)]
        if "dram__bytes_read.sum" in hdr:
            rec["dram_bytes"] = to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")]) + \
                to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
        out.append(rec)
    json.dump(out, open(dst, "w"), indent=1)
    return out


cv = summarise(os.path.join(G, "prof_cv_raw.csv"), os.path.join(P, f"{tag}_cost_volume_ncu.json"))
if "cost_volume_window_kernel" not in cv[0]["kernel"]:
    sys.exit(f"captured kernel {cv[0]['kernel']!r} is not the one bench.py reports the roofline of")
json.dump({"kernel": cv[0]["kernel"], "config": "cfg2", "dram_bytes_per_launch": cv[0]["dram_bytes"],
           "source": f"profiles/{tag}_cost_volume_ncu.json (ncu --set full, one launch at cfg2: dram__bytes_read.sum + "
                     "dram__bytes_write.sum)"},
          open(os.path.join(P, "cost_volume_traffic.json"), "w"), indent=1)
layers = ["3dconv0_1 + 3dconv1_0 (one launch)", "3dconv2_0", "3dconv3_0[0:32]", "3dconv3_0[32:64]", "3dconv1_1", "3dconv2_1",
          "3dconv3_1[0:32]", "3dconv3_1[32:64]", "3dconv4_0", "3dconv5_0", "3dconv6_0", "3dconv6_2"]
if os.path.exists(os.path.join(G, "prof_conv_raw.csv")):
    summarise(os.path.join(G, "prof_conv_raw.csv"), os.path.join(P, f"{tag}_conv3d_ncu.json"), layers)
if os.path.exists(os.path.join(G, "prof_tower_raw.csv")):
    summarise(os.path.join(G, "prof_tower_raw.csv"), os.path.join(P, f"{tag}_tower_layer_ncu.json"), ["a full-resolution 8 -> 8 layer"])
if os.path.exists(os.path.join(G, "tower_bench.json")):
    open(os.path.join(P, f"{tag}_tower_bench.json"), "w").write(open(os.path.join(G, "tower_bench.json")).read())
for f in ("prof_cv_details.txt", "prof_conv_details.txt", "prof_tower_details.txt"):
    if os.path.exists(os.path.join(G, f)):
        open(os.path.join(P, f"{tag}_{f}"), "w").write(open(os.path.join(G, f), errors="replace").read())
print("wrote summaries under profiles/")
