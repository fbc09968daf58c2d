"""Summarise ncu output for profiles/.

  python tools/ncu_summary.py launches <launches.csv> <out.md> [--title "..."]
      launch list of `ncu --metrics gpu__time_duration.sum --csv`: per-kernel launches / total time / share
  python tools/ncu_summary.py full <report.ncu-rep> <out.json> [--workload W]
      `ncu --set full` report: per kernel (aggregated over its launches) duration, DRAM bytes read/written per
      launch, tensor-pipe and issue utilisation, registers, occupancy limits; also updates profiles/ncu_traffic.json
      (DRAM bytes per launch by bench.py kernel class) used for `roofline.traffic`.
"""
import argparse
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"^void ", "", name)
    name = name.replace("amc::<unnamed>::", "").replace("amc::(anonymous namespace)::", "").replace("<unnamed>::", "")
    return name.replace("unnamed>::", "").strip()


# ncu kernel name -> bench.py kernel class (only where the mapping is one-to-one)
CLASS_OF = [
    (r"gemm_tc_kernel<\d+, *1, *\d+>", "gemm_wgrad"),
    (r"ln_bwd_vec_kernel", "ln_bwd"),
    (r"attn_(frames|mma|tile|tc|tc5)_bwd_kernel", "attn_bwd"),
    (r"attn_(frames|mma|tile|tc|tc5)_fwd_kernel", "attn_fwd"),
    (r"attn_long(_mma)?_fwd_kernel", "attn_fwd"),
    (r"patchify", "patchify"),
    (r"adamw_kernel", "adamw_clip"),
]


def launches(path, out, title):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi or "gpu__time_duration" not in ",".join(r):
            continue
        v = float(r[vi].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        a = agg.setdefault(short(r[ki]), [0, 0.0])
        a[0] += 1
        a[1] += v * scale
    tot = sum(a[1] for a in agg.values())
    n = sum(a[0] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# {title}\n# per-launch times are cold-cache and serialised: compare SHARES with bench.py's "
                f"kernel_ms_per_step, not absolutes\n# {n} launches, total {tot:.3f} ms\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {c} | {ms:.3f} | {100 * ms / tot:.1f}% |\n")
    print(f"{out}: {n} launches, {tot:.3f} ms")


KEEP = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active": "tensor_inst_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers",
    "launch__occupancy_limit_registers": "occ_limit_regs",
    "launch__occupancy_limit_shared_mem": "occ_limit_smem",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
}
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}
# newer ncu names the tcgen05 pipe separately
ALT = {"sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "hmma_pipe_active_pct",
       "sm__inst_executed_pipe_uniform.sum": None}


def full(rep, out, workload, frames=0):
    # a report, or the `ncu -i report --page raw --csv` text made from it on the GPU box (reports are too big to bring back)
    txt = open(rep).read() if rep.endswith(".csv") else \
        subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    tens = [i for i, h in enumerate(hdr) if "tensor" in h and h.endswith("pct_of_peak_sustained_active") and "avg" in h]
    agg = collections.OrderedDict()
    for r in rows[2:]:
        k = short(r[ki])
        a = agg.setdefault(k, {"launches": 0})
        a["launches"] += 1
        for i, h in enumerate(hdr):
            key = KEEP.get(h) or ALT.get(h)
            if not key or not r[i] or r[i] == "no data":
                continue
            v = float(r[i].replace(",", "")) * UNIT_SCALE.get(units[i], 1.0)
            a.setdefault(key + "_sum", 0.0)
            a[key + "_sum"] += v
        for i in tens:
            if r[i] and r[i] != "no data":
                a["max_tensor_metric_pct"] = max(a.get("max_tensor_metric_pct", 0.0), float(r[i].replace(",", "")))
    res = collections.OrderedDict()
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1].get("duration_us_sum", 0)):
        n = a["launches"]
        d = {"launches": n}
        for key, v in a.items():
            if key.endswith("_sum"):
                d[key[:-4] + "_per_launch"] = round(v / n, 3)
        if "max_tensor_metric_pct" in a:
            d["max_tensor_metric_pct"] = a["max_tensor_metric_pct"]
        if "dram_read_bytes_per_launch" in d:
            d["dram_bytes_per_launch"] = d["dram_read_bytes_per_launch"] + d.get("dram_write_bytes_per_launch", 0.0)
        res[k] = d
    json.dump({"report": os.path.basename(rep), "workload": workload, "kernels": res}, open(out, "w"), indent=1)
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    tj = json.load(open(tpath)) if os.path.exists(tpath) else {}
    wl = tj.setdefault(workload, {})
    for k, d in res.items():
        for pat, cls in CLASS_OF:
            if re.search(pat, k) and "dram_bytes_per_launch" in d:
                if wl.get(cls, {}).get("report") != os.path.basename(rep):
                    wl.pop(cls, None)           # a new capture replaces the class's entry (no mixing of rounds / batch sizes)
                e = wl.setdefault(cls, {"dram_bytes_per_launch": 0.0, "launches": 0, "kernels": [], "report": os.path.basename(rep),
                                        "frames": frames or None})
                if k in e["kernels"]:
                    continue
                tot = e["dram_bytes_per_launch"] * e["launches"] + d["dram_bytes_per_launch"] * d["launches"]
                e["launches"] += d["launches"]
                e["dram_bytes_per_launch"] = tot / e["launches"]
                e["kernels"].append(k)
    json.dump(tj, open(tpath, "w"), indent=1)
    print(f"{out}: {len(res)} kernels; {tpath} updated")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["launches", "full"])
    ap.add_argument("src")
    ap.add_argument("out")
    ap.add_argument("--title", default="ncu launch list")
    ap.add_argument("--workload", default="vit_p16_d256_L6")
    ap.add_argument("--batch", type=int, default=0, help="frames per launch of the capture (0 = the workload's bench batch)")
    a = ap.parse_args()
    if a.mode == "launches":
        launches(a.src, a.out, a.title)
    else:
        full(a.src, a.out, a.workload, a.batch)
