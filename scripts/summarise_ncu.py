#!/usr/bin/env python
"""Turns gpurun_out/launches_<tag>.csv (ncu --metrics gpu__time_duration.sum) and
gpurun_out/prof_*_<tag>.ncu-rep (ncu --set full) into small text summaries under profiles/.

usage: python scripts/summarise_ncu.py <tag> [<rep-stem> ...]
"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def launches(tag):
    path = os.path.join(OUT, f"launches_{tag}.csv")
    if not os.path.exists(path):
        return None
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
        k = row["Kernel Name"].split("(")[0]
        a = agg.setdefault(k, [0, 0.0, 1e30, 0.0])
        a[0] += 1; a[1] += v; a[2] = min(a[2], v); a[3] = max(a[3], v)
    tot = sum(a[1] for a in agg.values())
    out = [f"# ncu launch list, tag {tag}: `ncu --metrics gpu__time_duration.sum --clock-control none` "
           "(cold-cache, serialised: compare shares)", "",
           "| kernel | launches | total ms | avg ms | min ms | max ms | share |", "|---|---:|---:|---:|---:|---:|---:|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {a[0]} | {a[1]:.2f} | {a[1] / a[0]:.4f} | {a[2]:.4f} | {a[3]:.4f} | {a[1] / tot:.3f} |")
    return "\n".join(out) + "\n"


def full(stem):
    rep = os.path.join(OUT, stem + ".ncu-rep")
    if not os.path.exists(rep):
        return None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = [f"# ncu --set full capture `{stem}.ncu-rep` (per captured launch)", ""]
    for r in rows[2:]:
        out.append(f"## {r[idx['Kernel Name']].split('(')[0]}  (id {r[idx['ID']]})")
        for k in KEYS:
            if k in idx:
                out.append(f"- {k}: {r[idx[k]]} {units[idx[k]]}")
        out.append("")
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    tag = sys.argv[1]
    os.makedirs(PROF, exist_ok=True)
    s = launches(tag)
    if s:
        open(os.path.join(PROF, f"launches_{tag}.md"), "w").write(s)
        print(s)
    for stem in sys.argv[2:]:
        s = full(stem)
        if s:
            open(os.path.join(PROF, f"{stem}.md"), "w").write(s)
            print(s[:1500])
    for f in (f"bench_{tag}.json",):
        p = os.path.join(OUT, f)
        if os.path.exists(p) and os.path.getsize(p):
            open(os.path.join(PROF, f), "w").write(open(p).read())
