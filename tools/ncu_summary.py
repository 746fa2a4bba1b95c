#!/usr/bin/env python
"""Condenses ncu output brought back in gpurun_out/ into the tracked profiles/ folder:
  * <launches.csv>  (ncu --metrics gpu__time_duration.sum ... --csv)  -> per-kernel shares
  * <prof.ncu-rep>  (ncu --set full)                                   -> key counters per launch
usage: tools/ncu_summary.py TAG launches.csv [prof.ncu-rep]"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches = sys.argv[1], sys.argv[2]
rep = sys.argv[3] if len(sys.argv) > 3 else None
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

lines = []
if launches != "-":
    with open(launches) as f:
        lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").strip()
    v = float(row["Metric Value"].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values()) or 1.0
md = ["# ncu summary (%s)" % tag, ""]
if agg:
    md += ["Launch list: `%s` (`ncu --metrics gpu__time_duration.sum --clock-control none`; per-launch "
           "times are cold-cache and serialised: compare SHARES, not absolutes)." % os.path.basename(launches),
           "", "| kernel | launches | total us | avg us | share |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        md.append("| `%s` | %d | %.1f | %.2f | %.3f |" % (k, v[0], v[1], v[1] / v[0], v[1] / tot))
    md.append("")

if rep:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
            "launch__occupancy_limit_shared_mem",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum",
            "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "lts__t_sector_hit_rate.pct",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
    md += ["# ncu --set full, selected counters per captured launch", "",
           "Source: `%s`." % os.path.basename(rep), ""]
    min_us = float(os.environ.get("NCU_MIN_US", "0"))   # skip short launches in the detail part
    for r in rows[2:]:
        try:
            dur = float(r[idx["gpu__time_duration.sum"]].replace(",", ""))
            dur = {"ns": dur / 1e3, "us": dur, "ms": dur * 1e3}.get(units[idx["gpu__time_duration.sum"]], dur)
        except Exception:  # noqa: BLE001
            dur = 1e9
        if dur < min_us:
            continue
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "")
        md.append("## `%s`" % name)
        md.append("")
        md.append("| metric | value | unit |")
        md.append("|---|---:|---|")
        for w in want:
            if w in idx:
                md.append("| %s | %s | %s |" % (w, r[idx[w]], units[idx[w]]))
        md.append("")
with open(os.path.join(out_dir, "ncu_%s.md" % tag), "w") as f:
    f.write("\n".join(md))
print("\n".join(md[:30]))
