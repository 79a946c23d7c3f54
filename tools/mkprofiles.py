#!/usr/bin/env python
"""Turn the raw files of one tools/gpu_round2.sh visit (gpurun_out/<tag>_*) into the tracked summaries under profiles/:
   <tag>_launches_ncu.csv   the ncu launch list as captured (gpu__time_duration.sum per launch)
   <tag>_kernel_list.txt    per-kernel time / instructions / DRAM bytes / issue of two steady-state scans
   <tag>_dram_traffic.csv   DRAM read / write MB per kernel of ONE scan (bench.py sums it for roofline.traffic)
   <tag>_full_metrics.csv   selected metrics of the `ncu --set full` capture, one row per launch
Usage: python tools/mkprofiles.py <tag>"""
import csv
import os
import shutil
import subprocess
import sys

tag = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def short(name):
    n = name.split("(")[0]
    for pre in ("void ", "<unnamed>::"):
        n = n.replace(pre, "")
    return n.strip()


src = os.path.join(G, tag + "_launches.csv")
if os.path.exists(src):
    shutil.copy(src, os.path.join(P, tag + "_launches_ncu.csv"))
src = os.path.join(G, tag + "_list.txt")
if os.path.exists(src):
    shutil.copy(src, os.path.join(P, tag + "_kernel_list.txt"))

# one scan's worth of launches from the few-metric list: from one reg_reset_kernel to the next
lst = os.path.join(G, tag + "_list.csv")
if os.path.exists(lst):
    rows = [r for r in csv.reader(open(lst)) if len(r) > 10]
    h = rows[0]
    ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
    d = {}
    for r in rows[1:]:
        d.setdefault(int(r[ii]), {"name": short(r[ki])})[r[mi]] = float(r[vi].replace(",", ""))
    ids = sorted(d)
    starts = [i for i in ids if d[i]["name"] == "reg_reset_kernel"]
    if len(starts) >= 2:
        scan = [i for i in ids if starts[0] <= i < starts[1]]
        with open(os.path.join(P, tag + "_dram_traffic.csv"), "w") as f:
            f.write("# DRAM traffic per launch of one steady-state scan (default bench workload, 1 x B200), ncu capture %s\n" % tag)
            f.write("# (gpurun_out/%s_list.csv; dram__bytes_read.sum / dram__bytes_write.sum); bench.py sums the update kernels\n" % tag)
            f.write("kernel,dram_read_MB,dram_write_MB,time_us\n")
            for i in scan:
                m = d[i]
                f.write("%s,%.3f,%.3f,%.1f\n" % (m["name"], m.get("dram__bytes_read.sum", 0) / 1e6,
                                               m.get("dram__bytes_write.sum", 0) / 1e6, m.get("gpu__time_duration.sum", 0) / 1e3))

rep = os.path.join(G, tag + "_prof.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h = rows[0]
    want = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
            "smsp__inst_executed_op_global_red.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]
    idx = [h.index(w) for w in want if w in h]
    with open(os.path.join(P, tag + "_full_metrics.csv"), "w", newline="") as f:
        w = csv.writer(f)
        for r in rows:
            if len(r) >= len(h):
                w.writerow([r[i] for i in idx])
print("profiles written for", tag)
