#!/usr/bin/env python3
"""Summarises an ncu raw CSV (tools/ncu_capture.sh output) into the handful of numbers the
roofline discussion needs.   python tools/ncu_summary.py gpurun_out/<tag>.raw.csv.gz [units]"""
import csv
import gzip
import json
import sys

path = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
op = gzip.open if path.endswith(".gz") else open
rows = list(csv.reader(op(path, "rt")))
hdr, unit, val = rows[0], rows[1], rows[2]
get = {h: (unit[i], val[i]) for i, h in enumerate(hdr)}
keys = [
    "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct",
]
out = {}
for k in keys:
    if k in get:
        out[k] = f"{get[k][1]} {get[k][0]}".strip()
stalls = {h.split("issue_stalled_")[1]: float(v[1].replace(",", "")) for h, v in get.items()
          if "pcsamp_warps_issue_stalled_" in h and not h.endswith("_not_issued")}
tot = sum(stalls.values()) or 1.0
out["stalls_pct"] = {k: round(100 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1]) if v / tot > 0.01}
if units:
    inst = float(get["smsp__inst_executed.sum"][1].replace(",", ""))
    out["warp_inst_per_32_units"] = round(inst / (units / 32), 1)
print(json.dumps(out, indent=1))
