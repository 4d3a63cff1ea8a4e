#!/usr/bin/env python3
"""Collapses an ncu SASS source page (tools/ncu_capture.sh) into basic-block-like runs:
    python tools/ncu_hot.py gpurun_out/<tag>.source.csv.gz [min_share_pct]
Each line: start address offset, number of instructions, executions per instruction, average
active threads, share of all executed warp-instructions, first opcode(s)."""
import csv
import gzip
import sys

path = sys.argv[1]
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = list(csv.reader(gzip.open(path, "rt")))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
ia, isrc, iex, ith = ci["Address"], ci["Source"], ci["Instructions Executed"], ci["Avg. Threads Executed"]
ist = ci["# Samples"]
data = []
for r in rows[2:]:
    if len(r) <= iex:
        continue
    data.append((int(r[ia], 16), r[isrc].strip(), int(r[iex]), float(r[ith] or 0), int(r[ist] or 0)))
base = data[0][0]
total = sum(d[2] for d in data)
tot_samples = sum(d[4] for d in data) or 1
runs = []
for d in data:
    if runs and runs[-1]["count"] == d[2] and abs(runs[-1]["thr"] - d[3]) < 0.01:
        runs[-1]["n"] += 1
        runs[-1]["samples"] += d[4]
        runs[-1]["ops"].append(d[1].split()[0] if not d[1].startswith("@") else d[1].split()[1])
    else:
        runs.append({"addr": d[0] - base, "n": 1, "count": d[2], "thr": d[3], "samples": d[4],
                     "ops": [d[1].split()[0] if not d[1].startswith("@") else d[1].split()[1]]})
print(f"total warp-instructions {total}, instructions {len(data)}")
for r in runs:
    share = 100.0 * r["n"] * r["count"] / total
    if share >= min_share:
        print(f"{r['addr']:#7x} n={r['n']:4d} exec={r['count']:10d} thr={r['thr']:5.1f} "
              f"share={share:5.1f}% stall={100.0 * r['samples'] / tot_samples:5.1f}%  {' '.join(r['ops'][:8])}")
