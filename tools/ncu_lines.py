#!/usr/bin/env python3
"""Attributes an ncu SASS source page (tools/ncu_capture.sh) to CUDA source lines, using the line
table of the locally built library (same binary as on the GPU box):

    python tools/ncu_lines.py gpurun_out/<tag>.source.csv.gz <kernel-substring> [units] [min_per_unit]

Prints, per source line (file:line, incl. inlined callees), executed warp-instructions per 32
units, average active lanes and stall-sample share.  Needs cuobjdump / nvdisasm (CUDA toolkit)."""
import csv
import gzip
import os
import re
import subprocess
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path, kernel = sys.argv[1], sys.argv[2]
units = float(sys.argv[3]) if len(sys.argv) > 3 else None
min_per = float(sys.argv[4]) if len(sys.argv) > 4 else 2.0
lib = os.path.join(REPO, "pushworld_b200", "lib", "libpushworld_b200.so")

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, capture_output=True)
lines_of = None
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin"):
        continue
    sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur_fn, cur_line, table = None, None, {}
    for ln in sass.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            cur_fn = m.group(1)
            table.setdefault(cur_fn, [])
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            inl = " (inlined)" if "inlined" in m.group(3) else ""
            cur_line = f"{os.path.basename(m.group(1))}:{m.group(2)}"
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and cur_fn:
            table[cur_fn].append((int(m.group(1), 16), cur_line, m.group(2)))
    for fn, tab in table.items():
        if kernel in fn and tab:
            lines_of = tab
            break
    if lines_of:
        break
if not lines_of:
    raise SystemExit(f"kernel {kernel} not found in {lib}")

rows = list(csv.reader(gzip.open(path, "rt")))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > ci["Instructions Executed"]]
base = int(data[0][ci["Address"]], 16)
if len(data) != len(lines_of):
    print(f"warning: {len(data)} profiled instructions vs {len(lines_of)} in the local binary", file=sys.stderr)
by_off = {off: (line, text) for off, line, text in lines_of}
agg = {}
total = 0
tot_samples = 0
for r in data:
    off = int(r[ci["Address"]], 16) - base
    line = by_off.get(off, ("?", ""))[0]
    ex = int(r[ci["Instructions Executed"]])
    th = int(r[ci["Thread Instructions Executed"]])
    sm = int(r[ci["# Samples"]] or 0)
    a = agg.setdefault(line, [0, 0, 0, 0])
    a[0] += ex
    a[1] += th
    a[2] += sm
    a[3] += 1
    total += ex
    tot_samples += sm
scale = (units / 32.0) if units else 1.0
print(f"total warp-instructions {total}" + (f" = {total / scale:.1f} per 32 units" if units else ""))


def key(item):
    name = item[0]
    f, _, n = name.partition(":")
    return (f, int(n) if n.isdigit() else 0)


for line, (ex, th, sm, n) in sorted(agg.items(), key=key):
    per = ex / scale
    if units and per < min_per:
        continue
    print(f"{line:22s} sass={n:4d} exec/32u={per:8.1f} lanes={th / max(ex, 1):5.1f} stall={100.0 * sm / max(tot_samples, 1):5.1f}%")
