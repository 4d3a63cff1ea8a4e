#!/usr/bin/env python3
"""Prints the innermost backward-branch loop of a kernel's SASS that contains a given opcode
(default: the byte load of the action) and counts its instructions.

    cuobjdump -sass lib.so | python tools/sass_loop.py 'k_steps_boardILi4ELi10ELb1' [opcode-regex]
"""
import re
import sys

name = sys.argv[1]
needle = re.compile(sys.argv[2] if len(sys.argv) > 2 else r"LDG?\.E\.U8")
lines, on = [], False
for line in sys.stdin:
    if "Function :" in line:
        on = name in line
        continue
    if on:
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            lines.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(lines)}
best = None
for i, (a, ins) in enumerate(lines):
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)*(?:0x)([0-9a-f]+)", ins)
    if m:
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr:
            body = lines[addr[tgt]: i + 1]
            if any(needle.search(x) for _, x in body):
                if best is None or len(body) < len(best):
                    best = body
if best is None:
    print("no loop found")
    sys.exit(1)
quiet = "-q" in sys.argv
if not quiet:
    for a, ins in best:
        print(f"{a:05x}  {ins}")
ops = {}
for _, ins in best:
    op = ins.split()[1] if ins.startswith("@") else ins.split()[0]
    op = op.split(".")[0]
    ops[op] = ops.get(op, 0) + 1
print(f"# loop {best[0][0]:#x}..{best[-1][0]:#x}: {len(best)} instructions")
print("# " + ", ".join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])))
