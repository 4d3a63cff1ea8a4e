#!/usr/bin/env bash
# Usage (on the GPU box, from the repo root):  tools/ncu_capture.sh <tag> <kernel-regex> <skip> <cmd...>
# One `ncu --set full` capture of the first matching launch after <skip>; writes the compact,
# reviewable artefacts gpurun_out/<tag>.raw.csv.gz and gpurun_out/<tag>.source.csv.gz and drops
# the bulky .ncu-rep (gpurun only brings back 64 MiB).
set -u
tag=$1; kernel=$2; skip=$3; shift 3
mkdir -p gpurun_out
rep=/tmp/${tag}.ncu-rep
ncu --set full --clock-control none --import-source on -k "regex:${kernel}" -s "${skip}" -c 1 -f \
    -o "${rep%.ncu-rep}" "$@" > "gpurun_out/${tag}.log" 2>&1
ncu -i "${rep}" --page raw --csv 2>/dev/null | gzip -9 > "gpurun_out/${tag}.raw.csv.gz"
ncu -i "${rep}" --page source --csv 2>/dev/null | gzip -9 > "gpurun_out/${tag}.source.csv.gz"
ncu -i "${rep}" --page source --print-source cuda --csv 2>/dev/null | gzip -9 > "gpurun_out/${tag}.cuda.csv.gz"
ls -la "gpurun_out/${tag}".*
