#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python tools/occ_probe.py --out gpurun_out/occ_probe.json > gpurun_out/occ_probe.log 2>&1; cat gpurun_out/occ_probe.log | cut -c1-170
timeout 600 tools/ncu_capture.sh mixed_r2 k_rollout_occ 2 python tools/profile_kernels.py mixed
python tools/ncu_summary.py gpurun_out/mixed_r2.raw.csv.gz $((262144*256)) > gpurun_out/mixed_r2.summary.json; cat gpurun_out/mixed_r2.summary.json
