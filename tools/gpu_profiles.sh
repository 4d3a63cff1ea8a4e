#!/usr/bin/env bash
# Round-2 profile captures (one `ncu --set full` launch per kernel + the launch list of a bench run).
mkdir -p gpurun_out
tools/ncu_capture.sh mixed_r2 k_rollout_mask 2 python tools/profile_kernels.py mixed
python tools/ncu_summary.py gpurun_out/mixed_r2.raw.csv.gz $((262144*256)) > gpurun_out/mixed_r2.summary.json
python tools/ncu_lines.py gpurun_out/mixed_r2.source.csv.gz k_rollout_maskILb0ELb0 $((262144*256)) 1.0 > gpurun_out/mixed_r2.lines.txt
tools/ncu_capture.sh rollout_r2 k_rollout_pairs4 2 python bench.py --steps 2 --warmup 3 --no-configs
python tools/ncu_summary.py gpurun_out/rollout_r2.raw.csv.gz $((262144*1024)) > gpurun_out/rollout_r2.summary.json
python tools/ncu_lines.py gpurun_out/rollout_r2.source.csv.gz k_rollout_pairs4ILb0 $((262144*1024)) 0.3 > gpurun_out/rollout_r2.lines.txt
tools/ncu_capture.sh expand_r2 k_expand_uniform 2 python tools/profile_kernels.py expand
python tools/ncu_summary.py gpurun_out/expand_r2.raw.csv.gz $((2000000*4)) > gpurun_out/expand_r2.summary.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 > gpurun_out/launches_r2.bench.log 2>&1
cat gpurun_out/mixed_r2.summary.json gpurun_out/rollout_r2.summary.json | grep -E "Kernel|duration|inst_per|issue_active|dram__bytes"
