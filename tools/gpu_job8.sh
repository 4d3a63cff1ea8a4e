#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_table.py tests/test_gpu_fullsize.py tests/test_gpu_properties.py -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_configs.py --only cfg2 2>&1 | cut -c1-260
PW_NO_TABLE=1 timeout 600 python tools/bench_configs.py --only cfg2 2>&1 | head -1 | cut -c1-260
timeout 600 tools/ncu_capture.sh rollout_r2 k_rollout_table 2 python bench.py --steps 2 --warmup 3 --no-configs
python tools/ncu_summary.py gpurun_out/rollout_r2.raw.csv.gz $((262144*1024)) > gpurun_out/rollout_r2.summary.json; cat gpurun_out/rollout_r2.summary.json
