#!/usr/bin/env bash
# GPU job 1 (round 2): the whole GPU suite, the general-rule A/B probe, a short bench run, and
# ncu captures of the new kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python tools/occ_probe.py --out gpurun_out/occ_probe.json > gpurun_out/occ_probe.log 2>&1; tail -20 gpurun_out/occ_probe.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_job1.json 2> gpurun_out/bench_job1.err; tail -3 gpurun_out/bench_job1.err; head -c 3000 gpurun_out/bench_job1.json
timeout 600 tools/ncu_capture.sh mixed_r2 k_rollout_occ 2 python tools/profile_kernels.py mixed
python tools/ncu_summary.py gpurun_out/mixed_r2.raw.csv.gz $((262144*64)) > gpurun_out/mixed_r2.summary.json; cat gpurun_out/mixed_r2.summary.json
