#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_level0.py tests/test_gpu_dynamics.py tests/test_gpu_packed.py tests/test_gpu_vector_env.py -x -q 2>&1 | tail -5
timeout 600 python tools/level0_probe.py 2>&1 | tee gpurun_out/level0_probe.log | cut -c1-250
