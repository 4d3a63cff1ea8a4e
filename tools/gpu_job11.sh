#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dynamics.py tests/test_gpu_search.py tests/test_gpu_properties.py tests/test_gpu_fullsize.py::test_config4_expansion_matches_the_c_oracle -x -q 2>&1 | tail -4
timeout 600 python tools/bench_configs.py --only cfg4 2>&1 | cut -c1-300
timeout 600 tools/ncu_capture.sh expand_r2 k_expand_uniform 2 python tools/profile_kernels.py expand
python tools/ncu_summary.py gpurun_out/expand_r2.raw.csv.gz $((2000000*4)) > gpurun_out/expand_r2.summary.json; cat gpurun_out/expand_r2.summary.json
