#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python tools/table_probe.py 2>&1 | head -3 | tee gpurun_out/table_probe.log
timeout 900 python -m pytest tests/test_gpu_table.py tests/test_gpu_fullsize.py::test_config2_every_column_matches_the_c_oracle -x -q 2>&1 | tail -3
