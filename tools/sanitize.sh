#!/usr/bin/env bash
# compute-sanitizer over the small-size GPU suites: memcheck on everything that
# finishes in minutes under instrumentation, racecheck (shared-memory hazards) and synccheck on the
# suites that drive the shared-memory kernels (action tiles, tables, masks, render patterns).
# Usage (GPU box, repo root): tools/sanitize.sh ; logs go to gpurun_out/sanitizer_*.log
mkdir -p gpurun_out
CS=compute-sanitizer
SMALL="tests/test_gpu_packed.py tests/test_gpu_table.py tests/test_gpu_dynamics.py tests/test_gpu_dropin.py tests/test_gpu_vector_env.py tests/test_gpu_search.py tests/test_gpu_render.py"
timeout 1500 $CS --tool memcheck --error-exitcode 1 python -m pytest $SMALL -x -q > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck exit code $?" >> gpurun_out/sanitizer_memcheck.log
timeout 1200 $CS --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_packed.py tests/test_gpu_table.py -x -q > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck exit code $?" >> gpurun_out/sanitizer_racecheck.log
timeout 900 $CS --tool synccheck --error-exitcode 1 python -m pytest tests/test_gpu_packed.py tests/test_gpu_table.py -x -q > gpurun_out/sanitizer_synccheck.log 2>&1
echo "synccheck exit code $?" >> gpurun_out/sanitizer_synccheck.log
tail -n 4 gpurun_out/sanitizer_memcheck.log gpurun_out/sanitizer_racecheck.log gpurun_out/sanitizer_synccheck.log
