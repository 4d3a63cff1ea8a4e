#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_job10.json 2> gpurun_out/bench_job10.err; tail -3 gpurun_out/bench_job10.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_job10.json').read().strip().splitlines()[-1])
print("value %.3e  e2e %.3e  e2e_f32 %.3e  e2e_sum %.3e frac %.3f kernel_ms %.4f"%(d['value'], d['e2e']['value'], d['e2e_f32']['value'], d['e2e_summaries']['value'], d['roofline']['frac'], d['roofline']['kernel_ms']))
for k,v in d['configs'].items(): print(k, "%.3e"%v['units_per_s'], "ms %.3f"%v['kernel_ms'], "frac %.3f"%v['frac'], "8d %.3f"%v['survey_8d']['frac'])
print(d['per_step_latency'])
print(d['clocks'], d['cpu_baseline'])
PY
