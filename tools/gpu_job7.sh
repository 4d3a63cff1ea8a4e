#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python tools/occ_probe.py --out gpurun_out/occ_probe.json > gpurun_out/occ_probe.log 2>&1
python - <<'PY'
import json
for r in json.load(open('gpurun_out/occ_probe.json')):
    print("%-12s %-95s %6.2f ms %.3e %s"%(r['layout'], r['variant'][:95], r['kernel_ms'], r['env_steps_per_s'], r['equal_to_first_variant']))
PY
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_job7.json 2> gpurun_out/bench_job7.err; tail -3 gpurun_out/bench_job7.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_job7.json').read().strip().splitlines()[-1])
print("value %.3e  e2e %.3e  e2e_f32 %.3e  e2e_sum %.3e frac %.3f"%(d['value'], d['e2e']['value'], d['e2e_f32']['value'], d['e2e_summaries']['value'], d['roofline']['frac']))
for k,v in d['configs'].items(): print(k, "%.3e"%v['units_per_s'], "ms %.3f"%v['kernel_ms'], "frac %.3f"%v['frac'], "8d %.3f"%v['survey_8d']['frac'])
print(d['clocks'])
PY
