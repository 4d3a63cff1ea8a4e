#!/usr/bin/env bash
mkdir -p gpurun_out
for b in 12; do
PW_OCC_DEBUG=1 PW_OCC_BATCH=$b python - <<'PY' 2>&1 | grep pw_occ | head -4
import os, sys, torch
sys.path.insert(0, ".")
from pushworld_b200 import data
from pushworld_b200.engine import Engine
keys = data.benchmark_keys()
eng = Engine([data.puzzle_text(k) for k in keys])
n, T = 1 << 18, 256
ids = torch.arange(n, dtype=torch.int64) % len(keys)
for pid in (torch.sort(ids, stable=True).values.to(torch.int32).cuda(), ids.to(torch.int32).cuda()):
    actions = torch.randint(0, 4, (T, n), dtype=torch.uint8, device="cuda")
    reward = torch.empty((T, n), dtype=torch.float32, device="cuda")
    term = torch.empty((T, n), dtype=torch.uint8, device="cuda")
    xy = eng.new_state(n)
    eng.reset(xy, pid)
    eng.rollout(xy, pid, actions, reward=reward, terminated=term)
    torch.cuda.synchronize()
PY
done
