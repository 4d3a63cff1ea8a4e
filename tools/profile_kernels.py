#!/usr/bin/env python3
"""Runs one of the secondary workloads a few times so that ncu can capture its kernel:

    ncu --set full --clock-control none --import-source on -k regex:<kernel> -s 2 -c 1 \
        -o gpurun_out/prof_<name> python tools/profile_kernels.py <name>

names: step1 (pw_step T=1, Road Block 2^18), mixed (cfg3 rollout T=256), expand (cfg4, 2M states),
render_f32 / render_u8 / step_render_f32 (cfg5, 2^14 frames).
"""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from pushworld_b200 import data  # noqa: E402
from pushworld_b200.engine import Engine  # noqa: E402

name = sys.argv[1]
dev = torch.device("cuda", 0)
reps = 4
if name == "step1":
    eng = Engine([data.puzzle_text("level1/Road Block")], device=dev)
    n = 1 << 18
    xy = eng.new_state(n)
    eng.reset(xy, 0)
    act = torch.randint(0, 4, (n,), dtype=torch.uint8, device=dev)
    out = eng.step(xy, 0, act)
    for _ in range(reps):
        eng.step(xy, 0, act, out=out)
elif name == "mixed":
    keys = data.benchmark_keys()
    eng = Engine([data.puzzle_text(k) for k in keys], device=dev)
    n, T = 1 << 18, 256
    ids = torch.sort(torch.arange(n) % len(keys), stable=True).values.to(torch.int32).to(dev)
    actions = torch.randint(0, 4, (T, n), dtype=torch.uint8, device=dev)
    reward = torch.empty((T, n), dtype=torch.float32, device=dev)
    term = torch.empty((T, n), dtype=torch.uint8, device=dev)
    xy = eng.new_state(n)
    for _ in range(reps):
        eng.reset(xy, ids)
        eng.rollout(xy, ids, actions, reward=reward, terminated=term)
elif name == "expand":
    eng = Engine([data.puzzle_text("level2/Pull Dont Push")], device=dev)
    n = 2_000_000
    xy = eng.new_state(n)
    eng.reset(xy, 0)
    eng.rollout(xy, 0, torch.randint(0, 4, (64, n), dtype=torch.uint8, device=dev))
    for _ in range(reps):
        eng.expand(xy, 0)
elif name in ("render_f32", "render_u8", "step_render_f32"):
    eng = Engine([data.puzzle_text("level1/Road Block")], device=dev)
    n = 1 << 14
    xy = eng.new_state(n)
    eng.reset(xy, 0)
    act = torch.randint(0, 4, (n,), dtype=torch.uint8, device=dev)
    for _ in range(reps):
        if name == "step_render_f32":
            eng.step_render(xy, 0, act, dtype="float32")
        else:
            eng.render(xy, 0, dtype="float32" if name == "render_f32" else "uint8")
else:
    raise SystemExit(f"unknown workload {name}")
torch.cuda.synchronize()
print("done", name)
