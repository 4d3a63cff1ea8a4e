#!/usr/bin/env python3
"""Block-size sweep of the state-table rollout kernel on BASELINE config 2 (PW_TABLE_BLOCK)."""
import json, os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from pushworld_b200 import data
from pushworld_b200.engine import Engine

def time_ms(fn, warmup=3, reps=7):
    for _ in range(warmup): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best

eng = Engine([data.puzzle_text("level1/Road Block")])
n, T = 1 << 18, 1024
gen = torch.Generator(device="cuda").manual_seed(1234)
actions = torch.randint(0, 4, (T, n), dtype=torch.uint8, device="cuda", generator=gen)
reward = torch.empty((T, n), dtype=torch.float32, device="cuda")
term = torch.empty((T, n), dtype=torch.uint8, device="cuda")
record = torch.empty((T, n), dtype=torch.uint8, device="cuda")
xy = eng.new_state(n)
eng.reset(xy, 0)
xy0 = xy.clone()
for wide in (False, True):
    if wide: os.environ["PW_TABLE_WIDE"] = "1"
    for block in [0, 128, 256, 384, 448, 512, 640, 768, 896, 928, 960, 1024]:
        os.environ["PW_TABLE_BLOCK"] = str(block)
        def run():
            xy.copy_(xy0)
            eng.rollout(xy, 0, actions, reward=reward, terminated=term)
        def run_p():
            xy.copy_(xy0)
            eng.rollout_packed(xy, 0, actions, record)
        ms, msp = time_ms(run), time_ms(run_p)
        print(json.dumps({"wide": wide, "block": block, "plain_ms": round(ms, 4), "plain_frac": round(n*T*6.047/ms/1e6/6530.3, 3),
                          "packed_ms": round(msp, 4)}), flush=True)
