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
VARIANTS = {
    "quad": {},                                             # two-step table, four environments per thread
    "pairs": {"PW_NO_QUAD": "1"},                           # two-step table, one environment per thread
    "compact1": {"PW_NO_PAIRS": "1", "PW_TABLE_CTAS": "1"},  # one-step compact entries, 64 registers
    "compact2": {"PW_NO_PAIRS": "1", "PW_TABLE_CTAS": "2"},  # ... 32 registers, two CTAs per SM
    "wide": {"PW_TABLE_WIDE": "1"},                         # one-step 32-bit entries + reward table
}
only = sys.argv[1:] or list(VARIANTS)
for name in only:
    for k in ("PW_NO_QUAD", "PW_NO_PAIRS", "PW_TABLE_CTAS", "PW_TABLE_WIDE"):
        os.environ.pop(k, None)
    os.environ.update(VARIANTS[name])
    blocks = [0, 128, 192, 256, 320, 384, 448, 512] if name == "quad" else [0, 448, 512, 896, 1024]
    for block in blocks:
        os.environ["PW_TABLE_BLOCK"] = str(block)
        def run():
            xy.copy_(xy0)
            eng.rollout(xy, 0, actions, reward=reward, terminated=term)
        def run_p():
            xy.copy_(xy0)
            eng.rollout_packed(xy, 0, actions, record)
        ms, msp = time_ms(run), time_ms(run_p)
        print(json.dumps({"variant": name, "block": block, "plain_ms": round(ms, 4), "plain_frac": round(n*T*6.047/ms/1e6/6530.3, 3),
                          "packed_ms": round(msp, 4)}), flush=True)
