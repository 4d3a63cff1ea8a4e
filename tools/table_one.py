#!/usr/bin/env python3
"""A few launches of the state-table rollout on BASELINE config 2 for an ncu capture:
    tools/table_one.py plain|packed     (PW_TABLE_WIDE / PW_TABLE_CTAS / PW_TABLE_BLOCK select the variant)"""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from pushworld_b200 import data
from pushworld_b200.engine import Engine

mode = sys.argv[1] if len(sys.argv) > 1 else "plain"
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
for _ in range(4):
    xy.copy_(xy0)
    if mode == "packed":
        eng.rollout_packed(xy, 0, actions, record)
    else:
        eng.rollout(xy, 0, actions, reward=reward, terminated=term)
torch.cuda.synchronize()
