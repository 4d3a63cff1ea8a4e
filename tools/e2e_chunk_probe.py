#!/usr/bin/env python3
"""Chunk-size sweep of the host-buffer packed rollout (PushWorldVectorEnv.rollout_host_packed) on BASELINE config 2."""
import json, os, sys, time
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from pushworld_b200 import data
from pushworld_b200.vector_env import PushWorldVectorEnv

dev = torch.device("cuda", 0)
N, T = 1 << 18, 1024
env = PushWorldVectorEnv([data.puzzle_text("level1/Road Block")], N, puzzle_ids=0, device=dev)
h_act = torch.randint(0, 4, (T, N), dtype=torch.uint8).pin_memory()
h_rec = torch.empty((T, N), dtype=torch.uint8).pin_memory()
for chunk in (16, 32, 64, 128, 256, 512, 1024):
    def step():
        env.reset()
        env.rollout_host_packed(h_act, h_rec, chunk_steps=chunk)
    for _ in range(3): step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    a.record()
    for _ in range(reps): step()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print(json.dumps({"chunk_steps": chunk, "ms_per_rollout": round(ms, 3), "env_steps_per_s": N * T / ms * 1e3,
                      "GBps_each_way": round(N * T / ms / 1e6, 1)}), flush=True)
