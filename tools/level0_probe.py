#!/usr/bin/env python3
"""The reference's real level-0 set (15,400 puzzles in one engine), 2^18 environments with random
puzzle ids: plain rollout and on-device episodes, on the mask kernel and (PW_NO_OCC=1)
on round 1 scan-every-step kernels."""
import json, os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from pushworld_b200 import data
from pushworld_b200.engine import Engine

def time_ms(fn, warmup=2, reps=5):
    for _ in range(warmup): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best

keys = data.level0_keys()
eng = Engine([data.level0_text(k) for k in keys])
n, T = 1 << 18, 256
gen = torch.Generator(device="cuda").manual_seed(4321)
pid0 = torch.randint(0, len(keys), (n,), dtype=torch.int32, device="cuda", generator=gen)
actions = torch.randint(0, 4, (T, n), dtype=torch.uint8, device="cuda", generator=gen)
reward = torch.empty((T, n), dtype=torch.float32, device="cuda")
term = torch.empty((T, n), dtype=torch.uint8, device="cuda")
trunc = torch.empty((T, n), dtype=torch.uint8, device="cuda")
xy = eng.new_state(n)
res = {}
for label, env in (("mask kernel (k_rollout_mask)", {}), ("scan every step (round 1)", {"PW_NO_OCC": "1"})):
    os.environ.pop("PW_NO_OCC", None); os.environ.update(env)
    pid = pid0.clone()
    def plain():
        eng.reset(xy, pid); eng.rollout(xy, pid, actions, reward=reward, terminated=term)
    ms = time_ms(plain)
    dig = (int(reward.view(torch.int32).to(torch.int64).sum()), int(term.sum(dtype=torch.int64)), int(xy.to(torch.int64).sum()))
    steps = torch.zeros(n, dtype=torch.int32, device="cuda"); episode = torch.zeros(n, dtype=torch.int32, device="cuda")
    eng.reset(xy, pid, steps=steps)
    def episodes():
        eng.rollout_episodes(xy, pid, actions, steps, episode, seed=11, max_steps=100, reward=reward, terminated=term, truncated=trunc)
    ms_e = time_ms(episodes)
    def episodes_long():  # no truncation inside the rollout: only solved puzzles restart
        eng.rollout_episodes(xy, pid, actions, steps, episode, seed=11, max_steps=10**6, reward=reward, terminated=term, truncated=trunc)
    eng.reset(xy, pid, steps=steps)
    ms_l = time_ms(episodes_long)
    n_restarts = int(episode.sum())
    res[label] = dig
    print(json.dumps({"path": label, "plain_ms": round(ms, 3), "plain_steps_per_s": n*T/ms*1e3, "episodes_ms": round(ms_e, 3),
                      "episodes_steps_per_s": n*T/ms_e*1e3, "episodes_no_truncation_ms": round(ms_l, 3),
                      "restarts_seen_no_truncation": n_restarts, "digest": dig}), flush=True)
print("plain rollouts equal:", len(set(res.values())) == 1)
