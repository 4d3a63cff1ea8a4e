#!/usr/bin/env python3
"""Secondary measurements for BASELINE configs 2-5 and the T=1 step kernel (1 GPU).

    python tools/bench_configs.py [--only cfg3,cfg4] [--out gpurun_out/configs.json]

Prints / writes one JSON object per workload: units/s, kernel ms (CUDA events, best of 5 after
3 warm-ups), algorithmic bytes per unit as defined in SURVEY.md 8(d) / DESIGN.md, achieved GB/s
and the fraction of the measured HBM copy peak.  bench.py stays the graded headline; this tool
feeds DESIGN.md's table and profiles/.
"""

from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from pushworld_b200 import data  # noqa: E402
from pushworld_b200.engine import Engine  # noqa: E402


def peak_gbs() -> float:
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"])
    except (OSError, KeyError, TypeError, ValueError):
        return 6650.0


def time_ms(fn, warmup=3, reps=5) -> float:
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def report(name, units, ms, bytes_per_unit, extra=None):
    rate = units / (ms * 1e-3)
    gbs = rate * bytes_per_unit / 1e9
    rec = {"workload": name, "units": units, "kernel_ms": ms, "units_per_s": rate,
           "bytes_per_unit": bytes_per_unit, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peak_gbs()}
    if extra:
        rec.update(extra)
    print(json.dumps(rec), flush=True)
    return rec


def rollout_case(name, engine, puzzle, n, T, seed):
    dev = engine.device
    gen = torch.Generator(device=dev).manual_seed(seed)
    actions = torch.randint(0, 4, (T, n), dtype=torch.uint8, device=dev, generator=gen)
    reward = torch.empty((T, n), dtype=torch.float32, device=dev)
    term = torch.empty((T, n), dtype=torch.uint8, device=dev)
    xy = engine.new_state(n)

    def run():
        engine.reset(xy, puzzle)
        engine.rollout(xy, puzzle, actions, reward=reward, terminated=term)

    ms = time_ms(run)
    if isinstance(puzzle, torch.Tensor):
        n_obj = torch.tensor([p.num_objects for p in engine.parsed], device=dev)[puzzle.long()].float().mean().item()
    else:
        n_obj = engine.parsed[puzzle].num_objects
    fused = 6.0 + 16.0 * n_obj / T
    return report(name, n * T, ms, fused, {"unit": "env-steps", "mean_objects": n_obj,
                                           "survey_8d_bytes_per_unit": 16 * n_obj + 6})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    only = set(filter(None, args.only.split(",")))
    want = lambda k: not only or k in only
    results = []
    keys = data.benchmark_keys()
    dev = torch.device("cuda", 0)

    if want("cfg2"):
        eng = Engine([data.puzzle_text("level1/Road Block")], device=dev)
        results.append(rollout_case("cfg2 rollout: Road Block 2^18 x 1024", eng, 0, 1 << 18, 1024, 1234))
        # the same puzzle through the T=1 kernel (what a per-step vector env pays)
        n = 1 << 18
        xy = eng.new_state(n)
        eng.reset(xy, 0)
        act = torch.randint(0, 4, (n,), dtype=torch.uint8, device=dev)
        out = eng.step(xy, 0, act)
        ms = time_ms(lambda: eng.step(xy, 0, act, out=out))
        results.append(report("cfg2 single step (pw_step): Road Block 2^18", n, ms, 16 * 3 + 6,
                              {"unit": "env-steps", "note": "one Python call per step: launch-bound"}))
        stepper = eng.prepare_step(xy, 0, out=out)
        ms = time_ms(lambda: stepper(act))
        results.append(report("cfg2 single step through Engine.prepare_step: Road Block 2^18", n, ms, 16 * 3 + 6,
                              {"unit": "env-steps", "note": "arguments bound once; still one launch per Python call"}))
        # the kernel itself: 50 steps captured in one CUDA graph (no host work between the launches)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                for _ in range(50):
                    eng.step(xy, 0, act, out=out)
        ms = time_ms(graph.replay) / 50
        results.append(report("cfg2 single step (pw_step) in a CUDA graph: Road Block 2^18", n, ms, 16 * 3 + 6,
                              {"unit": "env-steps"}))

    if want("cfg3"):
        eng = Engine([data.puzzle_text(k) for k in keys], device=dev)
        n = 1 << 18
        ids = torch.arange(n, dtype=torch.int64) % len(keys)
        grouped = torch.sort(ids, stable=True).values.to(torch.int32).to(dev)
        results.append(rollout_case("cfg3 rollout: 223 puzzles grouped, 2^18 x 256", eng, grouped, n, 256, 1235))
        interleaved = ids.to(torch.int32).to(dev)
        results.append(rollout_case("cfg3 rollout: 223 puzzles interleaved (e mod 223), 2^18 x 256",
                                    eng, interleaved, n, 256, 1235))

    if want("level0"):
        # RL-style mixed batch: 2048 synthetic level-0-like puzzles (5x5 interior, 3 objects, as in the
        # reference's level0.zip), per-environment puzzle ids, on-device episode handling
        rng = np.random.default_rng(2026)
        texts = []
        while len(texts) < 2048:
            cells = [["." for _ in range(5)] for _ in range(5)]
            free = [(x, y) for x in range(5) for y in range(5)]
            rng.shuffle(free)
            for name, (x, y) in zip(["a", "m0", "g0", "m1", "w", "w"], free):
                cells[y][x] = name
            texts.append("\n".join(" ".join(r) for r in cells) + "\n")
        eng = Engine(texts, device=dev)
        n, T = 1 << 18, 256
        ids = torch.randint(0, len(texts), (n,), dtype=torch.int32, device=dev)
        gen = torch.Generator(device=dev).manual_seed(77)
        actions = torch.randint(0, 4, (T, n), dtype=torch.uint8, device=dev, generator=gen)
        reward = torch.empty((T, n), dtype=torch.float32, device=dev)
        term = torch.empty((T, n), dtype=torch.uint8, device=dev)
        xy = eng.new_state(n)

        def run_plain():
            eng.reset(xy, ids)
            eng.rollout(xy, ids, actions, reward=reward, terminated=term)

        ms = time_ms(run_plain)
        results.append(report("level0-like mixed batch (2048 puzzles, random ids), rollout 2^18 x 256",
                              n * T, ms, 6.0 + 16.0 * 3 / T, {"unit": "env-steps"}))
        steps = torch.zeros(n, dtype=torch.int32, device=dev)
        episode = torch.zeros(n, dtype=torch.int32, device=dev)
        pid = ids.clone()

        def run_episodes():
            eng.rollout_episodes(xy, pid, actions, steps, episode, seed=5, max_steps=100,
                                 reward=reward, terminated=term)

        eng.reset(xy, pid)
        ms = time_ms(run_episodes)
        results.append(report("level0-like mixed batch, rollout_episodes (max_steps=100, re-sampling)",
                              n * T, ms, 6.0 + 16.0 * 3 / T, {"unit": "env-steps"}))

    if want("cfg4"):
        key = "level2/Pull Dont Push"
        eng = Engine([data.puzzle_text(key)], device=dev)
        n = 10_000_000
        xy = eng.new_state(n)
        eng.reset(xy, 0)
        gen = torch.Generator(device=dev).manual_seed(1236)
        walk = torch.randint(0, 4, (64, n), dtype=torch.uint8, device=dev, generator=gen)
        eng.rollout(xy, 0, walk)  # frontier = 64-step seeded random walks
        del walk
        xy_out = torch.empty((2, eng.n_max, 4 * n), dtype=torch.int32, device=dev)
        moved = torch.empty(4 * n, dtype=torch.int32, device=dev)
        goal = torch.empty(4 * n, dtype=torch.uint8, device=dev)
        ms = time_ms(lambda: eng.expand(xy, 0, xy_out=xy_out, moved=moved, goal=goal))
        n_obj = eng.n_max
        results.append(report("cfg4 expand: Pull Dont Push, 10M-state frontier", 4 * n, ms,
                              10 * n_obj + 5, {"unit": "successors"}))
        del xy_out, moved, goal, xy

    if want("cfg5"):
        eng = Engine([data.puzzle_text("level1/Road Block")], device=dev)
        n = 1 << 16
        xy = eng.new_state(n)
        eng.reset(xy, 0)
        act = torch.randint(0, 4, (n,), dtype=torch.uint8, device=dev)
        for dtype, size in (("float32", 4), ("uint8", 1)):
            obs = torch.empty((n, 100, 160, 3), dtype=torch.float32 if size == 4 else torch.uint8, device=dev)
            ms = time_ms(lambda: eng.step_render(xy, 0, act, dtype=dtype, out=obs))
            results.append(report(f"cfg5 step + render (pw_step_render): Road Block 2^16, {dtype} obs (100,160,3)",
                                  n, ms, 100 * 160 * 3 * size + 54, {"unit": "frames"}))
            ms = time_ms(lambda: eng.render(xy, 0, dtype=dtype, out=obs))
            results.append(report(f"cfg5 render only: Road Block 2^16, {dtype}", n, ms,
                                  100 * 160 * 3 * size + 24, {"unit": "frames"}))
            del obs
        # standard padding (54 x 47 cells): mostly zero fill
        n = 2048
        xy = eng.new_state(n)
        eng.reset(xy, 0)
        obs = torch.empty((n, 1080, 940, 3), dtype=torch.float32, device=dev)
        ms = time_ms(lambda: eng.render(xy, 0, dtype="float32", pad_cells=(54, 47), out=obs))
        results.append(report("render standard padding: Road Block 2048, float32 (1080,940,3)", n, ms,
                              1080 * 940 * 3 * 4, {"unit": "frames"}))

    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
