#!/usr/bin/env python3
"""Golden vectors for the reference's generated LEVEL-0 puzzles (benchmark/puzzles/level0.zip),
produced by RUNNING THE REAL REFERENCE (build container only, like make_golden.py):

    level0_vectors.npz   for a fixed sample of 420 of the 15,400 puzzles (60 per category, taken
                         at regular intervals over the sorted keys): parse results and a seeded
                         200-step random walk (states, achieved goals, goal flags).
"""

import gzip
import json
import os
import random
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("PUSHWORLD_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(REF, "python3", "src"))

from pushworld.puzzle import PushWorldPuzzle  # noqa: E402

SEED = 20261017
PER_CATEGORY = 60


def sample_keys(all_keys):
    by_cat = {}
    for k in all_keys:
        by_cat.setdefault(k.split("/")[1], []).append(k)
    picked = []
    for cat in sorted(by_cat):
        keys = by_cat[cat]
        step = len(keys) // PER_CATEGORY
        picked += keys[::step][:PER_CATEGORY]
    return picked


def main():
    with gzip.open(os.path.join(REPO, "pushworld_b200", "data", "level0.json.gz")) as f:
        puzzles = json.loads(f.read().decode())["puzzles"]
    keys = sample_keys(sorted(puzzles))
    tmp = tempfile.mkdtemp()
    arrays, index = {}, []
    for pi, key in enumerate(keys):
        path = os.path.join(tmp, "p.pwp")
        with open(path, "w") as f:
            f.write(puzzles[key])
        puzzle = PushWorldPuzzle(path)
        n = len(puzzle.initial_state)
        rng = random.Random(f"{SEED}:{key}")
        actions = [rng.randrange(4) for _ in range(200)]
        state = puzzle.initial_state
        states = [[[int(v) for v in p] for p in state]]
        for a in actions:
            state = puzzle.get_next_state(state, a)
            states.append([[int(v) for v in p] for p in state])
        tup = [tuple(map(tuple, s)) for s in states]
        arrays[f"p{pi}_actions"] = np.array(actions, np.uint8)
        arrays[f"p{pi}_states"] = np.array(states, np.int8).reshape(len(states), n, 2)
        arrays[f"p{pi}_achieved"] = np.array([puzzle.count_achieved_goals(s) for s in tup], np.int8)
        arrays[f"p{pi}_isgoal"] = np.array([puzzle.is_goal_state(s) for s in tup], np.bool_)
        arrays[f"p{pi}_goal"] = np.array([[int(v) for v in p] for p in puzzle.goal_state], np.int8).reshape(-1, 2)
        index.append({"key": key, "width": puzzle.dimensions[0], "height": puzzle.dimensions[1], "n": n,
                      "g": len(puzzle.goal_state)})
        if pi % 50 == 0:
            print(f"[{pi + 1}/{len(keys)}] {key}", flush=True)
    arrays["index_json"] = np.frombuffer(json.dumps(index).encode(), np.uint8)
    np.savez_compressed(os.path.join(HERE, "level0_vectors.npz"), **arrays)
    print("wrote", len(keys), "puzzles")


if __name__ == "__main__":
    main()
