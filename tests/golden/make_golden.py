#!/usr/bin/env python3
"""Generates the golden vectors in this directory by RUNNING THE REAL REFERENCE.

Build-container only (needs /root/reference; takes ~3 min because the reference
builds O(cells^2) collision tables per puzzle).  The outputs are committed:

    reference_vectors.npz   states along solution replays and seeded random walks
                            for all 223 benchmark puzzles + the reference's own
                            test fixtures (python3/test/puzzles, cpp/test/puzzles)
    reference_renders.json  sha256 of PushWorldPuzzle.render / padded observations
    reference_gym.json      (reward, terminated, truncated) sequences from the
                            unmodified reference gym_env.PushWorldEnv

Nothing in tests/ reads /root/reference at run time; they read these files.
"""

import gzip
import hashlib
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
sys.path.append(os.path.join(REPO, "pushworld_b200", "compat", "shims"))

from pushworld.puzzle import Actions, PushWorldPuzzle  # noqa: E402
from pushworld.utils.env_utils import render_observation_padded  # noqa: E402
from pushworld.gym_env import PushWorldEnv  # noqa: E402

WALK_SEED = 20260117
RENDER_PUZZLES = [
    "level1/Road Block",
    "level1/Many Small Tools",
    "level2/Pull Dont Push",
    "level2/Clean Sweep",
    "level3/Yin Yang",
    "level4/Four Rooms",
]
RENDER_SETTINGS = [(20, 2), (5, 1), (7, 3), (12, 2), (3, 1)]


def as_int_state(state):
    return [[int(v) for v in p] for p in state]


def walk(puzzle, actions, start=None):
    state = puzzle.initial_state if start is None else start
    states = [as_int_state(state)]
    for a in actions:
        state = puzzle.get_next_state(state, int(a))
        states.append(as_int_state(state))
    return states


def sha(arr):
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()


def main():
    with gzip.open(os.path.join(REPO, "pushworld_b200", "data", "benchmark.json.gz")) as f:
        bundle = json.loads(f.read().decode())
    with open(os.path.join(HERE, "test_puzzles.json")) as f:
        fixtures = json.load(f)

    texts = {}
    for key, text in bundle["puzzles"].items():
        texts["benchmark/" + key] = text
    for name, text in fixtures["python3"].items():
        texts["python3/" + name] = text
    for name, text in fixtures["cpp"].items():
        if fixtures["python3"].get(name) != text:
            texts["cpp/" + name] = text

    arrays = {}
    index = []
    renders = {}
    tmp = tempfile.mkdtemp()
    loaded = {}
    for pi, key in enumerate(sorted(texts)):
        path = os.path.join(tmp, "p.pwp")
        with open(path, "w") as f:
            f.write(texts[key])
        puzzle = PushWorldPuzzle(path)
        loaded[key] = puzzle
        n = len(puzzle.initial_state)
        rng = random.Random(f"{WALK_SEED}:{key}")
        rec = {
            "key": key,
            "width": puzzle.dimensions[0],
            "height": puzzle.dimensions[1],
            "n": n,
            "g": len(puzzle.goal_state),
        }
        arrays[f"p{pi}_init"] = np.array(as_int_state(puzzle.initial_state), np.int16).reshape(n, 2)
        arrays[f"p{pi}_goal"] = np.array(as_int_state(puzzle.goal_state), np.int16).reshape(-1, 2)

        seqs = []
        if key.startswith("benchmark/"):
            plan = bundle["solutions"][key[len("benchmark/"):]]["plan"]
            seqs.append(("sol", [Actions.FROM_CHAR[c] for c in plan.upper()]))
            seqs.append(("walk", [rng.randrange(4) for _ in range(300)]))
        else:
            seqs.append(("walk", [rng.randrange(4) for _ in range(200)]))
        for tag, actions in seqs:
            states = walk(puzzle, actions)
            arrays[f"p{pi}_{tag}_actions"] = np.array(actions, np.uint8)
            arrays[f"p{pi}_{tag}_states"] = np.array(states, np.int16).reshape(len(states), n, 2)
            tup = [tuple(map(tuple, s)) for s in states]
            arrays[f"p{pi}_{tag}_achieved"] = np.array(
                [puzzle.count_achieved_goals(s) for s in tup], np.int8
            )
            arrays[f"p{pi}_{tag}_isgoal"] = np.array(
                [puzzle.is_goal_state(s) for s in tup], np.bool_
            )
            if tag == "sol":
                assert puzzle.is_valid_plan(actions), key
        index.append(rec)
        print(f"[{pi + 1}/{len(texts)}] {key} n={n}", flush=True)

    arrays["index_json"] = np.frombuffer(json.dumps(index).encode(), np.uint8)
    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **arrays)

    # ---------------------------------------------------------------- renders
    trivial = loaded["python3/trivial"]
    frames = trivial.render_plan([Actions.RIGHT, Actions.DOWN, Actions.RIGHT, Actions.UP])
    renders["trivial_plan_RDRU"] = {
        "sha256": [sha(img) for img in frames],
        # the same values test_puzzle.py:264-270 pins
        "python_tuple_hash": [hash(tuple(img.flat)) for img in frames],
        "shape": list(frames[0].shape),
    }
    cases = []
    keys = ["benchmark/" + k for k in RENDER_PUZZLES] + [
        "python3/trivial_tool",
        "python3/trivial_overlap",
        "python3/file_parsing",
        "python3/trivial",
    ]
    for key in keys:
        puzzle = loaded[key]
        rng = random.Random(f"render:{key}")
        state = puzzle.initial_state
        states = [state]
        for _ in range(3):
            for _ in range(25):
                state = puzzle.get_next_state(state, rng.randrange(4))
            states.append(state)
        for ppc, bw in RENDER_SETTINGS:
            if ppc < 1 + 2 * bw:
                continue
            for state in states:
                img = puzzle.render(state, border_width=bw, pixels_per_cell=ppc)
                case = {
                    "key": key,
                    "state": as_int_state(state),
                    "ppc": ppc,
                    "bw": bw,
                    "shape": list(img.shape),
                    "sha256_u8": sha(img),
                }
                # padded float observation: pad to (H+3, W+2) cells -> odd pixel
                # padding whenever ppc is odd, plus the benchmark standard (54, 47).
                for pad in ((puzzle.dimensions[1] + 3, puzzle.dimensions[0] + 2), (54, 47)):
                    if pad == (54, 47) and ppc > 7:
                        continue
                    obs = render_observation_padded(puzzle, state, pad[0], pad[1], ppc, bw)
                    case.setdefault("padded", []).append(
                        {"pad_cells": list(pad), "shape": list(obs.shape), "sha256_f32": sha(obs)}
                    )
                cases.append(case)
    renders["cases"] = cases
    with open(os.path.join(HERE, "reference_renders.json"), "w") as f:
        json.dump(renders, f, indent=1)

    # ---------------------------------------------------------------- gym env
    gym_cases = []
    for name, max_steps, seed in (
        ("multiple_goals", None, 1),
        ("multiple_goals", 7, 2),
        ("trivial", None, 3),
        ("transitive_pushing", 3, 4),
        ("trivial_obstacle", 25, 5),
    ):
        path = os.path.join(tmp, f"{name}.pwp")
        with open(path, "w") as f:
            f.write(fixtures["python3"][name])
        env = PushWorldEnv(path, max_steps=max_steps)
        rng = random.Random(seed)
        _, info = env.reset()
        actions = [rng.randrange(4) for _ in range(60)]
        if name == "multiple_goals" and max_steps is None:
            actions = [1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0] + actions
        out = []
        for a in actions:
            _, reward, terminated, truncated, info = env.step(a)
            out.append(
                [float(reward), bool(terminated), bool(truncated), as_int_state(info["puzzle_state"])]
            )
        gym_cases.append({"puzzle": name, "max_steps": max_steps, "actions": actions, "out": out})
    with open(os.path.join(HERE, "reference_gym.json"), "w") as f:
        json.dump(gym_cases, f)
    print("done")


if __name__ == "__main__":
    main()
