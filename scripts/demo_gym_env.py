#!/usr/bin/env python3
"""Headless counterpart of the reference's python3/scripts/demo_gym_env.py (BASELINE config 1):
creates a PushWorldEnv on one puzzle, resets it and takes 10 random actions -- every step and
observation is computed on the GPU.  The reference shows the frames with matplotlib (TkAgg); here
their shapes, value ranges and the step outputs are printed instead.

    python scripts/demo_gym_env.py [--path PUZZLE.pwp] [--steps 10]
"""

import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pushworld_b200 import data  # noqa: E402
from pushworld_b200.gym_env import PushWorldEnv  # noqa: E402
from pushworld_b200.puzzle import NUM_ACTIONS  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--path", default=None, help="puzzle file or directory (default: level1/Road Block)")
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    path = args.path or os.path.join(data.materialize(), "level1", "Road Block.pwp")
    env = PushWorldEnv(path)
    observation, info = env.reset()
    print(f"reset: observation {observation.shape} {observation.dtype} "
          f"range [{observation.min():.3f}, {observation.max():.3f}], state {info['puzzle_state']}")
    rng = np.random.default_rng(0)
    for i in range(args.steps):
        action = int(rng.integers(NUM_ACTIONS))
        observation, reward, terminated, truncated, info = env.step(action)
        print(f"step {i}: action {'LRUD'[action]} reward {reward:+.2f} terminated {terminated} "
              f"truncated {truncated} state {info['puzzle_state']}")


if __name__ == "__main__":
    main()
