#!/usr/bin/env python3
"""Solves a puzzle with the batched GPU breadth-first search (pushworld_b200.search.bfs_solve):
prints a shortest plan in the reference's plan notation (a string over L, R, U, D; compare
benchmark/solutions/*.yaml).

    python scripts/solve_bfs.py "level1/Road Block"      # a bundled benchmark puzzle
    python scripts/solve_bfs.py path/to/puzzle.pwp
"""

import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pushworld_b200 import data  # noqa: E402
from pushworld_b200.search import bfs_solve  # noqa: E402


def main():
    target = sys.argv[1] if len(sys.argv) > 1 else "level1/Road Block"
    max_states = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 25
    if os.path.exists(target):
        with open(target) as f:
            text = f.read()
    else:
        text = data.puzzle_text(target)
    t0 = time.time()
    plan, stats = bfs_solve(text, max_states=max_states, return_stats=True)
    dt = time.time() - t0
    print(f"plan: {plan if plan is not None else 'NO SOLUTION within the state budget'}")
    print(f"length {len(plan) if plan else '-'}, layers {stats['layers']}, generated {stats['generated']}, "
          f"unique {stats['unique']}, {dt:.2f} s")


if __name__ == "__main__":
    main()
