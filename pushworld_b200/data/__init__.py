"""The benchmark puzzle set (223 puzzles + their known solution plans) bundled as data.

``benchmark.json.gz`` is packed from the reference's ``benchmark/`` directory by
``tools/pack_reference_data.py``; keys are ``"level<k>/<Puzzle Name>"``.  The reference finds
puzzles by walking ``BENCHMARK_PUZZLES_PATH`` (python3/src/pushworld/config.py:27-29);
``materialize()`` recreates that directory layout for path-based APIs.
"""

from __future__ import annotations

import gzip
import json
import os
from functools import lru_cache
from typing import Dict, List, Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
BUNDLE_PATH = os.path.join(_HERE, "benchmark.json.gz")


@lru_cache(maxsize=1)
def _bundle() -> dict:
    with gzip.open(BUNDLE_PATH, "rb") as f:
        return json.loads(f.read().decode("utf-8"))


def benchmark_keys() -> List[str]:
    """All puzzle keys sorted by relative path (the order BASELINE config 3 uses)."""
    return sorted(_bundle()["puzzles"].keys())


def puzzle_text(key: str) -> str:
    return _bundle()["puzzles"][key]


def solution_plan(key: str) -> str:
    return _bundle()["solutions"][key]["plan"]


LEVEL0_PATH = os.path.join(_HERE, "level0.json.gz")


@lru_cache(maxsize=1)
def _level0() -> dict:
    with gzip.open(LEVEL0_PATH, "rb") as f:
        return json.loads(f.read().decode("utf-8"))["puzzles"]


def level0_keys(category: Optional[str] = None, split: Optional[str] = None) -> List[str]:
    """Keys (``level0/<category>/<train|test>/<name>``, sorted) of the reference's 15,400 generated
    level-0 puzzles (benchmark/puzzles/level0.zip, made by python3/src/pushworld/generate.py:136-259);
    categories: all, base, goals, obstacles, shapes, size, walls."""
    keys = sorted(_level0().keys())
    if category is not None:
        keys = [k for k in keys if k.split("/")[1] == category]
    if split is not None:
        keys = [k for k in keys if k.split("/")[2] == split]
    return keys


def level0_text(key: str) -> str:
    return _level0()[key]


def benchmark_texts() -> Dict[str, str]:
    return dict(_bundle()["puzzles"])


@lru_cache(maxsize=1)
def max_puzzle_dimensions():
    """(max height, max width) over the benchmark, incl. the wall border
    (utils/env_utils.py:25-41)."""
    max_h = max_w = 0
    for text in _bundle()["puzzles"].values():
        lines = text.split("\n")
        if lines and lines[-1] == "":
            lines.pop()
        max_h = max(max_h, len(lines) + 2)
        max_w = max(max_w, len(lines[0].strip().split()) + 2)
    return max_h, max_w


def materialize(directory: Optional[str] = None) -> str:
    """Writes the bundle out as ``<directory>/puzzles/level<k>/<name>.pwp`` (idempotent) and
    returns ``<directory>/puzzles``."""
    if directory is None:
        directory = os.path.join(_HERE, "_materialized")
    root = os.path.join(directory, "puzzles")
    for key, text in _bundle()["puzzles"].items():
        path = os.path.join(root, key + ".pwp")
        if not os.path.exists(path):
            os.makedirs(os.path.dirname(path), exist_ok=True)
            with open(path, "w") as f:
                f.write(text)
    return root
