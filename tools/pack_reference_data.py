#!/usr/bin/env python3
"""Packs the reference's *data* (puzzle grids + solution plans) into one bundle.

The benchmark puzzles are inputs, not code: BASELINE.json's configs are quoted on
them (``level1/Road Block.pwp``, "all benchmark/puzzles", ...), and
``/root/reference`` does not exist on the GPU box.  This script reads

    /root/reference/benchmark/puzzles/level{1..4}/*.pwp      (223 puzzles)
    /root/reference/benchmark/puzzles/level0.zip             (15,400 generated level-0 puzzles)
    /root/reference/benchmark/solutions/**/*.yaml            (223 plans)
    /root/reference/python3/test/puzzles/*.pwp               (14 fixtures)
    /root/reference/cpp/test/puzzles/*.pwp                   (15 fixtures)

and writes ``pushworld_b200/data/benchmark.json.gz``, ``pushworld_b200/data/level0.json.gz`` and
``tests/golden/test_puzzles.json``.  Run it in the build container only.
"""

import gzip
import json
import os
import sys

import yaml

REF = os.environ.get("PUSHWORLD_REFERENCE", "/root/reference")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def read(path):
    with open(path, "r") as f:
        return f.read()


def main():
    puzzles = {}
    root = os.path.join(REF, "benchmark", "puzzles")
    for level in sorted(os.listdir(root)):
        level_dir = os.path.join(root, level)
        if not os.path.isdir(level_dir):
            continue
        for name in sorted(os.listdir(level_dir)):
            if name.lower().endswith(".pwp"):
                puzzles[f"{level}/{name[:-4]}"] = read(os.path.join(level_dir, name))

    plans = {}
    sol_root = os.path.join(REF, "benchmark", "solutions")
    for parent, _, files in os.walk(sol_root):
        for name in files:
            if name.endswith(".yaml"):
                doc = yaml.safe_load(read(os.path.join(parent, name)))
                plans[doc["puzzle"]] = {"plan": doc["plan"], "planner": str(doc.get("planner"))}

    by_base = {key.split("/", 1)[1]: key for key in puzzles}
    solutions = {by_base[name]: rec for name, rec in plans.items() if name in by_base}
    missing = sorted(set(puzzles) - set(solutions))
    if missing:
        print("puzzles without a plan:", missing, file=sys.stderr)

    bundle = {
        "source": "google-deepmind/pushworld benchmark/ (Apache-2.0 data)",
        "puzzles": puzzles,
        "solutions": solutions,
    }
    out = os.path.join(REPO, "pushworld_b200", "data", "benchmark.json.gz")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with gzip.GzipFile(out, "wb", mtime=0) as f:
        f.write(json.dumps(bundle, sort_keys=True).encode())
    print(f"{out}: {len(puzzles)} puzzles, {len(solutions)} plans, {os.path.getsize(out)} bytes")

    # the generated "level 0" set (benchmark/puzzles/level0.zip: 7 categories x (2000 train + 200 test))
    import zipfile

    level0 = {}
    with zipfile.ZipFile(os.path.join(root, "level0.zip")) as z:
        for name in sorted(z.namelist()):
            if name.endswith(".pwp"):
                level0[name[:-4]] = z.read(name).decode("utf-8")
    out = os.path.join(REPO, "pushworld_b200", "data", "level0.json.gz")
    with gzip.GzipFile(out, "wb", mtime=0) as f:
        f.write(json.dumps({"source": "google-deepmind/pushworld benchmark/puzzles/level0.zip (Apache-2.0 data)",
                            "puzzles": level0}, sort_keys=True).encode())
    print(f"{out}: {len(level0)} level-0 puzzles, {os.path.getsize(out)} bytes")

    fixtures = {"python3": {}, "cpp": {}}
    for side, rel in (("python3", "python3/test/puzzles"), ("cpp", "cpp/test/puzzles")):
        d = os.path.join(REF, rel)
        for name in sorted(os.listdir(d)):
            if name.endswith(".pwp"):
                fixtures[side][name[:-4]] = read(os.path.join(d, name))
    out = os.path.join(REPO, "tests", "golden", "test_puzzles.json")
    with open(out, "w") as f:
        json.dump(fixtures, f, indent=1, sort_keys=True)
    print(f"{out}: {len(fixtures['python3'])} + {len(fixtures['cpp'])} fixtures")


if __name__ == "__main__":
    main()
