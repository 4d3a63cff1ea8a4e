import gzip
import json
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(REPO, "tests", "golden")
for p in (REPO, os.path.join(REPO, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


class Golden:
    """Vectors produced by the real reference (tests/golden/make_golden.py)."""

    def __init__(self):
        self.npz = np.load(os.path.join(GOLDEN, "reference_vectors.npz"))
        self.index = json.loads(bytes(self.npz["index_json"]).decode())
        with gzip.open(os.path.join(REPO, "pushworld_b200", "data", "benchmark.json.gz")) as f:
            self.bundle = json.loads(f.read().decode())
        with open(os.path.join(GOLDEN, "test_puzzles.json")) as f:
            self.fixtures = json.load(f)
        self.by_key = {rec["key"]: i for i, rec in enumerate(self.index)}

    def keys(self, prefix=""):
        return [rec["key"] for rec in self.index if rec["key"].startswith(prefix)]

    def text(self, key):
        group, name = key.split("/", 1)
        if group == "benchmark":
            return self.bundle["puzzles"][name]
        return self.fixtures[group][name]

    def rec(self, key):
        return self.index[self.by_key[key]]

    def arr(self, key, name):
        return self.npz[f"p{self.by_key[key]}_{name}"]

    def has(self, key, name):
        return f"p{self.by_key[key]}_{name}" in self.npz.files

    def plan(self, key):
        assert key.startswith("benchmark/")
        return self.bundle["solutions"][key[len("benchmark/"):]]["plan"]


@pytest.fixture(scope="session")
def golden():
    return Golden()
