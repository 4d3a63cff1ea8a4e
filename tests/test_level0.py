"""The reference's generated level-0 puzzles (benchmark/puzzles/level0.zip, shipped as
pushworld_b200/data/level0.json.gz): both oracles against vectors the REAL reference produced for a
fixed sample of 420 of the 15,400 puzzles (tests/golden/make_golden_level0.py).  No GPU."""

import json
import os

import numpy as np
import pytest

import c_oracle
import pushworld_oracle as po
from tests.conftest import GOLDEN


@pytest.fixture(scope="module")
def level0():
    npz = np.load(os.path.join(GOLDEN, "level0_vectors.npz"))
    index = json.loads(bytes(npz["index_json"]).decode())
    return npz, index


def test_bundle_is_complete():
    from pushworld_b200 import data

    keys = data.level0_keys()
    assert len(keys) == 15400
    cats = {k.split("/")[1] for k in keys}
    assert cats == {"all", "base", "goals", "obstacles", "shapes", "size", "walls"}
    assert len(data.level0_keys("walls", "test")) == 200 and len(data.level0_keys("all", "train")) == 2000


def test_oracles_match_the_reference_on_level0(level0):
    from pushworld_b200 import data

    npz, index = level0
    assert len(index) == 420
    n_steps = 0
    for pi, rec in enumerate(index):
        text = data.level0_text(rec["key"])
        py = po.OraclePuzzle(text)
        co = c_oracle.COraclePuzzle(text)
        assert (py.width, py.height, py.num_objects, py.num_goals) == (rec["width"], rec["height"], rec["n"], rec["g"])
        states = npz[f"p{pi}_states"].astype(np.int64)
        actions = npz[f"p{pi}_actions"]
        ach = npz[f"p{pi}_achieved"]
        isg = npz[f"p{pi}_isgoal"]
        assert tuple(map(tuple, states[0])) == py.initial_state
        assert tuple(map(tuple, npz[f"p{pi}_goal"].astype(np.int64))) == py.goal_state
        for t, a in enumerate(actions):
            s0 = tuple((int(x), int(y)) for x, y in states[t])
            s1 = tuple((int(x), int(y)) for x, y in states[t + 1])
            assert py.get_next_state(s0, int(a)) == s1, (rec["key"], t)
            assert co.step(s0, int(a))[0] == s1, (rec["key"], t)
            assert py.count_achieved_goals(s1) == ach[t + 1] and py.is_goal_state(s1) == bool(isg[t + 1])
            n_steps += 1
    assert n_steps == 420 * 200
