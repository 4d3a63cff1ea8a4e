"""The Python oracle (oracle/pushworld_oracle.py) against vectors produced by the real
reference: this is what pins the oracle (no GPU needed)."""

import hashlib
import json
import os

import numpy as np
import pytest

import pushworld_oracle as po
from tests.conftest import GOLDEN


def _states(arr):
    return [tuple((int(x), int(y)) for x, y in s) for s in arr]


def test_parse_matches_reference(golden):
    for key in golden.keys():
        rec = golden.rec(key)
        p = po.OraclePuzzle(golden.text(key))
        assert p.dimensions == (rec["width"], rec["height"]), key
        assert p.num_objects == rec["n"] and p.num_goals == rec["g"], key
        assert list(map(list, p.initial_state)) == golden.arr(key, "init").tolist(), key
        assert list(map(list, p.goal_state)) == golden.arr(key, "goal").tolist(), key


@pytest.mark.parametrize("tag", ["sol", "walk"])
def test_transitions_match_reference(golden, tag):
    checked = 0
    for key in golden.keys():
        if not golden.has(key, f"{tag}_actions"):
            continue
        p = po.OraclePuzzle(golden.text(key))
        actions = golden.arr(key, f"{tag}_actions")
        states = _states(golden.arr(key, f"{tag}_states"))
        achieved = golden.arr(key, f"{tag}_achieved")
        isgoal = golden.arr(key, f"{tag}_isgoal")
        for t, a in enumerate(actions):
            assert p.get_next_state(states[t], int(a)) == states[t + 1], (key, t)
            assert p.count_achieved_goals(states[t + 1]) == achieved[t + 1]
            assert p.is_goal_state(states[t + 1]) == bool(isgoal[t + 1])
            checked += 1
        if tag == "sol":
            assert p.is_valid_plan([int(a) for a in actions]), key
    assert checked > (18000 if tag == "sol" else 60000)


def test_render_matches_reference(golden):
    with open(os.path.join(GOLDEN, "reference_renders.json")) as f:
        renders = json.load(f)
    trivial = po.OraclePuzzle(golden.text("python3/trivial"))
    frames = trivial.render_plan([po.RIGHT, po.DOWN, po.RIGHT, po.UP])
    rec = renders["trivial_plan_RDRU"]
    assert list(frames[0].shape) == rec["shape"]
    # the very hashes python3/test/test_puzzle.py:264-270 asserts
    assert [hash(tuple(img.flat)) for img in frames] == [
        8141256401900123811,
        1770142108181252064,
        4744825492003518882,
        -7463149466192975143,
        -8235536721686713717,
    ]
    assert [hash(tuple(img.flat)) for img in frames] == rec["python_tuple_hash"]
    puzzles = {}
    for case in renders["cases"]:
        if case["ppc"] > 12 and case["shape"][0] > 400:
            continue  # keep the CPU suite short; the GPU suite covers every case
        p = puzzles.setdefault(case["key"], po.OraclePuzzle(golden.text(case["key"])))
        state = tuple(map(tuple, case["state"]))
        img = p.render(state, case["bw"], case["ppc"])
        assert list(img.shape) == case["shape"]
        assert hashlib.sha256(img.tobytes()).hexdigest() == case["sha256_u8"], case["key"]
        for pad in case.get("padded", []):
            if pad["pad_cells"] == [54, 47] and case["ppc"] > 3:
                continue
            obs = po.render_observation_padded(
                p, state, pad["pad_cells"][0], pad["pad_cells"][1], case["ppc"], case["bw"]
            )
            assert list(obs.shape) == pad["shape"]
            assert hashlib.sha256(obs.tobytes()).hexdigest() == pad["sha256_f32"]


def test_env_logic_matches_reference_gym(golden):
    with open(os.path.join(GOLDEN, "reference_gym.json")) as f:
        cases = json.load(f)
    for case in cases:
        p = po.OraclePuzzle(golden.text("python3/" + case["puzzle"]))
        env = po.OracleEnv([p], max_steps=case["max_steps"], render=False)
        env.reset()
        for a, (reward, term, trunc, state) in zip(case["actions"], case["out"]):
            _, r, te, tr, info = env.step(a)
            assert (r, te, tr) == (reward, term, trunc)
            assert list(map(list, info["puzzle_state"])) == state


def test_reference_known_answers():
    """python3/test/test_puzzle.py:69-98 restated against the oracle."""
    from tests import conftest

    g = conftest.Golden()
    p = po.OraclePuzzle(g.text("python3/pushing"))
    s = p.get_next_state(p.initial_state, po.DOWN)
    assert s == ((1, 2), (2, 1))
    s = p.get_next_state(p.initial_state, po.RIGHT)
    assert s == ((2, 1), (3, 1))
    s = p.get_next_state(s, po.RIGHT)
    assert s == ((3, 1), (4, 1))
    assert p.get_next_state(s, po.RIGHT) == ((3, 1), (4, 1))
    p = po.OraclePuzzle(g.text("python3/file_parsing"))
    assert p.dimensions == (12, 18)
    assert p.goal_state == ((6, 5), (3, 4))
    assert p.initial_state == ((1, 12), (6, 14), (1, 3), (4, 1), (2, 7), (3, 8))
    with pytest.raises(ValueError):
        po.OraclePuzzle(". .\n. . .\n")
    with pytest.raises(ValueError):
        po.OraclePuzzle(". m0\n. .\n")
    with pytest.raises(AssertionError):
        po.OraclePuzzle("a g1\n. m0\n")


def test_c_oracle_matches_reference(golden):
    """oracle/pw_oracle.c (table-driven, the reference's own algorithm) on every golden
    transition, plus the batched rollout entry point used as the CPU baseline."""
    import c_oracle

    checked = 0
    for key in golden.keys():
        cp = c_oracle.COraclePuzzle(golden.text(key))
        for tag in ("sol", "walk"):
            if not golden.has(key, f"{tag}_actions"):
                continue
            actions = golden.arr(key, f"{tag}_actions")
            states = _states(golden.arr(key, f"{tag}_states"))
            achieved = golden.arr(key, f"{tag}_achieved")
            for t, a in enumerate(actions):
                nxt, moved = cp.step(states[t], int(a))
                assert nxt == states[t + 1], (key, tag, t)
                changed = sum(1 << k for k in range(cp.n) if states[t][k] != states[t + 1][k])
                assert moved == changed
                checked += 1
            assert cp.achieved(states[-1]) == achieved[-1]
        if golden.has(key, "walk_actions"):
            acts = np.ascontiguousarray(np.repeat(golden.arr(key, "walk_actions")[:, None], 3, axis=1))
            init = golden.arr(key, "walk_states").astype(np.int32)[0]
            x = np.ascontiguousarray(np.repeat(init[:, 0:1], 3, axis=1))
            y = np.ascontiguousarray(np.repeat(init[:, 1:2], 3, axis=1))
            reward = np.zeros(acts.shape, np.float32)
            term = np.zeros(acts.shape, np.uint8)
            cp.rollout(x, y, acts, reward, term, threads=2)
            final = golden.arr(key, "walk_states").astype(np.int32)[-1]
            assert (x == final[:, 0:1]).all() and (y == final[:, 1:2]).all(), key
            assert np.array_equal(term[:, 0].astype(bool), golden.arr(key, "walk_isgoal")[1:])
    assert checked > 85000
