"""oracle/_ref -- the reference's own C++ simulator and novelty heuristic, compiled from the
reference sources (oracle/Makefile target _ref) -- against the vectors the PYTHON reference produced
(tests/golden) and against the C oracle.  Pins, with reference-built code:
  * the C oracle's moved-object masks (what pw_expand / pw_step report; the python reference has no
    such output, the C++ twin does: RelativeState.moved_object_indices, pushworld_puzzle.cc:439-452),
  * the python <-> C++ object-order mapping (puzzle.py:178-179,235-237 vs pushworld_puzzle.cc:274-315),
  * the novelty model the GPU novelty test uses (novelty.cc:30-77).
No GPU needed.  Skipped only if the library is neither prebuilt nor buildable (no reference sources)."""

import numpy as np
import pytest

import c_oracle
import ref_oracle

pytestmark = pytest.mark.skipif(not ref_oracle.available(), reason="oracle/_ref is not built and cannot be built here")


def _walks(golden, key):
    for tag in ("sol", "walk"):
        if golden.has(key, f"{tag}_actions"):
            yield golden.arr(key, f"{tag}_states").astype(np.int64), golden.arr(key, f"{tag}_actions")


def _tuple(state):
    return tuple((int(x), int(y)) for x, y in state)


def test_cpp_twin_reproduces_the_python_reference_on_all_golden_transitions(golden):
    """Every golden transition of every benchmark puzzle through the reference's C++ getNextState:
    same next state (after the order mapping), and its moved_object_indices equal the C oracle's
    moved mask and the set of objects whose position changed."""
    n_transitions = 0
    for key in golden.keys("benchmark/"):
        text = golden.text(key)
        ref = ref_oracle.RefPuzzle(text)
        oracle = c_oracle.COraclePuzzle(text)
        rec = golden.rec(key)
        assert ref.n == rec["n"]
        first = None
        for states, actions in _walks(golden, key):
            if first is None:
                first = _tuple(states[0])
            for t, a in enumerate(actions):
                s0, s1 = _tuple(states[t]), _tuple(states[t + 1])
                got, moved = ref.step(s0, int(a))
                assert got == s1, (key, t)
                changed = sum(1 << k for k in range(ref.n) if s0[k] != s1[k])
                assert moved == changed, (key, t)
                o_next, o_moved = oracle.step(s0, int(a))
                assert o_next == s1 and o_moved == moved, (key, t)
                n_transitions += 1
        assert ref.initial_state == oracle.py.initial_state
        assert ref.goal_state == oracle.py.goal_state
    assert n_transitions > 80000


def test_object_order_mapping_on_the_file_parsing_fixture(golden):
    """python order (A, M4, M1, M0, M2, M3) (test_puzzle.py:204-211) vs C++ order
    (A, M1, M4, M0, M2, M3) (test_pushworld_puzzle.cc:470-477)."""
    key = next(k for k in golden.keys() if k.endswith("file_parsing") and not k.startswith("benchmark/"))
    ref = ref_oracle.RefPuzzle(golden.text(key))
    assert ref.py.object_names == ["a", "m4", "m1", "m0", "m2", "m3"]
    assert ref_oracle.cpp_object_order(ref.py.object_names, ref.py.goal_names) == ["a", "m1", "m4", "m0", "m2", "m3"]
    assert ref.initial_state == ref.py.initial_state


def test_cpp_rollout_equals_c_oracle_rollout(golden):
    """The threaded rollout around the reference's getNextState (the `kind: reference` CPU baseline
    of bench.py) equals the C oracle's on seeded random actions."""
    rng = np.random.default_rng(5)
    for key in ("benchmark/level1/Road Block", "benchmark/level2/Pull Dont Push", "benchmark/level3/Yin Yang"):
        text = golden.text(key)
        ref = ref_oracle.RefPuzzle(text)
        oracle = c_oracle.COraclePuzzle(text)
        n_envs, T = 257, 300
        init = np.array(oracle.py.initial_state, dtype=np.int32)
        actions = rng.integers(0, 4, size=(T, n_envs), dtype=np.uint8)
        outs = []
        for impl in (ref, oracle):
            x = np.ascontiguousarray(np.repeat(init[:, 0:1], n_envs, axis=1))
            y = np.ascontiguousarray(np.repeat(init[:, 1:2], n_envs, axis=1))
            reward = np.empty((T, n_envs), np.float32)
            term = np.empty((T, n_envs), np.uint8)
            assert impl.rollout(x, y, actions, reward, term) == n_envs * T
            outs.append((x, y, reward, term))
        for a, b in zip(*outs):
            assert np.array_equal(a, b), key


class NoveltyModel:
    """cpp/src/heuristics/novelty.cc:30-77 restated with python sets (the model the GPU test of
    pw_novelty compares against; pinned here against the reference's own class)."""

    def __init__(self, n):
        self.singles = [set() for _ in range(n)]
        self.pairs = {}
        self.n = n

    def evaluate(self, state, moved_idx, insert=True):
        novelty = 3
        new_singles, new_pairs = [], []
        for i in moved_idx:
            if state[i] not in self.singles[i]:
                novelty = 1
                new_singles.append((i, state[i]))
            for j in range(self.n):
                if j == i:
                    continue
                a, b = (i, j) if i < j else (j, i)
                key = (a, b, state[a], state[b])
                if key not in self.pairs:
                    if novelty > 2:
                        novelty = 2
                    new_pairs.append(key)
        if insert:
            self.insert(new_singles, new_pairs)
        return novelty, new_singles, new_pairs

    def insert(self, singles, pairs):
        for i, p in singles:
            self.singles[i].add(p)
        for key in pairs:
            self.pairs[key] = True


def test_novelty_model_equals_the_reference_heuristic(golden):
    rng = np.random.default_rng(12)
    for key in ("benchmark/level2/Pull Dont Push", "benchmark/level1/Many Small Tools"):
        text = golden.text(key)
        oracle = c_oracle.COraclePuzzle(text)
        ref_h = ref_oracle.RefNovelty(oracle.n)
        model = NoveltyModel(oracle.n)
        state = oracle.py.initial_state
        seen = set()
        for _ in range(3000):
            nxt, moved = oracle.step(state, int(rng.integers(4)))
            moved_idx = [k for k in range(oracle.n) if (moved >> k) & 1]
            want = ref_h.evaluate(nxt, moved)
            got, _, _ = model.evaluate(nxt, moved_idx)
            assert got == want
            seen.add(want)
            state = nxt if rng.random() < 0.9 else oracle.py.initial_state
        assert seen == {1, 2, 3}
