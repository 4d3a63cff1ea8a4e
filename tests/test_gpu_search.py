"""GPU breadth-first search built on pw_expand_indexed + pw_visited_insert: exact
de-duplication and shortest plans, checked against a CPU BFS over the oracle."""

from collections import deque

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import pushworld_oracle as po  # noqa: E402


def _cpu_bfs(oracle, limit=200000):
    start = oracle.initial_state
    if oracle.is_goal_state(start):
        return 0, 1
    seen = {start}
    queue = deque([(start, 0)])
    while queue:
        state, depth = queue.popleft()
        for a in range(4):
            nxt = oracle.get_next_state(state, a)
            if nxt in seen:
                continue
            if oracle.is_goal_state(nxt):
                return depth + 1, len(seen) + 1
            seen.add(nxt)
            queue.append((nxt, depth + 1))
            if len(seen) > limit:
                return None, len(seen)
    return None, len(seen)


@pytest.mark.parametrize("key", [
    "python3/trivial", "python3/trivial_obstacle", "python3/transitive_pushing", "python3/pushing",
    "python3/multiple_goals", "python3/trivial_tool", "python3/necessary_transitive_pushing1",
    "python3/shortest_path_tool", "benchmark/level1/Road Block",
])
def test_bfs_finds_shortest_valid_plans(golden, key):
    from pushworld_b200.search import bfs_solve

    text = golden.text(key)
    oracle = po.OraclePuzzle(text)
    want_len, _ = _cpu_bfs(oracle)
    plan, stats = bfs_solve(text, max_states=1 << 21, return_stats=True)
    if want_len is None:
        assert plan is None
        return
    assert plan is not None and len(plan) == want_len, (key, plan, want_len)
    assert oracle.is_valid_plan([po.ACTION_FROM_CHAR[c] for c in plan])
    assert stats["unique"] <= stats["generated"]


def test_visited_insert_is_exact(golden):
    """Every state of a random walk inserted twice (and shuffled duplicates inside one call):
    exactly the distinct states are reported new."""
    from pushworld_b200.engine import Engine

    text = golden.text("benchmark/level2/Pull Dont Push")
    engine = Engine([text])
    oracle = po.OraclePuzzle(text)
    rng = np.random.default_rng(5)
    state, states = oracle.initial_state, []
    for _ in range(3000):
        state = oracle.get_next_state(state, int(rng.integers(4)))
        states.append(state)
    distinct = len(set(states))
    dup = states + states[::-1] + states[::7]
    arena = engine.state_from_tuples(dup)
    table = torch.zeros(1 << 15, dtype=torch.int64, device="cuda")
    is_new = torch.empty(len(dup), dtype=torch.uint8, device="cuda")
    engine.visited_insert(arena, 0, 0, len(dup), table, is_new)
    assert int(is_new.sum().item()) == distinct
    # what was reported new is one representative per distinct state
    new_states = {dup[i] for i in torch.nonzero(is_new).flatten().tolist()}
    assert new_states == set(states)
    # a second pass finds nothing new
    engine.visited_insert(arena, 0, 0, len(dup), table, is_new)
    assert int(is_new.sum().item()) == 0
    assert engine.poll_status() == (0, 0)
    # a table that is too small reports overflow instead of hanging
    tiny = torch.zeros(64, dtype=torch.int64, device="cuda")
    engine.visited_insert(arena, 0, 0, len(dup), tiny, is_new)
    assert engine.poll_status()[1] > 0


from tests.test_ref_oracle import NoveltyModel as _NoveltyModel  # noqa: E402  (pinned against oracle/_ref there)


def test_novelty_sequential_and_batched(golden):
    """pw_novelty with one state per call equals the reference's sequential algorithm; with a
    batch per call it equals 'judge the whole batch against the previous table, then insert'."""
    from pushworld_b200.engine import Engine

    text = golden.text("benchmark/level2/Pull Dont Push")
    engine = Engine([text])
    oracle = po.OraclePuzzle(text)
    rng = np.random.default_rng(12)
    # a stream of (state, moved objects) as a search would produce them
    stream = []
    state = oracle.initial_state
    for _ in range(1200):
        nxt, moved_idx = oracle.get_next_state_and_moved(state, int(rng.integers(4)))
        stream.append((nxt, moved_idx))
        state = nxt if rng.random() < 0.9 else oracle.initial_state
    arena = engine.state_from_tuples([s for s, _ in stream])
    moved = torch.tensor([sum(1 << k for k in m) for _, m in stream], dtype=torch.int32, device="cuda")

    # sequential: one state per call
    seen = engine.novelty_table(0)
    model = _NoveltyModel(oracle.num_objects)
    import ref_oracle

    # the reference's own NoveltyHeuristic (cpp/src/heuristics/novelty.cc, built into oracle/_ref)
    ref_h = ref_oracle.RefNovelty(oracle.num_objects) if ref_oracle.available() else None
    for t in range(400):
        got = int(engine.novelty(arena, 0, t, 1, moved[t:t + 1], seen).item())
        want, _, _ = model.evaluate(*stream[t])
        assert got == want, t
        if ref_h is not None:
            assert got == ref_h.evaluate(stream[t][0], sum(1 << k for k in stream[t][1])), t

    # batched: whole batches judged against the table before the call
    seen = engine.novelty_table(0)
    model = _NoveltyModel(oracle.num_objects)
    for lo in range(0, len(stream), 96):
        hi = min(len(stream), lo + 96)
        got = engine.novelty(arena, 0, lo, hi - lo, moved[lo:hi].contiguous(), seen).cpu().numpy()
        want, pending = [], []
        for s, m in stream[lo:hi]:
            nov, ns, nps = model.evaluate(s, m, insert=False)
            want.append(nov)
            pending.append((ns, nps))
        for ns, nps in pending:
            model.insert(ns, nps)
        assert got.tolist() == want, lo
    assert set(np.unique(got).tolist()) <= {1, 2, 3}
