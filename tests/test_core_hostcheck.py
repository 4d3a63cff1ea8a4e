"""The kernels' transition rule (pw_core.cuh, compiled for the CPU by tests/hostcheck) and
the host table compiler against the reference's golden vectors.  No GPU needed: this is
what lets kernel logic be debugged before spending GPU time; the GPU suite repeats the
same comparisons through the C-ABI."""

import numpy as np
import pytest

from pushworld_b200 import tables
from tests import hostcheck


def _check(golden, key, tag):
    text = golden.text(key)
    parsed = tables.parse_puzzle(text)
    rec = golden.rec(key)
    assert (parsed.width, parsed.height) == (rec["width"], rec["height"])
    assert list(map(list, parsed.initial_state)) == golden.arr(key, "init").tolist()
    assert list(map(list, parsed.goal_state)) == golden.arr(key, "goal").tolist()
    blob = tables.build_blob(parsed)
    actions = golden.arr(key, f"{tag}_actions")
    ref = golden.arr(key, f"{tag}_states").astype(np.int32)
    states, moved, ach = hostcheck.walk(blob, parsed.initial_state, actions)
    assert np.array_equal(states, ref[1:]), key
    # the table-driven rule (pw_fast.cuh) must agree with the geometric rule transition by transition
    states_t, moved_t, ach_t, table_bytes = hostcheck.walk_tab(blob, parsed.initial_state, actions)
    assert np.array_equal(states_t, states) and np.array_equal(moved_t, moved) and np.array_equal(ach_t, ach), key
    assert table_bytes < 64 * 1024, (key, table_bytes)
    assert np.array_equal(ach, golden.arr(key, f"{tag}_achieved")[1:].astype(np.int32)), key
    # the moved mask is exactly the set of objects whose position changed
    changed = (ref[1:] != ref[:-1]).any(axis=2)
    bits = (moved[:, None] >> np.arange(ref.shape[1])[None, :]) & 1
    assert np.array_equal(bits.astype(bool), changed), key
    return len(actions)


@pytest.mark.parametrize("tag", ["sol", "walk"])
def test_core_matches_reference_vectors(golden, tag):
    total = 0
    for key in golden.keys():
        if golden.has(key, f"{tag}_actions"):
            total += _check(golden, key, tag)
    assert total > (18000 if tag == "sol" else 60000)


def test_cell_code_matches_reference_renders(golden):
    """pw_render_core.cuh::cell_code (the kernels' per-cell closed form of render) against
    sha256 of the real reference's PushWorldPuzzle.render output."""
    import hashlib
    import json
    import os

    from tests.conftest import GOLDEN

    with open(os.path.join(GOLDEN, "reference_renders.json")) as f:
        renders = json.load(f)
    blobs = {}
    for case in renders["cases"]:
        key = case["key"]
        if key not in blobs:
            blobs[key] = tables.compile_text(golden.text(key))
        img = hostcheck.render_u8(blobs[key], case["state"], case["ppc"], case["bw"])
        assert list(img.shape) == case["shape"]
        assert hashlib.sha256(img.tobytes()).hexdigest() == case["sha256_u8"], (key, case["ppc"])
    assert len(renders["cases"]) > 100


def test_bitboard_rule_matches_reference_vectors(golden):
    """pw_board.cuh (whole-grid uint64 boards, used when width*height <= 64) on every golden
    transition of every eligible puzzle, and against the general rule on random walks."""
    eligible = n32 = nring = 0
    for key in golden.keys():
        rec = golden.rec(key)
        bits = 1 + (rec["height"] - 2) * (rec["width"] - 1)
        if bits > 64 or rec["n"] > 8:
            continue
        parsed = tables.parse_puzzle(golden.text(key))
        blob = tables.build_blob(parsed)
        for tag in ("sol", "walk"):
            if not golden.has(key, f"{tag}_actions"):
                continue
            actions = golden.arr(key, f"{tag}_actions")
            ref = golden.arr(key, f"{tag}_states").astype(np.int32)
            regular = np.frombuffer(blob, np.uint32)[3] & tables.FLAG_REGULAR
            outs = [hostcheck.board_walk(blob, parsed.initial_state, actions, 64)]
            if bits <= 32:
                outs.append(hostcheck.board_walk(blob, parsed.initial_state, actions, 32))
                n32 += 1
            if bits + rec["width"] - 1 <= 32:  # ring boards (word_bits 31 selects them)
                outs.append(hostcheck.board_walk(blob, parsed.initial_state, actions, 31))
                nring += 1
            if outs[0] is None:
                assert not regular, key
                continue
            for other in outs[1:]:
                assert all(np.array_equal(a, b) for a, b in zip(outs[0], other)), key
            states, moved, ach = outs[0]
            assert np.array_equal(states, ref[1:]), key
            assert np.array_equal(ach, golden.arr(key, f"{tag}_achieved")[1:].astype(np.int32)), key
            changed = (ref[1:] != ref[:-1]).any(axis=2)
            mbits = (moved[:, None] >> np.arange(ref.shape[1])[None, :]) & 1
            assert np.array_equal(mbits.astype(bool), changed), key
        eligible += 1
    assert eligible >= 10 and n32 >= 5 and nring >= 2, (eligible, n32, nring)


def test_bitboard_rule_equals_general_rule_on_random_small_puzzles():
    """Random small puzzles (multi-cell shapes, chains, agent walls): bitboard == general rule,
    including invalid actions."""
    rng = np.random.default_rng(2026)
    checked = ring_checked = 0
    for trial in range(300):
        w, h = int(rng.integers(3, 7)), int(rng.integers(2, 7))
        bits = 1 + h * (w + 1)
        if bits > 64:
            continue
        grid = [["." for _ in range(w)] for _ in range(h)]
        names = ["a"] + [f"m{k}" for k in range(int(rng.integers(1, 6)))]
        for name in names:
            for _ in range(int(rng.integers(1, 4))):  # a few cells each, maybe disconnected
                x, y = int(rng.integers(w)), int(rng.integers(h))
                if grid[y][x] == ".":
                    grid[y][x] = name
        for name in ("w", "aw", "g0"):
            x, y = int(rng.integers(w)), int(rng.integers(h))
            if grid[y][x] == ".":
                grid[y][x] = name
        text = "\n".join(" ".join(row) for row in grid) + "\n"
        try:
            parsed = tables.parse_puzzle(text)
        except (ValueError, AssertionError):
            continue
        blob = tables.build_blob(parsed)
        actions = rng.integers(0, 5, size=200).astype(np.uint8)  # 4 = invalid -> no-op
        ref = hostcheck.walk(blob, parsed.initial_state, actions)
        for word_bits in (31, 32, 64):
            if bits + (w + 1 if word_bits == 31 else 0) > max(word_bits, 32):
                continue
            out = hostcheck.board_walk(blob, parsed.initial_state, actions, word_bits)
            if out is None:
                continue
            for got, want in zip(out, ref):
                assert np.array_equal(got, want), text
            checked += 1
            ring_checked += word_bits == 31
    assert checked > 100 and ring_checked > 30, (checked, ring_checked)


def test_general_rule_matches_oracle_on_irregular_puzzles():
    """Random puzzles with overlapping objects ('+' cells), movables inside walls and agents on
    agent walls: the kernels' general rule against the Python oracle, state by state (objects
    may leave the grid here)."""
    import pushworld_oracle as po

    rng = np.random.default_rng(77)
    names_pool = ["a", "m0", "m1", "m2", "m3", "w", "aw", "g0", "g1"]
    checked = 0
    for trial in range(120):
        w, h = int(rng.integers(2, 8)), int(rng.integers(2, 7))
        grid = [[[] for _ in range(w)] for _ in range(h)]
        for name in names_pool:
            for _ in range(int(rng.integers(0, 4)) + (1 if name == "a" else 0)):
                x, y = int(rng.integers(w)), int(rng.integers(h))
                if name not in grid[y][x]:
                    grid[y][x].append(name)
        text = "\n".join(" ".join("+".join(c) if c else "." for c in row) for row in grid) + "\n"
        try:
            oracle = po.OraclePuzzle(text)
        except (ValueError, AssertionError):
            continue
        blob = tables.compile_text(text)
        actions = rng.integers(0, 4, size=120).astype(np.uint8)
        states, moved, ach = hostcheck.walk(blob, oracle.initial_state, actions)
        states_t, moved_t, ach_t, _ = hostcheck.walk_tab(blob, oracle.initial_state, actions)
        assert np.array_equal(states_t, states) and np.array_equal(moved_t, moved) and np.array_equal(ach_t, ach), text
        state = oracle.initial_state
        for t, a in enumerate(actions):
            state, moved_idx = oracle.get_next_state_and_moved(state, int(a))
            got = tuple((int(x), int(y)) for x, y in states[t])
            assert got == state, (text, t)
            assert int(moved[t]) == sum(1 << k for k in moved_idx), (text, t)
            assert int(ach[t]) == oracle.count_achieved_goals(state)
        checked += 1
    assert checked > 60
