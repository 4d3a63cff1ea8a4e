"""The reachable-state transition-table rollout (pushworld_b200/csrc/pw_table.cu) against the
kernels it replaces (an engine created with PW_NO_TABLE=1: ring / edge-mask boards, general rule)
and the C oracle: same rewards, flags, final states and counters -- also for environments loaded in
states the table does not hold, and for out-of-range actions."""

import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import c_oracle  # noqa: E402


def _engine(text, table):
    from pushworld_b200.engine import Engine

    old = os.environ.pop("PW_NO_TABLE", None)
    try:
        if not table:
            os.environ["PW_NO_TABLE"] = "1"
        return Engine([text])
    finally:
        os.environ.pop("PW_NO_TABLE", None)
        if old is not None:
            os.environ["PW_NO_TABLE"] = old


def _run(engine, xy0, actions, packed):
    from pushworld_b200.engine import record_to_outputs

    T, n = actions.shape
    xy = xy0.clone()
    counters = torch.zeros(4, dtype=torch.int64, device="cuda")
    if packed:
        record = torch.empty((T, n), dtype=torch.uint8, device="cuda")
        engine.rollout_packed(xy, 0, actions, record, counters=counters)
        reward, term, _, _ = record_to_outputs(record)
    else:
        reward = torch.empty((T, n), dtype=torch.float32, device="cuda")
        term = torch.empty((T, n), dtype=torch.uint8, device="cuda")
        engine.rollout(xy, 0, actions, reward=reward, terminated=term, counters=counters)
    return xy, reward, term, counters


@pytest.mark.parametrize("key", ["level1/Road Block", "level1/Obstacle Has Obstacle", "level1/Double Obstacle"])
def test_table_rollout_equals_the_other_kernels_and_the_oracle(key):
    from pushworld_b200 import data

    text = data.puzzle_text(key)
    tab, ref = _engine(text, True), _engine(text, False)
    n, T = 4096, 333
    gen = torch.Generator(device="cuda").manual_seed(41)
    actions = torch.randint(0, 4, (T, n), dtype=torch.uint8, device="cuda", generator=gen)
    xy0 = tab.new_state(n)
    tab.reset(xy0, 0)
    base = _run(ref, xy0, actions, False)
    for packed in (False, True):
        got = _run(tab, xy0, actions, packed)
        for a, b in zip(got, base):
            assert torch.equal(a, b), (key, packed)
    assert int(base[3][0]) == n * T and int(base[3][2]) > 0
    # ... and the C oracle
    oracle = c_oracle.COraclePuzzle(text)
    host = xy0.cpu().numpy()
    nobj = oracle.n
    ox, oy = np.ascontiguousarray(host[0, :nobj]), np.ascontiguousarray(host[1, :nobj])
    o_rew = np.empty((T, n), np.float32)
    o_term = np.empty((T, n), np.uint8)
    oracle.rollout(ox, oy, actions.cpu().numpy(), o_rew, o_term)
    assert np.array_equal(base[1].cpu().numpy(), o_rew) and np.array_equal(base[2].cpu().numpy(), o_term)
    final = base[0].cpu().numpy()
    assert np.array_equal(final[0, :nobj], ox) and np.array_equal(final[1, :nobj], oy)


def test_states_outside_the_table_and_invalid_actions():
    """Environments loaded in arbitrary (here: partly unreachable) states fall back to the general
    rule warp by warp; out-of-range actions are counted no-ops on the table path as everywhere."""
    from pushworld_b200 import data

    text = data.puzzle_text("level1/Road Block")
    tab, ref = _engine(text, True), _engine(text, False)
    n, T = 2048, 200
    gen = torch.Generator(device="cuda").manual_seed(43)
    actions = torch.randint(0, 4, (T, n), dtype=torch.uint8, device="cuda", generator=gen)
    actions[5, 100] = 7  # invalid: no-op
    actions[37, 64:96] = 200
    xy0 = tab.new_state(n)
    tab.reset(xy0, 0)
    # walk a little, then teleport the two movables of every third warp onto other free interior cells
    tab.rollout(xy0, 0, actions[:50].contiguous())
    assert tab.poll_invalid_actions() == 33
    host = xy0.cpu().numpy()
    rng = np.random.default_rng(3)
    for e in range(0, n, 96):
        ax, ay = host[0, 0, e], host[1, 0, e]
        cells = [(x, y) for x in range(1, 7) for y in range(1, 4) if (x, y) != (ax, ay)]
        (x1, y1), (x2, y2) = [cells[i] for i in rng.choice(len(cells), 2, replace=False)]
        host[0, 1, e], host[1, 1, e], host[0, 2, e], host[1, 2, e] = x1, y1, x2, y2
    xy0 = torch.from_numpy(host).cuda()
    base = _run(ref, xy0, actions, False)
    assert ref.poll_invalid_actions() == 33
    for packed in (False, True):
        got = _run(tab, xy0, actions, packed)
        assert tab.poll_invalid_actions() == 33
        for a, b in zip(got, base):
            assert torch.equal(a, b), packed


VARIANTS = {
    "quad": {},  # two-step table, four environments per thread (needs N % 128 == 0, aligned outputs)
    "pairs": {"PW_NO_QUAD": "1"},  # two-step table, one environment per thread
    "compact-1cta": {"PW_NO_PAIRS": "1", "PW_TABLE_CTAS": "1"},  # one-step compact entries
    "compact-2cta": {"PW_NO_PAIRS": "1", "PW_TABLE_CTAS": "2"},
    "wide": {"PW_TABLE_WIDE": "1"},  # one-step 32-bit entries (the form puzzles with several goals use)
}


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("n,T", [(4096, 77), (4096 + 32, 48), (128, 1), (256, 16)])
def test_every_table_kernel_variant(variant, n, T):
    """All table kernels (the default picks the first that fits the launch) against the engine without
    tables: odd / partial action tiles, out-of-range actions inside full and partial tiles, batch sizes
    that are not a multiple of 128 (the four-environments-per-thread kernel hands over to the
    one-environment one), outputs that are not 16-byte aligned."""
    from pushworld_b200 import data

    text = data.puzzle_text("level1/Road Block")
    tab, ref = _engine(text, True), _engine(text, False)
    gen = torch.Generator(device="cuda").manual_seed(1000 + n + T)
    actions = torch.randint(0, 4, (T, n), dtype=torch.uint8, device="cuda", generator=gen)
    bad = 0
    if T >= 16:
        actions[3, 5] = 4
        actions[T - 1, n - 1] = 255
        actions[T // 2, 64:96] = 9
        bad = 34
    xy0 = tab.new_state(n)
    tab.reset(xy0, 0)
    tab.rollout(xy0, 0, torch.randint(0, 4, (40, n), dtype=torch.uint8, device="cuda", generator=gen))  # spread out
    base = _run(ref, xy0, actions, False)
    assert ref.poll_invalid_actions() == bad
    keys = ("PW_NO_QUAD", "PW_NO_PAIRS", "PW_TABLE_CTAS", "PW_TABLE_WIDE")
    saved = {k: os.environ.pop(k, None) for k in keys}
    try:
        os.environ.update(VARIANTS[variant])
        for packed in (False, True):
            got = _run(tab, xy0, actions, packed)
            assert tab.poll_invalid_actions() == bad
            for a, b in zip(got, base):
                assert torch.equal(a, b), (variant, packed)
        # outputs at an odd element offset (not 16-byte aligned): the vector stores must not be used
        T2 = actions.shape[0]
        buf_r = torch.zeros(T2 * n + 1, dtype=torch.float32, device="cuda")
        buf_t = torch.zeros(T2 * n + 1, dtype=torch.uint8, device="cuda")
        reward, term = buf_r[1:].view(T2, n), buf_t[1:].view(T2, n)
        xy = xy0.clone()
        tab.rollout(xy, 0, actions, reward=reward, terminated=term)
        tab.poll_invalid_actions()
        assert torch.equal(reward, base[1]) and torch.equal(term, base[2]) and torch.equal(xy, base[0])
    finally:
        for k in keys:
            os.environ.pop(k, None)
            if saved[k] is not None:
                os.environ[k] = saved[k]


def test_every_benchmark_puzzle_with_and_without_tables():
    """All 223 benchmark puzzles, one engine each: whatever the engine tabulates (13 puzzles have few enough
    reachable states; one-goal ones get the two-step table, the others the one-step form) must roll out
    exactly like the engine without tables -- states numbered terminal-last, several goals, tall grids."""
    from pushworld_b200 import data

    n, T = 256, 48
    gen = torch.Generator(device="cuda").manual_seed(77)
    actions = torch.randint(0, 4, (T, n), dtype=torch.uint8, device="cuda", generator=gen)
    warm = torch.randint(0, 4, (24, n), dtype=torch.uint8, device="cuda", generator=gen)
    for key in data.benchmark_keys():
        text = data.puzzle_text(key)
        tab, ref = _engine(text, True), _engine(text, False)
        xy0 = tab.new_state(n)
        tab.reset(xy0, 0)
        ref.rollout(xy0, 0, warm)  # spread the environments out with the general rule
        base = _run(ref, xy0, actions, False)
        for packed in (False, True):
            got = _run(tab, xy0, actions, packed)
            for a, b in zip(got, base):
                assert torch.equal(a, b), (key, packed)
