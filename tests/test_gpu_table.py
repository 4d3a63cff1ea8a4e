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
