"""Parity against the C oracle (oracle/pw_oracle.c, OpenMP) at BASELINE.json's FULL sizes:

* config 2: ALL 2^18 `Road Block` environments x 1024 steps (2.7e8 transitions): per-step reward,
  terminated and the final state, for the plain outputs and for the packed per-step records;
* config 3: 2^18 environments over all 223 benchmark puzzles x 256 steps, grouped by puzzle and
  in BASELINE's literal `e mod 223` order;
* config 4: 1,000,000 frontier states -> 4,000,000 successors + moved masks + goal flags;
* config 5: 2^16 environments, step + render: every frame equals the oracle's frame of its state.
"""

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import c_oracle  # noqa: E402
import pushworld_oracle as po  # noqa: E402
from tests.oracle_batch import OracleSet  # noqa: E402


def _actions(T, n, seed):
    gen = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randint(0, 4, (T, n), dtype=torch.uint8, device="cuda", generator=gen)


def test_config2_every_column_matches_the_c_oracle():
    from pushworld_b200 import data
    from pushworld_b200.engine import Engine, record_to_outputs

    text = data.puzzle_text("level1/Road Block")
    engine = Engine([text])
    n, T = 1 << 18, 1024
    actions = _actions(T, n, 1234)
    xy = engine.new_state(n)
    engine.reset(xy, 0)
    reward = torch.empty((T, n), dtype=torch.float32, device="cuda")
    term = torch.empty((T, n), dtype=torch.uint8, device="cuda")
    engine.rollout(xy, 0, actions, reward=reward, terminated=term)
    # the packed-record launch of the same rollout
    xy_p = engine.new_state(n)
    engine.reset(xy_p, 0)
    record = torch.empty((T, n), dtype=torch.uint8, device="cuda")
    engine.rollout_packed(xy_p, 0, actions, record)
    r_p, t_p, tr_p, _ = record_to_outputs(record)
    assert torch.equal(r_p, reward) and torch.equal(t_p, term) and not bool(tr_p.any())
    assert torch.equal(xy_p, xy)
    del r_p, t_p, tr_p, record, xy_p

    oracle = c_oracle.COraclePuzzle(text)
    init = np.array(oracle.py.initial_state, dtype=np.int32)
    ox = np.ascontiguousarray(np.repeat(init[:, 0:1], n, axis=1))
    oy = np.ascontiguousarray(np.repeat(init[:, 1:2], n, axis=1))
    o_rew = np.empty((T, n), np.float32)
    o_term = np.empty((T, n), np.uint8)
    a_host = actions.cpu().numpy()
    done = oracle.rollout(ox, oy, a_host, o_rew, o_term)
    assert done == n * T
    assert np.array_equal(reward.cpu().numpy(), o_rew)
    assert np.array_equal(term.cpu().numpy(), o_term)
    got = xy.cpu().numpy()
    assert np.array_equal(got[0], ox) and np.array_equal(got[1], oy)
    assert o_term.any() and (o_rew == np.float32(10.0)).any()  # the sample does reach the goal


@pytest.mark.parametrize("order", ["grouped", "interleaved"])
def test_config3_mixed_batch_matches_the_c_oracle(order):
    from pushworld_b200 import data
    from pushworld_b200.engine import Engine, record_to_outputs

    keys = data.benchmark_keys()
    texts = [data.puzzle_text(k) for k in keys]
    engine = Engine(texts)
    n, T = 1 << 18, 256
    ids = np.arange(n, dtype=np.int64) % len(keys)
    if order == "grouped":
        ids = np.sort(ids, kind="stable")
    pid = torch.from_numpy(ids.astype(np.int32)).cuda()
    actions = _actions(T, n, 1235)
    xy = engine.new_state(n)
    engine.reset(xy, pid)
    reward = torch.empty((T, n), dtype=torch.float32, device="cuda")
    term = torch.empty((T, n), dtype=torch.uint8, device="cuda")
    counters = torch.zeros(4, dtype=torch.int64, device="cuda")
    engine.rollout(xy, pid, actions, reward=reward, terminated=term, counters=counters)
    xy_p = engine.new_state(n)
    engine.reset(xy_p, pid)
    record = torch.empty((T, n), dtype=torch.uint8, device="cuda")
    engine.rollout_packed(xy_p, pid, actions, record)
    r_p, t_p, _, _ = record_to_outputs(record)
    assert torch.equal(r_p, reward) and torch.equal(t_p, term) and torch.equal(xy_p, xy)
    del r_p, t_p, record, xy_p

    oracles = OracleSet(texts)
    ox, oy, o_rew, o_term = oracles.rollout(ids, actions.cpu().numpy(), engine.n_max)
    got = xy.cpu().numpy()
    assert np.array_equal(got[0], ox) and np.array_equal(got[1], oy)
    assert np.array_equal(reward.cpu().numpy(), o_rew)
    assert np.array_equal(term.cpu().numpy(), o_term)
    assert int(counters[0]) == n * T and int(counters[1]) == int(o_term.sum(dtype=np.int64))
    assert engine.poll_invalid_actions() == 0


def test_config4_expansion_matches_the_c_oracle():
    from pushworld_b200 import data
    from pushworld_b200.engine import Engine

    text = data.puzzle_text("level2/Pull Dont Push")
    engine = Engine([text])
    n = 1_000_000
    xy = engine.new_state(n)
    engine.reset(xy, 0)
    engine.rollout(xy, 0, _actions(64, n, 1236))  # frontier = 64-step seeded random walks
    xy_out, moved, goal = engine.expand(xy, 0)
    oracle = c_oracle.COraclePuzzle(text)
    host = xy.cpu().numpy()
    ox, oy, o_moved, o_goal = oracle.expand(np.ascontiguousarray(host[0]), np.ascontiguousarray(host[1]))
    got = xy_out.cpu().numpy()
    assert np.array_equal(got[0], ox) and np.array_equal(got[1], oy)
    assert np.array_equal(moved.cpu().numpy().astype(np.uint32), o_moved)
    assert np.array_equal(goal.cpu().numpy(), o_goal)
    assert len(np.unique(o_moved)) > 4  # the frontier does contain pushes of several objects


@pytest.mark.parametrize("dtype", ["float32", "uint8"])
def test_config5_every_frame_matches_the_oracle(dtype):
    from pushworld_b200 import data
    from pushworld_b200.engine import Engine

    text = data.puzzle_text("level1/Road Block")
    engine = Engine([text])
    oracle = po.OraclePuzzle(text)
    n = 1 << 16
    xy = engine.new_state(n)
    engine.reset(xy, 0)
    engine.rollout(xy, 0, _actions(48, n, 1237))
    before = xy.clone()
    act = _actions(1, n, 1238)[0]
    obs, out = engine.step_render(xy, 0, act, dtype=dtype)
    assert tuple(obs.shape) == (n, 100, 160, 3)
    # the step itself, against the C oracle
    c_puzzle = c_oracle.COraclePuzzle(text)
    hb = before.cpu().numpy()
    ox, oy = np.ascontiguousarray(hb[0]), np.ascontiguousarray(hb[1])
    o_rew = np.empty((1, n), np.float32)
    o_term = np.empty((1, n), np.uint8)
    c_puzzle.rollout(ox, oy, act.cpu().numpy()[None, :].copy(), o_rew, o_term)
    state = xy.cpu().numpy()
    assert np.array_equal(state[0], ox) and np.array_equal(state[1], oy)
    assert np.array_equal(out.reward.cpu().numpy(), o_rew[0])
    assert np.array_equal(out.terminated.cpu().numpy(), o_term[0])
    # every frame equals the oracle's frame of its state (frames are grouped by state: Road Block
    # has only a few thousand reachable states, each is rendered once on the CPU)
    cols = state.reshape(-1, n).T  # [n, 2 * n_max]
    uniq, inverse = np.unique(cols, axis=0, return_inverse=True)
    assert len(uniq) >= 256
    inverse = torch.from_numpy(inverse.reshape(-1)).cuda()
    n_obj = oracle.num_objects
    for u, row in enumerate(uniq):
        st = tuple((int(row[j]), int(row[engine.n_max + j])) for j in range(n_obj))
        if dtype == "float32":
            want = po.render_observation_padded(oracle, st, oracle.height, oracle.width, 20, 2)
        else:
            want = oracle.render(st, border_width=2, pixels_per_cell=20)
        members = torch.nonzero(inverse == u).reshape(-1)
        frames = obs[members]
        want_t = torch.from_numpy(np.ascontiguousarray(want)).cuda()
        assert bool((frames == want_t.unsqueeze(0)).all()), (u, st)
