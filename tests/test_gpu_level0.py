"""ONE engine over the reference's real level-0 set (all 15,400 puzzles of benchmark/puzzles/level0.zip;
13,321 of them fit the bitboard kernels, 2,079 do not, so the engine runs the general rule): golden
transitions, full rollouts against the C oracle, and on-device episode handling against a host
model -- the RL-style mixed batch SURVEY.md 8f-2 describes."""

import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import pushworld_oracle as po  # noqa: E402
from tests.conftest import GOLDEN  # noqa: E402
from tests.oracle_batch import OracleSet  # noqa: E402


@pytest.fixture(scope="module")
def world():
    from pushworld_b200 import data
    from pushworld_b200.engine import Engine

    keys = data.level0_keys()
    texts = [data.level0_text(k) for k in keys]
    return Engine(texts), keys, texts


def test_golden_transitions_in_one_mixed_launch(world):
    engine, keys, texts = world
    npz = np.load(os.path.join(GOLDEN, "level0_vectors.npz"))
    index = json.loads(bytes(npz["index_json"]).decode())
    pos = {k: i for i, k in enumerate(keys)}
    pid, s0, s1, act, dg, term = [], [], [], [], [], []
    for pi, rec in enumerate(index):
        states = npz[f"p{pi}_states"].astype(np.int32)
        n_t = len(states) - 1
        pid.append(np.full(n_t, pos[rec["key"]], np.int32))
        pad = np.zeros((n_t + 1, engine.n_max, 2), np.int32)
        pad[:, : rec["n"]] = states
        s0.append(pad[:-1])
        s1.append(pad[1:])
        act.append(npz[f"p{pi}_actions"])
        ach = npz[f"p{pi}_achieved"].astype(np.int32)
        dg.append(ach[1:] - ach[:-1])
        term.append(npz[f"p{pi}_isgoal"][1:])
    s0, s1 = np.concatenate(s0), np.concatenate(s1)
    xy = torch.from_numpy(np.ascontiguousarray(np.stack([s0[:, :, 0].T, s0[:, :, 1].T]))).cuda()
    out = engine.step(xy, torch.from_numpy(np.concatenate(pid)).cuda(), torch.from_numpy(np.concatenate(act)).cuda())
    want = np.stack([s1[:, :, 0].T, s1[:, :, 1].T])
    assert np.array_equal(xy.cpu().numpy(), want)
    dg, term = np.concatenate(dg), np.concatenate(term)
    assert np.array_equal(out.dgoals.cpu().numpy().astype(np.int32), dg)
    assert np.array_equal(out.terminated.cpu().numpy().astype(bool), term)
    want_reward = np.where(term, np.float32(10.0), (dg.astype(np.float64) - 0.01).astype(np.float32))
    assert np.array_equal(out.reward.cpu().numpy(), want_reward)
    assert len(dg) == 84000


def test_rollout_over_random_level0_puzzles_matches_the_c_oracle(world):
    from pushworld_b200.engine import record_to_outputs

    engine, keys, texts = world
    n, T = 1 << 15, 200
    rng = np.random.default_rng(77)
    subset = rng.choice(len(keys), 1500, replace=False)
    ids = subset[rng.integers(0, len(subset), size=n)].astype(np.int64)
    pid = torch.from_numpy(ids.astype(np.int32)).cuda()
    gen = torch.Generator(device="cuda").manual_seed(78)
    actions = torch.randint(0, 4, (T, n), dtype=torch.uint8, device="cuda", generator=gen)
    xy = engine.new_state(n)
    engine.reset(xy, pid)
    reward = torch.empty((T, n), dtype=torch.float32, device="cuda")
    term = torch.empty((T, n), dtype=torch.uint8, device="cuda")
    engine.rollout(xy, pid, actions, reward=reward, terminated=term)
    xy2 = engine.new_state(n)
    engine.reset(xy2, pid)
    record = torch.empty((T, n), dtype=torch.uint8, device="cuda")
    engine.rollout_packed(xy2, pid, actions, record)
    r2, t2, _, _ = record_to_outputs(record)
    assert torch.equal(r2, reward) and torch.equal(t2, term) and torch.equal(xy2, xy)
    ox, oy, o_rew, o_term = OracleSet(texts).rollout(ids, actions.cpu().numpy(), engine.n_max)
    got = xy.cpu().numpy()
    assert np.array_equal(got[0], ox) and np.array_equal(got[1], oy)
    assert np.array_equal(reward.cpu().numpy(), o_rew) and np.array_equal(term.cpu().numpy(), o_term)
    assert o_term.any()


@pytest.mark.parametrize("n", [64, 48])  # whole warps: the mask kernel's episode handling; 48: the per-thread rule
@pytest.mark.parametrize("packed", [False, True])
def test_episodes_on_the_real_level0_set(world, n, packed):
    """pw_rollout_episodes / pw_rollout_packed(RESAMPLE) on an engine that is NOT board-eligible:
    every step equals a host simulation with the oracle (draws from all 15,400 puzzles)."""
    from pushworld_b200.engine import draw_puzzle, record_to_outputs

    engine, keys, texts = world
    T, max_steps, seed = 120, 17, 0xABCDEF987
    rng = np.random.default_rng(5)
    actions = rng.integers(0, 4, size=(T, n), dtype=np.uint8)
    pid0 = rng.integers(0, len(texts), size=n).astype(np.int32)
    puzzle_id = torch.from_numpy(pid0.copy()).cuda()
    xy = engine.new_state(n)
    engine.reset(xy, puzzle_id)
    steps = torch.zeros(n, dtype=torch.int32, device="cuda")
    episode = torch.zeros(n, dtype=torch.int32, device="cuda")
    if packed:
        record = torch.empty((T, n), dtype=torch.uint8, device="cuda")
        engine.rollout_packed(xy, puzzle_id, torch.from_numpy(actions).cuda(), record, steps=steps,
                              max_steps=max_steps, episode=episode, seed=seed)
        reward, term, trunc, _ = record_to_outputs(record)
    else:
        reward = torch.empty((T, n), dtype=torch.float32, device="cuda")
        term = torch.empty((T, n), dtype=torch.uint8, device="cuda")
        trunc = torch.empty((T, n), dtype=torch.uint8, device="cuda")
        engine.rollout_episodes(xy, puzzle_id, torch.from_numpy(actions).cuda(), steps, episode, seed,
                                max_steps=max_steps, reward=reward, terminated=term, truncated=trunc)
    r, te, tr = reward.cpu().numpy(), term.cpu().numpy(), trunc.cpu().numpy()
    final, fpid = xy.cpu().numpy(), puzzle_id.cpu().numpy()
    fsteps, fep = steps.cpu().numpy(), episode.cpu().numpy()
    oracles = {}

    def oracle(p):
        if p not in oracles:
            oracles[p] = po.OraclePuzzle(texts[p])
        return oracles[p]

    n_resets = 0
    for e in range(n):
        pid, ep, st = int(pid0[e]), 0, 0
        state = oracle(pid).initial_state
        for t in range(T):
            nxt = oracle(pid).get_next_state(state, int(actions[t, e]))
            rew, terminated, _ = po.step_reward(oracle(pid), state, nxt)
            st += 1
            truncated = st >= max_steps
            assert np.float32(rew) == r[t, e] and terminated == bool(te[t, e]) and truncated == bool(tr[t, e]), (e, t)
            state = nxt
            if terminated or truncated:
                ep += 1
                pid = draw_puzzle(seed, e, ep, len(texts))
                state, st = oracle(pid).initial_state, 0
                n_resets += 1
        assert (fpid[e], fep[e], fsteps[e]) == (pid, ep, st), e
        got = tuple((int(final[0, j, e]), int(final[1, j, e])) for j in range(oracle(pid).num_objects))
        assert got == state, e
    assert n_resets >= n * 6
