"""GPU parity tests of the dynamics path, through the C ABI (ctypes -> libpushworld_b200.so).

Bit-exact comparisons (integer / index work) against
  * the golden vectors the real reference produced (tests/golden), and
  * the CPU oracle (oracle/pushworld_oracle.py) on seeded inputs.
Rewards are float32 and must equal np.float32(<the reference's python float>) exactly.
"""

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import pushworld_oracle as po  # noqa: E402


@pytest.fixture(scope="module")
def setup(golden):
    from pushworld_b200.engine import Engine

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    keys = golden.keys()
    engine = Engine([golden.text(k) for k in keys])
    return engine, keys


def _expected_reward(dg, term):
    # gym_env.py:212-221 evaluated in python floats, then rounded once to float32
    return np.where(term, np.float32(10.0), (dg.astype(np.float64) - 0.01).astype(np.float32))


def _transitions(golden, keys, tags=("sol", "walk")):
    """All golden (state, action, next_state) triples as flat arrays."""
    pid, s0, s1, act, a0, a1, g0, g1 = [], [], [], [], [], [], [], []
    for p, key in enumerate(keys):
        for tag in tags:
            if not golden.has(key, f"{tag}_actions"):
                continue
            states = golden.arr(key, f"{tag}_states").astype(np.int32)
            actions = golden.arr(key, f"{tag}_actions")
            ach = golden.arr(key, f"{tag}_achieved").astype(np.int32)
            isg = golden.arr(key, f"{tag}_isgoal")
            pid.append(np.full(len(actions), p, np.int32))
            s0.append(states[:-1])
            s1.append(states[1:])
            act.append(actions)
            a0.append(ach[:-1])
            a1.append(ach[1:])
            g1.append(isg[1:])
    return pid, s0, s1, act, a0, a1, g1


def _pack(states_list, n_max):
    total = sum(len(s) for s in states_list)
    out = np.zeros((2, n_max, total), np.int32)
    pos = 0
    for s in states_list:
        n = s.shape[1]
        out[0, :n, pos : pos + len(s)] = s[:, :, 0].T
        out[1, :n, pos : pos + len(s)] = s[:, :, 1].T
        pos += len(s)
    return out


def test_step_mixed_batch_matches_reference(setup, golden):
    """Every golden transition of every puzzle in ONE mixed-puzzle launch (85k environments)."""
    engine, keys = setup
    pid, s0, s1, act, a0, a1, g1 = _transitions(golden, keys)
    xy = torch.from_numpy(_pack(s0, engine.n_max)).cuda()
    want = _pack(s1, engine.n_max)
    puzzle_id = torch.from_numpy(np.concatenate(pid)).cuda()
    action = torch.from_numpy(np.concatenate(act)).cuda()
    n = action.numel()
    assert n > 80000
    steps = torch.zeros(n, dtype=torch.int32, device="cuda")
    out = engine.step(xy, puzzle_id, action, steps=steps, max_steps=1, want_moved=True)
    torch.cuda.synchronize()
    got = xy.cpu().numpy()
    assert np.array_equal(got, want)
    dg = np.concatenate(a1) - np.concatenate(a0)
    term = np.concatenate(g1)
    assert np.array_equal(out.dgoals.cpu().numpy().astype(np.int32), dg)
    assert np.array_equal(out.terminated.cpu().numpy().astype(bool), term)
    assert np.array_equal(out.reward.cpu().numpy(), _expected_reward(dg, term))
    assert out.truncated.cpu().numpy().all()  # max_steps = 1
    assert (steps.cpu().numpy() == 1).all()
    # moved mask == set of objects whose position changed
    before = _pack(s0, engine.n_max)
    changed = (before != want).any(axis=0)  # [n_max, N]
    moved = out.moved.cpu().numpy().astype(np.uint32)
    bits = ((moved[None, :] >> np.arange(engine.n_max, dtype=np.uint32)[:, None]) & 1).astype(bool)
    assert np.array_equal(bits, changed)
    assert engine.poll_invalid_actions() == 0


def test_step_uniform_launch_per_puzzle(setup, golden):
    """Same transitions, one single-puzzle launch per puzzle: exercises the shared-memory table
    path and all three state policies (registers n<=4, n<=8, shared memory)."""
    engine, keys = setup
    seen_policies = set()
    for p, key in enumerate(keys):
        pid, s0, s1, act, a0, a1, g1 = _transitions(golden, [key])
        n_obj = golden.rec(key)["n"]
        seen_policies.add(4 if n_obj <= 4 else (8 if n_obj <= 8 else 0))
        xy = torch.from_numpy(_pack(s0, engine.n_max)).cuda()
        action = torch.from_numpy(np.concatenate(act)).cuda()
        out = engine.step(xy, p, action)
        assert np.array_equal(xy.cpu().numpy(), _pack(s1, engine.n_max)), key
        dg = np.concatenate(a1) - np.concatenate(a0)
        term = np.concatenate(g1)
        assert np.array_equal(out.reward.cpu().numpy(), _expected_reward(dg, term)), key
        assert not out.truncated.cpu().numpy().any()  # max_steps None
    assert seen_policies == {0, 4, 8}


def test_rollout_replays_golden_walks(setup, golden):
    """pw_rollout with T = len(walk): column e replays the golden action sequence of puzzle e;
    per-step reward/terminated and the final state must match the reference."""
    engine, keys = setup
    bench = [k for k in keys if k.startswith("benchmark/")]
    ids = [keys.index(k) for k in bench]
    T = 300
    actions = np.stack([golden.arr(k, "walk_actions") for k in bench], axis=1)  # [T, N]
    assert actions.shape == (T, len(bench))
    puzzle_id = torch.tensor(ids, dtype=torch.int32, device="cuda")
    n = len(bench)
    xy = engine.new_state(n)
    engine.reset(xy, puzzle_id)
    init = _pack([golden.arr(k, "walk_states").astype(np.int32)[:1] for k in bench], engine.n_max)
    assert np.array_equal(xy.cpu().numpy(), init)
    reward = torch.empty((T, n), dtype=torch.float32, device="cuda")
    term = torch.empty((T, n), dtype=torch.uint8, device="cuda")
    dgo = torch.empty((T, n), dtype=torch.int8, device="cuda")
    trunc = torch.empty((T, n), dtype=torch.uint8, device="cuda")
    steps = torch.zeros(n, dtype=torch.int32, device="cuda")
    ret = torch.zeros(n, dtype=torch.float32, device="cuda")
    counters = torch.zeros(4, dtype=torch.int64, device="cuda")
    engine.rollout(xy, puzzle_id, torch.from_numpy(actions).cuda(), reward=reward, terminated=term,
                   dgoals=dgo, truncated=trunc, steps=steps, max_steps=250, return_sum=ret,
                   counters=counters)
    final = _pack([golden.arr(k, "walk_states").astype(np.int32)[-1:] for k in bench], engine.n_max)
    assert np.array_equal(xy.cpu().numpy(), final)
    ach = np.stack([golden.arr(k, "walk_achieved").astype(np.int32) for k in bench], axis=1)
    isg = np.stack([golden.arr(k, "walk_isgoal") for k in bench], axis=1)
    dg = ach[1:] - ach[:-1]
    assert np.array_equal(dgo.cpu().numpy().astype(np.int32), dg)
    assert np.array_equal(term.cpu().numpy().astype(bool), isg[1:])
    want_r = _expected_reward(dg, isg[1:])
    assert np.array_equal(reward.cpu().numpy(), want_r)
    want_trunc = np.arange(1, T + 1)[:, None] >= 250
    assert np.array_equal(trunc.cpu().numpy().astype(bool), np.broadcast_to(want_trunc, (T, n)))
    assert (steps.cpu().numpy() == T).all()
    # float32 accumulation in step order, like the kernel
    acc = np.zeros(n, np.float32)
    for t in range(T):
        acc = acc + want_r[t]
    assert np.array_equal(ret.cpu().numpy(), acc)
    c = counters.cpu().numpy()
    assert c[0] == T * n
    assert c[1] == isg[1:].sum()
    moved_steps = sum(
        int((np.diff(golden.arr(k, "walk_states").astype(np.int32), axis=0) != 0).any(axis=(1, 2)).sum())
        for k in bench
    )
    assert c[2] == moved_steps
    assert c[3] == np.maximum(dg, 0).sum()


def test_rollout_solution_plans_reach_goal_exactly_at_the_end(setup, golden):
    """python3/test/test_dataset.py restated: every benchmark plan is valid -- the goal is
    reached at the last step and not earlier (puzzle.py:413-424)."""
    engine, keys = setup
    bench = [k for k in keys if k.startswith("benchmark/")]
    lengths = [len(golden.plan(k)) for k in bench]
    T = max(lengths)
    n = len(bench)
    # pad shorter plans with an invalid action (255 = no-op by contract)
    actions = np.full((T, n), 255, np.uint8)
    for e, k in enumerate(bench):
        actions[: lengths[e], e] = [po.ACTION_FROM_CHAR[c] for c in golden.plan(k).upper()]
    puzzle_id = torch.tensor([keys.index(k) for k in bench], dtype=torch.int32, device="cuda")
    xy = engine.new_state(n)
    engine.reset(xy, puzzle_id)
    term = torch.empty((T, n), dtype=torch.uint8, device="cuda")
    engine.rollout(xy, puzzle_id, torch.from_numpy(actions).cuda(), terminated=term)
    term = term.cpu().numpy().astype(bool)
    for e in range(n):
        assert term[lengths[e] - 1, e], bench[e]
        assert not term[: lengths[e] - 1, e].any(), bench[e]
    padded = int((actions == 255).sum())
    assert engine.poll_invalid_actions() == padded
    assert sum(lengths) == 18143


def test_rollout_equals_repeated_steps_and_oracle(setup, golden):
    """Seeded random actions on a few puzzles: pw_rollout == T x pw_step == python oracle."""
    engine, keys = setup
    rng = np.random.default_rng(7)
    for key in ["benchmark/level1/Road Block", "benchmark/level2/Clean Sweep",
                "python3/trivial_tool", "python3/file_parsing", "benchmark/level4/Four Rooms"]:
        p = keys.index(key)
        oracle = po.OraclePuzzle(golden.text(key))
        n, T = 64, 96
        actions = rng.integers(0, 4, size=(T, n), dtype=np.uint8)
        xy_a = engine.new_state(n)
        engine.reset(xy_a, p)
        xy_b = xy_a.clone()
        reward = torch.empty((T, n), dtype=torch.float32, device="cuda")
        engine.rollout(xy_a, p, torch.from_numpy(actions).cuda(), reward=reward)
        step_rewards = []
        for t in range(T):
            out = engine.step(xy_b, p, torch.from_numpy(actions[t]).cuda())
            step_rewards.append(out.reward.cpu().numpy())
        assert torch.equal(xy_a, xy_b), key
        assert np.array_equal(reward.cpu().numpy(), np.stack(step_rewards)), key
        final = xy_a.cpu().numpy()
        for e in range(0, n, 8):
            state = oracle.initial_state
            for t in range(T):
                nxt = oracle.get_next_state(state, int(actions[t, e]))
                r, term, _ = po.step_reward(oracle, state, nxt)
                assert np.float32(r) == reward[t, e].item(), (key, e, t)
                state = nxt
            got = tuple((int(final[0, j, e]), int(final[1, j, e])) for j in range(oracle.num_objects))
            assert got == state, (key, e)


def test_autoreset(setup, golden):
    engine, keys = setup
    p = keys.index("python3/trivial")
    # RDRU solves trivial.pwp (test_puzzle.py:256); with autoreset the 5th step starts over
    plan = np.array([[1], [3], [1], [2], [1], [3], [1], [2]], np.uint8)
    xy = engine.new_state(1)
    engine.reset(xy, p)
    term = torch.empty((8, 1), dtype=torch.uint8, device="cuda")
    steps = torch.zeros(1, dtype=torch.int32, device="cuda")
    engine.rollout(xy, p, torch.from_numpy(plan).cuda(), terminated=term, steps=steps, autoreset=True)
    assert term.cpu().numpy()[:, 0].tolist() == [0, 0, 0, 1, 0, 0, 0, 1]
    assert steps.item() == 0
    assert xy.cpu().numpy()[:, :2, 0].T.tolist() == [[1, 2], [2, 2]]  # back at initial_state


def test_expand_matches_oracle(setup, golden):
    """pw_expand: 4 successors + moved mask (RelativeState.moved_object_indices,
    pushworld_puzzle.cc:439-452) + goal flag, against the oracle on golden walk states."""
    engine, keys = setup
    for key in ["benchmark/level2/Pull Dont Push", "benchmark/level1/Road Block",
                "benchmark/level2/Clean Sweep", "cpp/transitive_pushing", "python3/trivial"]:
        p = keys.index(key)
        oracle = po.OraclePuzzle(golden.text(key))
        states = golden.arr(key, "walk_states").astype(np.int32)[::3]
        n = len(states)
        xy = torch.from_numpy(_pack([states], engine.n_max)).cuda()
        xy_out, moved, goal = engine.expand(xy, p)
        assert torch.equal(xy, torch.from_numpy(_pack([states], engine.n_max)).cuda())  # input intact
        xo, mv, gl = xy_out.cpu().numpy(), moved.cpu().numpy().astype(np.uint32), goal.cpu().numpy()
        for e in range(n):
            state = tuple((int(a), int(b)) for a, b in states[e])
            for a in range(4):
                nxt, moved_idx = oracle.get_next_state_and_moved(state, a)
                s = a * n + e
                got = tuple((int(xo[0, j, s]), int(xo[1, j, s])) for j in range(oracle.num_objects))
                assert got == nxt, (key, e, a)
                assert mv[s] == sum(1 << k for k in moved_idx), (key, e, a)
                assert bool(gl[s]) == oracle.is_goal_state(nxt)


def test_irregular_puzzles_match_oracle(setup):
    """Puzzle files whose objects start overlapping walls / each other exercise the "not
    already overlapping" and "bounding box inside the grid" terms of puzzle.py:557-562,592;
    objects may even leave the grid.  The kernel must still agree with the oracle."""
    from pushworld_b200.engine import Engine

    texts = [
        "a m0+m1 . g0\n. . . .\n",  # two movables in one cell
        "a w+m0 . g0\n. m1 . .\n",  # movable inside a wall
        "a+aw m0 . .\n. . m1 g1\n",  # agent inside an agent wall
        "m2 m2 m2 .\nm2 a m2 .\nm2 m2 m2 m1\n",  # agent inside a ring
        "a m0 . w+m1\n. . . w+m1\n",  # movable partly inside a wall
        ". . . . .\n. m0 m0+m1 m1 .\n. a . . g0\n",
    ]
    engine = Engine(texts)
    rng = np.random.default_rng(11)
    T = 400
    for p, text in enumerate(texts):
        oracle = po.OraclePuzzle(text)
        n = 16
        actions = rng.integers(0, 4, size=(T, n), dtype=np.uint8)
        xy = engine.new_state(n)
        engine.reset(xy, p)
        dgo = torch.empty((T, n), dtype=torch.int8, device="cuda")
        term = torch.empty((T, n), dtype=torch.uint8, device="cuda")
        engine.rollout(xy, p, torch.from_numpy(actions).cuda(), dgoals=dgo, terminated=term)
        final = xy.cpu().numpy()
        for e in range(n):
            state = oracle.initial_state
            for t in range(T):
                nxt = oracle.get_next_state(state, int(actions[t, e]))
                _, te, dg = po.step_reward(oracle, state, nxt)
                assert dgo[t, e].item() == dg and bool(term[t, e].item()) == te, (p, e, t)
                state = nxt
            got = tuple((int(final[0, j, e]), int(final[1, j, e])) for j in range(oracle.num_objects))
            assert got == state, (p, e)


def test_argument_errors(setup):
    from pushworld_b200 import _lib

    engine, keys = setup
    xy = engine.new_state(4)
    with pytest.raises(ValueError):
        engine.step(xy, 0, torch.zeros(4, dtype=torch.int64, device="cuda"))
    with pytest.raises(ValueError):
        engine.step(xy, len(keys), torch.zeros(4, dtype=torch.uint8, device="cuda"))
    with pytest.raises(ValueError):
        engine.step(xy[:, :1], 0, torch.zeros(4, dtype=torch.uint8, device="cuda"))
    with pytest.raises(ValueError):
        engine.rollout(xy, 0, torch.zeros((3, 5), dtype=torch.uint8, device="cuda"))
    # empty batches are fine
    empty = engine.new_state(0)
    engine.reset(empty, 0)
    engine.step(empty, 0, torch.zeros(0, dtype=torch.uint8, device="cuda"))
    # invalid actions are no-ops and are counted
    engine.reset(xy, 0)
    before = xy.clone()
    engine.step(xy, 0, torch.full((4,), 7, dtype=torch.uint8, device="cuda"))
    assert torch.equal(xy, before)
    assert engine.poll_invalid_actions() == 4
    assert engine.poll_invalid_actions() == 0
    # the raw ABI reports errors through return codes + pw_last_error
    lib = _lib.load()
    rc = lib.pw_step(engine._h, 4, None, 10**6, xy[0].data_ptr(), xy[1].data_ptr(), 4, None,
                     None, None, None, None, None, 0, None, None)
    assert rc == _lib.PW_E_INVALID and b"puzzle_uniform" in lib.pw_last_error()


@pytest.mark.parametrize("names", [
    # general rule (file_parsing does not fit a 64-bit board)
    ["trivial", "multiple_goals", "transitive_pushing", "pushing", "trivial_obstacle", "trivial_tool",
     "file_parsing"],
    # every puzzle fits a uint64 board: per-thread bitboard kernel
    ["trivial", "multiple_goals", "transitive_pushing", "pushing", "trivial_obstacle", "trivial_tool"],
    # every puzzle fits a uint32 board
    ["trivial", "multiple_goals", "transitive_pushing", "pushing", "trivial_obstacle"],
])
@pytest.mark.parametrize("n", [48, 64])  # 64 = whole warps: the plain-rollout loop with episode handling
def test_rollout_episodes_resamples_puzzles_like_the_host_model(golden, names, n):
    """pw_rollout_episodes: on termination / truncation an environment draws a new puzzle with
    the documented counter-based generator and restarts; the whole trajectory must equal a host
    simulation with the oracle (reward, terminated, truncated per step; final puzzle, episode,
    step counter and state)."""
    from pushworld_b200.engine import Engine, draw_puzzle

    texts = [golden.text("python3/" + n) for n in names]
    engine = Engine(texts)
    oracles = [po.OraclePuzzle(t) for t in texts]
    T, max_steps, seed = 160, 9, 0xC0FFEE1234
    rng = np.random.default_rng(21)
    actions = rng.integers(0, 4, size=(T, n), dtype=np.uint8)
    pid0 = rng.integers(0, len(texts), size=n).astype(np.int32)
    puzzle_id = torch.from_numpy(pid0.copy()).cuda()
    xy = engine.new_state(n)
    engine.reset(xy, puzzle_id)
    steps = torch.zeros(n, dtype=torch.int32, device="cuda")
    episode = torch.zeros(n, dtype=torch.int32, device="cuda")
    reward = torch.empty((T, n), dtype=torch.float32, device="cuda")
    term = torch.empty((T, n), dtype=torch.uint8, device="cuda")
    trunc = torch.empty((T, n), dtype=torch.uint8, device="cuda")
    engine.rollout_episodes(xy, puzzle_id, torch.from_numpy(actions).cuda(), steps, episode, seed,
                            max_steps=max_steps, reward=reward, terminated=term, truncated=trunc)
    r, te, tr = reward.cpu().numpy(), term.cpu().numpy(), trunc.cpu().numpy()
    final, fpid = xy.cpu().numpy(), puzzle_id.cpu().numpy()
    fsteps, fep = steps.cpu().numpy(), episode.cpu().numpy()
    n_resets = 0
    for e in range(n):
        pid, ep, st = int(pid0[e]), 0, 0
        state = oracles[pid].initial_state
        for t in range(T):
            nxt = oracles[pid].get_next_state(state, int(actions[t, e]))
            rew, terminated, _ = po.step_reward(oracles[pid], state, nxt)
            st += 1
            truncated = st >= max_steps
            assert np.float32(rew) == r[t, e] and terminated == bool(te[t, e]) and truncated == bool(tr[t, e]), (e, t)
            state = nxt
            if terminated or truncated:
                ep += 1
                pid = draw_puzzle(seed, e, ep, len(texts))
                state, st = oracles[pid].initial_state, 0
                n_resets += 1
        assert (fpid[e], fep[e], fsteps[e]) == (pid, ep, st), e
        got = tuple((int(final[0, j, e]), int(final[1, j, e])) for j in range(oracles[pid].num_objects))
        assert got == state, e
    assert n_resets > n * 10
    assert len(set(fpid.tolist())) > 3  # the draw really spreads over the puzzles


def test_mixed_bitboard_kernel_matches_oracle_and_handles_irregular_inputs(golden):
    """An engine whose puzzles all fit one machine word runs per-environment puzzle ids on the
    per-thread bitboard kernel; plain rollouts, T=1 steps and states that are NOT regular (two
    objects on one cell: general-rule slow path inside the same launch) must all match the oracle."""
    from pushworld_b200.engine import Engine

    names = ["trivial", "multiple_goals", "transitive_pushing", "pushing", "trivial_obstacle"]
    texts = [golden.text("python3/" + n) for n in names]
    engine = Engine(texts)
    oracles = [po.OraclePuzzle(t) for t in texts]
    rng = np.random.default_rng(33)
    n, T = 96, 120
    pids = rng.integers(0, len(texts), size=n).astype(np.int32)
    puzzle_id = torch.from_numpy(pids).cuda()
    states = []
    for e in range(n):
        o = oracles[pids[e]]
        st = o.initial_state
        for _ in range(int(rng.integers(0, 30))):
            st = o.get_next_state(st, int(rng.integers(4)))
        if e % 7 == 3 and o.num_objects >= 3:  # irregular: object 2 dropped onto object 1
            st = st[:2] + (st[1],) + st[3:]
        states.append(st)
    actions = rng.integers(0, 5, size=(T, n)).astype(np.uint8)  # 4 = invalid -> no-op
    xy = engine.state_from_tuples(states)
    reward = torch.empty((T, n), dtype=torch.float32, device="cuda")
    term = torch.empty((T, n), dtype=torch.uint8, device="cuda")
    engine.rollout(xy, puzzle_id, torch.from_numpy(actions).cuda(), reward=reward, terminated=term)
    xy1 = engine.state_from_tuples(states)
    out1 = engine.step(xy1, puzzle_id, torch.from_numpy(actions[0]).cuda(), want_moved=True)
    final, r, te = xy.cpu().numpy(), reward.cpu().numpy(), term.cpu().numpy()
    first, moved1 = xy1.cpu().numpy(), out1.moved.cpu().numpy()
    for e in range(n):
        o = oracles[pids[e]]
        st = states[e]
        for t in range(T):
            a = int(actions[t, e])
            nxt, moved_idx = (st, []) if a > 3 else o.get_next_state_and_moved(st, a)
            rr, tt, _ = po.step_reward(o, st, nxt)
            assert np.float32(rr) == r[t, e] and tt == bool(te[t, e]), (e, t)
            if t == 0:
                got = tuple((int(first[0, j, e]), int(first[1, j, e])) for j in range(o.num_objects))
                assert got == nxt and int(moved1[e]) == sum(1 << k for k in moved_idx), e
            st = nxt
        got = tuple((int(final[0, j, e]), int(final[1, j, e])) for j in range(o.num_objects))
        assert got == st, e
    assert engine.poll_invalid_actions() == int((actions > 3).sum()) + int((actions[0] > 3).sum())


def _random_small_puzzle(rng, max_bits_plus_pitch=32):
    """A random regular puzzle whose board leaves a spare pitch in a uint32 (ring-board eligible)."""
    while True:
        w, h = int(rng.integers(1, 7)), int(rng.integers(1, 15))  # narrow and tall ones use the top bits
        if 1 + h * (w + 1) + (w + 1) > max_bits_plus_pitch:
            continue
        grid = [["." for _ in range(w)] for _ in range(h)]
        n_mov = int(rng.integers(1, 8))
        names = ["a"] + [f"m{k}" for k in range(n_mov)]
        for name in names:
            for _ in range(int(rng.integers(1, 3))):
                x, y = int(rng.integers(w)), int(rng.integers(h))
                if grid[y][x] == ".":
                    grid[y][x] = name
        extras = ["w", "aw"] + [f"g{k}" for k in range(int(rng.integers(1, min(n_mov, 3) + 1)))]
        for name in extras:
            x, y = int(rng.integers(w)), int(rng.integers(h))
            if grid[y][x] == ".":
                grid[y][x] = name
        text = "\n".join(" ".join(row) for row in grid) + "\n"
        try:
            oracle = po.OraclePuzzle(text)
        except (ValueError, AssertionError):
            continue
        if oracle.num_objects < 2 or len(oracle.goal_state) < 1:
            continue
        return text, oracle


def test_ring_rollout_kernel_matches_oracle(golden):
    """The ring-board rollout kernel (k_rollout_ring: plain rollouts of a single puzzle whose board
    leaves one spare pitch in a uint32): reward, terminated, final state and the counters against
    the oracle for benchmark puzzles and random small puzzles (2..8 objects, 1..3 goals), with
    invalid actions, step counts around the 4-step group size and irregular start states."""
    from pushworld_b200.engine import Engine

    rng = np.random.default_rng(4242)
    cases = [(golden.text(k), po.OraclePuzzle(golden.text(k)))
             for k in ("benchmark/level1/Road Block", "benchmark/level1/Obstacle Has Obstacle")]
    cases += [_random_small_puzzle(rng) for _ in range(60)]
    seen_n, seen_g, max_bits = set(), set(), 0
    for ci, (text, oracle) in enumerate(cases):
        assert 1 + (oracle.height - 2) * (oracle.width - 1) + oracle.width - 1 <= 32
        if len(oracle.goal_state) == 1:
            max_bits = max(max_bits, 1 + (oracle.height - 2) * (oracle.width - 1))
        engine = Engine([text])
        seen_n.add(oracle.num_objects)
        seen_g.add(len(oracle.goal_state))
        n = 96
        states = []
        for e in range(n):
            st = oracle.initial_state
            for _ in range(int(rng.integers(0, 25))):
                st = oracle.get_next_state(st, int(rng.integers(4)))
            if e % 11 == 5 and oracle.num_objects >= 3:  # irregular: object 2 dropped onto object 1
                st = st[:2] + (st[1],) + st[3:]
            states.append(st)
        for T in ([1, 3, 4, 5, 8, 41, 64, 100, 131] if ci < 4 else [int(rng.integers(1, 60))]):
            hi = 4 if ci % 3 else 6  # some cases with invalid actions (4, 5)
            actions = rng.integers(0, hi, size=(T, n)).astype(np.uint8)
            if ci % 5 == 0:
                actions[rng.integers(0, T), rng.integers(0, n)] = 255
            xy = engine.state_from_tuples(states)
            reward = torch.empty((T, n), dtype=torch.float32, device="cuda")
            term = torch.empty((T, n), dtype=torch.uint8, device="cuda")
            counters = torch.zeros(4, dtype=torch.int64, device="cuda")
            engine.rollout(xy, 0, torch.from_numpy(actions).cuda(), reward=reward, terminated=term,
                           counters=counters)
            final, r, te = xy.cpu().numpy(), reward.cpu().numpy(), term.cpu().numpy()
            n_term = n_moved = n_gain = 0
            for e in range(n):
                st = states[e]
                for t in range(T):
                    a = int(actions[t, e])
                    nxt = st if a > 3 else oracle.get_next_state(st, a)
                    rr, tt, _ = po.step_reward(oracle, st, nxt)
                    assert np.float32(rr) == r[t, e] and tt == bool(te[t, e]), (text, T, e, t)
                    n_term += tt
                    n_moved += nxt != st
                    n_gain += max(oracle.count_achieved_goals(nxt) - oracle.count_achieved_goals(st), 0)
                    st = nxt
                got = tuple((int(final[0, j, e]), int(final[1, j, e])) for j in range(oracle.num_objects))
                assert got == st, (text, T, e)
            assert counters.cpu().numpy().tolist() == [n * T, n_term, n_moved, n_gain], (text, T)
            assert engine.poll_invalid_actions() == int((actions > 3).sum())
    assert len(seen_n) >= 5 and len(seen_g) >= 2 and max_bits >= 27, (seen_n, seen_g, max_bits)


def test_puzzle_whose_tables_exceed_shared_memory():
    """A 60 x 1000 puzzle: its collision bit tables (~190 KB) do not fit the shared-memory staging
    limit, so single-puzzle launches read them through the global-table kernels.  Rollout, step and
    expansion against the oracle."""
    from pushworld_b200.engine import Engine

    rng = np.random.default_rng(99)
    w, h = 60, 1000
    grid = [["." for _ in range(w)] for _ in range(h)]
    grid[500][30] = "a"
    grid[499][30] = "a"
    for k, (x, y) in enumerate([(31, 500), (29, 500), (30, 498), (30, 501), (32, 500)]):
        grid[y][x] = f"m{k}"
    grid[500][34] = "w"
    grid[496][30] = "aw"
    grid[503][30] = "g0"
    text = "\n".join(" ".join(row) for row in grid) + "\n"
    oracle = po.OraclePuzzle(text)
    engine = Engine([text])
    n, T = 64, 60
    actions = rng.integers(0, 4, size=(T, n)).astype(np.uint8)
    xy = engine.new_state(n)
    engine.reset(xy, 0)
    reward = torch.empty((T, n), dtype=torch.float32, device="cuda")
    term = torch.empty((T, n), dtype=torch.uint8, device="cuda")
    engine.rollout(xy, 0, torch.from_numpy(actions).cuda(), reward=reward, terminated=term)
    final, r, te = xy.cpu().numpy(), reward.cpu().numpy(), term.cpu().numpy()
    states = []
    for e in range(n):
        st = oracle.initial_state
        for t in range(T):
            nxt = oracle.get_next_state(st, int(actions[t, e]))
            rr, tt, _ = po.step_reward(oracle, st, nxt)
            assert np.float32(rr) == r[t, e] and tt == bool(te[t, e]), (e, t)
            st = nxt
        got = tuple((int(final[0, j, e]), int(final[1, j, e])) for j in range(oracle.num_objects))
        assert got == st, e
        states.append(st)
    xy_out, moved, goal = engine.expand(xy, 0)
    xo, mo, go = xy_out.cpu().numpy(), moved.cpu().numpy(), goal.cpu().numpy()
    for e in range(n):
        for a in range(4):
            nxt, moved_idx = oracle.get_next_state_and_moved(states[e], a)
            s = a * n + e
            got = tuple((int(xo[0, j, s]), int(xo[1, j, s])) for j in range(oracle.num_objects))
            assert got == nxt and int(mo[s]) == sum(1 << k for k in moved_idx), (e, a)
            assert bool(go[s]) == oracle.is_goal_state(nxt)


def test_plain_rollouts_on_non_ring_board_puzzles(golden):
    """Plain rollouts (reward + terminated) of single puzzles that fit one machine word but not the
    ring condition run k_steps_board with the action tiles of run_board_lean (uint32 and uint64
    boards): against the oracle, with invalid actions, odd step counts and irregular start states."""
    from pushworld_b200.engine import Engine

    rng = np.random.default_rng(808)
    picked = []
    for key in golden.keys():
        if not key.startswith("benchmark/"):
            continue
        rec = golden.rec(key)
        bits = 1 + (rec["height"] - 2) * (rec["width"] - 1)
        if bits <= 64 and bits + rec["width"] - 1 > 32 and rec["n"] <= 8:
            picked.append((bits, key))
    picked.sort()
    assert len(picked) >= 6 and picked[0][0] <= 32 < picked[-1][0], picked  # both word sizes
    for bits, key in picked[:3] + picked[-3:]:
        text = golden.text(key)
        oracle = po.OraclePuzzle(text)
        engine = Engine([text])
        n = 64
        states = []
        for e in range(n):
            st = oracle.initial_state
            for _ in range(int(rng.integers(0, 25))):
                st = oracle.get_next_state(st, int(rng.integers(4)))
            if e % 13 == 7 and oracle.num_objects >= 3:  # irregular: object 2 dropped onto object 1
                st = st[:2] + (st[1],) + st[3:]
            states.append(st)
        for T in (5, 32, 71):
            actions = rng.integers(0, 5, size=(T, n)).astype(np.uint8)  # 4 = invalid -> no-op
            xy = engine.state_from_tuples(states)
            reward = torch.empty((T, n), dtype=torch.float32, device="cuda")
            term = torch.empty((T, n), dtype=torch.uint8, device="cuda")
            counters = torch.zeros(4, dtype=torch.int64, device="cuda")
            engine.rollout(xy, 0, torch.from_numpy(actions).cuda(), reward=reward, terminated=term,
                           counters=counters)
            final, r, te = xy.cpu().numpy(), reward.cpu().numpy(), term.cpu().numpy()
            n_term = n_moved = n_gain = 0
            for e in range(n):
                st = states[e]
                for t in range(T):
                    a = int(actions[t, e])
                    nxt = st if a > 3 else oracle.get_next_state(st, a)
                    rr, tt, _ = po.step_reward(oracle, st, nxt)
                    assert np.float32(rr) == r[t, e] and tt == bool(te[t, e]), (key, T, e, t)
                    n_term += tt
                    n_moved += nxt != st
                    n_gain += max(oracle.count_achieved_goals(nxt) - oracle.count_achieved_goals(st), 0)
                    st = nxt
                got = tuple((int(final[0, j, e]), int(final[1, j, e])) for j in range(oracle.num_objects))
                assert got == st, (key, T, e)
            assert counters.cpu().numpy().tolist() == [n * T, n_term, n_moved, n_gain], (key, T)
            assert engine.poll_invalid_actions() == int((actions > 3).sum())


def test_plain_rollouts_on_the_general_rule(setup, golden):
    """Plain rollouts (reward + terminated only, whole warps) take run_lean_sim with the table rule:
    the golden walks of all 223 benchmark puzzles in one mixed launch (224 environments), and
    single-puzzle launches of a few large puzzles (shared-memory tables), step by step against the
    reference's recorded trajectories."""
    engine, keys = setup
    bench = [k for k in keys if k.startswith("benchmark/")]
    bench = bench + [bench[0]]  # 224 = 7 warps
    ids = [keys.index(k) for k in bench]
    T = 300
    actions = np.stack([golden.arr(k, "walk_actions") for k in bench], axis=1)
    n = len(bench)
    puzzle_id = torch.tensor(ids, dtype=torch.int32, device="cuda")
    xy = engine.new_state(n)
    engine.reset(xy, puzzle_id)
    reward = torch.empty((T, n), dtype=torch.float32, device="cuda")
    term = torch.empty((T, n), dtype=torch.uint8, device="cuda")
    counters = torch.zeros(4, dtype=torch.int64, device="cuda")
    engine.rollout(xy, puzzle_id, torch.from_numpy(actions).cuda(), reward=reward, terminated=term,
                   counters=counters)
    final = _pack([golden.arr(k, "walk_states").astype(np.int32)[-1:] for k in bench], engine.n_max)
    assert np.array_equal(xy.cpu().numpy(), final)
    ach = np.stack([golden.arr(k, "walk_achieved").astype(np.int32) for k in bench], axis=1)
    isg = np.stack([golden.arr(k, "walk_isgoal") for k in bench], axis=1)
    dg = ach[1:] - ach[:-1]
    assert np.array_equal(term.cpu().numpy().astype(bool), isg[1:])
    assert np.array_equal(reward.cpu().numpy(), _expected_reward(dg, isg[1:]))
    c = counters.cpu().numpy()
    moved_steps = sum(
        int((np.diff(golden.arr(k, "walk_states").astype(np.int32), axis=0) != 0).any(axis=(1, 2)).sum())
        for k in bench
    )
    assert c.tolist() == [T * n, int(isg[1:].sum()), moved_steps, int(np.maximum(dg, 0).sum())]
    # single-puzzle launches: 32 copies of the same walk
    for key in ["benchmark/level4/Mind The Gap", "benchmark/level3/Crow Pulling", "benchmark/level2/Pull Dont Push"]:
        p = keys.index(key)
        acts = np.repeat(golden.arr(key, "walk_actions")[:, None], 32, axis=1)
        xy1 = engine.new_state(32)
        engine.reset(xy1, p)
        r1 = torch.empty((T, 32), dtype=torch.float32, device="cuda")
        t1 = torch.empty((T, 32), dtype=torch.uint8, device="cuda")
        engine.rollout(xy1, p, torch.from_numpy(np.ascontiguousarray(acts)).cuda(), reward=r1, terminated=t1)
        e = bench.index(key)
        assert np.array_equal(r1.cpu().numpy(), np.repeat(reward[:, e:e + 1].cpu().numpy(), 32, axis=1)), key
        assert np.array_equal(t1.cpu().numpy(), np.repeat(term[:, e:e + 1].cpu().numpy(), 32, axis=1)), key
        assert np.array_equal(xy1.cpu().numpy()[:, :, 0], xy.cpu().numpy()[:, :, e]), key


def test_out_of_range_puzzle_ids_are_counted_not_dereferenced(setup):
    """Direct Engine / C-ABI callers can pass any int32 as a puzzle id: the kernels run such an
    environment on puzzle 0 and count it (status word 2) instead of reading the tables out of bounds."""
    engine, keys = setup
    n = 96
    pid = torch.zeros(n, dtype=torch.int32, device="cuda")
    pid[5], pid[40], pid[77] = -3, len(keys), 2**30
    xy = engine.new_state(n)
    engine.reset(xy, pid)
    ok = engine.new_state(n)
    engine.reset(ok, torch.zeros(n, dtype=torch.int32, device="cuda"))
    assert torch.equal(xy, ok)
    assert engine.poll_bad_puzzle_ids() == 3
    actions = torch.randint(0, 4, (40, n), dtype=torch.uint8, device="cuda")
    reward = torch.empty((40, n), dtype=torch.float32, device="cuda")
    term = torch.empty((40, n), dtype=torch.uint8, device="cuda")
    engine.rollout(xy, pid, actions, reward=reward, terminated=term)
    engine.step(xy, pid, actions[0])
    engine.expand(xy, pid)
    engine.render(xy, pid)
    torch.cuda.synchronize()
    assert engine.poll_bad_puzzle_ids() == 12
    assert engine.poll_bad_puzzle_ids() == 0
