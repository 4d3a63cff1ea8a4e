"""PushWorldVectorEnv (the batched public API) against N independent reference-semantics
environments simulated with the oracle: observations, rewards, terminated / truncated flags, masked
resets, action validation, puzzle grouping, and the host-buffer rollout used by bench.py's e2e leg."""

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import pushworld_oracle as po  # noqa: E402


NAMES = ["trivial", "multiple_goals", "transitive_pushing", "trivial_obstacle", "trivial_tool"]


def _texts(golden):
    return [golden.text("python3/" + n) for n in NAMES]


def test_step_with_observations_matches_oracle_envs(golden):
    from pushworld_b200.vector_env import PushWorldVectorEnv

    texts = _texts(golden)
    n = 23
    env = PushWorldVectorEnv(texts, n, max_steps=7, observation="float32", pixels_per_cell=6,
                             border_width=1)
    assert env.observation_shape == (8 * 6, 11 * 6, 3)  # max puzzle dims: 8 rows x 11 columns
    # default assignment: environment e plays puzzle e % P, slots grouped by puzzle
    pids = env.puzzle_id.cpu().numpy()
    assert sorted(pids.tolist()) == pids.tolist()
    assert np.array_equal(pids, (env.env_index.cpu().numpy() % len(texts)))
    oracles = [po.OraclePuzzle(t) for t in texts]
    refs = [po.OracleEnv([oracles[p]], max_steps=7, border_width=1, pixels_per_cell=6, pad_cells=(8, 11))
            for p in pids]
    with pytest.raises(RuntimeError):
        env.step(torch.zeros(n, dtype=torch.uint8, device="cuda"))
    obs = env.reset()
    for e, ref in enumerate(refs):
        want, _ = ref.reset()
        assert np.array_equal(obs[e].cpu().numpy(), want)
    rng = np.random.default_rng(1)
    for t in range(12):
        actions = rng.integers(0, 4, size=n)
        obs, reward, terminated, truncated, info = env.step(actions)
        obs, reward = obs.cpu().numpy(), reward.cpu().numpy()
        terminated, truncated = terminated.cpu().numpy(), truncated.cpu().numpy()
        for e, ref in enumerate(refs):
            want_obs, r, te, tr, _ = ref.step(int(actions[e]))
            assert np.array_equal(obs[e], want_obs), (t, e)
            assert reward[e] == np.float32(r) and bool(terminated[e]) == te and bool(truncated[e]) == tr
    states = env.states()
    for e, ref in enumerate(refs):
        assert states[e] == ref.state
    # masked reset: only the truncated environments start over
    mask = torch.from_numpy(truncated.astype(np.uint8)).cuda()
    env.reset(mask=mask)
    states = env.states()
    for e, ref in enumerate(refs):
        want = ref.puzzle.initial_state if truncated[e] else ref.state
        assert states[e] == want
    with pytest.raises(ValueError, match="not in the action space"):
        env.step(np.full(n, 4))


def test_uniform_puzzle_rollout_and_host_pipeline(golden):
    from pushworld_b200.vector_env import PushWorldVectorEnv

    text = golden.text("benchmark/level1/Road Block")
    n, T = 512, 200
    env = PushWorldVectorEnv([text], n, puzzle_ids=0)
    env.reset()
    gen = torch.Generator(device="cuda").manual_seed(3)
    actions = torch.randint(0, 4, (T, n), dtype=torch.uint8, device="cuda", generator=gen)
    out = env.rollout(actions, want_return=True)
    final_a = env.xy.clone()
    # the same rollout through pinned host buffers (3-stream pipeline, chunks that do not divide T)
    env.reset()
    h_act = actions.cpu().pin_memory()
    h_rew = torch.empty((T, n), dtype=torch.float32).pin_memory()
    h_term = torch.empty((T, n), dtype=torch.uint8).pin_memory()
    env.rollout_host(h_act, h_rew, h_term, chunk_steps=48)
    assert torch.equal(env.xy, final_a)
    assert torch.equal(h_rew, out.reward.cpu()) and torch.equal(h_term, out.terminated.cpu())
    assert int(out.counters[0].item()) == n * T
    assert int(out.counters[1].item()) == int(out.terminated.sum().item())
    # oracle on a few columns
    oracle = po.OraclePuzzle(text)
    a = actions.cpu().numpy()
    for e in range(0, n, 97):
        state = oracle.initial_state
        total = np.float32(0)
        for t in range(T):
            nxt = oracle.get_next_state(state, int(a[t, e]))
            r, te, _ = po.step_reward(oracle, state, nxt)
            assert np.float32(r) == h_rew[t, e].item() and te == bool(h_term[t, e].item())
            total = np.float32(total + np.float32(r))
            state = nxt
        assert out.return_sum[e].item() == total


def test_constructor_errors(golden, tmp_path):
    from pushworld_b200.vector_env import PushWorldVectorEnv

    texts = _texts(golden)
    with pytest.raises(ValueError):
        PushWorldVectorEnv(str(tmp_path), 4)  # no puzzles (gym_env.py:71-72)
    with pytest.raises(ValueError):
        PushWorldVectorEnv(texts, 4, border_width=0)
    with pytest.raises(ValueError):
        PushWorldVectorEnv(texts, 4, pixels_per_cell=2)
    with pytest.raises(ValueError):
        PushWorldVectorEnv(texts, 4, puzzle_ids=[0, 1, 2, 99])
    env = PushWorldVectorEnv(texts, 4, standard_padding=True, observation="uint8", pixels_per_cell=3,
                             border_width=1)
    assert env.observation_shape == (54 * 3, 47 * 3, 3)  # benchmark maximum (env_utils.py:25-41)
    assert env.reset().shape == (4, 162, 141, 3)
    # a directory of .pwp files works like PushWorldEnv's puzzle_path
    for i, t in enumerate(texts):
        (tmp_path / f"p{i}.pwp").write_text(t)
    env = PushWorldVectorEnv(str(tmp_path), 10)
    assert env.engine.num_puzzles == len(texts)


def test_rollout_host_returns_matches_dense_rollout(golden):
    from pushworld_b200.vector_env import PushWorldVectorEnv

    text = golden.text("benchmark/level1/Road Block")
    n, T = 300, 150
    env = PushWorldVectorEnv([text], n, puzzle_ids=0)
    env.reset()
    gen = torch.Generator(device="cuda").manual_seed(8)
    actions = torch.randint(0, 4, (T, n), dtype=torch.uint8, device="cuda", generator=gen)
    dense = env.rollout(actions, want_return=True)
    final = env.xy.clone()
    env.reset()
    h_act = actions.cpu().pin_memory()
    h_ret = torch.empty(n, dtype=torch.float32).pin_memory()
    h_nt = torch.empty(n, dtype=torch.int32).pin_memory()
    env.rollout_host_returns(h_act, h_ret, h_nt, chunk_steps=32)
    assert torch.equal(env.xy, final)
    assert torch.equal(h_nt.to(torch.int64), dense.terminated.sum(dim=0, dtype=torch.int64).cpu())
    # float32 return sums: each 32-step chunk launch sums its rewards in step order and adds the
    # partial sum to the running total
    acc = np.zeros(n, np.float32)
    r = dense.reward.cpu().numpy()
    for lo in range(0, T, 32):
        part = np.zeros(n, np.float32)
        for t in range(lo, min(T, lo + 32)):
            part = part + r[t]
        acc = acc + part
    assert np.array_equal(h_ret.numpy(), acc)


def test_prepared_step_equals_step(golden):
    """Engine.prepare_step (arguments bound once) and PushWorldVectorEnv(reuse_step_outputs=True) give
    the same results as the plain step path."""
    from pushworld_b200.vector_env import PushWorldVectorEnv

    text = golden.text("benchmark/level2/Pull Dont Push")
    n, T = 64, 40
    rng = np.random.default_rng(5)
    actions = rng.integers(0, 4, size=(T, n)).astype(np.uint8)
    a = PushWorldVectorEnv([text], n, puzzle_ids=0, max_steps=25)
    b = PushWorldVectorEnv([text], n, puzzle_ids=0, max_steps=25, reuse_step_outputs=True)
    a.reset()
    b.reset()
    first = None
    for t in range(T):
        act = torch.from_numpy(actions[t]).cuda()
        _, r_a, te_a, tr_a, info_a = a.step(act)
        _, r_b, te_b, tr_b, info_b = b.step(act)
        assert torch.equal(r_a, r_b) and torch.equal(te_a, te_b) and torch.equal(tr_a, tr_b)
        assert torch.equal(info_a["dgoals"], info_b["dgoals"])
        if first is None:
            first = r_b
        assert r_b is first  # the same tensor every step
    assert torch.equal(a.xy, b.xy) and torch.equal(a.steps, b.steps)


def test_graphed_step_equals_plain_steps():
    """PushWorldVectorEnv.capture_step(): a step replayed as a CUDA graph gives the same rewards,
    flags, observations and states as step() (gym_env.py:188-226 semantics, incl. max_steps)."""
    import torch

    from pushworld_b200 import data
    from pushworld_b200.vector_env import PushWorldVectorEnv

    texts = [data.puzzle_text(k) for k in ("level1/Road Block", "level1/Many Small Tools", "level2/Pull Dont Push")]
    n = 96
    for observation in (None, "uint8"):
        a = PushWorldVectorEnv(texts, n, max_steps=7, observation=observation)
        b = PushWorldVectorEnv(texts, n, max_steps=7, observation=observation)
        a.reset()
        b.reset()
        step = b.capture_step()
        gen = torch.Generator(device="cuda").manual_seed(3)
        for t in range(12):
            actions = torch.randint(0, 4, (n,), dtype=torch.uint8, device="cuda", generator=gen)
            obs_a, rew_a, term_a, trunc_a, _ = a.step(actions)
            step.actions.copy_(actions)
            obs_b, rew_b, term_b, trunc_b = step.replay()
            torch.cuda.synchronize()
            assert torch.equal(rew_a, rew_b) and torch.equal(term_a, term_b) and torch.equal(trunc_a, trunc_b), t
            assert torch.equal(a.xy, b.xy) and torch.equal(a.steps, b.steps)
            if observation is not None:
                assert torch.equal(obs_a, obs_b)
        assert bool(trunc_a.any())
