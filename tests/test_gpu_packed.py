"""pw_rollout_packed: the one-byte-per-step records expand to exactly the arrays pw_rollout /
pw_rollout_episodes write (reward float32 bit for bit, terminated, truncated, dgoals), on every
kernel family: ring boards, edge-mask boards, per-thread boards, the general rule (uniform and
mixed), whole-warp and ragged batch sizes, with and without truncation / autoreset / episodes."""

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _actions(T, n, seed):
    gen = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randint(0, 4, (T, n), dtype=torch.uint8, device="cuda", generator=gen)


def _both(engine, puzzle, n, T, seed, max_steps=None, autoreset=False, use_steps=False):
    from pushworld_b200.engine import record_to_outputs

    actions = _actions(T, n, seed)
    outs = []
    for packed in (False, True):
        xy = engine.new_state(n)
        steps = torch.zeros(n, dtype=torch.int32, device="cuda") if use_steps else None
        engine.reset(xy, puzzle, steps=steps)
        counters = torch.zeros(4, dtype=torch.int64, device="cuda")
        if packed:
            record = torch.empty((T, n), dtype=torch.uint8, device="cuda")
            engine.rollout_packed(xy, puzzle, actions, record, steps=steps, max_steps=max_steps,
                                  counters=counters, autoreset=autoreset)
            reward, term, trunc, dg = record_to_outputs(record)
        else:
            reward = torch.empty((T, n), dtype=torch.float32, device="cuda")
            term = torch.empty((T, n), dtype=torch.uint8, device="cuda")
            lean = not use_steps and max_steps is None and not autoreset
            trunc = None if lean else torch.empty((T, n), dtype=torch.uint8, device="cuda")
            dg = None if lean else torch.empty((T, n), dtype=torch.int8, device="cuda")
            engine.rollout(xy, puzzle, actions, reward=reward, terminated=term, truncated=trunc,
                           dgoals=dg, steps=steps, max_steps=max_steps, counters=counters,
                           autoreset=autoreset)
        outs.append((xy, reward, term, trunc, dg, counters, steps))
    a, b = outs
    assert torch.equal(a[0], b[0])
    assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    if a[3] is not None:
        assert torch.equal(a[3], b[3]) and torch.equal(a[4], b[4])
    else:
        assert not bool(b[3].any())
    assert torch.equal(a[5], b[5])
    if use_steps:
        assert torch.equal(a[6], b[6])
    return b


@pytest.mark.parametrize("n", [4096, 4096 + 17])
@pytest.mark.parametrize("key", ["level1/Road Block", "level1/Obstacle Has Obstacle",
                                 "level1/Many Small Tools", "level3/Yin Yang"])
def test_packed_equals_plain_uniform(key, n):
    from pushworld_b200 import data
    from pushworld_b200.engine import Engine

    engine = Engine([data.puzzle_text(key)])
    out = _both(engine, 0, n, 200, 11)
    assert int(out[5][0]) == n * 200
    _both(engine, 0, n, 130, 12, max_steps=50, use_steps=True)
    _both(engine, 0, n, 130, 13, max_steps=40, autoreset=True, use_steps=True)


@pytest.mark.parametrize("n", [8192, 8192 + 5])
def test_packed_equals_plain_mixed_benchmark(n):
    from pushworld_b200 import data
    from pushworld_b200.engine import Engine

    keys = data.benchmark_keys()
    engine = Engine([data.puzzle_text(k) for k in keys])
    pid = (torch.arange(n, device="cuda") % len(keys)).to(torch.int32)
    _both(engine, pid, n, 96, 21)
    grouped = torch.sort(pid.long(), stable=True).values.to(torch.int32)
    _both(engine, grouped, n, 96, 22)
    _both(engine, grouped, n, 96, 23, max_steps=30, autoreset=True, use_steps=True)


def _small_puzzles(count, seed):
    rng = np.random.default_rng(seed)
    texts = []
    while len(texts) < count:
        cells = [["." for _ in range(5)] for _ in range(5)]
        free = [(x, y) for x in range(5) for y in range(5)]
        rng.shuffle(free)
        for name, (x, y) in zip(["a", "m0", "g0", "m1", "w", "w"], free):
            cells[y][x] = name
        texts.append("\n".join(" ".join(r) for r in cells) + "\n")
    return texts


@pytest.mark.parametrize("n", [4096, 4096 + 9])
def test_packed_equals_plain_board_mixed_and_episodes(n):
    from pushworld_b200.engine import Engine, record_to_outputs

    texts = _small_puzzles(64, 5)
    engine = Engine(texts)
    pid = torch.randint(0, len(texts), (n,), dtype=torch.int32, device="cuda",
                        generator=torch.Generator(device="cuda").manual_seed(3))
    _both(engine, pid, n, 128, 31)
    # episode handling on the device: packed records vs the separate arrays of pw_rollout_episodes
    T = 160
    actions = _actions(T, n, 32)
    res = []
    for packed in (False, True):
        xy = engine.new_state(n)
        ids = pid.clone()
        steps = torch.zeros(n, dtype=torch.int32, device="cuda")
        episode = torch.zeros(n, dtype=torch.int32, device="cuda")
        engine.reset(xy, ids, steps=steps)
        if packed:
            record = torch.empty((T, n), dtype=torch.uint8, device="cuda")
            engine.rollout_packed(xy, ids, actions, record, steps=steps, max_steps=25,
                                  episode=episode, seed=77)
            reward, term, trunc, _ = record_to_outputs(record)
        else:
            reward = torch.empty((T, n), dtype=torch.float32, device="cuda")
            term = torch.empty((T, n), dtype=torch.uint8, device="cuda")
            trunc = torch.empty((T, n), dtype=torch.uint8, device="cuda")
            engine.rollout_episodes(xy, ids, actions, steps, episode, seed=77, max_steps=25,
                                    reward=reward, terminated=term, truncated=trunc)
        res.append((xy, ids, steps, episode, reward, term, trunc))
    for a, b in zip(*res):
        assert torch.equal(a, b)
    assert bool(res[0][6].any()) and int(res[0][3].max()) >= 6


def test_mask_kernel_with_cta_wide_closure_equals_lock_step():
    """PW_MASK_CTA=1 (a measured alternative of the general-rule kernel, DESIGN.md 3.5: prefilter
    candidates of a whole CTA resolved by its first threads between two barriers) against the default
    lock-step loop: plain and packed outputs, grouped and interleaved puzzle ids, episodes."""
    import os

    from pushworld_b200 import data
    from pushworld_b200.engine import Engine

    keys = data.benchmark_keys()
    engine = Engine([data.puzzle_text(k) for k in keys])
    n = 4096
    pid = (torch.arange(n, device="cuda") % len(keys)).to(torch.int32)
    grouped = torch.sort(pid.long(), stable=True).values.to(torch.int32)
    cases = [(pid, 96, 41, {}), (grouped, 96, 42, {}),
             (grouped, 96, 43, dict(max_steps=30, autoreset=True, use_steps=True))]
    assert "PW_MASK_CTA" not in os.environ
    base = [_both(engine, ids, n, T, seed, **kw) for ids, T, seed, kw in cases]
    os.environ["PW_MASK_CTA"] = "1"
    try:
        alt = [_both(engine, ids, n, T, seed, **kw) for ids, T, seed, kw in cases]
    finally:
        del os.environ["PW_MASK_CTA"]
    for a, b in zip(base, alt):
        for x, y in zip(a, b):
            assert (x is None and y is None) or torch.equal(x, y)
