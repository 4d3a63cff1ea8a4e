"""Size-independent properties checked at BASELINE's FULL sizes (where the oracle would take
hours) and metamorphic checks under the 8 grid symmetries:

* config 2 (2^18 Road Block environments x 1024 steps): a rollout split into chunks equals the
  rollout in one launch; permuting environments permutes the outputs; counters equal reductions of
  the per-step outputs; the ring-board kernel and the edge-mask board kernel agree on all 2.7e8
  steps; a sample of columns equals the oracle;
* config 4 (expansion): successors of an action-major expansion equal single steps;
* symmetries (python3/src/pushworld/transform.py + test_transform.py restated): rotating /
  flipping a puzzle and mapping the actions accordingly gives the rotated / flipped trajectory.
"""

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import pushworld_oracle as po  # noqa: E402


def test_config2_full_size_rollout_properties():
    from pushworld_b200 import data
    from pushworld_b200.engine import Engine

    text = data.puzzle_text("level1/Road Block")
    engine = Engine([text])
    n, T = 1 << 18, 1024
    gen = torch.Generator(device="cuda").manual_seed(1234)
    actions = torch.randint(0, 4, (T, n), dtype=torch.uint8, device="cuda", generator=gen)

    def run(acts, chunks):
        xy = engine.new_state(acts.shape[1])
        engine.reset(xy, 0)
        reward = torch.empty(acts.shape, dtype=torch.float32, device="cuda")
        term = torch.empty(acts.shape, dtype=torch.uint8, device="cuda")
        counters = torch.zeros(4, dtype=torch.int64, device="cuda")
        lo = 0
        for c in chunks:
            engine.rollout(xy, 0, acts[lo:lo + c], reward=reward[lo:lo + c], terminated=term[lo:lo + c],
                           counters=counters)
            lo += c
        assert lo == acts.shape[0]
        return xy, reward, term, counters

    xy_a, rew_a, term_a, cnt_a = run(actions, [T])
    # 1. composition: chunked rollouts (uneven chunks, incl. T=1 launches) == one launch
    xy_b, rew_b, term_b, cnt_b = run(actions, [1, 3, 508, 1, 511])
    assert torch.equal(xy_a, xy_b) and torch.equal(rew_a, rew_b) and torch.equal(term_a, term_b)
    assert torch.equal(cnt_a, cnt_b)
    del xy_b, rew_b, term_b
    # 2. counters are reductions of the outputs
    assert int(cnt_a[0]) == n * T
    assert int(cnt_a[1]) == int(term_a.sum(dtype=torch.int64))
    # goals gained = transitions into the (single-goal) terminated state
    gained = term_a[0].to(torch.int64).sum() + ((term_a[1:] == 1) & (term_a[:-1] == 0)).sum(dtype=torch.int64)
    assert int(cnt_a[3]) == int(gained)
    # rewards only take the reference's values
    vals = torch.unique(rew_a).cpu().numpy()
    assert set(vals.tolist()) <= {float(np.float32(-0.01)), float(np.float32(0.99)),
                                  float(np.float32(-1.01)), 10.0}
    # 2b. two implementations, one answer: asking for one more output (dgoals) takes the launch off
    # the ring-board kernel and onto the edge-mask board kernel (pw_board.cuh: board_step)
    xy_c = engine.new_state(n)
    engine.reset(xy_c, 0)
    rew_c = torch.empty_like(rew_a)
    term_c = torch.empty_like(term_a)
    dg_c = torch.empty((T, n), dtype=torch.int8, device="cuda")
    engine.rollout(xy_c, 0, actions, reward=rew_c, terminated=term_c, dgoals=dg_c)
    assert torch.equal(xy_a, xy_c) and torch.equal(rew_a, rew_c) and torch.equal(term_a, term_c)
    assert int(cnt_a[3]) == int(dg_c.clamp(min=0).sum(dtype=torch.int64))
    del xy_c, rew_c, term_c, dg_c
    # 3. permutation equivariance on a slice of the batch
    m = 1 << 14
    perm = torch.randperm(m, device="cuda", generator=gen)
    xy_p, rew_p, term_p, _ = run(actions[:, :m][:, perm].contiguous(), [T])
    assert torch.equal(rew_p, rew_a[:, :m][:, perm]) and torch.equal(term_p, term_a[:, :m][:, perm])
    assert torch.equal(xy_p, xy_a[:, :, :m][:, :, perm])
    # 4. 64 sampled columns against the oracle, all 1024 steps
    oracle = po.OraclePuzzle(text)
    cols = np.linspace(0, n - 1, 64).astype(np.int64)
    a = actions[:, cols].cpu().numpy()
    r = rew_a[:, cols].cpu().numpy()
    te = term_a[:, cols].cpu().numpy()
    final = xy_a[:, :, cols].cpu().numpy()
    for c in range(len(cols)):
        state = oracle.initial_state
        for t in range(T):
            nxt = oracle.get_next_state(state, int(a[t, c]))
            rr, tt, _ = po.step_reward(oracle, state, nxt)
            assert np.float32(rr) == r[t, c] and tt == bool(te[t, c]), (c, t)
            state = nxt
        got = tuple((int(final[0, j, c]), int(final[1, j, c])) for j in range(oracle.num_objects))
        assert got == state


def test_config4_expansion_equals_single_steps():
    from pushworld_b200 import data
    from pushworld_b200.engine import Engine

    engine = Engine([data.puzzle_text("level2/Pull Dont Push")])
    n = 1_000_000
    xy = engine.new_state(n)
    engine.reset(xy, 0)
    gen = torch.Generator(device="cuda").manual_seed(1236)
    engine.rollout(xy, 0, torch.randint(0, 4, (64, n), dtype=torch.uint8, device="cuda", generator=gen))
    xy_out, moved, goal = engine.expand(xy, 0)
    for a in range(4):
        stepped = xy.clone()
        out = engine.step(stepped, 0, torch.full((n,), a, dtype=torch.uint8, device="cuda"), want_moved=True)
        assert torch.equal(xy_out[:, :, a * n:(a + 1) * n], stepped)
        assert torch.equal(moved[a * n:(a + 1) * n], out.moved)
        assert torch.equal(goal[a * n:(a + 1) * n], out.terminated)


def _transform(text, rotations, flipped):
    """rot90 clockwise `rotations` times, after an optional top-bottom flip (transform.py:21-48)."""
    rows = [line.split() for line in text.splitlines() if line.strip()]
    grid = np.array(rows, dtype=object)
    if flipped:
        grid = grid[::-1]
    grid = np.rot90(grid, k=-rotations)
    return "\n".join("  ".join(r) for r in grid.tolist()) + "\n"


def _map_action(a, rotations, flipped):
    # test_transform.py:44-55: flip swaps UP/DOWN; a clockwise quarter turn maps U->R->D->L->U
    flip = {po.UP: po.DOWN, po.DOWN: po.UP, po.LEFT: po.LEFT, po.RIGHT: po.RIGHT}
    rot = {po.UP: po.RIGHT, po.RIGHT: po.DOWN, po.DOWN: po.LEFT, po.LEFT: po.UP}
    if flipped:
        a = flip[a]
    for _ in range(rotations):
        a = rot[a]
    return a


@pytest.mark.parametrize("key", ["level1/Many Small Tools", "level2/Pull Dont Push", "level3/Yin Yang"])
def test_grid_symmetries_commute_with_the_dynamics(key):
    """Transforming the puzzle and the actions must transform rewards / termination identically:
    the per-step reward and terminated flags of all 8 symmetric variants are equal."""
    from pushworld_b200 import data
    from pushworld_b200.engine import Engine

    text = data.puzzle_text(key)
    variants = [(r, f) for f in (False, True) for r in range(4)]
    texts = [_transform(text, r, f) for r, f in variants]
    engine = Engine(texts)
    n, T = 256, 300
    rng = np.random.default_rng(9)
    base_actions = rng.integers(0, 4, size=(T, n))
    outs = []
    for v, (r, f) in enumerate(variants):
        mapped = np.vectorize(lambda a: _map_action(int(a), r, f))(base_actions).astype(np.uint8)
        xy = engine.new_state(n)
        engine.reset(xy, v)
        reward = torch.empty((T, n), dtype=torch.float32, device="cuda")
        term = torch.empty((T, n), dtype=torch.uint8, device="cuda")
        moved_any = torch.zeros(4, dtype=torch.int64, device="cuda")
        engine.rollout(xy, v, torch.from_numpy(mapped).cuda(), reward=reward, terminated=term,
                       counters=moved_any)
        outs.append((reward, term, moved_any))
    for reward, term, counters in outs[1:]:
        assert torch.equal(reward, outs[0][0]) and torch.equal(term, outs[0][1])
        assert torch.equal(counters, outs[0][2])
