"""GPU parity tests of the observation path (pw_render / pw_step_render) through the C ABI:
byte-exact against sha256 digests of the real reference's PushWorldPuzzle.render and
render_observation_padded outputs (tests/golden/reference_renders.json) and against the
CPU oracle on seeded states."""

import hashlib
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import pushworld_oracle as po  # noqa: E402
from tests.conftest import GOLDEN  # noqa: E402


@pytest.fixture(scope="module")
def renders():
    with open(os.path.join(GOLDEN, "reference_renders.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def setup(golden, renders):
    from pushworld_b200.engine import Engine

    keys = sorted({c["key"] for c in renders["cases"]})
    return Engine([golden.text(k) for k in keys]), keys


def _sha(t):
    return hashlib.sha256(np.ascontiguousarray(t.cpu().numpy()).tobytes()).hexdigest()


def test_reference_test_rendering_hashes(setup, golden, renders):
    """python3/test/test_puzzle.py:249-271: trivial.pwp, plan RDRU, the five pinned hashes."""
    engine, keys = setup
    p = keys.index("python3/trivial")
    xy = engine.new_state(1)
    engine.reset(xy, p)
    frames = [engine.render(xy, p)[0].cpu().numpy()]
    for a in (po.RIGHT, po.DOWN, po.RIGHT, po.UP):
        engine.step(xy, p, torch.tensor([a], dtype=torch.uint8, device="cuda"))
        frames.append(engine.render(xy, p)[0].cpu().numpy())
    assert frames[0].shape == (100, 100, 3) and frames[0].dtype == np.uint8
    assert [hash(tuple(img.flat)) for img in frames] == [
        8141256401900123811,
        1770142108181252064,
        4744825492003518882,
        -7463149466192975143,
        -8235536721686713717,
    ]


def test_render_matches_reference_digests(setup, golden, renders):
    engine, keys = setup
    n_checked = 0
    for case in renders["cases"]:
        p = keys.index(case["key"])
        xy = engine.state_from_tuples([case["state"]])
        img = engine.render(xy, p, dtype="uint8", pixels_per_cell=case["ppc"], border_width=case["bw"])
        assert list(img.shape[1:]) == case["shape"]
        assert _sha(img[0]) == case["sha256_u8"], (case["key"], case["ppc"], case["bw"])
        for pad in case.get("padded", []):
            obs = engine.render(xy, p, dtype="float32", pixels_per_cell=case["ppc"],
                                border_width=case["bw"], pad_cells=pad["pad_cells"])
            assert list(obs.shape[1:]) == pad["shape"]
            assert _sha(obs[0]) == pad["sha256_f32"], (case["key"], case["ppc"], pad["pad_cells"])
            n_checked += 1
    assert n_checked > 200


def test_batched_mixed_render_matches_oracle(setup, golden):
    """One launch, several puzzles of different sizes padded to a common frame, both dtypes."""
    engine, keys = setup
    rng = np.random.default_rng(3)
    oracles = [po.OraclePuzzle(golden.text(k)) for k in keys]
    states, pids = [], []
    for p, o in enumerate(oracles):
        s = o.initial_state
        for _ in range(3):
            for _ in range(10):
                s = o.get_next_state(s, int(rng.integers(4)))
            states.append(s)
            pids.append(p)
    xy = engine.state_from_tuples(states)
    puzzle_id = torch.tensor(pids, dtype=torch.int32, device="cuda")
    pad = (engine.max_height + 1, engine.max_width + 2)
    for ppc, bw in ((5, 1), (6, 2)):
        f32 = engine.render(xy, puzzle_id, dtype="float32", pixels_per_cell=ppc, border_width=bw,
                            pad_cells=pad).cpu().numpy()
        u8 = engine.render(xy, puzzle_id, dtype="uint8", pixels_per_cell=ppc, border_width=bw,
                           pad_cells=pad).cpu().numpy()
        for e, (s, p) in enumerate(zip(states, pids)):
            want = po.render_observation_padded(oracles[p], s, pad[0], pad[1], ppc, bw)
            assert np.array_equal(f32[e], want), (keys[p], ppc)
            # uint8 padded frame: same pixels, integer valued
            assert np.array_equal(u8[e].astype(np.float32) / 255, want), (keys[p], ppc)


def test_step_render_fused_equals_step_then_render(setup, golden):
    engine, keys = setup
    rng = np.random.default_rng(5)
    for key in ("benchmark/level1/Road Block", "benchmark/level2/Clean Sweep", "python3/trivial_tool"):
        p = keys.index(key)
        n = 33
        xy_a = engine.new_state(n)
        engine.reset(xy_a, p)
        xy_b = xy_a.clone()
        steps_a = torch.zeros(n, dtype=torch.int32, device="cuda")
        steps_b = torch.zeros(n, dtype=torch.int32, device="cuda")
        for t in range(12):
            action = torch.from_numpy(rng.integers(0, 4, size=n, dtype=np.uint8)).cuda()
            obs, out = engine.step_render(xy_a, p, action, dtype="float32", steps=steps_a, max_steps=9)
            ref_out = engine.step(xy_b, p, action, steps=steps_b, max_steps=9)
            ref_obs = engine.render(xy_b, p, dtype="float32")
            assert torch.equal(xy_a, xy_b)
            assert torch.equal(obs, ref_obs)
            assert torch.equal(out.reward, ref_out.reward)
            assert torch.equal(out.terminated, ref_out.terminated)
            assert torch.equal(out.truncated, ref_out.truncated)
            assert torch.equal(out.dgoals, ref_out.dgoals)
        assert torch.equal(steps_a, steps_b)
    # uint8 fused output, mixed puzzles
    pids = torch.tensor([0, 1, 2, 3] * 4, dtype=torch.int32, device="cuda")
    xy = engine.new_state(16)
    engine.reset(xy, pids)
    xy2 = xy.clone()
    action = torch.from_numpy(rng.integers(0, 4, size=16, dtype=np.uint8)).cuda()
    obs, _ = engine.step_render(xy, pids, action, dtype="uint8", pixels_per_cell=7, border_width=3)
    engine.step(xy2, pids, action)
    assert torch.equal(obs, engine.render(xy2, pids, dtype="uint8", pixels_per_cell=7, border_width=3))


def test_render_argument_errors(setup):
    engine, keys = setup
    xy = engine.new_state(1)
    engine.reset(xy, 0)
    from pushworld_b200._lib import PushWorldNativeError

    with pytest.raises(PushWorldNativeError, match="border_width must be >= 1"):
        engine.render(xy, 0, border_width=0)  # puzzle.py:444-445
    with pytest.raises(PushWorldNativeError, match="pixels_per_cell must be >= 1"):
        engine.render(xy, 0, pixels_per_cell=4, border_width=2)  # puzzle.py:447-448
    with pytest.raises(PushWorldNativeError, match="smaller than the puzzle"):
        engine.render(xy, 0, pad_cells=(2, 2))
    with pytest.raises(ValueError):
        engine.render(xy, 0, dtype="float16")


def test_render_fuzz_alignment_paths(golden):
    """Random small puzzles, odd pixels_per_cell / border widths / paddings and batches whose frames
    start at odd byte offsets: exercises the 16-byte, 4-byte and byte-wise row copies and the
    persistent-CTA frame loop; every frame must equal the oracle's padded observation."""
    from pushworld_b200.engine import Engine

    rng = np.random.default_rng(404)
    texts = []
    while len(texts) < 6:
        w, h = int(rng.integers(2, 7)), int(rng.integers(2, 6))
        grid = [["." for _ in range(w)] for _ in range(h)]
        for name in ["a", "m0", "g0", "m1", "m1", "w", "aw", "m2", "m2"]:
            x, y = int(rng.integers(w)), int(rng.integers(h))
            if grid[y][x] == ".":
                grid[y][x] = name
        text = "\n".join(" ".join(r) for r in grid) + "\n"
        try:
            po.OraclePuzzle(text)
            texts.append(text)
        except (ValueError, AssertionError):
            pass
    engine = Engine(texts)
    oracles = [po.OraclePuzzle(t) for t in texts]
    n = 37
    pids = rng.integers(0, len(texts), size=n)
    states = []
    for p in pids:
        s = oracles[p].initial_state
        for _ in range(int(rng.integers(0, 12))):
            s = oracles[p].get_next_state(s, int(rng.integers(4)))
        states.append(s)
    xy = engine.state_from_tuples(states)
    puzzle_id = torch.tensor(pids, dtype=torch.int32, device="cuda")
    for ppc, bw, extra in ((3, 1, (0, 0)), (5, 2, (1, 0)), (7, 1, (0, 1)), (4, 1, (2, 3)), (9, 4, (1, 1))):
        pad = (engine.max_height + extra[0], engine.max_width + extra[1])
        for dtype in ("uint8", "float32"):
            got = engine.render(xy, puzzle_id, dtype=dtype, pixels_per_cell=ppc, border_width=bw,
                                pad_cells=pad).cpu().numpy()
            for e in range(n):
                want = po.render_observation_padded(oracles[pids[e]], states[e], pad[0], pad[1], ppc, bw)
                if dtype == "uint8":
                    assert np.array_equal(got[e].astype(np.float32) / 255, want), (ppc, bw, e)
                else:
                    assert np.array_equal(got[e], want), (ppc, bw, e)
