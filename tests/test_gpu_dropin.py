"""The reference's own wrapper / simulator tests restated against the drop-in classes
(pushworld_b200.puzzle / gym_env / dm_env), which run every step and image on the GPU.

Each test names the reference test it mirrors (python3/test/*.py).  Cases that poke the
reference's private collision tables are restated geometrically: e.g. test_agent_movement
injects entries into ``_agent_collision_map``; here the same situation is produced with an
``aw`` cell in the puzzle text.  Expected values are the reference's (also cross-checked against
tests/golden/reference_gym.json, produced by the unmodified reference)."""

import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from tests.conftest import GOLDEN  # noqa: E402


@pytest.fixture(scope="module")
def puzzle_dir(tmp_path_factory, golden):
    root = tmp_path_factory.mktemp("puzzles")
    for name, text in golden.fixtures["python3"].items():
        (root / f"{name}.pwp").write_text(text)
    return str(root)


def _path(puzzle_dir, name):
    return os.path.join(puzzle_dir, name + ".pwp")


# --------------------------------------------------------------------------- test_puzzle.py
def test_agent_movement(puzzle_dir):
    """test_puzzle.py:28-66."""
    from pushworld_b200.puzzle import Actions, PushWorldPuzzle

    puzzle = PushWorldPuzzle(_path(puzzle_dir, "agent_movement"))
    s0 = puzzle.initial_state
    assert puzzle.get_next_state(s0, Actions.LEFT)[0] == (1, 2)
    assert puzzle.get_next_state(s0, Actions.RIGHT)[0] == (3, 2)
    assert puzzle.get_next_state(s0, Actions.UP)[0] == (2, 1)
    assert puzzle.get_next_state(s0, Actions.DOWN)[0] == (2, 3)
    # "add a left agent wall" etc.: same situations, built with aw cells around the agent
    walled = {
        Actions.LEFT: " .  .  .\nAW  A  .\n .  .  .\n",
        Actions.RIGHT: " .  .  .\n .  A AW\n .  .  .\n",
        Actions.UP: " . AW  .\n .  A  .\n .  .  .\n",
        Actions.DOWN: " .  .  .\n .  A  .\n . AW  .\n",
    }
    for blocked, text in walled.items():
        p = PushWorldPuzzle(text=text)
        for a in range(4):
            moved = p.get_next_state(p.initial_state, a)[0] != (2, 2)
            assert moved == (a != blocked)


def test_pushing(puzzle_dir):
    """test_puzzle.py:69-80."""
    from pushworld_b200.puzzle import Actions, PushWorldPuzzle

    puzzle = PushWorldPuzzle(_path(puzzle_dir, "pushing"))
    assert puzzle.get_next_state(puzzle.initial_state, Actions.DOWN) == ((1, 2), (2, 1))
    s = puzzle.get_next_state(puzzle.initial_state, Actions.RIGHT)
    assert s == ((2, 1), (3, 1))
    s = puzzle.get_next_state(s, Actions.RIGHT)
    assert s == ((3, 1), (4, 1))
    assert puzzle.get_next_state(s, Actions.RIGHT) == ((3, 1), (4, 1))  # transitive stopping


def test_transitive_pushing(puzzle_dir):
    """test_puzzle.py:83-98."""
    from pushworld_b200.puzzle import Actions, PushWorldPuzzle

    puzzle = PushWorldPuzzle(_path(puzzle_dir, "transitive_pushing"))
    s0 = puzzle.initial_state
    assert puzzle.get_next_state(s0, Actions.DOWN) == ((1, 2), (5, 1), (3, 1))
    s = puzzle.get_next_state(s0, Actions.RIGHT)
    assert s == ((2, 1), (5, 1), (3, 1))
    s = puzzle.get_next_state(s, Actions.RIGHT)
    assert s == ((3, 1), (5, 1), (4, 1))
    s = puzzle.get_next_state(s, Actions.RIGHT)
    assert s == ((4, 1), (6, 1), (5, 1))
    s = puzzle.get_next_state(s, Actions.RIGHT)
    assert s == ((4, 1), (6, 1), (5, 1))
    assert puzzle.get_next_state(s, Actions.DOWN) == ((4, 2), (6, 1), (5, 1))


def test_goal_states(puzzle_dir):
    """test_puzzle.py:101-123."""
    from pushworld_b200.puzzle import PushWorldPuzzle

    puzzle = PushWorldPuzzle(_path(puzzle_dir, "is_goal_state"))
    cases = [
        (((5, 1), (3, 6), (2, 5)), True, 2),
        (((2, 8), (3, 6), (2, 5)), True, 2),
        (((1, 1), (3, 3), (2, 5)), False, 1),
        (((1, 1), (3, 6), (2, 2)), False, 1),
        (((1, 1), (3, 4), (1, 5)), False, 0),
    ]
    for state, is_goal, count in cases:
        assert puzzle.is_goal_state(state) == is_goal
        assert puzzle.count_achieved_goals(state) == count


def test_trivial_file_loading(puzzle_dir):
    """test_puzzle.py:126-194 (trajectory part; the private collision sets do not exist here)."""
    from pushworld_b200.puzzle import Actions, PushWorldPuzzle

    puzzle = PushWorldPuzzle(_path(puzzle_dir, "trivial"))
    assert puzzle.goal_state == ((3, 1),)
    assert puzzle.initial_state == ((1, 2), (2, 2))
    state = puzzle.initial_state
    assert not puzzle.is_goal_state(state)
    script = [
        (Actions.LEFT, ((1, 2), (2, 2)), False),
        (Actions.UP, ((1, 2), (2, 2)), False),
        (Actions.DOWN, ((1, 2), (2, 2)), False),  # agent wall
        (Actions.RIGHT, ((2, 2), (3, 2)), False),
        (Actions.RIGHT, ((2, 2), (3, 2)), False),  # transitive stopping
        (Actions.DOWN, ((2, 3), (3, 2)), False),
        (Actions.DOWN, ((2, 3), (3, 2)), False),
        (Actions.RIGHT, ((3, 3), (3, 2)), False),
        (Actions.RIGHT, ((3, 3), (3, 2)), False),
        (Actions.UP, ((3, 2), (3, 1)), True),
        (Actions.UP, ((3, 2), (3, 1)), True),
    ]
    for action, want, is_goal in script:
        state = puzzle.get_next_state(state, action)
        assert state == want
        assert puzzle.is_goal_state(state) == is_goal


def test_file_parsing(puzzle_dir):
    """test_puzzle.py:197-246 (public properties; object order (A, M4, M1, M0, M2, M3))."""
    from pushworld_b200.puzzle import PushWorldPuzzle

    puzzle = PushWorldPuzzle(_path(puzzle_dir, "file_parsing"))
    assert puzzle.dimensions == (12, 18)
    assert puzzle.goal_state == ((6, 5), (3, 4))
    assert puzzle.initial_state == ((1, 12), (6, 14), (1, 3), (4, 1), (2, 7), (3, 8))
    assert len(puzzle.movable_objects) == 6
    assert puzzle.movable_objects[4].cells == {(0, 0), (1, 0), (2, 0), (0, 1), (2, 1), (0, 2), (1, 2), (2, 2)}
    assert (9, 1) in puzzle.wall_positions and (0, 0) in puzzle.wall_positions
    assert (8, 8) in puzzle.agent_wall_positions and (0, 0) in puzzle.agent_wall_positions
    with pytest.raises(ValueError):
        PushWorldPuzzle(text=". .\n. . .\n")
    with pytest.raises(ValueError):
        PushWorldPuzzle(text=". m0\n. .\n")
    with pytest.raises(AssertionError):
        PushWorldPuzzle(text="a g1\n. m0\n")


def test_rendering(puzzle_dir):
    """test_puzzle.py:249-271, including the five pinned image hashes."""
    from pushworld_b200.puzzle import Actions, PushWorldPuzzle

    puzzle = PushWorldPuzzle(_path(puzzle_dir, "trivial"))
    initial_image = puzzle.render(puzzle.initial_state)
    assert initial_image.shape == (100, 100, 3)
    plan = [Actions.RIGHT, Actions.DOWN, Actions.RIGHT, Actions.UP]
    plan_images = puzzle.render_plan(plan)
    assert (initial_image == plan_images[0]).all()
    assert [hash(tuple(image.flat)) for image in plan_images] == [
        8141256401900123811,
        1770142108181252064,
        4744825492003518882,
        -7463149466192975143,
        -8235536721686713717,
    ]
    with pytest.raises(ValueError):
        puzzle.render(puzzle.initial_state, border_width=0)
    with pytest.raises(ValueError):
        puzzle.render(puzzle.initial_state, border_width=2, pixels_per_cell=4)
    assert puzzle.is_valid_plan(plan)
    assert not puzzle.is_valid_plan(plan + [Actions.UP])  # goal reached before the plan ended
    assert not puzzle.is_valid_plan(plan[:-1])


# --------------------------------------------------------------------------- test_gym_env.py
def _gym_env(*args, **kwargs):
    from pushworld_b200.gym_env import PushWorldEnv

    return PushWorldEnv(*args, **kwargs)


def test_gym_observations_and_renderings(puzzle_dir):
    """test_gym_env.py:28-44."""
    from pushworld_b200.puzzle import Actions

    env = _gym_env(_path(puzzle_dir, "trivial"))
    observation, info = env.reset()
    assert observation in env.observation_space
    image = env.current_puzzle.render(info["puzzle_state"])
    assert (image == observation * 255).all()
    assert (image == env.render()).all()
    observation, _, _, _, info = env.step(Actions.RIGHT)
    assert observation in env.observation_space
    image = env.current_puzzle.render(info["puzzle_state"])
    assert (image == observation * 255).all()
    assert (image == env.render()).all()


def test_gym_standard_padding(puzzle_dir):
    """test_gym_env.py:47-64."""
    env = _gym_env(_path(puzzle_dir, "trivial"), standard_padding=False)
    o1, _ = env.reset()
    assert o1 in env.observation_space
    env = _gym_env(_path(puzzle_dir, "trivial"), standard_padding=True)
    o2, _ = env.reset()
    assert o2 in env.observation_space
    assert o2.shape == (54 * 20, 47 * 20, 3)
    assert o1.shape != o2.shape
    assert np.count_nonzero(o1.sum(axis=0)) == np.count_nonzero(o2.sum(axis=0))


def test_gym_reward(puzzle_dir):
    """test_gym_env.py:67-84: -0.01 / 0.99 / -1.01 / 10 as python floats."""
    from pushworld_b200.puzzle import Actions

    env = _gym_env(_path(puzzle_dir, "multiple_goals"))
    env.reset()
    assert env.step(Actions.RIGHT)[1] == -0.01
    assert env.step(Actions.RIGHT)[1] == 1 + -0.01
    assert env.step(Actions.RIGHT)[1] == -1 + -0.01
    env.reset()
    for a in (Actions.RIGHT, Actions.RIGHT, Actions.LEFT, Actions.LEFT, Actions.LEFT):
        env.step(a)
    assert env.step(Actions.LEFT)[1] == 10


@pytest.mark.parametrize("standard_padding", [True, False])
def test_gym_all_goals_achieved(puzzle_dir, standard_padding):
    """test_gym_env.py:87-102."""
    from pushworld_b200.puzzle import Actions

    env = _gym_env(_path(puzzle_dir, "trivial"), standard_padding=standard_padding)
    env.reset()
    env.step(Actions.RIGHT)
    env.step(Actions.DOWN)
    env.step(Actions.RIGHT)
    obs, reward, terminated, truncated, _ = env.step(Actions.UP)
    assert obs in env.observation_space
    assert reward == 10 and terminated is True and truncated is False
    # stepping after termination keeps working (no implicit reset)
    _, reward, terminated, _, _ = env.step(Actions.LEFT)
    assert terminated is True and reward == 10


def test_gym_truncation_and_termination(puzzle_dir):
    """test_gym_env.py:105-152."""
    from pushworld_b200.puzzle import Actions

    env = _gym_env(_path(puzzle_dir, "transitive_pushing"), max_steps=3)
    env.reset()
    assert env.step(Actions.LEFT)[3] is False
    assert env.step(Actions.LEFT)[3] is False
    assert env.step(Actions.LEFT)[3] is True
    env = _gym_env(_path(puzzle_dir, "transitive_pushing"))
    env.reset()
    env.step(Actions.RIGHT)
    env.step(Actions.RIGHT)
    _, reward, terminated, truncated, _ = env.step(Actions.RIGHT)
    assert reward == 10 and terminated is True and truncated is False
    env = _gym_env(_path(puzzle_dir, "multiple_goals"))
    for action in (Actions.LEFT, Actions.RIGHT):
        env.reset()
        env.step(action)
        _, reward, terminated, _, _ = env.step(action)
        assert reward > 0 and terminated is False


def test_gym_reset_and_errors(puzzle_dir, tmp_path):
    """test_gym_env.py:155-169 + the error conditions of gym_env.py:71-76,195-199."""
    from pushworld_b200.puzzle import Actions

    env = _gym_env(_path(puzzle_dir, "trivial"))
    with pytest.raises(RuntimeError):
        env.step(Actions.RIGHT)
    o1, _ = env.reset()
    o2 = env.step(Actions.RIGHT)[0]
    o3, _ = env.reset()
    assert not (o1 == o2).all()
    assert (o1 == o3).all()
    for bad in (4, -1, 1.0, "1", None):
        with pytest.raises(ValueError):
            env.step(bad)
    env = _gym_env(puzzle_dir)
    initial_states = set(tuple(env.reset()[1]["puzzle_state"]) for _ in range(100))
    assert len(initial_states) > 1
    with pytest.raises(ValueError):
        _gym_env(str(tmp_path))  # no puzzles
    with pytest.raises(ValueError):
        _gym_env(_path(puzzle_dir, "trivial"), border_width=0)
    with pytest.raises(ValueError):
        _gym_env(_path(puzzle_dir, "trivial"), pixels_per_cell=2)


def test_gym_matches_unmodified_reference_sequences(puzzle_dir):
    """(reward, terminated, truncated, state) of the unmodified reference gym env on seeded
    action sequences (tests/golden/reference_gym.json)."""
    with open(os.path.join(GOLDEN, "reference_gym.json")) as f:
        cases = json.load(f)
    for case in cases:
        env = _gym_env(_path(puzzle_dir, case["puzzle"]), max_steps=case["max_steps"])
        env.reset()
        for a, (reward, term, trunc, state) in zip(case["actions"], case["out"]):
            _, r, te, tr, info = env.step(a)
            assert (r, te, tr) == (reward, term, trunc), (case["puzzle"], a)
            assert [list(p) for p in info["puzzle_state"]] == state


# --------------------------------------------------------------------------- test_dm_env.py
def test_dm_env(puzzle_dir):
    """test_dm_env.py: observations, rewards (TimeStep[1]), FIRST / MID / LAST."""
    import dm_env

    from pushworld_b200.dm_env import PushWorldEnv
    from pushworld_b200.puzzle import Actions

    env = PushWorldEnv(_path(puzzle_dir, "trivial"))
    observation = env.reset().observation
    env.observation_spec().validate(observation)
    image = env.current_puzzle.render(env.current_state)
    assert (image == observation * 255).all()
    assert (image == env.render() * 255).all()
    observation = env.step(Actions.RIGHT).observation
    image = env.current_puzzle.render(env.current_state)
    assert (image == observation * 255).all()

    env = PushWorldEnv(_path(puzzle_dir, "multiple_goals"))
    env.reset()
    assert env.step(Actions.RIGHT)[1] == -0.01
    assert env.step(Actions.RIGHT)[1] == 1 + -0.01
    assert env.step(Actions.RIGHT)[1] == -1 + -0.01

    env = PushWorldEnv(_path(puzzle_dir, "transitive_pushing"), max_steps=3)
    assert env.reset().step_type == dm_env.StepType.FIRST
    assert env.step(Actions.LEFT).step_type == dm_env.StepType.MID
    assert env.step(Actions.LEFT).step_type == dm_env.StepType.MID
    timestep = env.step(Actions.LEFT)
    assert timestep.step_type == dm_env.StepType.LAST and timestep.reward != 10.0

    env = PushWorldEnv(_path(puzzle_dir, "trivial"), standard_padding=True)
    env.reset()
    for a in (Actions.RIGHT, Actions.DOWN, Actions.RIGHT):
        env.step(a)
    timestep = env.step(Actions.UP)
    env.observation_spec().validate(timestep.observation)
    assert timestep.reward == 10.0 and timestep.step_type == dm_env.StepType.LAST
    with pytest.raises(ValueError):
        env.step(7)


# --------------------------------------------------------------------------- demo_gym_env.py
def test_demo_gym_env_plumbing():
    """BASELINE config 1: python3/scripts/demo_gym_env.py:31-48 headless -- reset + 10 random
    steps on `Road Block`; observation (100, 160, 3) float32 in [0, 1]."""
    from pushworld_b200 import data
    from pushworld_b200.gym_env import PushWorldEnv
    from pushworld_b200.puzzle import NUM_ACTIONS

    root = data.materialize()
    env = PushWorldEnv(os.path.join(root, "level1", "Road Block.pwp"))
    obs, info = env.reset()
    assert obs.shape == (100, 160, 3) and obs.dtype == np.float32
    rng = np.random.default_rng(0)
    for _ in range(10):
        obs, reward, terminated, truncated, info = env.step(int(rng.integers(NUM_ACTIONS)))
        assert obs.shape == (100, 160, 3) and 0.0 <= obs.min() and obs.max() <= 1.0
        assert isinstance(reward, float)
