"""Drop-in for ``pushworld.puzzle`` (python3/src/pushworld/puzzle.py) backed by the CUDA engine.

``PushWorldPuzzle`` keeps the reference's public surface -- ``initial_state``, ``goal_state``,
``dimensions``, ``wall_positions``, ``agent_wall_positions``, ``movable_objects``,
``get_next_state``, ``count_achieved_goals``, ``is_goal_state``, ``is_valid_plan``, ``render``,
``render_plan`` (puzzle.py:100-128) -- but every transition, goal count and image is computed by
the sm_100a kernels through the C ABI (a batch of one environment; plans are replayed as ONE
fused rollout launch).  There is no CPU implementation of the dynamics in this package: without a
CUDA device construction fails.

The reference's private collision tables (``_agent_collision_map`` ...) do not exist here: the
kernels test geometry directly.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, List, Optional, Set, Tuple

import numpy as np
import torch

from . import tables
from .engine import Engine

# puzzle.py:21-29
DEFAULT_BORDER_WIDTH = 2
DEFAULT_PIXELS_PER_CELL = 20
NUM_ACTIONS = 4
AGENT_IDX = 0

Point = Tuple[int, int]
State = Tuple[Point, ...]
Color = Tuple[int, int, int]


class Actions:
    """puzzle.py:32-50."""

    LEFT, RIGHT, UP, DOWN = range(NUM_ACTIONS)
    FROM_CHAR = {"L": LEFT, "R": RIGHT, "U": UP, "D": DOWN}
    DISPLACEMENTS = np.array([(-1, 0), (1, 0), (0, -1), (0, 1)])


def hex_to_rgb(hex_string: str) -> Color:
    return tuple(int(hex_string[i : i + 2], 16) for i in (0, 2, 4))


class Colors:
    """puzzle.py:65-79 (the kernels use the same values through tables.PALETTE)."""

    AGENT = tables.PALETTE[tables.PAL_AGENT]
    AGENT_BORDER = tables.PALETTE[tables.PAL_AGENT_BORDER]
    AGENT_WALL = tables.PALETTE[tables.PAL_AGENT_WALL]
    AGENT_WALL_BORDER = tables.PALETTE[tables.PAL_AGENT_WALL_BORDER]
    GOAL = None
    GOAL_BORDER = tables.PALETTE[tables.PAL_GOAL_BORDER]
    GOAL_OBJECT = tables.PALETTE[tables.PAL_GOAL_OBJECT]
    GOAL_OBJECT_BORDER = tables.PALETTE[tables.PAL_GOAL_OBJECT_BORDER]
    MOVABLE = tables.PALETTE[tables.PAL_MOVABLE]
    MOVABLE_BORDER = tables.PALETTE[tables.PAL_MOVABLE_BORDER]
    WALL = tables.PALETTE[tables.PAL_WALL]
    WALL_BORDER = tables.PALETTE[tables.PAL_WALL_BORDER]


@dataclass(frozen=True)
class PushWorldObject:
    """puzzle.py:82-97."""

    position: Point
    fill_color: Optional[Color]
    border_color: Color
    cells: Set[Point]


def _as_state(state) -> State:
    return tuple((int(p[0]), int(p[1])) for p in state)


class PushWorldPuzzle:
    """A puzzle in the PushWorld environment, stepped on the GPU.

    Args:
        file_path: path of a ``.pwp`` file (or pass ``text=`` with the file's content).
        device: CUDA device (default: current device).
    """

    def __init__(self, file_path: Optional[str] = None, *, text: Optional[str] = None,
                 device=None, engine: Optional[Engine] = None, index: int = 0) -> None:
        if engine is not None:
            self._engine, self._index = engine, index
        else:
            if text is None:
                if file_path is None:
                    raise ValueError("file_path or text is required")
                with open(file_path, "r") as f:
                    text = f.read()
            self._engine, self._index = Engine([text], device=device), 0
        p = self._engine.parsed[self._index]
        self._parsed = p
        self._width, self._height = p.width, p.height
        self.num_movables = p.num_objects
        self._initial_state: State = tuple(p.initial_state)
        self._goal_state: Tuple[Point, ...] = tuple(p.goal_state)
        n_goals = p.num_goals
        self._movable_objects = [
            PushWorldObject(
                position=p.initial_state[k],
                fill_color=Colors.AGENT if k == 0 else (Colors.GOAL_OBJECT if k <= n_goals else Colors.MOVABLE),
                border_color=Colors.AGENT_BORDER if k == 0 else (
                    Colors.GOAL_OBJECT_BORDER if k <= n_goals else Colors.MOVABLE_BORDER),
                cells=set(p.shapes[k]),
            )
            for k in range(p.num_objects)
        ]
        self._xy = self._engine.new_state(1)
        self._action = torch.zeros(1, dtype=torch.uint8, device=self._engine.device)

    # ---- properties (puzzle.py:313-346) ----------------------------------------------
    @property
    def initial_state(self) -> State:
        return self._initial_state

    @property
    def goal_state(self) -> Tuple[Point, ...]:
        return self._goal_state

    @property
    def dimensions(self) -> Tuple[int, int]:
        return (self._width, self._height)

    @property
    def wall_positions(self) -> Set[Point]:
        return set(self._parsed.wall_cells)

    @property
    def agent_wall_positions(self) -> Set[Point]:
        # after construction the reference's set holds aw UNION w (puzzle.py:253,273)
        return set(self._parsed.agent_wall_cells)

    @property
    def movable_objects(self) -> List[PushWorldObject]:
        return self._movable_objects

    @property
    def engine(self) -> Engine:
        return self._engine

    # ---- device helpers ------------------------------------------------------------------
    def _upload(self, state: State) -> None:
        if len(state) != self.num_movables:
            raise ValueError(f"a state of this puzzle has {self.num_movables} positions")
        host = np.zeros((2, self._engine.n_max, 1), dtype=np.int32)
        for j, (x, y) in enumerate(state):
            host[0, j, 0], host[1, j, 0] = x, y
        self._xy.copy_(torch.from_numpy(host))

    def _download(self) -> State:
        host = self._xy.cpu().numpy()
        return tuple((int(host[0, j, 0]), int(host[1, j, 0])) for j in range(self.num_movables))

    # ---- dynamics (puzzle.py:348-424) -----------------------------------------------------
    def get_next_state(self, state: State, action: int) -> State:
        """The state that results from performing ``action`` in ``state`` (pw_step, N = 1)."""
        if isinstance(action, (bool, np.bool_)) or int(action) != action or not 0 <= int(action) < NUM_ACTIONS:
            raise IndexError(f"action {action!r} is not one of 0..{NUM_ACTIONS - 1}")
        self._upload(_as_state(state))
        self._action.fill_(int(action))
        self._engine.step(self._xy, self._index, self._action)
        return self._download()

    def _step_outputs(self, state: State, action: int):
        """(next_state, reward, terminated, delta_goals) of one step; used by the env wrappers."""
        self._upload(_as_state(state))
        self._action.fill_(int(action))
        out = self._engine.step(self._xy, self._index, self._action)
        return (self._download(), bool(out.terminated.item()), int(out.dgoals.item()))

    def count_achieved_goals(self, state: State) -> int:
        """Number of goal objects on their goal positions (pw_goals, N = 1)."""
        self._upload(_as_state(state))
        achieved, _ = self._engine.goals(self._xy, self._index)
        return int(achieved.item())

    def is_goal_state(self, state: State) -> bool:
        self._upload(_as_state(state))
        _, is_goal = self._engine.goals(self._xy, self._index)
        return bool(is_goal.item())

    def is_valid_plan(self, plan: Iterable[int]) -> bool:
        """True iff the plan reaches the goal at its LAST step and not earlier
        (puzzle.py:413-424); replayed as one fused rollout launch."""
        plan = [int(a) for a in plan]
        if self.is_goal_state(self._initial_state):
            return len(plan) == 0
        if len(plan) == 0:
            return False
        dev = self._engine.device
        self._upload(self._initial_state)
        actions = torch.tensor(plan, dtype=torch.uint8, device=dev).reshape(-1, 1)
        term = torch.empty((len(plan), 1), dtype=torch.uint8, device=dev)
        self._engine.rollout(self._xy, self._index, actions, terminated=term)
        term = term[:, 0].cpu().numpy().astype(bool)
        return bool(term[-1]) and not term[:-1].any()

    # ---- rendering (puzzle.py:426-506) -------------------------------------------------
    def _check_render_args(self, border_width: int, pixels_per_cell: int) -> None:
        if border_width < 1:
            raise ValueError("border_width must be >= 1")
        if pixels_per_cell < 1 + 2 * border_width:
            raise ValueError("pixels_per_cell must be >= 1 + 2*border_width")

    def render(self, state: State, border_width: int = DEFAULT_BORDER_WIDTH,
               pixels_per_cell: int = DEFAULT_PIXELS_PER_CELL) -> np.ndarray:
        """RGB uint8 image ``(height*ppc, width*ppc, 3)`` of ``state`` (pw_render)."""
        self._check_render_args(border_width, pixels_per_cell)
        self._upload(_as_state(state))
        img = self._engine.render(self._xy, self._index, dtype="uint8",
                                  pixels_per_cell=pixels_per_cell, border_width=border_width)
        return img[0].cpu().numpy()

    def render_padded(self, state: State, max_cell_height: int, max_cell_width: int,
                      pixels_per_cell: int, border_width: int) -> np.ndarray:
        """float32 observation, zero padded and centred (utils/env_utils.py:44-91)."""
        self._check_render_args(border_width, pixels_per_cell)
        self._upload(_as_state(state))
        obs = self._engine.render(self._xy, self._index, dtype="float32",
                                  pixels_per_cell=pixels_per_cell, border_width=border_width,
                                  pad_cells=(max_cell_height, max_cell_width))
        return obs[0].cpu().numpy()

    def render_plan(self, plan: Iterable[int], border_width: int = DEFAULT_BORDER_WIDTH,
                    pixels_per_cell: int = DEFAULT_PIXELS_PER_CELL) -> List[np.ndarray]:
        """Images of the initial state and of the state after every action of ``plan``: the
        plan is expanded into a batch of len(plan)+1 states rendered by one launch."""
        self._check_render_args(border_width, pixels_per_cell)
        plan = [int(a) for a in plan]
        states = [self._initial_state]
        state = self._initial_state
        for a in plan:
            state = self.get_next_state(state, a)
            states.append(state)
        xy = self._engine.state_from_tuples(states)
        imgs = self._engine.render(xy, self._index, dtype="uint8", pixels_per_cell=pixels_per_cell,
                                   border_width=border_width).cpu().numpy()
        return [imgs[i] for i in range(len(states))]
