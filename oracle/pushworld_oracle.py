"""CPU oracle for the PushWorld hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module.  Nothing under
``pushworld_b200/`` imports it: the product path is CUDA-only and fails loudly
when the extension is missing.

This is a restatement (not a copy) of the algorithm in the reference
``python3/src/pushworld/puzzle.py`` using *direct cell-set geometry* instead of
the reference's precomputed collision hash tables.  The equivalence

    rel in _movable_collision_map[a][i][j]
        <=> overlaps(C_i + p_i + d, C_j + p_j) and not overlaps(C_i + p_i, C_j + p_j)

follows from ``_populate_dynamic_collisions`` (puzzle.py:567-593) and
``_populate_static_collisions`` (puzzle.py:522-564; the latter additionally
requires the object's bounding box to lie inside the grid, puzzle.py:557-561).

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the real
reference from ``/root/reference/python3/src`` and this oracle is checked
against (a) the reference's own known-answer tests (test_puzzle.py,
test_gym_env.py) and (b) golden vectors produced by the reference for all 223
benchmark puzzles (solution replays + seeded random walks + renders), committed
under ``tests/golden``.

Reference citations are ``file:line`` relative to ``/root/reference/``.
"""

from __future__ import annotations

import random
from typing import Dict, FrozenSet, Iterable, List, Optional, Sequence, Tuple

import numpy as np

Point = Tuple[int, int]
State = Tuple[Point, ...]

NUM_ACTIONS = 4
LEFT, RIGHT, UP, DOWN = range(NUM_ACTIONS)  # puzzle.py:35
# puzzle.py:43-50 -- y grows downwards, so UP is -y.
DISPLACEMENTS: Tuple[Point, ...] = ((-1, 0), (1, 0), (0, -1), (0, 1))
ACTION_FROM_CHAR = {"L": LEFT, "R": RIGHT, "U": UP, "D": DOWN}  # puzzle.py:37-42

# puzzle.py:65-79
COLOR_AGENT = (0x00, 0xDC, 0x00)
COLOR_AGENT_BORDER = (0x00, 0x6E, 0x00)
COLOR_AGENT_WALL = (0xFA, 0xC7, 0x1E)
COLOR_AGENT_WALL_BORDER = (0x7D, 0x64, 0x0F)
COLOR_GOAL_BORDER = (0xB9, 0x00, 0x00)
COLOR_GOAL_OBJECT = (0xDC, 0x00, 0x00)
COLOR_GOAL_OBJECT_BORDER = (0x6E, 0x00, 0x00)
COLOR_MOVABLE = (0x46, 0x9B, 0xFF)
COLOR_MOVABLE_BORDER = (0x23, 0x48, 0x7F)
COLOR_WALL = (0x0A, 0x0A, 0x0A)
COLOR_WALL_BORDER = (0x05, 0x05, 0x05)


def _shift(cells: Iterable[Point], dx: int, dy: int) -> FrozenSet[Point]:
    return frozenset((x + dx, y + dy) for x, y in cells)


class OraclePuzzle:
    """Geometry-based restatement of ``PushWorldPuzzle`` (puzzle.py:100-506)."""

    def __init__(self, text: str) -> None:
        # ---- parse (puzzle.py:131-157) ----
        cells: Dict[str, set] = {}
        n_cols = None
        n_rows = 0
        # the reference iterates over file lines; a line is whatever "for line in
        # file" yields, i.e. split on "\n" keeping a possibly empty last piece out.
        lines = text.split("\n")
        if lines and lines[-1] == "":
            lines = lines[:-1]
        for r, line in enumerate(lines):
            tokens = line.split()
            if n_cols is None:
                n_cols = len(tokens)
            elif len(tokens) != n_cols:
                raise ValueError(
                    f"Row {r + 1} does not have the same number of elements as "
                    "the first row."
                )
            n_rows = r + 1
            for c, token in enumerate(tokens):
                for name in token.split("+"):
                    name = name.lower()
                    if name != ".":
                        cells.setdefault(name, set()).add((c + 1, r + 1))
        if "a" not in cells:
            raise ValueError("Every puzzle must have an agent object, indicated by 'a'.")

        # ---- border walls (puzzle.py:160-168) ----
        self.width = n_cols + 2
        self.height = n_rows + 2
        wall = cells.setdefault("w", set())
        for x in range(self.width):
            wall.add((x, 0))
            wall.add((x, self.height - 1))
        for y in range(self.height):
            wall.add((0, y))
            wall.add((self.width - 1, y))

        # ---- object order (puzzle.py:170-237) ----
        names_desc = sorted(cells.keys(), reverse=True)  # python string order
        origin: Dict[str, Point] = {}
        shape: Dict[str, FrozenSet[Point]] = {}
        for name in names_desc:
            if name in ("w", "aw"):
                origin[name] = (0, 0)  # puzzle.py:184-185
            else:
                origin[name] = (
                    min(x for x, _ in cells[name]),
                    min(y for _, y in cells[name]),
                )
            ox, oy = origin[name]
            shape[name] = _shift(cells[name], -ox, -oy)

        order = ["a"]
        goal_names: List[str] = []
        for name in names_desc:
            if name[0] == "g":
                movable = "m" + name[1:]
                assert movable in cells, f"Goal has no associated movable object: {movable}"
                goal_names.append(name)
                order.append(movable)
        for name in cells:  # first-appearance order (dict insertion), puzzle.py:235-237
            if name[0] == "m" and name not in order:
                order.append(name)

        self.object_names: List[str] = order
        self.goal_names: List[str] = goal_names
        self.num_objects = len(order)
        self.num_goals = len(goal_names)
        self.shapes: List[FrozenSet[Point]] = [shape[name] for name in order]
        self.goal_shapes: List[FrozenSet[Point]] = [shape[name] for name in goal_names]
        self.initial_state: State = tuple(origin[name] for name in order)
        self.goal_state: Tuple[Point, ...] = tuple(origin[name] for name in goal_names)
        self.has_agent_walls = "aw" in cells
        self.walls: FrozenSet[Point] = frozenset(cells["w"])
        # puzzle.py:273 -- the agent is stopped by aw UNION w.
        self.agent_walls: FrozenSet[Point] = frozenset(cells.get("aw", set())) | self.walls
        self._sizes = [
            (max(x for x, _ in s) + 1, max(y for _, y in s) + 1) for s in self.shapes
        ]

    # ------------------------------------------------------------------ helpers
    @property
    def dimensions(self) -> Tuple[int, int]:
        return (self.width, self.height)  # puzzle.py:328-330

    def _placed(self, idx: int, pos: Point) -> FrozenSet[Point]:
        return _shift(self.shapes[idx], pos[0], pos[1])

    def _statically_blocked(self, idx: int, pos: Point, d: Point, obstacles) -> bool:
        """puzzle.py:522-564: a position is a collision position iff the bounding
        box is inside the grid, the moved object hits an obstacle cell and the
        object does not already overlap the obstacles."""
        w, h = self._sizes[idx]
        if not (0 <= pos[0] <= self.width - w and 0 <= pos[1] <= self.height - h):
            return False
        here = self._placed(idx, pos)
        there = _shift(here, d[0], d[1])
        return bool(there & obstacles) and not (here & obstacles)

    def _pushes(self, i: int, pi: Point, j: int, pj: Point, d: Point) -> bool:
        """puzzle.py:567-593 evaluated directly."""
        ci = self._placed(i, pi)
        cj = self._placed(j, pj)
        return bool(_shift(ci, d[0], d[1]) & cj) and not (ci & cj)

    # ------------------------------------------------------------------ dynamics
    def get_next_state_and_moved(self, state: State, action: int):
        """puzzle.py:348-394.  Returns (next_state, moved_indices) where
        moved_indices mirrors the C++ ``RelativeState.moved_object_indices``
        (pushworld_puzzle.cc:439-452): [] if nothing moved, else [0, pushed...]
        in ascending index order."""
        d = DISPLACEMENTS[action]
        if self._statically_blocked(0, state[0], d, self.agent_walls):
            return state, []
        pushed = [False] * self.num_objects
        pushed[0] = True
        frontier = [0]
        while frontier:
            i = frontier.pop()
            for j in range(1, self.num_objects):
                if pushed[j]:
                    continue
                if not self._pushes(i, state[i], j, state[j], d):
                    continue
                if self._statically_blocked(j, state[j], d, self.walls):
                    return state, []  # transitive stopping, puzzle.py:376-379
                pushed[j] = True
                frontier.append(j)
        nxt = tuple(
            (p[0] + d[0], p[1] + d[1]) if pushed[k] else p for k, p in enumerate(state)
        )
        return nxt, [k for k in range(self.num_objects) if pushed[k]]

    def get_next_state(self, state: State, action: int) -> State:
        return self.get_next_state_and_moved(state, action)[0]

    def count_achieved_goals(self, state: State) -> int:  # puzzle.py:396-407
        return sum(
            1 for k, g in enumerate(self.goal_state) if tuple(state[1 + k]) == tuple(g)
        )

    def is_goal_state(self, state: State) -> bool:  # puzzle.py:409-411
        return tuple(map(tuple, state[1 : 1 + self.num_goals])) == self.goal_state

    def is_valid_plan(self, plan: Iterable[int]) -> bool:  # puzzle.py:413-424
        state = self.initial_state
        for action in plan:
            if self.is_goal_state(state):
                return False
            state = self.get_next_state(state, action)
        return self.is_goal_state(state)

    # ------------------------------------------------------------------ rendering
    def _layers(self, state: State):
        """Bottom-to-top layer list of puzzle.py:453-458 as
        (absolute cell set, fill colour or None, border colour)."""
        layers = []
        if self.has_agent_walls:
            # Quirk pinned by test_puzzle.py:264-271: the agent-wall object's cell
            # set is the same set object that puzzle.py:273 later extends with all
            # wall cells, so this layer covers aw UNION w.
            layers.append((self.agent_walls, COLOR_AGENT_WALL, COLOR_AGENT_WALL_BORDER))
        layers.append((self.walls, COLOR_WALL, COLOR_WALL_BORDER))
        for k in range(self.num_objects):
            if k == 0:
                fill, border = COLOR_AGENT, COLOR_AGENT_BORDER
            elif k <= self.num_goals:
                fill, border = COLOR_GOAL_OBJECT, COLOR_GOAL_OBJECT_BORDER
            else:
                fill, border = COLOR_MOVABLE, COLOR_MOVABLE_BORDER
            layers.append((self._placed(k, tuple(state[k])), fill, border))
        for k in range(self.num_goals):
            layers.append(
                (_shift(self.goal_shapes[k], *self.goal_state[k]), None, COLOR_GOAL_BORDER)
            )
        return layers

    def render(self, state: State, border_width: int = 2, pixels_per_cell: int = 20) -> np.ndarray:
        """puzzle.py:426-469 + _draw_object (puzzle.py:596-638), restated per cell:
        a cell's fill is painted first, then a border strip for each of the 8
        neighbours that is not part of the same object."""
        if border_width < 1:
            raise ValueError("border_width must be >= 1")
        if pixels_per_cell < 1 + 2 * border_width:
            raise ValueError("pixels_per_cell must be >= 1 + 2*border_width")
        ppc, bw = pixels_per_cell, border_width
        img = np.full((self.height * ppc, self.width * ppc, 3), 255, np.uint8)
        for cellset, fill, border in self._layers(state):
            for cx, cy in cellset:
                x0, y0 = cx * ppc, cy * ppc
                if x0 < 0 or y0 < 0:
                    # numpy negative-index slicing in the reference would wrap;
                    # reachable states never get here (objects stay in the grid).
                    raise ValueError("cell outside the image")
                if fill is not None:
                    img[y0 : y0 + ppc, x0 : x0 + ppc] = fill
                for ny in (-1, 0, 1):
                    for nx in (-1, 0, 1):
                        if (nx or ny) and (cx + nx, cy + ny) not in cellset:
                            ys = y0 + (ppc - bw if ny > 0 else 0)
                            ye = ys + (ppc if ny == 0 else bw)
                            xs = x0 + (ppc - bw if nx > 0 else 0)
                            xe = xs + (ppc if nx == 0 else bw)
                            img[ys:ye, xs:xe] = border
        return img

    def render_plan(self, plan, border_width: int = 2, pixels_per_cell: int = 20):
        state = self.initial_state
        images = [self.render(state, border_width, pixels_per_cell)]
        for action in plan:
            state = self.get_next_state(state, action)
            images.append(self.render(state, border_width, pixels_per_cell))
        return images


def render_observation_padded(
    puzzle: OraclePuzzle,
    state: State,
    max_cell_height: int,
    max_cell_width: int,
    pixels_per_cell: int,
    border_width: int,
) -> np.ndarray:
    """utils/env_utils.py:44-91: uint8/255 as float32, zero padded and centred
    (the top/left side gets floor(pad/2) pixels)."""
    img = puzzle.render(state, border_width, pixels_per_cell).astype(np.float32) / 255
    pad_h = max_cell_height * pixels_per_cell - img.shape[0]
    pad_w = max_cell_width * pixels_per_cell - img.shape[1]
    out = np.zeros((img.shape[0] + pad_h, img.shape[1] + pad_w, 3), np.float32)
    out[pad_h // 2 : pad_h // 2 + img.shape[0], pad_w // 2 : pad_w // 2 + img.shape[1]] = img
    return out


class OracleEnv:
    """Step/reset logic of gym_env.py:150-226 (dm_env.py:150-234 is the same body)."""

    def __init__(
        self,
        puzzles: Sequence[OraclePuzzle],
        max_steps: Optional[int] = None,
        border_width: int = 2,
        pixels_per_cell: int = 20,
        pad_cells: Optional[Tuple[int, int]] = None,
        render: bool = True,
    ) -> None:
        self.puzzles = list(puzzles)
        self.max_steps = max_steps
        self.bw = border_width
        self.ppc = pixels_per_cell
        if pad_cells is None:
            pad_cells = (
                max(p.height for p in self.puzzles),
                max(p.width for p in self.puzzles),
            )
        self.pad_cells = pad_cells
        self.do_render = render
        self.rng = random.Random(123)  # gym_env.py:109
        self.puzzle: Optional[OraclePuzzle] = None
        self.state: Optional[State] = None
        self.steps = 0

    def _obs(self):
        if not self.do_render:
            return None
        return render_observation_padded(
            self.puzzle, self.state, self.pad_cells[0], self.pad_cells[1], self.ppc, self.bw
        )

    def reset(self, seed: Optional[int] = None):
        if seed is not None:
            self.rng = random.Random(seed)
        self.puzzle = self.rng.choice(self.puzzles)
        self.state = self.puzzle.initial_state
        self.steps = 0
        return self._obs(), {"puzzle_state": self.state}

    def step(self, action: int):
        if action not in (0, 1, 2, 3):
            raise ValueError("The provided action is not in the action space.")
        if self.state is None:
            raise RuntimeError("reset() must be called before step() can be called.")
        self.steps += 1
        prev = self.state
        self.state = self.puzzle.get_next_state(prev, action)
        obs = self._obs()
        terminated = self.puzzle.is_goal_state(self.state)
        if terminated:
            reward = 10.0
        else:
            reward = (
                self.puzzle.count_achieved_goals(self.state)
                - self.puzzle.count_achieved_goals(prev)
                - 0.01
            )
        truncated = False if self.max_steps is None else self.steps >= self.max_steps
        return obs, reward, terminated, truncated, {"puzzle_state": self.state}


def step_reward(puzzle: OraclePuzzle, prev: State, nxt: State) -> Tuple[float, bool, int]:
    """(reward, terminated, delta_goals) as in gym_env.py:210-221."""
    terminated = puzzle.is_goal_state(nxt)
    delta = puzzle.count_achieved_goals(nxt) - puzzle.count_achieved_goals(prev)
    return (10.0 if terminated else delta - 0.01), terminated, delta
