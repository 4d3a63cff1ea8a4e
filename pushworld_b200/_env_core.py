"""Shared body of the two single-environment drop-ins (``gym_env.PushWorldEnv`` and
``dm_env.PushWorldEnv``): the reference duplicates this logic (python3/src/pushworld/
gym_env.py:57-226 and dm_env.py:60-234); here both wrappers delegate to one class that drives
the CUDA engine with a batch of ONE environment (one step + render call)."""

from __future__ import annotations

import random
from typing import List, Optional, Tuple

import numpy as np
import torch

from . import data as _data
from .engine import Engine
from .puzzle import DEFAULT_BORDER_WIDTH, DEFAULT_PIXELS_PER_CELL, NUM_ACTIONS, PushWorldPuzzle, State
from .utils.filesystem import iter_files_with_extension

PUZZLE_EXTENSION = ".pwp"


class SingleEnvCore:
    def __init__(self, puzzle_path: str, max_steps: Optional[int], border_width: int,
                 pixels_per_cell: int, standard_padding: bool, device=None) -> None:
        texts: List[str] = []
        for path in iter_files_with_extension(puzzle_path, PUZZLE_EXTENSION):  # gym_env.py:66-69
            with open(path, "r") as f:
                texts.append(f.read())
        if len(texts) == 0:
            raise ValueError(f"No PushWorld puzzles found in: {puzzle_path}")
        if border_width < 1:
            raise ValueError("border_width must be >= 1")
        if pixels_per_cell < 3:
            raise ValueError("pixels_per_cell must be >= 3")
        self.engine = Engine(texts, device=device)
        self.puzzles = [PushWorldPuzzle(engine=self.engine, index=i) for i in range(len(texts))]
        self.max_steps = max_steps
        self.pixels_per_cell = pixels_per_cell
        self.border_width = border_width
        self.max_cell_width = self.engine.max_width
        self.max_cell_height = self.engine.max_height
        if standard_padding:  # gym_env.py:86-105
            std_h, std_w = _data.max_puzzle_dimensions()
            if std_h < self.max_cell_height:
                raise ValueError(
                    "`standard_padding` is True, but the maximum puzzle height in "
                    "BENCHMARK_PUZZLES_PATH is less than the height of the puzzle(s) "
                    "in the given `puzzle_path`.")
            if std_w < self.max_cell_width:
                raise ValueError(
                    "`standard_padding` is True, but the maximum puzzle width in "
                    "BENCHMARK_PUZZLES_PATH is less than the width of the puzzle(s) "
                    "in the given `puzzle_path`.")
            self.max_cell_height, self.max_cell_width = std_h, std_w
        self.rng = random.Random(123)  # gym_env.py:107-109
        self.current_index: Optional[int] = None
        self.current_state: Optional[State] = None
        self.steps = 0
        dev = self.engine.device
        self._xy = self.engine.new_state(1)
        self._pid = torch.zeros(1, dtype=torch.int32, device=dev)
        self._action = torch.zeros(1, dtype=torch.uint8, device=dev)
        self._steps_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.observation_shape = (self.max_cell_height * pixels_per_cell,
                                  self.max_cell_width * pixels_per_cell, 3)

    @property
    def current_puzzle(self) -> Optional[PushWorldPuzzle]:
        return None if self.current_index is None else self.puzzles[self.current_index]

    def _pad(self) -> Tuple[int, int]:
        return (self.max_cell_height, self.max_cell_width)

    def _state_tuple(self) -> State:
        n = self.puzzles[self.current_index].num_movables
        host = self._xy.cpu().numpy()
        return tuple((int(host[0, j, 0]), int(host[1, j, 0])) for j in range(n))

    def reset(self, seed: Optional[int] = None) -> np.ndarray:
        if seed is not None:
            self.rng = random.Random(seed)
        self.current_index = self.rng.choice(range(len(self.puzzles)))  # gym_env.py:174
        self._pid.fill_(self.current_index)
        self.engine.reset(self._xy, self._pid, steps=self._steps_dev)
        self.steps = 0
        self.current_state = self._state_tuple()
        obs = self.engine.render(self._xy, self._pid, dtype="float32",
                                 pixels_per_cell=self.pixels_per_cell,
                                 border_width=self.border_width, pad_cells=self._pad())
        return obs[0].cpu().numpy()

    def step(self, action: int):
        """-> (observation, reward (python float), terminated, truncated)"""
        if self.current_state is None:
            raise RuntimeError("reset() must be called before step() can be called.")
        self.steps += 1
        self._action.fill_(int(action))
        obs, out = self.engine.step_render(
            self._xy, self._pid, self._action, dtype="float32",
            pixels_per_cell=self.pixels_per_cell, border_width=self.border_width,
            pad_cells=self._pad(), steps=self._steps_dev, max_steps=self.max_steps)
        terminated = bool(out.terminated.item())
        # gym_env.py:212-221 in python floats, from the kernel's exact integer goal delta
        reward = 10.0 if terminated else int(out.dgoals.item()) - 0.01
        truncated = bool(out.truncated.item())
        self.current_state = self._state_tuple()
        return obs[0].cpu().numpy(), reward, terminated, truncated

    def render_uint8(self) -> np.ndarray:
        return self.puzzles[self.current_index].render(
            self.current_state, border_width=self.border_width,
            pixels_per_cell=self.pixels_per_cell)


def valid_action(action) -> bool:
    if isinstance(action, (bool, np.bool_)):
        return False
    if isinstance(action, (int, np.integer)):
        return 0 <= int(action) < NUM_ACTIONS
    if isinstance(action, np.ndarray) and action.shape == () and np.issubdtype(action.dtype, np.integer):
        return 0 <= int(action) < NUM_ACTIONS
    return False
