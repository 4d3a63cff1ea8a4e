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
        n_max = self.engine.n_max
        self.observation_shape = (self.max_cell_height * pixels_per_cell,
                                  self.max_cell_width * pixels_per_cell, 3)
        # Everything a step returns besides the observation lives in ONE small device buffer --
        # state (int32 [2, n_max, 1]) | reward f32 | steps i32 | dgoals i8 | terminated u8 | truncated u8 --
        # so that a step reads it back with one copy into pinned host memory; the observation is
        # rendered into a second buffer with its own pinned twin.  One stream synchronisation per step.
        state_bytes = 8 * n_max
        self._meta_dev = torch.zeros(state_bytes + 12, dtype=torch.uint8, device=dev)
        self._meta_host = torch.zeros(state_bytes + 12, dtype=torch.uint8).pin_memory()
        self._xy = self._meta_dev[:state_bytes].view(torch.int32).view(2, n_max, 1)
        self._steps_dev = self._meta_dev[state_bytes + 4:state_bytes + 8].view(torch.int32)
        from .engine import StepOutput

        self._out = StepOutput(
            reward=self._meta_dev[state_bytes:state_bytes + 4].view(torch.float32),
            dgoals=self._meta_dev[state_bytes + 8:state_bytes + 9].view(torch.int8),
            terminated=self._meta_dev[state_bytes + 9:state_bytes + 10],
            truncated=self._meta_dev[state_bytes + 10:state_bytes + 11],
        )
        self._state_bytes = state_bytes
        self._meta_np = self._meta_host.numpy()
        self._obs_dev = torch.empty((1,) + self.observation_shape, dtype=torch.float32, device=dev)
        self._obs_host = torch.empty(self.observation_shape, dtype=torch.float32).pin_memory()
        self._pid = torch.zeros(1, dtype=torch.int32, device=dev)
        self._action = torch.zeros(1, dtype=torch.uint8, device=dev)
        self._action_host = torch.zeros(1, dtype=torch.uint8).pin_memory()
        self._graph = None       # CUDA graph of one step: action in, step + render, both read-backs
        self._graph_failed = False
        self._stream = torch.cuda.Stream(dev)

    @property
    def current_puzzle(self) -> Optional[PushWorldPuzzle]:
        return None if self.current_index is None else self.puzzles[self.current_index]

    def _pad(self) -> Tuple[int, int]:
        return (self.max_cell_height, self.max_cell_width)

    def _state_tuple(self) -> State:
        """The current state from the pinned copy of the meta buffer (no device access)."""
        n = self.puzzles[self.current_index].num_movables
        n_max = self.engine.n_max
        xy = self._meta_np[:self._state_bytes].view(np.int32)
        return tuple((int(xy[j]), int(xy[n_max + j])) for j in range(n))

    def reset(self, seed: Optional[int] = None) -> np.ndarray:
        if seed is not None:
            self.rng = random.Random(seed)
        self.current_index = self.rng.choice(range(len(self.puzzles)))  # gym_env.py:174
        with torch.cuda.stream(self._stream):
            self._pid.fill_(self.current_index)
            self.engine.reset(self._xy, self._pid, steps=self._steps_dev)
            self.engine.render(self._xy, self._pid, dtype="float32", pixels_per_cell=self.pixels_per_cell,
                               border_width=self.border_width, pad_cells=self._pad(), out=self._obs_dev)
            self._meta_host.copy_(self._meta_dev, non_blocking=True)
            self._obs_host.copy_(self._obs_dev[0], non_blocking=True)
        self._stream.synchronize()
        self.steps = 0
        self.current_state = self._state_tuple()
        return self._obs_host.numpy().copy()

    def _enqueue_step(self) -> None:
        """Action in, step + render (pw_step_render), both read-backs: all asynchronous."""
        self._action.copy_(self._action_host, non_blocking=True)
        self.engine.step_render(
            self._xy, self._pid, self._action, dtype="float32",
            pixels_per_cell=self.pixels_per_cell, border_width=self.border_width,
            pad_cells=self._pad(), steps=self._steps_dev, max_steps=self.max_steps,
            out=self._obs_dev, step_out=self._out)
        self._meta_host.copy_(self._meta_dev, non_blocking=True)
        self._obs_host.copy_(self._obs_dev[0], non_blocking=True)

    def _capture(self) -> None:
        """Captures one step as a CUDA graph (replayed by every later step: one launch call)."""
        try:
            with torch.cuda.stream(self._stream):
                self._enqueue_step()  # warm-up outside the capture (lazy initialisation, allocator)
            self._stream.synchronize()
            # the warm-up advanced the environment: put the state and the step counter back
            self._meta_dev.copy_(self._saved_meta)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=self._stream):
                self._enqueue_step()
            self._stream.synchronize()
            self._meta_dev.copy_(self._saved_meta)
            torch.cuda.current_stream(self.engine.device).synchronize()
            self._graph = graph
        except Exception:  # capture not possible here: keep launching the same work directly (still on the GPU)
            self._graph_failed = True
            self._meta_dev.copy_(self._saved_meta)
            torch.cuda.current_stream(self.engine.device).synchronize()

    def step(self, action: int):
        """-> (observation, reward (python float), terminated, truncated)"""
        if self.current_state is None:
            raise RuntimeError("reset() must be called before step() can be called.")
        self.steps += 1
        self._action_host[0] = int(action)
        if self._graph is None and not self._graph_failed:
            self._saved_meta = self._meta_dev.clone()
            torch.cuda.current_stream(self.engine.device).synchronize()
            self._capture()
        with torch.cuda.stream(self._stream):
            if self._graph is not None:
                self._graph.replay()
            else:
                self._enqueue_step()
        self._stream.synchronize()
        sb = self._state_bytes
        terminated = bool(self._meta_np[sb + 9])
        # gym_env.py:212-221 in python floats, from the kernel's exact integer goal delta
        reward = 10.0 if terminated else int(self._meta_np[sb + 8:sb + 9].view(np.int8)[0]) - 0.01
        truncated = bool(self._meta_np[sb + 10])
        self.current_state = self._state_tuple()
        return self._obs_host.numpy().copy(), reward, terminated, truncated

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
