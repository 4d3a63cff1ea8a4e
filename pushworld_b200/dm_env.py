"""Drop-in for ``pushworld.dm_env`` (python3/src/pushworld/dm_env.py): a ``dm_env.Environment``
whose ``step`` runs on the GPU; returns ``TimeStep``s with FIRST / MID / LAST step types."""

from __future__ import annotations

from typing import Optional

import numpy as np

from .compat import ensure_dm_env

dm_env = ensure_dm_env()
from dm_env import specs  # noqa: E402

from ._env_core import SingleEnvCore  # noqa: E402
from .puzzle import DEFAULT_BORDER_WIDTH, DEFAULT_PIXELS_PER_CELL, NUM_ACTIONS, PushWorldPuzzle, State  # noqa: E402


class PushWorldEnv(dm_env.Environment):
    """See ``pushworld.dm_env.PushWorldEnv`` (dm_env.py:35-58) for the argument semantics."""

    def __init__(self, puzzle_path: str, max_steps: Optional[int] = None,
                 border_width: int = DEFAULT_BORDER_WIDTH,
                 pixels_per_cell: int = DEFAULT_PIXELS_PER_CELL,
                 standard_padding: bool = False, device=None) -> None:
        self._core = SingleEnvCore(puzzle_path, max_steps, border_width, pixels_per_cell,
                                   standard_padding, device=device)
        self._action_space = specs.DiscreteArray(num_values=NUM_ACTIONS, dtype=int, name="action")
        self._observation_space = specs.BoundedArray(
            shape=self._core.observation_shape, dtype=np.float32, name="board",
            minimum=0.0, maximum=1.0)

    def observation_spec(self):
        return self._observation_space

    def action_spec(self):
        return self._action_space

    @property
    def current_puzzle(self) -> Optional[PushWorldPuzzle]:
        return self._core.current_puzzle

    @property
    def current_state(self) -> Optional[State]:
        return self._core.current_state

    def reset(self, seed: Optional[int] = None):
        return dm_env.restart(self._core.reset(seed))

    def step(self, action: int):
        try:
            self._action_space.validate(action)
        except ValueError:
            raise ValueError("The provided action is not in the action space.")
        observation, reward, terminated, truncated = self._core.step(int(action))
        if terminated or truncated:  # dm_env.py:230-234
            return dm_env.termination(reward, observation)
        return dm_env.transition(reward, observation)

    def render(self, mode="rgb_array") -> np.ndarray:
        assert mode == "rgb_array", "mode must be rgb_array."
        return self._core.render_uint8().astype(np.float32) / 255
