"""Drop-in for ``pushworld.gym_env`` (python3/src/pushworld/gym_env.py): an OpenAI-Gym (0.26 API)
environment whose ``step`` runs on the GPU.  Same constructor arguments, same return values
``(observation, reward, terminated, truncated, {"puzzle_state": state})``, same exceptions."""

from __future__ import annotations

from typing import Any, Dict, Optional, Tuple

import numpy as np

from .compat import ensure_gym

gym = ensure_gym()

from ._env_core import SingleEnvCore  # noqa: E402
from .puzzle import DEFAULT_BORDER_WIDTH, DEFAULT_PIXELS_PER_CELL, NUM_ACTIONS, PushWorldPuzzle  # noqa: E402


class PushWorldEnv(gym.Env):
    """See ``pushworld.gym_env.PushWorldEnv`` (gym_env.py:32-64) for the argument semantics."""

    def __init__(self, puzzle_path: str, max_steps: Optional[int] = None,
                 border_width: int = DEFAULT_BORDER_WIDTH,
                 pixels_per_cell: int = DEFAULT_PIXELS_PER_CELL,
                 standard_padding: bool = False, device=None) -> None:
        self._core = SingleEnvCore(puzzle_path, max_steps, border_width, pixels_per_cell,
                                   standard_padding, device=device)
        self._action_space = gym.spaces.Discrete(NUM_ACTIONS)
        self._observation_space = gym.spaces.Box(
            low=0.0, high=1.0, shape=self._core.observation_shape, dtype=np.float32)

    @property
    def action_space(self):
        return self._action_space

    @property
    def observation_space(self):
        return self._observation_space

    @property
    def metadata(self) -> Dict[str, Any]:
        return {"render_modes": ["rgb_array"]}

    @property
    def render_mode(self) -> str:
        return "rgb_array"

    @property
    def current_puzzle(self) -> Optional[PushWorldPuzzle]:
        return self._core.current_puzzle

    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None) -> Tuple[np.ndarray, dict]:
        observation = self._core.reset(seed)
        return observation, {"puzzle_state": self._core.current_state}

    def step(self, action: int):
        if not self._action_space.contains(action):
            raise ValueError("The provided action is not in the action space.")
        observation, reward, terminated, truncated = self._core.step(action)
        return observation, reward, terminated, truncated, {"puzzle_state": self._core.current_state}

    def render(self, mode="rgb_array") -> np.ndarray:
        assert mode == "rgb_array", "mode must be rgb_array."
        return self._core.render_uint8()
