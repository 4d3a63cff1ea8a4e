"""Drop-in for ``pushworld.utils.env_utils`` (python3/src/pushworld/utils/env_utils.py)."""

from typing import Tuple

import numpy as np

from .. import data as _data


def get_max_puzzle_dimensions() -> Tuple[int, int]:
    """(max height, max width) of the benchmark puzzles incl. the wall border
    (env_utils.py:25-41); computed from the bundled benchmark data."""
    return _data.max_puzzle_dimensions()


def render_observation_padded(puzzle, state, max_cell_height: int, max_cell_width: int,
                              pixels_per_cell: int, border_width: int) -> np.ndarray:
    """float32 RGB observation in [0, 1], zero padded and centred to
    ``(max_cell_height*ppc, max_cell_width*ppc, 3)`` (env_utils.py:44-91); rendered by the
    pw_render kernel (``puzzle`` is a ``pushworld_b200.puzzle.PushWorldPuzzle``)."""
    return puzzle.render_padded(state, max_cell_height, max_cell_width, pixels_per_cell, border_width)
