"""Constants mirroring python3/src/pushworld/config.py:20-33 for the parts this package covers.
``BENCHMARK_PUZZLES_PATH`` is materialised from the bundled data on first use."""

from . import data as _data

PUZZLE_EXTENSION = ".pwp"


def benchmark_puzzles_path() -> str:
    return _data.materialize()
