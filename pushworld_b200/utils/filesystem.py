"""Puzzle file discovery with the reference's semantics (python3/src/pushworld/utils/
filesystem.py:21-61,132-159): a path is either one file with the right extension or a
directory walked with ``os.walk``; the walk order defines the puzzle order that
``PushWorldEnv.reset`` samples from."""

import os
from typing import Dict, Iterator


def iter_files_with_extension(file_or_directory_path: str, extension: str) -> Iterator[str]:
    extension = extension.lower()
    path = file_or_directory_path.rstrip(os.path.sep)
    if os.path.isfile(path):
        if not path.lower().endswith(extension):
            raise ValueError(
                f"The given file does not have the expected extension ({extension}): {path}"
            )
        yield path
        return
    for parent, _, filenames in os.walk(path):
        for filename in filenames:
            if filename.lower().endswith(extension):
                yield os.path.join(parent, filename)


def get_puzzle_file_paths(puzzle_path: str, extension: str = ".pwp") -> Dict[str, str]:
    """Puzzle name -> file path; raises ValueError on duplicate names."""
    found: Dict[str, str] = {}
    for path in iter_files_with_extension(puzzle_path, extension):
        name = os.path.splitext(os.path.basename(path))[0]
        if name in found:
            raise ValueError(f'Two puzzles have the same name "{name}": {path} and {found[name]}')
        found[name] = path
    return found
