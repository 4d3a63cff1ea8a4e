"""Host-side ``.pwp`` parser and table compiler.

Mirrors the parsing half of the reference's ``PushWorldPuzzle.__init__``
(python3/src/pushworld/puzzle.py:130-257): the object ORDER it produces is part of the
state contract.  Instead of the reference's collision hash tables (puzzle.py:259-308,
O(cells^2) to build) it emits the row-bitmask geometry the CUDA kernels test directly;
the binary layout is documented in ``include/pushworld_b200.h``.

This module is pure host logic (no device work); the native twin is
``pw_compile_puzzle`` in ``csrc/pw_compile.cpp`` and the two are tested to produce
identical blobs.
"""

from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Set, Tuple

import numpy as np

Point = Tuple[int, int]

MAGIC = 0x31545750
HEADER_WORDS = 16
FLAG_HAS_AW = 1
FLAG_REGULAR = 2
NO_GOAL = 0x80008000
MAX_OBJECTS = 32
MAX_WIDTH = 64
MAX_HEIGHT = 4096

# palette indices shared with csrc/pw_render.cuh; colours from puzzle.py:65-79
PALETTE = [
    (255, 255, 255),  # 0 background (puzzle.py:451)
    (0x00, 0xDC, 0x00),  # 1 AGENT
    (0x00, 0x6E, 0x00),  # 2 AGENT_BORDER
    (0xFA, 0xC7, 0x1E),  # 3 AGENT_WALL
    (0x7D, 0x64, 0x0F),  # 4 AGENT_WALL_BORDER
    (0xB9, 0x00, 0x00),  # 5 GOAL_BORDER
    (0xDC, 0x00, 0x00),  # 6 GOAL_OBJECT
    (0x6E, 0x00, 0x00),  # 7 GOAL_OBJECT_BORDER
    (0x46, 0x9B, 0xFF),  # 8 MOVABLE
    (0x23, 0x48, 0x7F),  # 9 MOVABLE_BORDER
    (0x0A, 0x0A, 0x0A),  # 10 WALL
    (0x05, 0x05, 0x05),  # 11 WALL_BORDER
    (0, 0, 0),  # 12 padding (env_utils.py:84-91 pads with zeros)
]
PAL_BACKGROUND, PAL_AGENT, PAL_AGENT_BORDER = 0, 1, 2
PAL_AGENT_WALL, PAL_AGENT_WALL_BORDER, PAL_GOAL_BORDER = 3, 4, 5
PAL_GOAL_OBJECT, PAL_GOAL_OBJECT_BORDER, PAL_MOVABLE, PAL_MOVABLE_BORDER = 6, 7, 8, 9
PAL_WALL, PAL_WALL_BORDER, PAL_PAD, PAL_NONE = 10, 11, 12, 15


class PuzzleLimitError(ValueError):
    """The puzzle is valid PushWorld but exceeds the table format's limits."""


@dataclass
class ParsedPuzzle:
    width: int
    height: int
    object_names: List[str]  # state order: 'a', goal movables, other movables
    goal_names: List[str]
    shapes: List[Set[Point]]  # cells relative to the object's bbox-min corner
    goal_shapes: List[Set[Point]]
    initial_state: Tuple[Point, ...]
    goal_state: Tuple[Point, ...]
    wall_cells: Set[Point]
    agent_wall_cells: Set[Point]  # aw UNION w (puzzle.py:273)
    has_agent_walls: bool
    text: str = field(default="", repr=False)

    @property
    def num_objects(self) -> int:
        return len(self.object_names)

    @property
    def num_goals(self) -> int:
        return len(self.goal_names)


def parse_puzzle(text: str) -> ParsedPuzzle:
    """puzzle.py:131-257 without the collision tables."""
    lines = text.split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    cells: Dict[str, Set[Point]] = {}
    n_cols = -1
    for row, line in enumerate(lines):
        tokens = line.split()
        if row == 0:
            n_cols = len(tokens)
        elif len(tokens) != n_cols:
            raise ValueError(
                f"Row {row + 1} does not have the same number of elements as the first row."
            )
        for col, token in enumerate(tokens):
            for name in token.split("+"):
                name = name.lower()
                if name != ".":
                    cells.setdefault(name, set()).add((col + 1, row + 1))
    if "a" not in cells:
        raise ValueError("Every puzzle must have an agent object, indicated by 'a'.")
    if n_cols <= 0:
        raise ValueError("The puzzle has no cells.")

    width, height = n_cols + 2, len(lines) + 2  # puzzle.py:160-161
    has_aw = "aw" in cells
    walls = cells.setdefault("w", set())
    walls.update((x, 0) for x in range(width))
    walls.update((x, height - 1) for x in range(width))
    walls.update((0, y) for y in range(height))
    walls.update((width - 1, y) for y in range(height))

    origin: Dict[str, Point] = {}
    shape: Dict[str, Set[Point]] = {}
    for name, pts in cells.items():
        if name in ("w", "aw"):
            origin[name] = (0, 0)
        else:
            origin[name] = (min(p[0] for p in pts), min(p[1] for p in pts))
        ox, oy = origin[name]
        shape[name] = {(x - ox, y - oy) for x, y in pts}

    order = ["a"]
    goals: List[str] = []
    for name in sorted(cells, reverse=True):  # puzzle.py:178-179
        if name[0] == "g":
            movable = "m" + name[1:]
            assert movable in cells, f"Goal has no associated movable object: {movable}"
            goals.append(name)
            order.append(movable)
    for name in cells:  # first appearance, puzzle.py:235-237
        if name[0] == "m" and name not in order:
            order.append(name)

    return ParsedPuzzle(
        width=width,
        height=height,
        object_names=order,
        goal_names=goals,
        shapes=[shape[n] for n in order],
        goal_shapes=[shape[n] for n in goals],
        initial_state=tuple(origin[n] for n in order),
        goal_state=tuple(origin[n] for n in goals),
        wall_cells=set(walls),
        agent_wall_cells=set(cells.get("aw", set())) | set(walls),
        has_agent_walls=has_aw,
        text=text,
    )


def _rows(cells: Set[Point], n_rows: int) -> List[int]:
    rows = [0] * n_rows
    for x, y in cells:
        rows[y] |= 1 << x
    return rows


def _is_regular(p: ParsedPuzzle) -> bool:
    """No movable overlaps another movable or a wall and the agent does not overlap
    aw UNION w.  Steps preserve this, so the "not already overlapping" term of
    puzzle.py:562,592 is vacuous for every reachable state of a regular puzzle."""
    seen: Set[Point] = set()
    for k, (cells, pos) in enumerate(zip(p.shapes, p.initial_state)):
        placed = {(x + pos[0], y + pos[1]) for x, y in cells}
        if placed & seen:
            return False
        if placed & (p.agent_wall_cells if k == 0 else p.wall_cells):
            return False
        seen |= placed
    return True


def build_blob(p: ParsedPuzzle) -> bytes:
    """Serialises the geometry in the layout of include/pushworld_b200.h."""
    n, g = p.num_objects, p.num_goals
    if n > MAX_OBJECTS:
        raise PuzzleLimitError(f"{n} movable objects; the table format allows {MAX_OBJECTS}")
    if p.width > MAX_WIDTH or p.height > MAX_HEIGHT:
        raise PuzzleLimitError(
            f"grid {p.width}x{p.height} exceeds the {MAX_WIDTH}x{MAX_HEIGHT} table limit"
        )
    rows: List[int] = []
    n_desc = n + g
    obj_off = HEADER_WORDS // 2  # u64 units
    pool_off = obj_off + n_desc * 2

    def push_rows(rr: Sequence[int]) -> int:
        off = pool_off + len(rows)
        rows.extend(rr)
        return off

    wall_off = push_rows(_rows(p.wall_cells, p.height))
    awall_off = push_rows(_rows(p.agent_wall_cells, p.height))
    desc: List[int] = []
    for k in range(n_desc):
        if k < n:
            cells, pos = p.shapes[k], p.initial_state[k]
            goal = p.goal_state[k - 1] if 1 <= k <= g else None
            if k == 0:
                pal = (PAL_AGENT, PAL_AGENT_BORDER)
            elif k <= g:
                pal = (PAL_GOAL_OBJECT, PAL_GOAL_OBJECT_BORDER)
            else:
                pal = (PAL_MOVABLE, PAL_MOVABLE_BORDER)
        else:
            cells, pos = p.goal_shapes[k - n], p.goal_state[k - n]
            goal = None
            pal = (PAL_NONE, PAL_GOAL_BORDER)
        bw = max(x for x, _ in cells) + 1
        bh = max(y for _, y in cells) + 1
        off = push_rows(_rows(cells, bh))
        if off > 0xFFFF:
            raise PuzzleLimitError("puzzle geometry exceeds the 16-bit row offset")
        if bh > 255 or bw > MAX_WIDTH:
            raise PuzzleLimitError(f"an object's bounding box exceeds {MAX_WIDTH} columns x 255 rows")
        desc.append(off | (bw << 16) | (bh << 24))
        desc.append((pos[0] & 0xFFFF) | ((pos[1] & 0xFFFF) << 16))
        desc.append(NO_GOAL if goal is None else (goal[0] & 0xFFFF) | ((goal[1] & 0xFFFF) << 16))
        desc.append(pal[0] | (pal[1] << 8))
    size_u64 = pool_off + len(rows)
    flags = (FLAG_HAS_AW if p.has_agent_walls else 0) | (FLAG_REGULAR if _is_regular(p) else 0)
    header = [
        MAGIC,
        p.width | (p.height << 16),
        n | (g << 16),
        flags,
        wall_off,
        awall_off,
        obj_off,
        size_u64,
    ] + [0] * (HEADER_WORDS - 8)
    blob = struct.pack(f"<{HEADER_WORDS}I", *header)
    blob += struct.pack(f"<{len(desc)}I", *desc)
    blob += struct.pack(f"<{len(rows)}Q", *rows)
    assert len(blob) == size_u64 * 8
    return blob


def compile_text(text: str) -> bytes:
    return build_blob(parse_puzzle(text))


@dataclass
class TableSet:
    """Concatenated blobs of several puzzles, ready for pw_engine_create."""

    blob: np.ndarray  # uint8
    offsets_u64: np.ndarray  # uint32[num_puzzles]
    puzzles: List[ParsedPuzzle]

    @property
    def n_max(self) -> int:
        return max(p.num_objects for p in self.puzzles)


def build_table_set(puzzles: Sequence[ParsedPuzzle]) -> TableSet:
    chunks, offsets, pos = [], [], 0
    for p in puzzles:
        b = build_blob(p)
        offsets.append(pos // 8)
        chunks.append(b)
        pos += len(b)
    return TableSet(
        blob=np.frombuffer(b"".join(chunks), dtype=np.uint8).copy(),
        offsets_u64=np.array(offsets, dtype=np.uint32),
        puzzles=list(puzzles),
    )
