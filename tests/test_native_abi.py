"""Host-side checks of the native library that need no GPU: it loads, exports every symbol
include/pushworld_b200.h declares, and its .pwp compiler agrees byte for byte with the
Python table compiler (and raises the reference's exception types)."""

import os
import re

import pytest

from pushworld_b200 import _lib, tables
from tests.conftest import REPO


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(REPO, "include", "pushworld_b200.h")).read()
    declared = set(re.findall(r"PW_API\s+[\w\s\*]+?\b(pw_\w+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.pw_abi_version() == 1
    assert lib.pw_last_error() in (b"", None)


def test_native_compiler_matches_python_compiler(golden):
    for key in golden.keys():
        text = golden.text(key)
        assert _lib.compile_puzzle_text(text) == tables.compile_text(text), key


def test_table_header_fields(golden):
    import numpy as np

    blob = _lib.compile_puzzle_text(golden.text("python3/file_parsing"))
    w = np.frombuffer(blob, np.uint32)
    assert w[0] == tables.MAGIC
    assert (w[1] & 0xFFFF, w[1] >> 16) == (12, 18)  # test_puzzle.py:201
    assert (w[2] & 0xFFFF, w[2] >> 16) == (6, 2)
    assert w[3] & tables.FLAG_HAS_AW and w[3] & tables.FLAG_REGULAR
    assert w[7] * 8 == len(blob)
    # trivial_tool has AW+M1 in one cell: still regular (aw only stops the agent)
    w = np.frombuffer(_lib.compile_puzzle_text(golden.text("python3/trivial_tool")), np.uint32)
    assert w[3] & tables.FLAG_REGULAR
    # a movable on top of a wall is not
    w = np.frombuffer(_lib.compile_puzzle_text("a . w+m0\n. . .\n"), np.uint32)
    assert not (w[3] & tables.FLAG_REGULAR)


@pytest.mark.parametrize(
    "text, exc",
    [
        (". .\n. . .\n", ValueError),  # ragged rows, puzzle.py:141-145
        (". m0\n. .\n", ValueError),  # no agent, puzzle.py:154-157
        ("a g1\n. m0\n", AssertionError),  # goal without movable, puzzle.py:230-232
        ("a .\n. .\n\n", ValueError),  # trailing blank line = ragged row
        (" ".join(["."] * 63 + ["a"]) + "\n", ValueError),  # 66 columns > PW_MAX_WIDTH
        # a movable whose bounding box spans 300 rows: the 8-bit height field cannot hold it
        ("m1 a\n" + ". .\n" * 298 + "m1 .\n", ValueError),
    ],
)
def test_compile_errors(text, exc):
    with pytest.raises(exc):
        _lib.compile_puzzle_text(text)
    with pytest.raises(exc):
        tables.compile_text(text)


def test_line_endings_and_case():
    ref = _lib.compile_puzzle_text("a m0 .\n. . g0\n")
    assert _lib.compile_puzzle_text("A M0 .\r\n. . G0\r\n") == ref
    assert _lib.compile_puzzle_text("a m0 .\n. . g0") == ref
    assert _lib.compile_puzzle_text("  a\tm0   .  \n .  . g0 \n") == ref


def test_record_expansion_matches_the_reference_reward_arithmetic():
    """The packed per-step record (pw_rollout_packed) expands to the float32 image of the python
    float the reference computes (gym_env.py:212-221)."""
    import numpy as np

    from pushworld_b200.engine import record_to_outputs

    codes = np.arange(256, dtype=np.uint8)
    reward, term, trunc, dg = record_to_outputs(codes)
    for c in range(256):
        d = (c & 0x3F) - 32
        want = 10.0 if c & 0x80 else float(d) - 0.01
        assert reward[c] == np.float32(want) and term[c] == (c >> 7) and trunc[c] == ((c >> 6) & 1)
        assert dg[c] == d
    import torch

    r_t, t_t, tr_t, dg_t = record_to_outputs(torch.from_numpy(codes))
    assert np.array_equal(r_t.numpy(), reward) and np.array_equal(dg_t.numpy(), dg)
    assert np.array_equal(t_t.numpy(), term) and np.array_equal(tr_t.numpy(), trunc)
