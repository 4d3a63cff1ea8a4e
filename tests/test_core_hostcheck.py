"""The kernels' transition rule (pw_core.cuh, compiled for the CPU by tests/hostcheck) and
the host table compiler against the reference's golden vectors.  No GPU needed: this is
what lets kernel logic be debugged before spending GPU time; the GPU suite repeats the
same comparisons through the C-ABI."""

import numpy as np
import pytest

from pushworld_b200 import tables
from tests import hostcheck


def _check(golden, key, tag):
    text = golden.text(key)
    parsed = tables.parse_puzzle(text)
    rec = golden.rec(key)
    assert (parsed.width, parsed.height) == (rec["width"], rec["height"])
    assert list(map(list, parsed.initial_state)) == golden.arr(key, "init").tolist()
    assert list(map(list, parsed.goal_state)) == golden.arr(key, "goal").tolist()
    blob = tables.build_blob(parsed)
    actions = golden.arr(key, f"{tag}_actions")
    ref = golden.arr(key, f"{tag}_states").astype(np.int32)
    states, moved, ach = hostcheck.walk(blob, parsed.initial_state, actions)
    assert np.array_equal(states, ref[1:]), key
    assert np.array_equal(ach, golden.arr(key, f"{tag}_achieved")[1:].astype(np.int32)), key
    # the moved mask is exactly the set of objects whose position changed
    changed = (ref[1:] != ref[:-1]).any(axis=2)
    bits = (moved[:, None] >> np.arange(ref.shape[1])[None, :]) & 1
    assert np.array_equal(bits.astype(bool), changed), key
    return len(actions)


@pytest.mark.parametrize("tag", ["sol", "walk"])
def test_core_matches_reference_vectors(golden, tag):
    total = 0
    for key in golden.keys():
        if golden.has(key, f"{tag}_actions"):
            total += _check(golden, key, tag)
    assert total > (18000 if tag == "sol" else 60000)


def test_cell_code_matches_reference_renders(golden):
    """pw_render_core.cuh::cell_code (the kernels' per-cell closed form of render) against
    sha256 of the real reference's PushWorldPuzzle.render output."""
    import hashlib
    import json
    import os

    from tests.conftest import GOLDEN

    with open(os.path.join(GOLDEN, "reference_renders.json")) as f:
        renders = json.load(f)
    blobs = {}
    for case in renders["cases"]:
        key = case["key"]
        if key not in blobs:
            blobs[key] = tables.compile_text(golden.text(key))
        img = hostcheck.render_u8(blobs[key], case["state"], case["ppc"], case["bw"])
        assert list(img.shape) == case["shape"]
        assert hashlib.sha256(img.tobytes()).hexdigest() == case["sha256_u8"], (key, case["ppc"])
    assert len(renders["cases"]) > 100
