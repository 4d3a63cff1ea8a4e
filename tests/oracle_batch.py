"""Batched use of the C oracle (oracle/pw_oracle.c) by the full-size parity tests: rollouts of
mixed-puzzle batches are checked by running the oracle once per puzzle on that puzzle's columns."""

import numpy as np

import c_oracle


class OracleSet:
    """The C oracle of every puzzle of an engine (built lazily, kept for the session)."""

    def __init__(self, texts):
        self.texts = list(texts)
        self._puzzles = {}

    def puzzle(self, p):
        if p not in self._puzzles:
            self._puzzles[p] = c_oracle.COraclePuzzle(self.texts[p])
        return self._puzzles[p]

    def rollout(self, puzzle_ids, actions, n_max, x0=None, y0=None):
        """puzzle_ids int [N], actions uint8 [T, N] -> (x [n_max, N], y [n_max, N], reward f32
        [T, N], terminated u8 [T, N]) from each puzzle's initial state (or x0 / y0)."""
        T, N = actions.shape
        x = np.zeros((n_max, N), np.int32)
        y = np.zeros((n_max, N), np.int32)
        reward = np.empty((T, N), np.float32)
        term = np.empty((T, N), np.uint8)
        order = np.argsort(puzzle_ids, kind="stable")
        sorted_ids = puzzle_ids[order]
        bounds = np.flatnonzero(np.diff(sorted_ids)) + 1
        starts = np.concatenate(([0], bounds))
        ends = np.concatenate((bounds, [N]))
        for lo, hi in zip(starts, ends):
            p = int(sorted_ids[lo])
            cols = order[lo:hi]
            puzzle = self.puzzle(p)
            n = puzzle.n
            if x0 is None:
                init = np.array(puzzle.py.initial_state, dtype=np.int32)
                px = np.ascontiguousarray(np.repeat(init[:, 0:1], len(cols), axis=1))
                py = np.ascontiguousarray(np.repeat(init[:, 1:2], len(cols), axis=1))
            else:
                px = np.ascontiguousarray(x0[:n, cols])
                py = np.ascontiguousarray(y0[:n, cols])
            acts = np.ascontiguousarray(actions[:, cols])
            r = np.empty(acts.shape, np.float32)
            te = np.empty(acts.shape, np.uint8)
            puzzle.rollout(px, py, acts, r, te)
            x[:n, cols] = px
            y[:n, cols] = py
            reward[:, cols] = r
            term[:, cols] = te
        return x, y, reward, term
