// pw_board.cuh -- whole-grid bitboard form of the transition rule (host/device code).
//
// For puzzles whose grid has at most 64 cells (width * height <= 64 incl. the wall border, e.g.
// BASELINE config 2's `Road Block`, 8x5) every object is ONE uint64 occupancy board, cell (x, y)
// at bit y * W + x, and a move is a shift by 1 (LEFT/RIGHT) or W (UP/DOWN) bits.
//
// The board form is only used for REGULAR states: every object inside the grid, no movable
// overlapping another movable or a wall, the agent clear of aw UNION w.  Then
//   * the "not already overlapping" terms of puzzle.py:562,592 are vacuous, so
//       pushes(i, j)  <=>  shift(B_i) & B_j != 0          blocked(j, S)  <=>  shift(B_j) & S != 0
//   * no object touches a border cell (the border is wall), so shifts never wrap around or fall
//     off the 64-bit word, and
//   * a step maps regular states to regular states (anything a moved object would overlap is
//     pushed along or stops the move), so the check is needed once, when the state is loaded.
// With U = union of the shifted boards of the pushed set P, j joins P iff U & B_j != 0
// (puzzle.py:359-382); the sweep below repeats until P stops growing.  Irregular states are
// never handled here: board_load() reports them and the caller runs the general rule
// (pw_core.cuh) for that environment instead.
#pragma once
#include <stdint.h>

#include "pw_core.cuh"

namespace pw {

template <int NMAX>
struct Board {
  uint64_t b[NMAX];     // occupancy of object k (0 for k >= n)
  uint64_t goal[NMAX];  // goal[k] = board of object k+1 sitting on its goal (~0 when none)
  uint64_t wall, awall;
  int W, n, g;
};

PW_HD int pw_ctz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
  return __ffsll((long long)v) - 1;
#else
  return __builtin_ctzll(v);
#endif
}

// Places shape `d` at (x, y); returns false if it does not lie inside the grid.
template <class Table>
PW_HD bool place_shape(const Table& t, uint32_t d, int x, int y, int W, int H, uint64_t* out) {
  const int w = desc_w(d), h = desc_h(d);
  if (x < 0 || y < 0 || x + w > W || y + h > H) return false;
  uint64_t board = 0;
  const int o = desc_off(d);
  for (int r = 0; r < h; ++r) board |= t.row(o + r) << ((y + r) * W + x);
  *out = board;
  return true;
}

// Static part (walls, goal boards); valid when W * H <= 64 and n <= NMAX.
template <int NMAX, class Table>
PW_HD void board_init(Board<NMAX>& s, const Table& t) {
  const uint32_t dims = t.word(kHdrDims), counts = t.word(kHdrCounts);
  const int W = (int)(dims & 0xffffu), H = (int)(dims >> 16);
  s.W = W;
  s.n = (int)(counts & 0xffffu);
  s.g = (int)(counts >> 16);
  const int wo = (int)t.word(kHdrWallOff), ao = (int)t.word(kHdrAwallOff);
  s.wall = s.awall = 0;
  for (int y = 0; y < H; ++y) {
    s.wall |= t.row(wo + y) << (y * W);
    s.awall |= t.row(ao + y) << (y * W);
  }
  const int desc_word = (int)t.word(kHdrObjOff) * 2;
#pragma unroll
  for (int k = 0; k < NMAX; ++k) {
    s.b[k] = 0;
    s.goal[k] = ~0ull;
    if (k + 1 < s.n && k < s.g) {
      const uint32_t gw = t.word(desc_word + 4 * (k + 1) + 2);
      uint64_t board;
      if (place_shape(t, t.word(desc_word + 4 * (k + 1)), lo16(gw), hi16(gw), W, H, &board))
        s.goal[k] = board;  // a goal the object cannot reach inside the grid never matches
    }
  }
}

// Builds the boards from positions; false if the state is not regular (see file comment).
template <int NMAX, class Table, class GetX, class GetY>
PW_HD bool board_load(Board<NMAX>& s, const Table& t, GetX get_x, GetY get_y) {
  const uint32_t dims = t.word(kHdrDims);
  const int W = (int)(dims & 0xffffu), H = (int)(dims >> 16);
  const int desc_word = (int)t.word(kHdrObjOff) * 2;
  uint64_t seen = 0;
  bool ok = true;
#pragma unroll
  for (int k = 0; k < NMAX; ++k) {
    s.b[k] = 0;
    if (k < s.n) {
      uint64_t board = 0;
      ok = ok && place_shape(t, t.word(desc_word + 4 * k), get_x(k), get_y(k), W, H, &board);
      ok = ok && (board & seen) == 0 && (board & (k == 0 ? s.awall : s.wall)) == 0;
      seen |= board;
      s.b[k] = board;
    }
  }
  return ok;
}

// Position (bbox-min corner) of a non-empty board.
PW_HD void board_position(uint64_t board, int W, int* x, int* y) {
  const int first = pw_ctz64(board);
  *y = first / W;
  uint64_t cols = 0;
  const uint64_t row_mask = (W >= 64) ? ~0ull : ((1ull << W) - 1ull);
  for (uint64_t v = board >> (*y * W); v; v >>= W) cols |= v & row_mask;
  *x = pw_ctz64(cols);
}

// One transition on boards.  Returns the moved-object bitmask (0 = nothing moved).
// Straight-line formulation: every board is shifted once up front (m[k]); a sweep adds m[j] to
// the union U for every object U touches (re-testing an already pushed object is idempotent, so
// no "already pushed" bookkeeping is needed) and sweeps repeat until the pushed set stops
// growing.  Walls hit by any pushed object are accumulated and tested once at the end.
PW_HD uint32_t pick_byte(uint32_t packed, uint32_t index) {  // byte `index` (0..3) of `packed`
#if defined(__CUDA_ARCH__)
  return __byte_perm(packed, 0u, index | 0x4440u);
#else
  return (packed >> (8u * index)) & 0xffu;
#endif
}

template <int NMAX>
PW_HD uint32_t board_step(Board<NMAX>& s, int action) {
  const uint32_t au = (uint32_t)action;
  const uint32_t a2 = au & 3u;
  const uint32_t W = (uint32_t)s.W;
  // RIGHT / DOWN move towards higher bits, LEFT / UP towards lower bits (puzzle.py:43-50)
  const uint32_t sl = pick_byte((1u << 8) | (W << 24), a2);
  const uint32_t sr = pick_byte(1u | (W << 16), a2);
  uint64_t m[NMAX], mw[NMAX];
#pragma unroll
  for (int k = 0; k < NMAX; ++k) {
    m[k] = (s.b[k] << sl) >> sr;
    mw[k] = m[k] & s.wall;
  }
  uint64_t u = m[0], hit_wall = 0;
  uint32_t pushed = 1u;
  bool grow = au <= 3u && (m[0] & s.awall) == 0;  // puzzle.py:353-354
  const bool agent_ok = grow;
  while (grow) {
    const uint32_t before = pushed;
#pragma unroll
    for (int j = 1; j < NMAX; ++j) {
      if ((u & s.b[j]) != 0) {  // b[j] == 0 for j >= n
        u |= m[j];
        hit_wall |= mw[j];  // transitive stopping, puzzle.py:376-379
        pushed |= 1u << j;
      }
    }
    grow = pushed != before;
  }
  pushed = (agent_ok && hit_wall == 0) ? pushed : 0u;
#pragma unroll
  for (int k = 0; k < NMAX; ++k) s.b[k] = ((pushed >> k) & 1u) ? m[k] : s.b[k];
  return pushed;
}

template <int NMAX>
PW_HD int board_achieved(const Board<NMAX>& s) {  // puzzle.py:396-407
  int count = 0;
#pragma unroll
  for (int k = 0; k + 1 < NMAX; ++k) count += (s.b[k + 1] == s.goal[k]) ? 1 : 0;
  return count;
}

}  // namespace pw
