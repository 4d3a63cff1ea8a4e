// pw_board.cuh -- compact bitboard form of the transition rule (host/device code).
//
// For small puzzles every object is ONE machine word of occupancy bits and a move is a shift.
// Only the INTERIOR of the grid is stored (the 1-cell border is wall by construction,
// puzzle.py:160-168): with pitch P = W - 1, interior cell (x, y), 1 <= x <= W-2, 1 <= y <= H-2,
// lives at bit 1 + (y-1) * P + (x-1).  Column W-1 of every row (bit ... + P - 1) and bit 0 are
// GUARD bits that are set in the wall board, so a horizontal move off either side of a row lands
// on a wall bit; vertical moves off the first / last interior row are caught by an edge mask
// instead of guard rows.  A board therefore needs 1 + (H-2) * (W-1) bits: 22 for BASELINE config
// 2's `Road Block` (8x5) -> uint32 boards; up to 64 bits -> uint64 boards.
//
// The board form is only used for REGULAR states: every object inside the interior, no movable
// overlapping another movable or a wall, the agent clear of aw UNION w.  Then
//   * the "not already overlapping" terms of puzzle.py:562,592 are vacuous, so
//       pushes(i, j)  <=>  shift(B_i) & B_j != 0
//       blocked(j, S) <=>  shift(B_j) & S != 0  or  B_j touches the edge row it moves off
//   * a step maps regular states to regular states (anything a moved object would overlap is
//     pushed along or stops the move), so the check is needed once, when the state is loaded.
// With U = union of the shifted boards of the pushed set P, j joins P iff U & B_j != 0
// (puzzle.py:359-382); sweeps repeat until P stops growing.  Irregular states are never handled
// here: board_load() reports them and the caller runs the general rule (pw_core.cuh) for that
// environment instead.
#pragma once
#include <stdint.h>

#include "pw_core.cuh"

namespace pw {

template <int NMAX, typename Word>
struct Board {
  Word b[NMAX];              // occupancy of object k (0 for k >= n)
  Word goal[NMAX];           // goal[k] = board of object k+1 sitting on its goal (all ones: none)
  Word wall, awall;          // interior walls + guard bits; awall = aw UNION w
  Word edge_up, edge_down;   // first / last interior row
  int P, n, g;               // pitch = W - 1
};

// bits a puzzle needs in board form
PW_HD int board_bits(int W, int H) { return (W < 3 || H < 3) ? 1 << 30 : 1 + (H - 2) * (W - 1); }

PW_HD int pw_ctz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
  return __ffsll((long long)v) - 1;
#else
  return __builtin_ctzll(v);
#endif
}

// interior part of a full-width grid row mask, as board bits of interior row r
template <typename Word>
PW_HD Word board_row(uint64_t grid_row, int r, int W) {
  const uint64_t interior = (grid_row >> 1) & ((1ull << (W - 2)) - 1ull);
  return (Word)(interior << (1 + r * (W - 1)));
}

// Places shape `d` at grid position (x, y); false if it does not lie inside the interior.
template <typename Word, class Table>
PW_HD bool place_shape(const Table& t, uint32_t d, int x, int y, int W, int H, Word* out) {
  const int w = desc_w(d), h = desc_h(d);
  if (x < 1 || y < 1 || x + w > W - 1 || y + h > H - 1) return false;
  Word board = 0;
  const int o = desc_off(d);
  for (int r = 0; r < h; ++r)
    board |= (Word)((uint64_t)t.row(o + r) << (1 + (y - 1 + r) * (W - 1) + (x - 1)));
  *out = board;
  return true;
}

// Static part (walls, guards, edges, goal boards); needs board_bits(W, H) <= bits of Word.
template <int NMAX, typename Word, class Table>
PW_HD void board_init(Board<NMAX, Word>& s, const Table& t) {
  const uint32_t dims = t.word(kHdrDims), counts = t.word(kHdrCounts);
  const int W = (int)(dims & 0xffffu), H = (int)(dims >> 16);
  const int P = W - 1;
  s.P = P;
  s.n = (int)(counts & 0xffffu);
  s.g = (int)(counts >> 16);
  const int wo = (int)t.word(kHdrWallOff), ao = (int)t.word(kHdrAwallOff);
  Word guard = 1;  // bit 0
  s.wall = s.awall = 0;
  for (int r = 0; r < H - 2; ++r) {
    guard |= (Word)((Word)1 << (1 + r * P + (P - 1)));
    s.wall |= board_row<Word>(t.row(wo + r + 1), r, W);
    s.awall |= board_row<Word>(t.row(ao + r + 1), r, W);
  }
  s.wall |= guard;
  s.awall |= guard;
  const Word row_bits = (Word)((((uint64_t)1 << (W - 2)) - 1ull) << 1);
  s.edge_up = row_bits;
  s.edge_down = (Word)(row_bits << ((H - 3) * P));
  const int desc_word = (int)t.word(kHdrObjOff) * 2;
#pragma unroll
  for (int k = 0; k < NMAX; ++k) {
    s.b[k] = 0;
    s.goal[k] = (Word) ~(Word)0;
    if (k + 1 < s.n && k < s.g) {
      const uint32_t gw = t.word(desc_word + 4 * (k + 1) + 2);
      Word board;
      if (place_shape<Word>(t, t.word(desc_word + 4 * (k + 1)), lo16(gw), hi16(gw), W, H, &board))
        s.goal[k] = board;  // a goal the object cannot reach inside the interior never matches
    }
  }
}

// Builds the boards from positions; false if the state is not regular (see file comment).
template <int NMAX, typename Word, class Table, class GetX, class GetY>
PW_HD bool board_load(Board<NMAX, Word>& s, const Table& t, GetX get_x, GetY get_y) {
  const uint32_t dims = t.word(kHdrDims);
  const int W = (int)(dims & 0xffffu), H = (int)(dims >> 16);
  const int desc_word = (int)t.word(kHdrObjOff) * 2;
  Word seen = 0;
  bool ok = true;
#pragma unroll
  for (int k = 0; k < NMAX; ++k) {
    s.b[k] = 0;
    if (k < s.n) {
      Word board = 0;
      ok = ok && place_shape<Word>(t, t.word(desc_word + 4 * k), get_x(k), get_y(k), W, H, &board);
      ok = ok && (board & seen) == 0 && (board & (k == 0 ? s.awall : s.wall)) == 0;
      seen |= board;
      s.b[k] = board;
    }
  }
  return ok;
}

// Grid position (bbox-min corner) of a non-empty board with pitch P.
template <typename Word>
PW_HD void board_position(Word board, int P, int* x, int* y) {
  const uint64_t v = (uint64_t)board >> 1;  // drop the guard bit 0
  const int r = pw_ctz64(v) / P;
  uint64_t cols = 0;
  const uint64_t row_mask = (1ull << P) - 1ull;
  for (uint64_t rest = v >> (r * P); rest; rest >>= P) cols |= rest & row_mask;
  *x = pw_ctz64(cols) + 1;
  *y = r + 1;
}

PW_HD uint32_t pick_byte(uint32_t packed, uint32_t index) {  // byte `index` (0..3) of `packed`
#if defined(__CUDA_ARCH__)
  return __byte_perm(packed, 0u, index | 0x4440u);
#else
  return (packed >> (8u * index)) & 0xffu;
#endif
}

// Per-action constants, looked up by action index (kept in shared memory by the kernels):
// shift amounts (RIGHT / DOWN move towards higher bits, LEFT / UP towards lower bits,
// puzzle.py:43-50) and the edge row a vertical move would leave the interior through.
template <typename Word>
struct alignas(16) BoardAction {
  Word edge;
  uint32_t sl, sr;
};

template <int NMAX, typename Word>
PW_HD void board_action_table(const Board<NMAX, Word>& s, BoardAction<Word>* lut /*[4]*/) {
  const uint32_t P = (uint32_t)s.P;
  lut[0].sl = 0, lut[0].sr = 1, lut[0].edge = 0;            // LEFT
  lut[1].sl = 1, lut[1].sr = 0, lut[1].edge = 0;            // RIGHT
  lut[2].sl = 0, lut[2].sr = P, lut[2].edge = s.edge_up;    // UP
  lut[3].sl = P, lut[3].sr = 0, lut[3].edge = s.edge_down;  // DOWN
}

// One transition on boards.  Returns the moved-object bitmask (0 = nothing moved).
// Straight-line formulation: every board is shifted once up front (m[k]); a sweep adds m[j] to
// the union U for every object U touches (re-testing an already pushed object is idempotent, so
// no "already pushed" bookkeeping is needed) and sweeps repeat until the pushed set stops
// growing.  Obstacles hit by any pushed object are accumulated and tested once at the end.
template <int NMAX, typename Word>
PW_HD uint32_t board_step(Board<NMAX, Word>& s, int action, const BoardAction<Word> act);

// The per-action constants computed from the board itself (no table): used when every thread
// plays a different puzzle and a shared look-up table is not available.
template <int NMAX, typename Word>
PW_HD BoardAction<Word> board_action_of(const Board<NMAX, Word>& s, uint32_t a2) {
  const uint32_t P = (uint32_t)s.P;
  BoardAction<Word> act;
  act.sl = pick_byte((1u << 8) | (P << 24), a2);
  act.sr = pick_byte(1u | (P << 16), a2);
  act.edge = a2 == 2u ? s.edge_up : (a2 == 3u ? s.edge_down : (Word)0);
  return act;
}

template <int NMAX, typename Word>
PW_HD uint32_t board_step(Board<NMAX, Word>& s, int action, const BoardAction<Word>* lut) {
  return board_step(s, action, lut[(uint32_t)action & 3u]);
}

template <int NMAX, typename Word>
PW_HD uint32_t board_step(Board<NMAX, Word>& s, int action, const BoardAction<Word> act) {
  const uint32_t au = (uint32_t)action;
  const uint32_t sl = act.sl, sr = act.sr;
  const Word edge = act.edge;
  Word m[NMAX], stop[NMAX];
#pragma unroll
  for (int k = 0; k < NMAX; ++k) {
    m[k] = (Word)((Word)(s.b[k] << sl) >> sr);
    stop[k] = (Word)((m[k] & (k == 0 ? s.awall : s.wall)) | (s.b[k] & edge));
  }
  const bool agent_ok = au <= 3u && stop[0] == 0;  // puzzle.py:353-354
  uint32_t pushed = 1u;
  Word hit = 0;
  if (NMAX == 2) {
    // one pushable object: it is pushed iff the moved agent touches it
    const bool p1 = (m[0] & s.b[1]) != 0;
    hit = p1 ? stop[1] : (Word)0;
    pushed |= p1 ? 2u : 0u;
  } else if (NMAX == 3) {
    // two pushable objects: direct pushes plus one level of chaining closes the set
    const bool d1 = (m[0] & s.b[1]) != 0, d2 = (m[0] & s.b[2]) != 0;
    const bool p2 = d2 || (d1 && (m[1] & s.b[2]) != 0);
    const bool p1 = d1 || (p2 && (m[2] & s.b[1]) != 0);
    hit = (Word)((p1 ? stop[1] : (Word)0) | (p2 ? stop[2] : (Word)0));
    pushed |= (p1 ? 2u : 0u) | (p2 ? 4u : 0u);
  } else {
    Word u = m[0];
    bool grow = agent_ok;
    while (grow) {
      const uint32_t before = pushed;
#pragma unroll
      for (int j = 1; j < NMAX; ++j) {
        if ((u & s.b[j]) != 0) {  // b[j] == 0 for j >= n
          u |= m[j];
          hit |= stop[j];  // transitive stopping, puzzle.py:376-379
          pushed |= 1u << j;
        }
      }
      grow = pushed != before;
    }
  }
  pushed = (agent_ok && hit == 0) ? pushed : 0u;
#pragma unroll
  for (int k = 0; k < NMAX; ++k) s.b[k] = ((pushed >> k) & 1u) ? m[k] : s.b[k];
  return pushed;
}

// ---- ring boards: uint32 boards with P spare bits -----------------------------------------
// When bits + P <= 32 the unused top of the word becomes one GUARD ZONE [bits, 32) that is set in
// both wall boards, and a move is a single ROTATE:
//   * RIGHT / DOWN rotate left by 1 / P: nothing wraps (the highest cell bit is bits - 1, and
//     bits - 1 + P <= 31); a cell leaving the last interior row lands in the guard zone.
//   * LEFT / UP rotate left by 31 / 32 - P: objects never own bit 0 or a guard column, so the only
//     bits that wrap are the cells of the first interior row moving UP (bits 1 .. P-1), and they
//     land on 33 - P .. 31 >= bits: again the guard zone.
// The guard zone therefore replaces the edge-row tests of board_step, and a regular state stays
// free of guard bits because a move that puts any bit there is blocked.  One SHF per object and
// one AND per wall test instead of two and two.
PW_HD uint32_t rotl32(uint32_t v, uint32_t s) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_l(v, v, s);  // wraps: uses s & 31
#else
  s &= 31u;
  return s ? (v << s) | (v >> (32u - s)) : v;
#endif
}

PW_HD bool ring_fits(int bits, int P) { return bits + P <= 32; }
PW_HD uint32_t ring_guard(int bits) { return ~((1u << bits) - 1u); }
// rotate-left amount of action a in byte a (puzzle.py:43-50: LEFT, RIGHT, UP, DOWN)
PW_HD uint32_t ring_amounts(int P) {
  return 31u | (1u << 8) | ((32u - (uint32_t)P) << 16) | ((uint32_t)P << 24);
}
PW_HD uint32_t ring_amount_of(uint32_t amounts, uint32_t action) {
#if defined(__CUDA_ARCH__)
  uint32_t amount;  // raw prmt: no masking of the selector; bytes 1..3 are ignored by rotl32
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(amount) : "r"(amounts), "r"(0u), "r"(action));
  return amount;
#else
  return (amounts >> (8u * (action & 3u))) & 0xffu;
#endif
}

// One transition on ring boards; `wall` / `awall` include the guard zone, `valid` = action <= 3
// (an invalid action is a no-op, like a blocked agent).  Returns the moved-object bitmask when
// kMask, else 1 / 0 for moved / unchanged.  Same rule as board_step.
template <int NMAX, bool kMask>
PW_HD uint32_t ring_step(uint32_t (&b)[NMAX], const uint32_t wall, const uint32_t awall,
                         const uint32_t amount, const bool valid) {
  uint32_t m[NMAX];
#pragma unroll
  for (int k = 0; k < NMAX; ++k) m[k] = rotl32(b[k], amount);
  bool blocked = !valid | ((m[0] & awall) != 0);  // puzzle.py:353-354
  if (NMAX <= 3) {
    bool p1 = false, p2 = false;
    if (NMAX == 2) {
      p1 = (m[0] & b[1]) != 0;
    } else if (NMAX == 3) {
      // two pushable objects: direct pushes plus one level of chaining closes the set
      const bool d1 = (m[0] & b[1]) != 0, d2 = (m[0] & b[NMAX - 1]) != 0;
      p2 = d2 | (d1 & ((m[1] & b[NMAX - 1]) != 0));
      p1 = d1 | (p2 & ((m[NMAX - 1] & b[1]) != 0));
      blocked |= p2 & ((m[NMAX - 1] & wall) != 0);
    }
    if (NMAX >= 2) blocked |= p1 & ((m[1] & wall) != 0);  // transitive stopping, puzzle.py:376-379
    const bool go = !blocked, go1 = go & p1, go2 = go & p2;
    b[0] = go ? m[0] : b[0];
    if (NMAX >= 2) b[1] = go1 ? m[1] : b[1];
    if (NMAX == 3) b[NMAX - 1] = go2 ? m[NMAX - 1] : b[NMAX - 1];
    if (kMask) return go ? (1u | (p1 ? 2u : 0u) | (p2 ? 4u : 0u)) : 0u;
    return go ? 1u : 0u;
  } else {
    uint32_t u = m[0], hit = 0, pushed = 1u;
    bool grow = !blocked;
    while (grow) {
      const uint32_t before = pushed;
#pragma unroll
      for (int j = 1; j < NMAX; ++j) {
        if ((u & b[j]) != 0) {  // b[j] == 0 for j >= n
          u |= m[j];
          hit |= m[j] & wall;
          pushed |= 1u << j;
        }
      }
      grow = pushed != before;
    }
    pushed = (blocked | (hit != 0)) ? 0u : pushed;
#pragma unroll
    for (int k = 0; k < NMAX; ++k) b[k] = ((pushed >> k) & 1u) ? m[k] : b[k];
    return kMask ? pushed : (pushed != 0 ? 1u : 0u);
  }
}

template <int NMAX, typename Word>
PW_HD int board_achieved(const Board<NMAX, Word>& s) {  // puzzle.py:396-407
  int count = 0;
#pragma unroll
  for (int k = 0; k + 1 < NMAX; ++k) {
    if (k >= s.g) break;  // uniform across the launch: a branch, not a wasted compare
    count += (s.b[k + 1] == s.goal[k]) ? 1 : 0;
  }
  return count;
}

}  // namespace pw
