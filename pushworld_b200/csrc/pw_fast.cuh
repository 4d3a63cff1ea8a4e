// pw_fast.cuh -- collision bit tables and the transition rule that reads them.
//
// The reference never evaluates geometry inside get_next_state: PushWorldPuzzle.__init__
// precomputes, per action, the positions at which an object is stopped by the walls and the
// relative positions at which one movable pushes another (puzzle.py:259-308, 522-593), and a step
// is a handful of set look-ups (puzzle.py:348-394).  This file is that design with bitmaps instead
// of hash sets.  The tables are DERIVED from the geometry blob at engine creation by evaluating
// the exact predicates of pw_core.cuh (static_blocked, shapes_overlap) at every position of a
// finite window outside of which they are provably false, so
//     step_tab(env, a) == step(env, a)        for every state, on or off the grid.
// All 223 benchmark puzzles together need 0.63 MB of tables (largest: 23 KB).
//
// Layout of one puzzle's table F (u64 units, offsets relative to F):
//   F[0]                 n | offset of the initial-state section << 32 (see the end of this list)
//   F[1 + i*n + j]       pair (pusher i, pushee j >= 1, i != j):
//                          low word  = offset of the bitmap
//                          high word = Ww | Hw << 8 | w_j << 17 | h_j << 24
//                        with Ww = w_i + w_j + 1, Hw = h_i + h_j + 1: the window of relative
//                        positions r = p_j - p_i, -w_j <= rx <= w_i, -h_j <= ry <= h_i; bit
//                        (a * Hw + ry + h_j) * Ww + rx + w_j  <=>  i, moved by action a, pushes j
//   F[1 + n*n + j]       walls vs object j (the agent against aw UNION w, puzzle.py:273):
//                          low word  = offset of the bitmap
//                          high word = xmax | ymax << 6 | W << 18   (xmax = W - w_j, ymax = H - h_j)
//                        bit ((y * W + x) << 2 | a)  <=>  j at (x, y) is stopped when moved by a;
//                        positions whose bounding box leaves the grid are never stopped (:557-561)
//   initial-state section (what an episode restart of pw_occ.cu's mask kernel needs, precomputed):
//     [0]                low word = 1 if the initial state qualifies for the mask layout (grid of at
//                        most 32 x 32 cells, every movable inside it and off the walls), bits 8..15 =
//                        goals achieved in the initial state (puzzle.py:396-407)
//     [1 + (j - 1)]      object j >= 1 in its initial position: row mask | column mask << 32
#pragma once
#include <stdint.h>

#include "pw_core.cuh"

namespace pw {

// bit (index & 31) of `word`, as a test against a shifted one (the shift wraps: no masking needed)
PW_HD bool tab_bit(uint32_t word, uint32_t index) {
#if defined(__CUDA_ARCH__)
  return (word & __funnelshift_l(1u, 1u, index)) != 0u;  // 1 << (index & 31): bit 0 wraps around
#else
  return ((word >> (index & 31u)) & 1u) != 0u;
#endif
}

// `env.fast(k)` = F[k]; `env.fast32(k)` = 32-bit word k of F (little endian: bit b of the bitmap at
// u64 offset o is bit b & 31 of 32-bit word 2 * o + (b >> 5)).  Is object j at (x, y) stopped by the static cells when moved by action a?
template <class Env>
PW_HD bool tab_wall(const Env& env, uint32_t nn1, uint32_t j, uint32_t a, int x, int y) {
  const uint64_t wd = env.fast(nn1 + j);
  const uint32_t hi = (uint32_t)(wd >> 32);
  if ((uint32_t)x > (hi & 63u) || (uint32_t)y > ((hi >> 6) & 0xfffu)) return false;
  const uint32_t idx = (((uint32_t)y * (hi >> 18) + (uint32_t)x) << 2) | a;
  return tab_bit(env.fast32(((uint32_t)wd << 1) + (idx >> 5)), idx);
}

// Does pusher i (descriptor index base = 1 + i*n), moved by action a, push j at relative position
// (rx, ry) = p_j - p_i?
template <class Env>
PW_HD bool tab_pushes(const Env& env, uint32_t pair_index, uint32_t a, int rx, int ry) {
  const uint64_t pd = env.fast(pair_index);
  const uint32_t hi = (uint32_t)(pd >> 32);
  const uint32_t Ww = hi & 0xffu, Hw = (hi >> 8) & 0x1ffu;
  const uint32_t ux = (uint32_t)(rx + (int)((hi >> 17) & 0x7fu)), uy = (uint32_t)(ry + (int)(hi >> 24));
  if (ux >= Ww || uy >= Hw) return false;
  const uint32_t idx = (a * Hw + uy) * Ww + ux;
  return tab_bit(env.fast32(((uint32_t)pd << 1) + (idx >> 5)), idx);
}

// tab_pushes for all four actions at once: bit a of the result <=> pusher, moved by action a, pushes
// j.  The descriptor load and the window test do not depend on the action (the bitmap is indexed
// (a * Hw + uy) * Ww + ux): successor expansion asks the question for LEFT, RIGHT, UP and DOWN.
template <class Env>
PW_HD uint32_t tab_pushes4(const Env& env, uint32_t pair_index, int rx, int ry) {
  const uint64_t pd = env.fast(pair_index);
  const uint32_t hi = (uint32_t)(pd >> 32);
  const uint32_t Ww = hi & 0xffu, Hw = (hi >> 8) & 0x1ffu;
  const uint32_t ux = (uint32_t)(rx + (int)((hi >> 17) & 0x7fu)), uy = (uint32_t)(ry + (int)(hi >> 24));
  if (ux >= Ww || uy >= Hw) return 0u;
  const uint32_t word0 = (uint32_t)pd << 1, plane = Hw * Ww;
  uint32_t idx = uy * Ww + ux, out = 0u;
#pragma unroll
  for (uint32_t a = 0; a < 4u; ++a, idx += plane) out |= tab_bit(env.fast32(word0 + (idx >> 5)), idx) ? (1u << a) : 0u;
  return out;
}

// The pushed set of one transition, read from the tables: bit k = object k moves (bit 0 = agent),
// 0 if the action is invalid or anything in the set is stopped.  Positions are not modified.
// The scan of a pusher over the other objects is branch-free per object (one pair look-up each);
// the wall look-ups of the objects it hits come after the scan.
// `agent_hits` (optional, with have_agent_hits): the objects the agent pushes directly under this
// action, already known to the caller (successor expansion computes them for all four actions with
// tab_pushes4); the agent's own scan is then skipped.
template <class Env>
PW_HD uint32_t pushed_set_tab(const Env& env, int action, bool have_agent_hits = false, uint32_t agent_hits = 0u) {
  if ((unsigned)action > 3u) return 0;  // invalid actions are no-ops (counted by the caller)
  const uint32_t a = (uint32_t)action;
  const int n = env.n();
  const uint32_t nn1 = 1u + (uint32_t)(n * n);
  if (tab_wall(env, nn1, 0u, a, env.x(0), env.y(0))) return 0;  // puzzle.py:353-354
  const int jend = Env::kStaticN ? Env::kStaticN : n;
#ifndef PW_SCAN_UNROLL
#define PW_SCAN_UNROLL 4
#endif
  // register policies unroll the scan completely; the shared-memory policy keeps several objects'
  // independent look-up chains in flight per iteration
  constexpr int kUnroll = Env::kStaticN ? Env::kStaticN : PW_SCAN_UNROLL;
  uint32_t pushed = 1u, frontier = 1u;
  while (frontier) {
    const int i = pw_ffs(frontier) - 1;
    frontier &= frontier - 1;
    const int xi = env.x_dyn(i), yi = env.y_dyn(i);
    const uint32_t base = 1u + (uint32_t)(i * n);
    // object j's verdict enters at bit 31 and is shifted down once per later object: after the
    // n - 1 objects, bit 32 - n + j; the final shift by 32 - n puts it at bit j
    uint32_t hits = 0;
    if (have_agent_hits && i == 0) {
      hits = agent_hits;
    } else {
#pragma unroll kUnroll
      for (int j = 1; j < jend; ++j) {
        if (Env::kStaticN && j >= n) break;
        // (the descriptor of the pair (i, i) is zero: an empty window)
        const bool push = tab_pushes(env, base + (uint32_t)j, a, env.x(j) - xi, env.y(j) - yi);
        hits = (hits >> 1) | (push ? 0x80000000u : 0u);
      }
      hits = n > 1 ? hits >> (32 - n) : 0u;
    }
    hits &= ~pushed;
    pushed |= hits;
    frontier |= hits;
    while (hits) {  // transitive stopping, puzzle.py:376-379
      const int j = pw_ffs(hits) - 1;
      hits &= hits - 1;
      if (tab_wall(env, nn1, (uint32_t)j, a, env.x_dyn(j), env.y_dyn(j))) return 0;
    }
  }
  return pushed;
}

// One transition on the tables.  Same contract as step() (pw_core.cuh): returns the moved-object
// bitmask, 0 if nothing moved.
template <class Env>
PW_HD uint32_t step_tab(Env& env, int action) {
  const uint32_t pushed = pushed_set_tab(env, action);
  if (pushed) env.translate(pushed, action_dx(action), action_dy(action));
  return pushed;
}

}  // namespace pw

// ---- host-side construction (engine creation, tests/hostcheck) -----------------------------
#include <vector>

namespace pw {

// Geometry view of a blob for the exact predicates of pw_core.cuh.
struct FastBuildEnv {
  const uint64_t* base;
  int w_, h_;
  PW_HD uint64_t row(int off) const { return base[off]; }
  PW_HD int width() const { return w_; }
  PW_HD int height() const { return h_; }
};

// Appends the table of the puzzle in `blob` to `out` (its offsets are relative to its own start).
inline void build_fast_table(const uint64_t* blob, std::vector<uint64_t>* out) {
  const uint32_t* words = reinterpret_cast<const uint32_t*>(blob);
  const int W = (int)(words[kHdrDims] & 0xffffu), H = (int)(words[kHdrDims] >> 16);
  const int n = (int)(words[kHdrCounts] & 0xffffu);
  const int wall_off = (int)words[kHdrWallOff], awall_off = (int)words[kHdrAwallOff];
  const int desc_word = (int)words[kHdrObjOff] * 2;
  const FastBuildEnv env{blob, W, H};
  const size_t start = out->size();
  out->resize(start + 1 + (size_t)n * n + n, 0ull);
  (*out)[start] = (uint64_t)n;
  auto set_bit = [&](size_t off, uint32_t idx) { (*out)[start + off + (idx >> 6)] |= 1ull << (idx & 63u); };
  // walls
  for (int j = 0; j < n; ++j) {
    const uint32_t dj = words[desc_word + 4 * j];
    const int w = desc_w(dj), h = desc_h(dj);
    const int xmax = W - w, ymax = H - h;
    const size_t off = out->size() - start;
    if (xmax < 0 || ymax < 0) {  // larger than the grid: never inside it, never stopped
      out->resize(out->size() + 1, 0ull);
      (*out)[start + 1 + (size_t)n * n + j] = (uint64_t)off;  // xmax = ymax = 0, W = 0 -> one cell, all zero
      continue;
    }
    const uint32_t bits = (uint32_t)W * (uint32_t)H * 4u;
    out->resize(out->size() + (bits + 63u) / 64u, 0ull);
    for (int y = 0; y <= ymax; ++y)
      for (int x = 0; x <= xmax; ++x)
        for (int a = 0; a < 4; ++a)
          if (static_blocked(env, dj, x, y, action_dx(a), action_dy(a), j == 0 ? awall_off : wall_off))
            set_bit(off, (uint32_t)((y * W + x) << 2 | a));
    (*out)[start + 1 + (size_t)n * n + j] =
        (uint64_t)off | ((uint64_t)((uint32_t)xmax | ((uint32_t)ymax << 6) | ((uint32_t)W << 18)) << 32);
  }
  // pairs
  for (int i = 0; i < n; ++i) {
    const uint32_t di = words[desc_word + 4 * i];
    const int wi = desc_w(di), hi = desc_h(di);
    for (int j = 1; j < n; ++j) {
      if (j == i) continue;
      const uint32_t dj = words[desc_word + 4 * j];
      const int wj = desc_w(dj), hj = desc_h(dj);
      const int Ww = wi + wj + 1, Hw = hi + hj + 1;
      const size_t off = out->size() - start;
      const uint32_t bits = 4u * (uint32_t)Ww * (uint32_t)Hw;
      out->resize(out->size() + (bits + 63u) / 64u, 0ull);
      for (int a = 0; a < 4; ++a) {
        const int dx = action_dx(a), dy = action_dy(a);
        for (int uy = 0; uy < Hw; ++uy)
          for (int ux = 0; ux < Ww; ++ux) {
            const int rx = ux - wj, ry = uy - hj;  // p_j - p_i
            // puzzle.py:567-593: the moved pusher overlaps j and the unmoved one does not
            if (shapes_overlap(env, di, dj, rx - dx, ry - dy) && !shapes_overlap(env, di, dj, rx, ry))
              set_bit(off, (uint32_t)((a * Hw + uy) * Ww + ux));
          }
      }
      (*out)[start + 1 + (size_t)i * n + j] =
          (uint64_t)off | ((uint64_t)((uint32_t)Ww | ((uint32_t)Hw << 8) | ((uint32_t)wj << 17) |
                                      ((uint32_t)hj << 24))
                           << 32);
    }
  }
  // initial-state section
  {
    const int g = (int)(words[kHdrCounts] >> 16);
    const size_t off = out->size() - start;
    out->resize(out->size() + (size_t)n, 0ull);  // 1 + (n - 1) words
    bool ok = W <= 32 && H <= 32;
    int achieved = 0;
    for (int j = 1; j < n; ++j) {
      const uint32_t dj = words[desc_word + 4 * j], init = words[desc_word + 4 * j + 1];
      const int w = desc_w(dj), h = desc_h(dj), o = desc_off(dj), x = lo16(init), y = hi16(init);
      if (j <= g && init == words[desc_word + 4 * j + 2]) ++achieved;
      if (x < 0 || y < 0 || x + w > W || y + h > H || x + w > 32 || y + h > 32) {
        ok = false;
        continue;
      }
      for (int r = 0; r < h; ++r) ok = ok && ((blob[o + r] << x) & blob[wall_off + y + r]) == 0ull;
      const uint64_t rm = (0xffffffffull >> (32 - h)) << y, cm = (0xffffffffull >> (32 - w)) << x;
      (*out)[start + off + 1 + (size_t)(j - 1)] = (rm & 0xffffffffull) | ((cm & 0xffffffffull) << 32);
    }
    (*out)[start + off] = (ok ? 1ull : 0ull) | ((uint64_t)achieved << 8);
    (*out)[start] |= (uint64_t)off << 32;
  }
}

}  // namespace pw
