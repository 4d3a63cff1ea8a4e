// pw_core.cuh -- the PushWorld transition rule on row bitmasks.
//
// One (environment, action) -> next state, written once as host/device code so the very
// same source runs inside the sm_100a kernels (pw_kernels.cu) and inside the CPU-side
// check harness that tests/ builds with g++ (tests/hostcheck).  Nothing here touches
// memory except through the `Env` policy object it is given.
//
// What it computes (reference: python3/src/pushworld/puzzle.py:348-394, twin
// cpp/src/pushworld_puzzle.cc:386-460), restated on geometry instead of the reference's
// precomputed hash sets:
//   blocked(i, S)  := bbox(i) inside the grid  AND  overlaps(C_i + p_i + d, S)
//                     AND NOT overlaps(C_i + p_i, S)              (puzzle.py:522-564)
//   pushes(i, j)   := overlaps(C_i + p_i + d, C_j + p_j) AND NOT overlaps(C_i + p_i, C_j + p_j)
//                                                                  (puzzle.py:567-593)
//   1. blocked(agent, aw U w)            -> state unchanged        (puzzle.py:353-354)
//   2. P := least fixed point of {0} under pushes(i in P, j notin P, j >= 1)   (:359-382)
//   3. any j in P \ {0} with blocked(j, w) -> state unchanged ("transitive stopping", :376-379)
//   4. every k in P moves by d                                     (:384-394)
// The closure is order independent, so the LIFO order of the reference is irrelevant.
//
// Shapes are bbox-relative row masks (bit c of row r = cell (c, r)); walls are one u64 per
// grid row.  All pair tests are done in relative coordinates, so they are exact for ANY
// int32 positions (also off-grid ones that only arise from irregular puzzle files).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PW_HD __host__ __device__ __forceinline__
#else
#define PW_HD inline
#endif

namespace pw {

// ---- table blob accessors (layout: include/pushworld_b200.h) -------------------------
enum : int {
  kHdrMagic = 0,
  kHdrDims = 1,
  kHdrCounts = 2,
  kHdrFlags = 3,
  kHdrWallOff = 4,
  kHdrAwallOff = 5,
  kHdrObjOff = 6,
  kHdrSizeU64 = 7,
  kHeaderWords = 16,
};
constexpr uint32_t kNoGoal = 0x80008000u;
constexpr int kMaxObjects = 32;

PW_HD int pw_ffs(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __ffs((int)v);
#else
  return __builtin_ffs((int)v);
#endif
}

PW_HD int desc_w(uint32_t d) { return (int)((d >> 16) & 0xffu); }
PW_HD int desc_h(uint32_t d) { return (int)(d >> 24); }
PW_HD int desc_off(uint32_t d) { return (int)(d & 0xffffu); }
PW_HD int lo16(uint32_t v) { return (int)(int16_t)(v & 0xffffu); }
PW_HD int hi16(uint32_t v) { return (int)(int16_t)(v >> 16); }

PW_HD int action_dx(int a) { return a == 0 ? -1 : (a == 1 ? 1 : 0); }  // puzzle.py:43-50
PW_HD int action_dy(int a) { return a == 2 ? -1 : (a == 3 ? 1 : 0); }

// Does shape I placed at the origin overlap shape J placed at (sx, sy)?
template <class Env>
PW_HD bool shapes_overlap(const Env& env, uint32_t di, uint32_t dj, int sx, int sy) {
  const int wi = desc_w(di), hi = desc_h(di), wj = desc_w(dj), hj = desc_h(dj);
  if (sx >= wi || -sx >= wj || sy >= hi || -sy >= hj) return false;  // bounding boxes apart
  if ((di >> 16) == 0x0101u && (dj >> 16) == 0x0101u) return true;  // 1x1 vs 1x1 cell
  const int r0 = sy > 0 ? sy : 0;
  const int r1 = (sy + hj) < hi ? (sy + hj) : hi;
  const int oi = desc_off(di), oj = desc_off(dj) - sy;
  uint64_t acc = 0;
  for (int r = r0; r < r1; ++r) {
    const uint64_t a = env.row(oi + r);
    const uint64_t b = env.row(oj + r);
    acc |= a & (sx >= 0 ? (b << sx) : (b >> (-sx)));  // |sx| < 64 by the bbox test
  }
  return acc != 0;
}

// puzzle.py:522-564 evaluated directly: is the object stopped by the static rows at `soff`?
template <class Env>
PW_HD bool static_blocked(const Env& env, uint32_t dj, int x, int y, int dx, int dy, int soff) {
  const int w = desc_w(dj), h = desc_h(dj);
  const int W = env.width(), H = env.height();
  if (x < 0 || y < 0 || x + w > W || y + h > H) return false;  // puzzle.py:557-561
  const int o = desc_off(dj);
  const int xs = x + dx;  // >= -1
  uint64_t hit = 0;
  for (int r = 0; r < h; ++r) {
    const int yy = y + dy + r;
    if (yy < 0 || yy >= H) continue;  // moved cell outside the grid: no obstacle there
    const uint64_t a = env.row(o + r);
    // xs == 64 only for a 1-wide object in column 63 of a 64-wide grid moving right
    const uint64_t m = xs < 0 ? (a >> 1) : (xs > 63 ? 0ull : (a << xs));
    hit |= m & env.row(soff + yy);
  }
  if (hit == 0) return false;
  uint64_t cur = 0;  // "not already overlapping" term, puzzle.py:562
  for (int r = 0; r < h; ++r) cur |= (env.row(o + r) << x) & env.row(soff + y + r);
  return cur == 0;
}

// One transition.  Returns the moved-object bitmask (bit 0 = agent), 0 if nothing moved.
// `Env` provides: n(), width(), height(), wall_off(), awall_off(), row(off), desc(j),
// x(j), y(j), and the *_dyn(j) variants for a runtime index, translate(mask, dx, dy),
// plus kStaticN (compile-time loop bound, 0 = use n()).
//
// For every pusher i the not-yet-pushed objects are handled in two phases so that the lanes of
// a warp (one environment each) stay converged:
//   1. a cheap bounding-box test of the MOVED pusher against every candidate j -> bit mask;
//   2. only for the (few) candidates the exact row-mask tests and the wall test of the pushee.
// Without the split, lanes whose pushes happen at different j would each execute the heavy
// path on their own.
template <class Env>
PW_HD uint32_t step(Env& env, int action) {
  if ((unsigned)action > 3u) return 0;  // invalid actions are no-ops (counted by the caller)
  const int dx = action_dx(action), dy = action_dy(action);
  if (static_blocked(env, env.desc(0), env.x(0), env.y(0), dx, dy, env.awall_off())) return 0;

  const int n = env.n();
  const int jend = Env::kStaticN ? Env::kStaticN : n;
  uint32_t pushed = 1u, frontier = 1u;
  while (frontier) {
    const int i = pw_ffs(frontier) - 1;
    frontier &= frontier - 1;
    const uint32_t di = env.desc_dyn(i);
    const int xm = env.x_dyn(i) + dx, ym = env.y_dyn(i) + dy;  // origin of the moved pusher
    const int wi = desc_w(di), hi = desc_h(di);
    // phase 1: -w_j < x_j - xm < w_i  and  -h_j < y_j - ym < h_i, as two unsigned range checks
    uint32_t cand = 0;
#pragma unroll
    for (int j = 1; j < jend; ++j) {
      if (Env::kStaticN && j >= n) break;
      const uint32_t dj = env.desc(j);
      const int wj1 = desc_w(dj) - 1, hj1 = desc_h(dj) - 1;
      const bool near_x = (unsigned)(env.x(j) - xm + wj1) < (unsigned)(wi + wj1);
      const bool near_y = (unsigned)(env.y(j) - ym + hj1) < (unsigned)(hi + hj1);
      cand |= (near_x && near_y) ? (1u << j) : 0u;
    }
    cand &= ~pushed;
    // phase 2: exact tests for the candidates
    while (cand) {
      const int j = pw_ffs(cand) - 1;
      cand &= cand - 1;
      const uint32_t dj = env.desc_dyn(j);
      const int xj = env.x_dyn(j), yj = env.y_dyn(j);
      if (!shapes_overlap(env, di, dj, xj - xm, yj - ym)) continue;            // would not touch
      if (shapes_overlap(env, di, dj, xj - xm + dx, yj - ym + dy)) continue;   // already overlapping
      if (static_blocked(env, dj, xj, yj, dx, dy, env.wall_off())) return 0;
      pushed |= 1u << j;
      frontier |= 1u << j;
    }
  }
  env.translate(pushed, dx, dy);
  return pushed;
}

// puzzle.py:396-407: number of goal objects (state index 1..g) that sit on their goal.
template <class Env>
PW_HD int achieved_goals(const Env& env) {
  const int g = env.g();
  const int kend = Env::kStaticN ? Env::kStaticN - 1 : g;
  int count = 0;
#pragma unroll
  for (int k = 0; k < kend; ++k) {
    if (Env::kStaticN && k >= g) break;
    count += (env.x(k + 1) == env.goal_x(k) && env.y(k + 1) == env.goal_y(k)) ? 1 : 0;
  }
  return count;
}

}  // namespace pw
