// pw_occ.cu -- the general-rule rollout kernel (BASELINE config 3: mixed batch over all puzzles).
//
// What a random-policy step of the 223 benchmark puzzles looks like (measured with the C oracle,
// 7.3e6 steps): 84 % the agent walks into free space, 10 % it is stopped by a wall, 5 % it touches
// a movable (2 % end in transitive stopping, 3 % push 1.35 objects on average).  The table rule of
// pw_fast.cuh pays a scan over all n-1 other objects on EVERY step to find that out.  This kernel
// splits a step into the part every environment needs and the part few need:
//
//  * ONE THREAD PER ENVIRONMENT for the common step.  Per environment, the union occupancy of its
//    movables is kept as one uint32 per grid row in shared memory (puzzles up to 32 x 32 cells
//    incl. the border: 211 of the 223; the north star's occupancy bitboards).  The step is then ONE
//    collision-table look-up for the walls (pw_fast.cuh's agent table, exact incl. the "not already
//    overlapping" clause of puzzle.py:562) and a test of the moved agent's rows against the
//    occupancy: if it touches no movable cell, pushes(agent, j) is false for every j
//    (puzzle.py:567-593 needs the moved pusher to overlap j) and the agent just moves.  Only the
//    agent's position lives in registers; the fast path is branch-free, 32 environments per warp.
//  * ONE WARP PER ENVIRONMENT for the push closure (puzzle.py:359-394), the layout BASELINE.json's
//    north star names: the lanes whose agent touches something are served one after the other by
//    the whole warp, lane j <-> object j.  A round (one pusher) is: the pusher's displaced rows
//    against the occupancy (lane r <-> row r, one __reduce_or_sync) -- no hit, no push, next pusher
//    --, then ONE pair-table look-up per lane (tab_pushes, only objects spanning a hit row) and one
//    __ballot_sync, one wall look-up per hit object (tab_wall) and one ballot for transitive
//    stopping (puzzle.py:376-379).  Pushed objects leave the occupancy and re-enter it displaced
//    (shared-memory atomics: several lanes may touch one row word).  The cost of a closure does not
//    depend on the number of objects.
//  * Environments without occupancy rows (grid wider or taller than 32, an agent taller than 4
//    rows, or an irregular state: a movable outside the grid / overlapping another movable or a
//    wall) take the closure on every step the walls do not stop -- warp-cooperatively when they are
//    few in a warp, with the per-thread table rule of pw_fast.cuh (pushed_set_tab) when they are
//    many (a warp full of 47 x 54 puzzles runs exactly the round-1 rule).
//  * actions come from cp.async tiles ([32 steps][32 envs] per warp, double buffered); reward f32 +
//    terminated u8 (pw_rollout) or one record byte (pw_rollout_packed) go out coalesced per step.
//
// Results do not depend on the schedule: every step of an environment is executed by exactly one
// of the exact paths, all of which read the reference's collision relations from the same tables.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>

#include "pw_env.cuh"
#include "pw_fast.cuh"
#include "pw_internal.h"
#include "pw_steps.cuh"

namespace pw {
namespace {

constexpr int kOccBlock = 128;     // threads (= environments) per CTA
constexpr int kOccRows = 32;       // occupancy rows per environment (uint32 each)
constexpr int kPosPitch = kOccBlock + 1;  // odd pitch: a column (one environment, all objects) is conflict free too
constexpr int kTileBytes = 32 * 32;
constexpr int kGenericLanes = 6;   // this many closure-every-step lanes in a warp: serve them per thread

// State policy (see pw_env.cuh for the interface): the agent's position lives in registers (it
// changes in most steps), the other objects' positions in shared memory, column `threadIdx.x` of an
// [n_max - 1][kPosPitch] int32 array per coordinate.  Tables are read from global memory through
// the read-only path.
struct OccEnv : EnvBase<GlobalTable> {
  static constexpr int kStaticN = 0;
  static constexpr int kUnroll = 1;
  int ax, ay;
  int* sx;  // object j >= 1 at sx[(j - 1) * kPosPitch]
  int* sy;
  // masks: sy holds, instead of y, the ROW MASK of the object -- bit r set for every grid row
  // y .. y + h - 1 it spans (y = lowest set bit): the candidate filter of the closure.  Only for
  // environments with occupancy rows (see build_occupancy).
  bool masks = false;
  __device__ __forceinline__ void init(GlobalTable table, int* x_col, int* y_col) {
    this->init_header(table);
    sx = x_col;
    sy = y_col;
  }
  __device__ __forceinline__ void load(const int32_t* x, const int32_t* y, int64_t stride, int64_t e) {
    ax = x[e];
    ay = y[e];
    for (int j = 1; j < this->n_; ++j) {
      sx[(j - 1) * kPosPitch] = x[j * stride + e];
      sy[(j - 1) * kPosPitch] = y[j * stride + e];
    }
  }
  __device__ __forceinline__ void store(int32_t* x, int32_t* y, int64_t stride, int64_t e) const {
    x[e] = ax;
    y[e] = ay;
    for (int j = 1; j < this->n_; ++j) {
      x[j * stride + e] = sx[(j - 1) * kPosPitch];
      y[j * stride + e] = this->y(j);
    }
  }
  __device__ __forceinline__ uint32_t desc(int j) const { return this->table_desc(j, 0); }
  __device__ __forceinline__ uint32_t desc_dyn(int j) const { return this->table_desc(j, 0); }
  // j >= 1 in every loop over pushees / goal objects; index 0 only through x(0) / x_dyn(pusher)
  __device__ __forceinline__ int x(int j) const { return j == 0 ? ax : sx[(j - 1) * kPosPitch]; }
  __device__ __forceinline__ int y(int j) const {
    if (j == 0) return ay;
    const int v = sy[(j - 1) * kPosPitch];
    return masks ? __ffs(v) - 1 : v;
  }
  __device__ __forceinline__ int x_dyn(int j) const { return x(j); }
  __device__ __forceinline__ int y_dyn(int j) const { return y(j); }
  __device__ __forceinline__ int goal_x(int k) const { return lo16(this->table_desc(k + 1, 2)); }
  __device__ __forceinline__ int goal_y(int k) const { return hi16(this->table_desc(k + 1, 2)); }
  __device__ __forceinline__ void translate(uint32_t mask, int dx, int dy) {  // plain positions only
    if (mask & 1u) {
      ax += dx;
      ay += dy;
    }
    mask &= ~1u;
    while (mask) {
      const int j = __ffs((int)mask) - 1;
      mask &= mask - 1;
      if (dx) sx[(j - 1) * kPosPitch] += dx;
      if (dy) sy[(j - 1) * kPosPitch] += dy;
    }
  }
};

// Builds the occupancy rows of this thread's environment from the positions of objects 1..n-1
// (occ: &s_occ[0][tid], row r at occ[r * kOccBlock]) and reports whether the environment can use
// them: every movable inside the grid, no two movables overlapping, none overlapping a wall.  Such a
// state stays that way: a movable that does not overlap a wall is stopped by the border walls
// (puzzle.py:160-168, 284-293), and a movable that would be moved onto another one pushes it
// (:359-382), so the occupancy can be maintained incrementally.
__device__ __forceinline__ bool build_occupancy(const OccEnv& env, uint32_t* occ, int W, int H) {
  for (int r = 0; r < H; ++r) occ[r * kOccBlock] = 0u;
  const int n = env.n(), wall_off = env.wall_off();
  bool ok = true;
  for (int j = 1; j < n; ++j) {
    const uint32_t d = env.desc(j);
    const int x = env.x(j), y = env.y(j);
    const int w = desc_w(d), h = desc_h(d), o = desc_off(d);
    if (x < 0 || y < 0 || x + w > W || y + h > H) {
      ok = false;
      continue;
    }
    for (int r = 0; r < h; ++r) {
      const uint64_t bits = env.row(o + r) << x;
      const uint32_t have = occ[(y + r) * kOccBlock];
      ok = ok && (have & (uint32_t)bits) == 0u && (env.row(wall_off + y + r) & bits) == 0ull;
      occ[(y + r) * kOccBlock] = have | (uint32_t)bits;
    }
  }
  return ok;
}

// The push closure of one step (puzzle.py:359-394) for an environment with occupancy rows, by the
// thread that owns it:
//   * a pusher whose displaced cells touch no occupied cell pushes nothing (puzzle.py:567-593 needs
//     the moved pusher to overlap the pushee), so its scan over the other objects is skipped;
//   * otherwise only objects that span one of the rows with a hit are tested, with the exact
//     collision-table look-ups of pw_fast.cuh (tab_pushes, tab_wall);
//   * objects that join the pushed set leave the occupancy and re-enter it displaced (or where they
//     were, on transitive stopping, puzzle.py:376-379).
// `members` = the lanes of the warp that execute this call together: every loop whose trip count
// depends on the environment is headed by a vote over them, so the lanes stay converged.
// Returns the moved-object mask (0 = nothing moves); positions and occupancy are updated.
__device__ __forceinline__ uint32_t closure_thread(OccEnv& env, uint32_t* occ, uint32_t a, int H, uint32_t nn1,
                                                   const uint32_t members) {
  const int n = env.n();
  const int dx = action_dx((int)a), dy = action_dy((int)a);
  // (the caller has already found that the walls do not stop the agent, puzzle.py:353-354)
  uint32_t pushed = 1u, frontier = 1u;
  bool stop = false;
  while (__any_sync(members, frontier != 0u)) {
    uint32_t cand = 0u;
    int i = 0, xi = 0, yi = 0;
    if (frontier != 0u) {
      i = __ffs((int)frontier) - 1;
      frontier &= frontier - 1;
      const uint32_t di = env.desc(i);
      xi = env.x(i);
      yi = env.y(i);
      const int hi = desc_h(di), oi = desc_off(di);
      const int xm = xi + dx, ym = yi + dy;
      uint32_t hitrows = 0u;
      if (xm > -64 && xm < 32) {  // (only the agent can be outside the grid)
        for (int r = 0; r < hi; ++r) {
          const int yy = ym + r;
          if ((unsigned)yy >= (unsigned)H) continue;
          const uint64_t row = env.row(oi + r);
          const uint32_t bits = xm >= 0 ? (uint32_t)(row << xm) : (uint32_t)(row >> (-xm));
          hitrows |= (bits & occ[yy * kOccBlock]) ? (1u << yy) : 0u;
        }
      }
      if (hitrows != 0u) {
#pragma unroll 4
        for (int j = 1; j < n; ++j) cand |= ((uint32_t)env.sy[(j - 1) * kPosPitch] & hitrows) ? (1u << j) : 0u;
        cand &= ~pushed;
      }
    }
    while (__any_sync(members, cand != 0u)) {
      if (cand != 0u) {
        const int j = __ffs((int)cand) - 1;
        cand &= cand - 1;
        const int xj = env.x(j), yj = env.y(j);
        if (tab_pushes(env, 1u + (uint32_t)(i * n + j), a, xj - xi, yj - yi)) {
          if (tab_wall(env, nn1, (uint32_t)j, a, xj, yj)) {  // transitive stopping
            stop = true;
            cand = 0u;
            frontier = 0u;
          } else {
            pushed |= 1u << j;
            frontier |= 1u << j;
            const uint32_t dj = env.desc(j);
            const int hj = desc_h(dj), oj = desc_off(dj);
            for (int r = 0; r < hj; ++r) occ[(yj + r) * kOccBlock] &= ~(uint32_t)(env.row(oj + r) << xj);
          }
        }
      }
    }
  }
  uint32_t rest = pushed & ~1u;
  while (__any_sync(members, rest != 0u)) {
    if (rest != 0u) {
      const int j = __ffs((int)rest) - 1;
      rest &= rest - 1;
      const uint32_t dj = env.desc(j);
      const int hj = desc_h(dj), oj = desc_off(dj);
      const int nx = env.x(j) + (stop ? 0 : dx), ny = env.y(j) + (stop ? 0 : dy);
      for (int r = 0; r < hj; ++r) occ[(ny + r) * kOccBlock] |= (uint32_t)(env.row(oj + r) << nx);
      if (!stop) {
        env.sx[(j - 1) * kPosPitch] = nx;
        const uint32_t m = (uint32_t)env.sy[(j - 1) * kPosPitch];
        env.sy[(j - 1) * kPosPitch] = (int)(dy > 0 ? m << 1 : (dy < 0 ? m >> 1 : m));
      }
    }
  }
  if (stop) return 0u;
  env.ax += dx;
  env.ay += dy;
  return pushed;
}

__device__ __forceinline__ const uint64_t* shfl_ptr(const uint64_t* p, int src) {
  const unsigned long long v = reinterpret_cast<unsigned long long>(p);
  const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src), hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
  return reinterpret_cast<const uint64_t*>(((unsigned long long)hi << 32) | lo);
}

// What one closure tells the lane that owns the environment.
struct ClosureResult {
  uint32_t moved;  // moved-object mask, 0 = nothing moves
  int achieved;    // goals achieved afterwards (valid if a movable moved)
};

// The push closure of ONE environment -- the one of lane `owner` -- executed by the whole warp:
// lane j <-> object j (PW_MAX_OBJECTS = 32).  Called by all 32 lanes with the same `owner`.
//   env        this lane's own environment (the owner's table pointers etc. are broadcast from it)
//   col        the owner's column in the shared-memory arrays (s_occ, positions)
//   use_occ    the owner's environment has occupancy rows (and row masks instead of y)
// Updates the owner's positions / occupancy in shared memory; the owner's agent position is
// updated by the caller from the returned mask.
__device__ __forceinline__ ClosureResult closure_warp(const OccEnv& env, const int owner, const uint32_t a,
                                                      uint32_t* s_occ, int* s_posx, int* s_posy, const int col,
                                                      const bool use_occ, const uint32_t lane) {
  // the owner's puzzle, as a table view every lane can read
  EnvBase<GlobalTable> we;
  we.t.base = shfl_ptr(env.t.base, owner);
  we.ft.base = shfl_ptr(env.ft.base, owner);
  const int n = __shfl_sync(0xffffffffu, env.n_, owner);
  const int g = __shfl_sync(0xffffffffu, env.g_, owner);
  const int H = __shfl_sync(0xffffffffu, env.h_, owner);
  we.desc_word_ = __shfl_sync(0xffffffffu, env.desc_word_, owner);
  we.n_ = n;
  const uint32_t nn1 = 1u + (uint32_t)(n * n);
  const int dx = action_dx((int)a), dy = action_dy((int)a);
  // lane j: object j (lane 0: the agent)
  const bool live = lane >= 1u && (int)lane < n;
  const uint32_t dj = (int)lane < n ? we.table_desc((int)lane, 0) : 0u;
  int xj = __shfl_sync(0xffffffffu, env.ax, owner), yj = __shfl_sync(0xffffffffu, env.ay, owner);
  uint32_t rowmask = 0u;
  if (live) {
    xj = s_posx[(lane - 1u) * kPosPitch + col];
    const int v = s_posy[(lane - 1u) * kPosPitch + col];
    rowmask = (uint32_t)v;
    yj = use_occ ? __ffs(v) - 1 : v;
  }
  const int hj = desc_h(dj), oj = desc_off(dj);
  uint32_t pushed = 1u, frontier = 1u;
  bool stop = false;
  while (frontier != 0u) {  // warp-uniform
    const int i = __ffs((int)frontier) - 1;
    frontier &= frontier - 1;
    const int xi = __shfl_sync(0xffffffffu, xj, i), yi = __shfl_sync(0xffffffffu, yj, i);
    uint32_t hitrows = 0xffffffffu;
    if (use_occ) {
      // the displaced pusher against the occupancy: lane r <-> its shape row r (h_i <= 32 rows)
      const uint32_t di = __shfl_sync(0xffffffffu, dj, i);
      const int hi = desc_h(di), xm = xi + dx, yy = yi + dy + (int)lane;
      uint32_t mine = 0u;
      if ((int)lane < hi && (unsigned)yy < (unsigned)H && xm > -64 && xm < 32) {  // (only the agent can leave the grid)
        const uint64_t row = we.row(desc_off(di) + (int)lane);
        const uint32_t bits = xm >= 0 ? (uint32_t)(row << xm) : (uint32_t)(row >> (-xm));
        mine = (bits & s_occ[yy * kOccBlock + col]) ? (1u << yy) : 0u;
      }
      hitrows = __reduce_or_sync(0xffffffffu, mine);
      if (hitrows == 0u) continue;  // touches nothing: pushes nothing (puzzle.py:567-593)
    }
    const bool candidate = live && !((pushed >> lane) & 1u) && (!use_occ || (rowmask & hitrows) != 0u);
    const bool push = candidate && tab_pushes(we, 1u + (uint32_t)(i * n) + lane, a, xj - xi, yj - yi);
    const uint32_t hits = __ballot_sync(0xffffffffu, push);
    if (hits == 0u) continue;
    if (__ballot_sync(0xffffffffu, push && tab_wall(we, nn1, lane, a, xj, yj)) != 0u) {
      stop = true;  // transitive stopping, puzzle.py:376-379
      break;
    }
    pushed |= hits;
    frontier |= hits;
    if (use_occ && push)  // the object leaves the occupancy: it moves along (or is put back below)
      for (int r = 0; r < hj; ++r) atomicAnd(&s_occ[(yj + r) * kOccBlock + col], ~(uint32_t)(we.row(oj + r) << xj));
  }
  const bool mine_moved = live && ((pushed >> lane) & 1u);
  if (mine_moved) {
    const int nx = xj + (stop ? 0 : dx), ny = yj + (stop ? 0 : dy);
    if (use_occ)
      for (int r = 0; r < hj; ++r) atomicOr(&s_occ[(ny + r) * kOccBlock + col], (uint32_t)(we.row(oj + r) << nx));
    if (!stop) {
      xj = nx;
      yj = ny;
      s_posx[(lane - 1u) * kPosPitch + col] = nx;
      s_posy[(lane - 1u) * kPosPitch + col] =
          use_occ ? (int)(dy > 0 ? rowmask << 1 : (dy < 0 ? rowmask >> 1 : rowmask)) : ny;
    }
  }
  ClosureResult res;
  res.moved = stop ? 0u : pushed;
  res.achieved = 0;
  if (res.moved & ~1u) {  // puzzle.py:396-407: goal objects 1 .. g on their goal positions
    const bool goal_lane = lane >= 1u && (int)lane <= g;
    const uint32_t gw = goal_lane ? we.table_desc((int)lane, 2) : kNoGoal;
    res.achieved = __popc(__ballot_sync(0xffffffffu, goal_lane && xj == lo16(gw) && yj == hi16(gw)));
  }
  __syncwarp();  // the owner's shared-memory state is complete before anyone reads it again
  return res;
}

template <bool kPacked>
__global__ void __launch_bounds__(kOccBlock)
    k_rollout_occ(const __grid_constant__ StepParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ uint32_t s_counters[4];
  // dynamic shared memory: occupancy | positions x | positions y | action tiles | reward table
  uint32_t* s_occ = reinterpret_cast<uint32_t*>(smem_raw);
  const int pos_rows = p.n_max > 1 ? p.n_max - 1 : 1;
  int* s_posx = reinterpret_cast<int*>(s_occ + kOccRows * kOccBlock);
  int* s_posy = s_posx + pos_rows * kPosPitch;
  uint8_t* s_act = reinterpret_cast<uint8_t*>(s_posy + pos_rows * kPosPitch);
  s_act += (16 - (reinterpret_cast<uintptr_t>(s_act) & 15u)) & 15u;  // cp.async needs 16-byte aligned tiles
  float* s_lut = reinterpret_cast<float*>(s_act + (kOccBlock / 32) * 2 * kTileBytes);
  if (threadIdx.x < 4) s_counters[threadIdx.x] = 0;
  if (!kPacked) {
    // record byte -> reward (gym_env.py:212-221): 10 if terminated else (double)dg - 0.01, rounded
    // to float32 once; IEEE arithmetic, so identical to the host-side table of the other kernels
    for (int c = threadIdx.x; c < 256; c += kOccBlock)
      s_lut[c] = (c & 0x80) ? 10.0f : (float)((double)((c & 0x3f) - kMaxObjects) - 0.01);
  }
  __syncthreads();

  const int tid = threadIdx.x;
  const uint32_t lane = tid & 31u, warp = tid >> 5;
  const int64_t e = (int64_t)blockIdx.x * kOccBlock + tid;
  StepCounts c;
  if (e < p.N) {  // whole warps (the launcher requires N % 32 == 0)
    const int pid = p.puzzle_id ? checked_puzzle_id(p.puzzle_id[e], p.num_puzzles, p.status) : p.puzzle_uniform;
    OccEnv env;
    env.init(GlobalTable{p.blob + p.puzzle_off[pid]}, s_posx + tid, s_posy + tid);
    env.load(p.x, p.y, p.stride, e);
    const int n = env.n(), g = env.g(), W = env.width(), H = env.height();
    // the agent's wall table (pw_fast.cuh: tab_wall), descriptor cached in registers
    const uint64_t wd = env.fast(1u + (uint32_t)(n * n));
    const uint32_t aw_hi = (uint32_t)(wd >> 32);
    const uint32_t aw_xmax = aw_hi & 63u, aw_ymax = (aw_hi >> 6) & 0xfffu, aw_w = aw_hi >> 18;
    const uint32_t aw_word0 = (uint32_t)wd << 1;
    const uint32_t d0 = env.desc(0);
    const int h0 = desc_h(d0), o0 = desc_off(d0);
    // the agent's shape rows in registers (bounding boxes of up to 4 rows: every benchmark agent)
    uint32_t arow[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) arow[r] = r < h0 ? (uint32_t)env.row(o0 + r) : 0u;
    uint32_t* const occ = s_occ + tid;
    // occupancy rows: puzzles of at most 32 x 32 cells whose state is regular (build_occupancy)
    const bool fast_ok = W <= 32 && H <= kOccRows && h0 <= 4 && build_occupancy(env, occ, W, H);
    if (fast_ok) {  // row masks instead of y for objects 1 .. n-1 (see OccEnv::masks)
      for (int j = 1; j < n; ++j) {
        const int h = desc_h(env.desc(j));
        env.sy[(j - 1) * kPosPitch] = (int)((0xffffffffu >> (32 - h)) << env.sy[(j - 1) * kPosPitch]);
      }
      env.masks = true;
    }
    __syncwarp();
    int& ax = env.ax;
    int& ay = env.ay;
    int ach = achieved_goals(env);
    uint32_t code_idle = record_byte(0, ach == g, false);  // the record of a step that moves no movable

    const int T = p.T;
    uint8_t* const tiles = s_act + warp * (2 * kTileBytes);
    const uint32_t tiles_s = (uint32_t)__cvta_generic_to_shared(tiles);
    const uint8_t* const warp_actions = p.actions + (e - (int64_t)lane);
    auto issue_block = [&](const int block) {  // lane l copies the warp's 32 actions of step 32*block + l
      const int row = block * 32 + (int)lane;
      const uint8_t* src = warp_actions + (size_t)(row < T ? row : T - 1) * (size_t)p.N;  // clamped: never used
      const uint32_t dst = tiles_s + (uint32_t)(block & 1) * kTileBytes + lane * 32u;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u), "l"(src + 16) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // output pointers of this lane's column, advanced by N per step
    float* rew = kPacked ? nullptr : p.reward + e;
    uint8_t* out8 = kPacked ? p.record + e : p.terminated + e;
    uint32_t n_term = 0, n_moved = 0, n_gain = 0, n_invalid = 0;
    // actions of a step as packed displacement bytes: LEFT, RIGHT, UP, DOWN (puzzle.py:43-50)
    constexpr uint32_t kDxBytes = 0x000001ffu, kDyBytes = 0x01ff0000u;
    const int blocks = (T + 31) / 32;
    issue_block(0);
    for (int blk = 0; blk < blocks; ++blk) {
      __syncwarp();          // every lane is done with the tile the next block lands in
      issue_block(blk + 1);  // past the end: re-reads the last row, never used
      asm volatile("cp.async.wait_group 1;" ::: "memory");
      __syncwarp();  // block `blk` is visible to the whole warp
      const uint8_t* mine = tiles + (blk & 1) * kTileBytes + lane;
      const int steps = T - blk * 32 < 32 ? T - blk * 32 : 32;
#pragma unroll 1
      for (int k = 0; k < steps; ++k, mine += 32) {
        // ---- the common step: one thread per environment, branch-free ---------------------------
        const uint32_t a = mine[0];
        const bool valid = a <= 3u;
        const uint32_t sel = a * 0x1111u + 0x8880u;  // byte a, sign extended
        // (raw prmt: __byte_perm masks the selector and would drop the sign-replication bits)
        int dx, dy;
        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(dx) : "r"(kDxBytes), "r"(0u), "r"(sel));
        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(dy) : "r"(kDyBytes), "r"(0u), "r"(sel));
        // walls (puzzle.py:353-354): one bit of the agent's collision table
        const bool here = (uint32_t)ax <= aw_xmax && (uint32_t)ay <= aw_ymax;
        const uint32_t idx = here ? ((((uint32_t)ay * aw_w + (uint32_t)ax) << 2) | (a & 3u)) : 0u;
        const bool stopped = here && tab_bit(env.fast32(aw_word0 + (idx >> 5)), idx);
        const int x = ax + dx, y = ay + dy;
        // does the moved agent (inside the grid: rows y .. y + h0 - 1 exist) touch a movable?
        const bool there = fast_ok && (uint32_t)x <= aw_xmax && (uint32_t)y <= aw_ymax;
        // (rows beyond the agent's bounding box have arow = 0: their clamped reads do not matter)
        const int y0 = there ? y : 0;
        uint32_t hit = 0u;
#pragma unroll
        for (int r = 0; r < 4; ++r) hit |= (arow[r] << (x & 31)) & occ[min(y0 + r, kOccRows - 1) * kOccBlock];
        const bool go = valid && !stopped;
        const bool free_move = go && there && hit == 0u;
        n_invalid += valid ? 0u : 1u;
        n_moved += free_move ? 1u : 0u;
        ax = free_move ? x : ax;
        ay = free_move ? y : ay;
        uint32_t code = code_idle;
        // ---- the push closure for the environments that need it --------------------------------
        const bool wants = go && !free_move;
        uint32_t need = __ballot_sync(0xffffffffu, wants);
        if (need != 0u) {
          const uint32_t need_generic = __ballot_sync(0xffffffffu, wants && !fast_ok);
          uint32_t moved_mask = 0u;
          int now = ach;
          if (p.occ_warp_closure == 0 || __popc(need_generic) >= kGenericLanes) {
            // environments without occupancy rows: the per-thread table rule, together
            if (wants && !fast_ok) {
              moved_mask = pushed_set_tab(env, (int)a);
              if (moved_mask) env.translate(moved_mask, action_dx((int)a), action_dy((int)a));
            }
            need &= ~need_generic;
            __syncwarp();
          }
          if (p.occ_warp_closure == 0) {
            // ... and those with occupancy rows: the per-thread closure, lanes converged
            if (wants && fast_ok) moved_mask = closure_thread(env, occ, a, H, 1u + (uint32_t)(n * n), need);
            if (wants && (moved_mask & ~1u)) now = achieved_goals(env);
            __syncwarp();
          } else {
            if (wants && !fast_ok && (moved_mask & ~1u)) now = achieved_goals(env);
            while (need != 0u) {  // one warp per environment, one after the other (PW_OCC_WARP_CLOSURE=1)
              const int owner = __ffs((int)need) - 1;
              need &= need - 1;
              const uint32_t a_owner = __shfl_sync(0xffffffffu, a, owner);
              const bool occ_owner = __shfl_sync(0xffffffffu, (int)fast_ok, owner) != 0;
              const ClosureResult res = closure_warp(env, owner, a_owner, s_occ, s_posx, s_posy,
                                                     (int)(warp * 32u) + owner, occ_owner, lane);
              if ((int)lane == owner) {
                moved_mask = res.moved;
                if (moved_mask) {
                  ax = x;
                  ay = y;
                }
                if (moved_mask & ~1u) now = res.achieved;
              }
            }
          }
          if (wants) {
            n_moved += moved_mask ? 1u : 0u;
            const int dg = now - ach;
            ach = now;
            const bool term = ach == g;  // puzzle.py:409-411
            code = record_byte(dg, term, false);
            code_idle = record_byte(0, term, false);
            n_gain += (uint32_t)max(dg, 0);
          }
        }
        n_term += code >> 7;
        if (kPacked) {
          *out8 = (uint8_t)code;
        } else {
          *rew = s_lut[code];
          *out8 = (uint8_t)(code >> 7);
          rew += p.N;
        }
        out8 += p.N;
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    c.terminated += n_term;
    c.moved += n_moved;
    c.dgoals += n_gain;
    c.invalid += n_invalid;
    env.store(p.x, p.y, p.stride, e);
  }
  flush_counts(c, p, s_counters);
}

// ------------------------------------------------------------------------------------------
// The same step rule with the lanes of a warp RUNNING AHEAD of each other (k_rollout_occ_ra).
// In the lock-step kernel above some lane needs the closure in 3 of 4 steps, so its cost is paid
// almost every step for one or two environments.  Here every lane has its own step index inside a
// window of 64 steps (two 32-step tiles): a lane whose agent touches something PARKS that step and
// the others go on; when enough lanes are parked -- or nobody else can advance, or somebody waits at
// the window limit for a parked lane -- the parked lanes run the closure TOGETHER.  Results are
// staged as one record byte per step in a shared-memory ring and leave it coalesced whenever the
// whole warp has left a tile.  An action found to move nothing is remembered until the
// environment's state changes (blocked_actions): an agent that keeps bumping into a stuck object
// does not park again.
// ------------------------------------------------------------------------------------------

// Pushed set of one step for parked lanes (same contract as closure_thread): per pusher the
// branch-free scan of pw_fast.cuh's pushed_set_tab (one pair look-up per object), skipped when the
// pusher's displaced rows touch no occupied cell.
__device__ __forceinline__ uint32_t closure_scan(OccEnv& env, uint32_t* occ, uint32_t a, int H, uint32_t nn1,
                                                 const uint32_t members) {
  const int n = env.n();
  const int dx = action_dx((int)a), dy = action_dy((int)a);
  uint32_t pushed = 1u, frontier = 1u;
  bool stop = false;
  while (__any_sync(members, frontier != 0u)) {
    uint32_t hits = 0u;
    const bool active = frontier != 0u;
    const int i = active ? __ffs((int)frontier) - 1 : 0;
    frontier &= frontier - 1;
    const int xi = env.x(i), yi = env.y(i);
    bool touches = false;
    if (active) {
      const uint32_t di = env.desc(i);
      const int hi = desc_h(di), oi = desc_off(di);
      const int xm = xi + dx, ym = yi + dy;
      if (xm > -64 && xm < 32) {  // (only the agent can be outside the grid)
        for (int r = 0; r < hi; ++r) {
          const int yy = ym + r;
          if ((unsigned)yy >= (unsigned)H) continue;
          const uint64_t row = env.row(oi + r);
          const uint32_t bits = xm >= 0 ? (uint32_t)(row << xm) : (uint32_t)(row >> (-xm));
          touches = touches || (bits & occ[yy * kOccBlock]) != 0u;
        }
      }
    }
    if (__any_sync(members, touches)) {
      if (touches) {
        const uint32_t base = 1u + (uint32_t)(i * n);
#pragma unroll 4
        for (int j = 1; j < n; ++j) {
          const bool push = tab_pushes(env, base + (uint32_t)j, a, env.sx[(j - 1) * kPosPitch] - xi, env.y(j) - yi);
          hits = (hits >> 1) | (push ? 0x80000000u : 0u);
        }
        hits = n > 1 ? hits >> (32 - n) : 0u;
        hits &= ~pushed;
      }
      uint32_t check = hits;
      while (__any_sync(members, check != 0u)) {
        if (check != 0u) {
          const int j = __ffs((int)check) - 1;
          check &= check - 1;
          const int xj = env.x(j), yj = env.y(j);
          if (tab_wall(env, nn1, (uint32_t)j, a, xj, yj)) {  // transitive stopping, puzzle.py:376-379
            stop = true;
            check = 0u;
            hits = 0u;
            frontier = 0u;
          } else {  // the object leaves the occupancy: it moves along (or is put back below)
            const uint32_t dj = env.desc(j);
            const int hj = desc_h(dj), oj = desc_off(dj);
            for (int r = 0; r < hj; ++r) occ[(yj + r) * kOccBlock] &= ~(uint32_t)(env.row(oj + r) << xj);
            pushed |= 1u << j;
          }
        }
      }
      if (!stop) frontier |= hits;
    }
  }
  uint32_t rest = pushed & ~1u;
  while (__any_sync(members, rest != 0u)) {
    if (rest != 0u) {
      const int j = __ffs((int)rest) - 1;
      rest &= rest - 1;
      const uint32_t dj = env.desc(j);
      const int hj = desc_h(dj), oj = desc_off(dj);
      const int nx = env.x(j) + (stop ? 0 : dx), ny = env.y(j) + (stop ? 0 : dy);
      for (int r = 0; r < hj; ++r) occ[(ny + r) * kOccBlock] |= (uint32_t)(env.row(oj + r) << nx);
      if (!stop) {
        env.sx[(j - 1) * kPosPitch] = nx;
        const uint32_t m = (uint32_t)env.sy[(j - 1) * kPosPitch];
        env.sy[(j - 1) * kPosPitch] = (int)(dy > 0 ? m << 1 : (dy < 0 ? m >> 1 : m));
      }
    }
  }
  if (stop) return 0u;
  env.ax += dx;
  env.ay += dy;
  return pushed;
}

constexpr int kActTiles = 3;  // action tiles per warp: two readable + one in flight

template <bool kPacked>
__global__ void __launch_bounds__(kOccBlock)
    k_rollout_occ_ra(const __grid_constant__ StepParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ uint32_t s_counters[4];
  // dynamic shared memory: occupancy | positions x | positions y | action tiles | result ring | reward table
  uint32_t* s_occ = reinterpret_cast<uint32_t*>(smem_raw);
  const int pos_rows = p.n_max > 1 ? p.n_max - 1 : 1;
  int* s_posx = reinterpret_cast<int*>(s_occ + kOccRows * kOccBlock);
  int* s_posy = s_posx + pos_rows * kPosPitch;
  uint8_t* s_act = reinterpret_cast<uint8_t*>(s_posy + pos_rows * kPosPitch);
  s_act += (16 - (reinterpret_cast<uintptr_t>(s_act) & 15u)) & 15u;  // cp.async needs 16-byte aligned tiles
  uint8_t* s_out = s_act + (kOccBlock / 32) * kActTiles * kTileBytes;
  float* s_lut = reinterpret_cast<float*>(s_out + (kOccBlock / 32) * 2 * kTileBytes);
  if (threadIdx.x < 4) s_counters[threadIdx.x] = 0;
  if (!kPacked) {
    for (int c = threadIdx.x; c < 256; c += kOccBlock)
      s_lut[c] = (c & 0x80) ? 10.0f : (float)((double)((c & 0x3f) - kMaxObjects) - 0.01);
  }
  __syncthreads();

  const int tid = threadIdx.x;
  const uint32_t lane = tid & 31u, warp = tid >> 5;
  const int64_t e = (int64_t)blockIdx.x * kOccBlock + tid;
  StepCounts c;
  if (e < p.N) {  // whole warps (the launcher requires N % 32 == 0)
    const int pid = p.puzzle_id ? checked_puzzle_id(p.puzzle_id[e], p.num_puzzles, p.status) : p.puzzle_uniform;
    OccEnv env;
    env.init(GlobalTable{p.blob + p.puzzle_off[pid]}, s_posx + tid, s_posy + tid);
    env.load(p.x, p.y, p.stride, e);
    const int n = env.n(), g = env.g(), W = env.width(), H = env.height();
    const uint32_t nn1 = 1u + (uint32_t)(n * n);
    const uint64_t wd = env.fast(nn1);
    const uint32_t aw_hi = (uint32_t)(wd >> 32);
    const uint32_t aw_xmax = aw_hi & 63u, aw_ymax = (aw_hi >> 6) & 0xfffu, aw_w = aw_hi >> 18;
    const uint32_t aw_word0 = (uint32_t)wd << 1;
    const uint32_t d0 = env.desc(0);
    const int h0 = desc_h(d0), o0 = desc_off(d0);
    uint32_t arow[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) arow[r] = r < h0 ? (uint32_t)env.row(o0 + r) : 0u;
    uint32_t* const occ = s_occ + tid;
    const bool fast_ok = W <= 32 && H <= kOccRows && h0 <= 4 && build_occupancy(env, occ, W, H);
    if (fast_ok) {
      for (int j = 1; j < n; ++j) {
        const int h = desc_h(env.desc(j));
        env.sy[(j - 1) * kPosPitch] = (int)((0xffffffffu >> (32 - h)) << env.sy[(j - 1) * kPosPitch]);
      }
      env.masks = true;
    }
    int& ax = env.ax;
    int& ay = env.ay;
    int ach = achieved_goals(env);
    uint32_t code_idle = record_byte(0, ach == g, false);

    const int T = p.T;
    uint8_t* const act_ring = s_act + warp * (kActTiles * kTileBytes);
    uint8_t* const out_ring = s_out + warp * (2 * kTileBytes);
    const uint32_t act_ring_s = (uint32_t)__cvta_generic_to_shared(act_ring);
    const uint8_t* const warp_actions = p.actions + (e - (int64_t)lane);
    auto issue_tile = [&](const int tile) {  // lane l copies the warp's 32 actions of step 32*tile + l
      const int row = tile * 32 + (int)lane;
      const uint8_t* src = warp_actions + (size_t)(row < T ? row : T - 1) * (size_t)p.N;  // clamped: never used
      const uint32_t dst = act_ring_s + (uint32_t)(tile % kActTiles) * kTileBytes + lane * 32u;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u), "l"(src + 16) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto flush_tile = [&](const int tile) {  // results of tile `tile` leave the ring, coalesced
      const int rows = T - tile * 32 < 32 ? T - tile * 32 : 32;
      const uint8_t* src = out_ring + (tile & 1) * kTileBytes;
      const size_t first = (size_t)tile * 32u * (size_t)p.N + (size_t)(e - (int64_t)lane);
      if (kPacked) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {  // 32 rows x 32 bytes = 64 chunks of 16 bytes, two per lane
          const int chunk = (int)lane + 32 * k, row = chunk >> 1;
          if (row < rows) {
            const uint4 v = *reinterpret_cast<const uint4*>(src + chunk * 16);
            *reinterpret_cast<uint4*>(p.record + first + (size_t)row * (size_t)p.N + (chunk & 1) * 16) = v;
          }
        }
      } else {
        float* rew = p.reward + first + lane;
        uint8_t* term = p.terminated + first + lane;
        const uint8_t* mine = src + lane;
#pragma unroll 4
        for (int r = 0; r < rows; ++r) {
          const uint32_t code = mine[r * 32];
          rew[(size_t)r * (size_t)p.N] = s_lut[code];
          term[(size_t)r * (size_t)p.N] = (uint8_t)(code >> 7);
        }
      }
    };

    const int n_tiles = (T + 31) / 32;
    issue_tile(0);
    issue_tile(1);
    issue_tile(2);
    asm volatile("cp.async.wait_group 1;" ::: "memory");  // tiles 0 and 1 have landed
    __syncwarp();
    int base = 0;            // the window: tiles base, base + 1 (both landed)
    int t = 0;               // this lane's step index
    bool parked = false;     // this lane's step t needs the closure
    uint32_t parked_action = 0;
    uint32_t blocked_actions = 0;  // bit a: action a moves nothing in the current state
    uint32_t n_term = 0, n_moved = 0, n_gain = 0, n_invalid = 0;
    constexpr uint32_t kDxBytes = 0x000001ffu, kDyBytes = 0x01ff0000u;
    const int batch = p.occ_batch;
    while (base < n_tiles) {
      const int lim = min(T, 32 * (base + 2));
      if (!parked && t < lim) {
        // ---- the common step (branch-free) ---------------------------------------------------
        const uint32_t a = act_ring[(uint32_t)((t >> 5) % kActTiles) * kTileBytes + ((uint32_t)(t & 31) << 5) + lane];
        const bool valid = a <= 3u;
        const uint32_t sel = a * 0x1111u + 0x8880u;
        int dx, dy;
        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(dx) : "r"(kDxBytes), "r"(0u), "r"(sel));
        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(dy) : "r"(kDyBytes), "r"(0u), "r"(sel));
        const bool here = (uint32_t)ax <= aw_xmax && (uint32_t)ay <= aw_ymax;
        const uint32_t idx = here ? ((((uint32_t)ay * aw_w + (uint32_t)ax) << 2) | (a & 3u)) : 0u;
        const bool stopped = (here && tab_bit(env.fast32(aw_word0 + (idx >> 5)), idx)) || ((blocked_actions >> (a & 3u)) & 1u);
        const int x = ax + dx, y = ay + dy;
        const bool there = fast_ok && (uint32_t)x <= aw_xmax && (uint32_t)y <= aw_ymax;
        const int y0 = there ? y : 0;
        uint32_t hit = 0u;
#pragma unroll
        for (int r = 0; r < 4; ++r) hit |= (arow[r] << (x & 31)) & occ[min(y0 + r, kOccRows - 1) * kOccBlock];
        const bool go = valid && !stopped;
        const bool free_move = go && there && hit == 0u;
        n_invalid += valid ? 0u : 1u;
        n_moved += free_move ? 1u : 0u;
        ax = free_move ? x : ax;
        ay = free_move ? y : ay;
        blocked_actions = free_move ? 0u : blocked_actions;
        if (go && !free_move) {
          parked = true;
          parked_action = a;
        } else {
          out_ring[((uint32_t)(t & 63) << 5) + lane] = (uint8_t)code_idle;
          n_term += code_idle >> 7;
          ++t;
        }
      }
      const uint32_t m_parked = __ballot_sync(0xffffffffu, parked);
      const uint32_t m_parked_occ = __ballot_sync(0xffffffffu, parked && fast_ok);
      const uint32_t m_ready = __ballot_sync(0xffffffffu, !parked && t < lim);
      const uint32_t m_idle = __ballot_sync(0xffffffffu, !parked && t >= lim && t < T);
      if (m_parked != 0u && (__popc(m_parked) >= batch || m_ready == 0u || m_idle != 0u)) {
        if (parked) {
          uint32_t pushed;
          if (fast_ok) {
            pushed = closure_scan(env, occ, parked_action, H, nn1, m_parked_occ);
          } else {
            pushed = pushed_set_tab(env, (int)parked_action);
            if (pushed) env.translate(pushed, action_dx((int)parked_action), action_dy((int)parked_action));
          }
          int now = ach;
          if (pushed) {
            ++n_moved;
            blocked_actions = 0u;
            if (pushed & ~1u) now = achieved_goals(env);
          } else {
            blocked_actions |= 1u << parked_action;
          }
          const int dg = now - ach;
          ach = now;
          const bool term = ach == g;  // puzzle.py:409-411
          out_ring[((uint32_t)(t & 63) << 5) + lane] = (uint8_t)record_byte(dg, term, false);
          code_idle = record_byte(0, term, false);
          n_term += term ? 1u : 0u;
          n_gain += (uint32_t)max(dg, 0);
          ++t;
          parked = false;
        }
      }
      // every lane has left tile `base`: flush its results, refill its action slot, move the window
      if (__all_sync(0xffffffffu, t >= min(T, 32 * (base + 1)))) {
        __syncwarp();
        flush_tile(base);
        asm volatile("cp.async.wait_group 0;" ::: "memory");  // tile base + 2 has landed
        __syncwarp();
        issue_tile(base + 3);  // into the slot of tile `base`
        ++base;
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    c.terminated += n_term;
    c.moved += n_moved;
    c.dgoals += n_gain;
    c.invalid += n_invalid;
    env.store(p.x, p.y, p.stride, e);
  }
  flush_counts(c, p, s_counters);
}

// ------------------------------------------------------------------------------------------
// k_rollout_mask: lock-step, one thread per environment, with a BOUNDING-BOX MASK PREFILTER in front
// of the collision-table look-ups.  For puzzles of at most 32 x 32 cells every object j >= 1 is
// kept in shared memory as two words instead of (x, y): its ROW mask (bit r set for the grid rows
// y .. y + h - 1 it spans) and its COLUMN mask (x .. x + w - 1); x and y are their lowest set bits.
// A pusher displaced by the action can only push j if the bounding boxes overlap, i.e. if
// (rows_i' & rows_j) != 0 and (cols_i' & cols_j) != 0 with the pusher's masks shifted by the
// action (puzzle.py:567-593 needs the moved pusher to OVERLAP the pushee): one 64-bit shared-memory
// load and two ANDs per object instead of a descriptor load, a window test and a bit look-up.
// Only the few candidates that pass are decided by the exact look-ups (tab_pushes, tab_wall of
// pw_fast.cuh).  Every loop whose trip count depends on the environment is headed by a warp vote,
// so the 32 environments of a warp stay converged: all lanes do useful work in every step, which
// is what the run-ahead / warp-cooperative variants above cannot offer.
// Environments that do not qualify (bigger grid, a movable outside the grid or on a wall) keep
// plain positions and run pushed_set_tab, the round-1 rule, in the same loop.
// ------------------------------------------------------------------------------------------
// Row pitch of the per-object mask rows in shared memory, in uint2: one more than the CTA size, so that the
// lanes of a warp reading ONE environment's objects (the cooperative prefilter below) hit different banks.
constexpr int kMaskPitch = kOccBlock + 1;
constexpr int kCoopCost = 3;  // a cooperative prefilter pass costs about as much as three objects of the per-lane loop

struct MaskEnv : EnvBase<GlobalTable> {
  static constexpr int kStaticN = 0;
  static constexpr int kUnroll = 1;
  int ax, ay;
  uint2* sm;  // object j >= 1 at sm[(j - 1) * kMaskPitch]: masks (x = rows, y = columns) or plain (x, y)
  bool masks = false;
  __device__ __forceinline__ void init(GlobalTable table, uint2* col) {
    this->init_header(table);
    sm = col;
  }
  __device__ __forceinline__ void load(const int32_t* x, const int32_t* y, int64_t stride, int64_t e) {
    ax = x[e];
    ay = y[e];
    for (int j = 1; j < this->n_; ++j)
      sm[(j - 1) * kMaskPitch] = make_uint2((uint32_t)x[j * stride + e], (uint32_t)y[j * stride + e]);
  }
  __device__ __forceinline__ void store(int32_t* x, int32_t* y, int64_t stride, int64_t e) const {
    x[e] = ax;
    y[e] = ay;
    for (int j = 1; j < this->n_; ++j) {
      x[j * stride + e] = this->x(j);
      y[j * stride + e] = this->y(j);
    }
  }
  __device__ __forceinline__ void load_initial() {  // plain layout
    masks = false;
    const uint32_t v0 = this->table_desc(0, 1);
    ax = lo16(v0);
    ay = hi16(v0);
    for (int j = 1; j < this->n_; ++j) {
      const uint32_t v = this->table_desc(j, 1);
      sm[(j - 1) * kMaskPitch] = make_uint2((uint32_t)lo16(v), (uint32_t)hi16(v));
    }
  }
  __device__ __forceinline__ uint32_t desc(int j) const { return this->table_desc(j, 0); }
  __device__ __forceinline__ uint32_t desc_dyn(int j) const { return this->table_desc(j, 0); }
  // plain layout: (x, y); mask layout: (row mask, column mask)
  __device__ __forceinline__ int x(int j) const {
    if (j == 0) return ax;
    const uint2 v = sm[(j - 1) * kMaskPitch];
    return masks ? __ffs((int)v.y) - 1 : (int)v.x;
  }
  __device__ __forceinline__ int y(int j) const {
    if (j == 0) return ay;
    const uint2 v = sm[(j - 1) * kMaskPitch];
    return masks ? __ffs((int)v.x) - 1 : (int)v.y;
  }
  __device__ __forceinline__ int x_dyn(int j) const { return x(j); }
  __device__ __forceinline__ int y_dyn(int j) const { return y(j); }
  __device__ __forceinline__ int goal_x(int k) const { return lo16(this->table_desc(k + 1, 2)); }
  __device__ __forceinline__ int goal_y(int k) const { return hi16(this->table_desc(k + 1, 2)); }
  __device__ __forceinline__ void translate(uint32_t mask, int dx, int dy) {  // plain layout only
    if (mask & 1u) {
      ax += dx;
      ay += dy;
    }
    mask &= ~1u;
    while (mask) {
      const int j = __ffs((int)mask) - 1;
      mask &= mask - 1;
      uint2 v = sm[(j - 1) * kMaskPitch];
      v.x += (uint32_t)dx;
      v.y += (uint32_t)dy;
      sm[(j - 1) * kMaskPitch] = v;
    }
  }
};

__device__ __forceinline__ uint32_t shift_mask(uint32_t m, int d) { return d > 0 ? m << 1 : (d < 0 ? m >> 1 : m); }

// kEpisodes: step counter, truncation (gym_env.py:223) and -- after a terminated or truncated step -- a
// restart from the initial state, of a freshly drawn puzzle under pw_rollout_episodes' flag
// (gym_env.py:171-176 on the device); the per-puzzle constants below are then derived again.
//
// kCta (PW_MASK_CTA=1, a measured alternative -- slower): the closure is resolved by
// the CTA instead of the warp.  In lock-step, ~99 % of the warp-steps
// run the exact look-ups for the ~4 lanes that have a prefilter candidate and 62 % run a second closure
// round for 1-2 lanes.  Here a step has three phases separated by two barriers: (A) every thread does
// the common part and the agent's prefilter; the 13 % of the environments with a candidate post
// themselves (candidates, action, agent position) in a shared-memory list; (B) the first `count`
// threads of the CTA take one posted environment each -- its masks are in shared memory, its
// tables behind a pointer posted at set-up -- and resolve its closure completely (exact look-ups,
// further rounds, displacement of the masks, goal delta), all lanes of the resolving warp busy;
// (C) the owners read the verdict back and write the step's outputs.
struct FastView {  // the collision tables of a posted environment (what tab_pushes / tab_wall read)
  const uint64_t* ft;
  __device__ __forceinline__ uint64_t fast(uint32_t k) const { return __ldg(ft + k); }
  __device__ __forceinline__ uint32_t fast32(uint32_t k) const { return __ldg(reinterpret_cast<const uint32_t*>(ft) + k); }
};

// (capped at 72 registers -- 7 CTAs per SM; uncapped, the episode variants take 80 and the others up to 91 -- which
// the compiler meets without spills: 3.4e10 -> 4.4e10 env-steps/s on the real level-0 set; a cap of 64 registers
// spills in the plain variants and loses a quarter)
template <bool kPacked, bool kEpisodes, bool kCta>
__global__ void __launch_bounds__(kOccBlock, 7)
    k_rollout_mask(const __grid_constant__ StepParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ uint32_t s_counters[4];
  __shared__ uint32_t s_posted[2];  // kCta: environments posted in the current / next step
  // dynamic shared memory: object masks / positions | action tiles | reward table | kCta: mail boxes, lists
  uint2* s_obj = reinterpret_cast<uint2*>(smem_raw);
  const int pos_rows = p.n_max > 1 ? p.n_max - 1 : 1;
  uint8_t* s_act = smem_raw + (((size_t)pos_rows * kMaskPitch * sizeof(uint2) + 15) & ~(size_t)15);
  float* s_lut = reinterpret_cast<float*>(s_act + (kOccBlock / 32) * 2 * kTileBytes);
  uint4* s_mail = reinterpret_cast<uint4*>(s_lut + 256);  // in: candidates, action, ax, ay; out: pushed set, goal delta
  uint4* s_cst = s_mail + kOccBlock;                      // per environment: n, g, descriptor word offset, -
  const uint64_t** s_ft = reinterpret_cast<const uint64_t**>(s_cst + kOccBlock);  // its collision tables
  const uint64_t** s_tb = s_ft + kOccBlock;                                       // its blob (goal positions)
  uint16_t* s_list = reinterpret_cast<uint16_t*>(s_tb + kOccBlock);               // [2][kOccBlock] posted environments
  if (threadIdx.x < 4) s_counters[threadIdx.x] = 0;
  if (threadIdx.x < 2) s_posted[threadIdx.x] = 0;
  if (!kPacked) {
    // record byte -> reward (gym_env.py:212-221): 10 if terminated else (double)dg - 0.01, rounded
    // to float32 once; IEEE arithmetic, so identical to the host-side table of the other kernels
    for (int c = threadIdx.x; c < 256; c += kOccBlock)
      s_lut[c] = (c & 0x80) ? 10.0f : (float)((double)((c & 0x3f) - kMaxObjects) - 0.01);
  }
  __syncthreads();

  const int tid = threadIdx.x;
  const uint32_t lane = tid & 31u, warp = tid >> 5;
  const int64_t e = (int64_t)blockIdx.x * kOccBlock + tid;
  StepCounts c;
  if (e < p.N) {  // whole warps (the launcher requires N % 32 == 0)
    int pid = p.puzzle_id ? checked_puzzle_id(p.puzzle_id[e], p.num_puzzles, p.status) : p.puzzle_uniform;
    MaskEnv env;
    env.init(GlobalTable{p.blob + p.puzzle_off[pid]}, s_obj + tid);
    env.load(p.x, p.y, p.stride, e);
    // per-puzzle constants (derived again when an episode restarts on another puzzle)
    int n, g, w0, h0, ach;
    uint32_t nn1, aw_xmax, aw_ymax, aw_w, aw_word0, agent_rows, agent_cols, code_idle;
    bool mask_ok;
    auto setup_puzzle = [&]() {  // what depends on the puzzle only
      n = env.n();
      g = env.g();
      nn1 = 1u + (uint32_t)(n * n);
      // the agent's wall table (pw_fast.cuh: tab_wall), descriptor cached in registers
      const uint64_t wd = env.fast(nn1);
      const uint32_t aw_hi = (uint32_t)(wd >> 32);
      aw_xmax = aw_hi & 63u;
      aw_ymax = (aw_hi >> 6) & 0xfffu;
      aw_w = aw_hi >> 18;
      aw_word0 = (uint32_t)wd << 1;
      const uint32_t d0 = env.desc(0);
      w0 = desc_w(d0);
      h0 = desc_h(d0);
      agent_rows = 0xffffffffu >> (32 - min(h0, 32));
      agent_cols = 0xffffffffu >> (32 - min(w0, 32));
      if (kCta) {
        s_cst[tid] = make_uint4((uint32_t)n, (uint32_t)g, (uint32_t)env.desc_word_, 0u);
        s_ft[tid] = env.ft.base;
        s_tb[tid] = env.t.base;
      }
    };
    auto setup_state = [&]() {  // expects plain positions in env.sm
      const int W = env.width(), H = env.height();
      env.masks = false;
      // mask layout: grid of at most 32 x 32 cells, every movable inside it and off the walls (such a
      // movable is stopped by the border walls, puzzle.py:160-168, 284-293, so it stays inside)
      mask_ok = W <= 32 && H <= 32;
      if (mask_ok) {
        const int wall_off = env.wall_off();
        for (int j = 1; j < n; ++j) {
          const uint32_t d = env.desc(j);
          const int x = env.x(j), y = env.y(j), w = desc_w(d), h = desc_h(d), o = desc_off(d);
          if (x < 0 || y < 0 || x + w > W || y + h > H) {
            mask_ok = false;
            continue;
          }
          for (int r = 0; r < h; ++r) mask_ok = mask_ok && ((env.row(o + r) << x) & env.row(wall_off + y + r)) == 0ull;
        }
      }
      ach = achieved_goals(env);
      if (mask_ok) {
        for (int j = 1; j < n; ++j) {
          const uint32_t d = env.desc(j);
          const uint2 v = env.sm[(j - 1) * kMaskPitch];
          env.sm[(j - 1) * kMaskPitch] = make_uint2((0xffffffffu >> (32 - desc_h(d))) << v.y,
                                                   (0xffffffffu >> (32 - desc_w(d))) << v.x);
        }
        env.masks = true;
      }
      code_idle = record_byte(0, ach == g, false);  // the record of a step that moves no movable
    };
    // a new episode: the initial state, from the section of the puzzle's table that holds it in mask
    // form (pw_fast.cuh: build_fast_table) when it qualifies -- a restart is then n loads
    auto restart_episode = [&]() {
      const uint32_t section = (uint32_t)(env.fast(0u) >> 32);
      const uint32_t head = (uint32_t)env.fast(section);
      if (head & 1u) {
        const uint32_t v0 = env.table_desc(0, 1);
        env.ax = lo16(v0);
        env.ay = hi16(v0);
        for (int j = 1; j < n; ++j) {
          const uint64_t m = env.fast(section + (uint32_t)j);
          env.sm[(j - 1) * kMaskPitch] = make_uint2((uint32_t)m, (uint32_t)(m >> 32));
        }
        env.masks = true;
        mask_ok = true;
        ach = (int)((head >> 8) & 0xffu);
        code_idle = record_byte(0, ach == g, false);
      } else {
        env.load_initial();
        setup_state();
      }
    };
    setup_puzzle();
    setup_state();
    // rows past this environment's last object hold empty masks: the prefilter loop runs to the warp's
    // largest object count without a per-object bound test
    auto clear_unused_rows = [&]() {
      for (int j = n; j <= pos_rows; ++j) env.sm[(j - 1) * kMaskPitch] = make_uint2(0u, 0u);
    };
    clear_unused_rows();
    int& ax = env.ax;
    int& ay = env.ay;
    int ep_steps = kEpisodes ? p.steps[e] : 0;

    const int T = p.T;
    uint8_t* const tiles = s_act + warp * (2 * kTileBytes);
    const uint32_t tiles_s = (uint32_t)__cvta_generic_to_shared(tiles);
    const uint8_t* const warp_actions = p.actions + (e - (int64_t)lane);
    auto issue_block = [&](const int block) {  // lane l copies the warp's 32 actions of step 32*block + l
      const int row = block * 32 + (int)lane;
      const uint8_t* src = warp_actions + (size_t)(row < T ? row : T - 1) * (size_t)p.N;  // clamped: never used
      const uint32_t dst = tiles_s + (uint32_t)(block & 1) * kTileBytes + lane * 32u;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u), "l"(src + 16) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // outputs of step t at element offset t * N + e: ONE running 64-bit offset instead of three running pointers
    int64_t out_off = e;
    const bool want_trunc = kEpisodes && !kPacked && p.truncated != nullptr;
    uint32_t n_term = 0, n_moved = 0, n_gain = 0, n_invalid = 0;
    constexpr uint32_t kDxBytes = 0x000001ffu, kDyBytes = 0x01ff0000u;  // LEFT, RIGHT, UP, DOWN (puzzle.py:43-50)
    uint32_t all_mask_ok = __ballot_sync(0xffffffffu, mask_ok);
    int n_warp = __reduce_max_sync(0xffffffffu, n);  // the warp's largest object count
    uint32_t parity = 0;  // kCta: which of the two posted lists this step fills
    const int blocks = (T + 31) / 32;
    issue_block(0);
    for (int blk = 0; blk < blocks; ++blk) {
      __syncwarp();          // every lane is done with the tile the next block lands in
      issue_block(blk + 1);  // past the end: re-reads the last row, never used
      asm volatile("cp.async.wait_group 1;" ::: "memory");
      __syncwarp();  // block `blk` is visible to the whole warp
      const uint8_t* mine = tiles + (blk & 1) * kTileBytes + lane;
      const int steps = T - blk * 32 < 32 ? T - blk * 32 : 32;
#pragma unroll 1
      for (int k = 0; k < steps; ++k, mine += 32) {
        const uint32_t a = mine[0];
        const bool valid = a <= 3u;
        const uint32_t sel = a * 0x1111u + 0x8880u;  // byte a, sign extended
        // (raw prmt: __byte_perm masks the selector and would drop the sign-replication bits)
        int dx, dy;
        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(dx) : "r"(kDxBytes), "r"(0u), "r"(sel));
        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(dy) : "r"(kDyBytes), "r"(0u), "r"(sel));
        // walls (puzzle.py:353-354): one bit of the agent's collision table
        const bool here = (uint32_t)ax <= aw_xmax && (uint32_t)ay <= aw_ymax;
        const uint32_t idx = here ? ((((uint32_t)ay * aw_w + (uint32_t)ax) << 2) | (a & 3u)) : 0u;
        const bool go = valid && !(here && tab_bit(env.fast32(aw_word0 + (idx >> 5)), idx));
        n_invalid += valid ? 0u : 1u;
        uint32_t pushed = 0u;
        int now = ach;  // goals achieved after this step
        bool posted = false;
        if (kCta) {
          // ---- phase A: the agent's prefilter; environments with a candidate are posted ------------
          if (mask_ok && go) {
            const int x = ax + dx, y = ay + dy;
            const bool framed = (uint32_t)x <= 32u - (uint32_t)min(w0, 32) && (uint32_t)y <= 32u - (uint32_t)min(h0, 32);
            const uint32_t prows = framed ? agent_rows << y : 0xffffffffu;
            const uint32_t pcols = framed ? agent_cols << x : 0xffffffffu;
            uint32_t cand = 0u;
#pragma unroll 4
            for (int j = 1; j < n; ++j) {
              const uint2 v = env.sm[(j - 1) * kMaskPitch];
              const bool near = (v.x & prows) != 0u && (v.y & pcols) != 0u;
              cand = (cand >> 1) | (near ? 0x80000000u : 0u);
            }
            cand = n > 1 ? cand >> (32 - n) : 0u;
            if (cand == 0u) {  // nothing in the way: the agent moves alone
              pushed = 1u;
              ax += dx;
              ay += dy;
            } else {
              posted = true;
              s_mail[tid] = make_uint4(cand, a, (uint32_t)ax, (uint32_t)ay);
            }
          }
          const uint32_t votes = __ballot_sync(0xffffffffu, posted);
          if (votes != 0u) {
            uint32_t first = 0;
            if (lane == 0) first = atomicAdd(&s_posted[parity], (uint32_t)__popc(votes));
            first = __shfl_sync(0xffffffffu, first, 0);
            if (posted) s_list[parity * kOccBlock + first + __popc(votes & ((1u << lane) - 1u))] = (uint16_t)tid;
          }
        } else
        if (all_mask_ok != 0u) {
          // ---- mask layout: prefilter + exact look-ups, lanes converged ------------------------
          const bool mine_masks = mask_ok && go;
          uint32_t frontier = mine_masks ? 1u : 0u;
          pushed = frontier;
          bool stop = false;
          while (__any_sync(0xffffffffu, frontier != 0u)) {
            const bool active = frontier != 0u;
            const int i = active ? __ffs((int)frontier) - 1 : 0;
            frontier &= frontier - 1;
            int xi = ax, yi = ay;
            uint32_t prows, pcols;  // the displaced pusher's masks
            if (i == 0) {
              // an agent (partly) outside the 32 x 32 frame cannot be put in masks: test everything
              const int x = ax + dx, y = ay + dy;
              const bool framed = (uint32_t)x <= 32u - (uint32_t)min(w0, 32) && (uint32_t)y <= 32u - (uint32_t)min(h0, 32);
              prows = framed ? agent_rows << y : 0xffffffffu;
              pcols = framed ? agent_cols << x : 0xffffffffu;
            } else {
              const uint2 v = env.sm[(i - 1) * kMaskPitch];
              xi = __ffs((int)v.y) - 1;
              yi = __ffs((int)v.x) - 1;
              prows = shift_mask(v.x, dy);
              pcols = shift_mask(v.y, dx);
            }
            // the trip count is the warp's largest object count (lanes with fewer objects, and lanes without
            // a pusher, test nothing): a uniform loop instead of one whose bound differs per lane
            uint32_t cand = 0u;
            uint32_t pushers = __ballot_sync(0xffffffffu, active);
            if (__popc(pushers) * kCoopCost >= n_warp) {
              const uint32_t my_rows = active ? prows : 0u;  // (rows past the lane's own objects are empty masks)
#pragma unroll 4
              for (int j = 1; j < n_warp; ++j) {
                const uint2 v = env.sm[(j - 1) * kMaskPitch];
                const bool near = (v.x & my_rows) != 0u && (v.y & pcols) != 0u;
                cand = (cand >> 1) | (near ? 0x80000000u : 0u);
              }
              cand = n_warp > 1 ? cand >> (32 - n_warp) : 0u;
              cand &= ~pushed;
            } else {
              // Later closure rounds have a pusher in one or two lanes only: instead of every lane walking the
              // loop above for them, the WARP tests one such environment's objects in parallel -- lane L takes
              // object L of environment q (its masks are in shared memory; the row pitch keeps the 32 reads on
              // different banks) -- and the ballot is q's candidate set.
              __syncwarp();  // the owners' mask writes (displacement, restarts) before other lanes read them
              while (pushers != 0u) {
                const int q = __ffs((int)pushers) - 1;
                pushers &= pushers - 1;
                const uint32_t q_rows = __shfl_sync(0xffffffffu, prows, q), q_cols = __shfl_sync(0xffffffffu, pcols, q);
                const int q_n = __shfl_sync(0xffffffffu, n, q);
                bool near = false;
                if (lane >= 1u && (int)lane < q_n) {
                  const uint2 v = env.sm[((int)lane - 1) * kMaskPitch + (q - (int)lane)];
                  near = (v.x & q_rows) != 0u && (v.y & q_cols) != 0u;
                }
                const uint32_t found = __ballot_sync(0xffffffffu, near);
                if ((int)lane == q) cand = found & ~pushed;
              }
              __syncwarp();  // ... and those reads before the owners write again
            }
            while (__any_sync(0xffffffffu, cand != 0u)) {
              if (cand != 0u) {
                const int j = __ffs((int)cand) - 1;
                cand &= cand - 1;
                const uint2 v = env.sm[(j - 1) * kMaskPitch];
                const int xj = __ffs((int)v.y) - 1, yj = __ffs((int)v.x) - 1;
                if (tab_pushes(env, 1u + (uint32_t)(i * n + j), a, xj - xi, yj - yi)) {
                  if (tab_wall(env, nn1, (uint32_t)j, a, xj, yj)) {  // transitive stopping, puzzle.py:376-379
                    stop = true;
                    cand = 0u;
                    frontier = 0u;
                  } else {
                    pushed |= 1u << j;
                    frontier |= 1u << j;
                  }
                }
              }
            }
          }
          pushed = stop ? 0u : pushed;
          if (pushed != 0u) {
            ax += dx;
            ay += dy;
          }
          // displace the pushed objects; goal objects (1 .. g) among them change the achieved count
          // (puzzle.py:396-407), which is kept incrementally: only pushed objects can change it
          uint32_t rest = pushed & ~1u;
          while (__any_sync(0xffffffffu, rest != 0u)) {
            if (rest != 0u) {
              const int j = __ffs((int)rest) - 1;
              rest &= rest - 1;
              const uint2 v = env.sm[(j - 1) * kMaskPitch];
              const uint2 w = make_uint2(shift_mask(v.x, dy), shift_mask(v.y, dx));
              env.sm[(j - 1) * kMaskPitch] = w;
              if (j <= g) {
                const uint32_t gw = env.table_desc(j, 2);
                const int gx = lo16(gw), gy = hi16(gw);
                const bool was = __ffs((int)v.y) - 1 == gx && __ffs((int)v.x) - 1 == gy;
                const bool is = __ffs((int)w.y) - 1 == gx && __ffs((int)w.x) - 1 == gy;
                now += (is ? 1 : 0) - (was ? 1 : 0);
              }
            }
          }
        }
        if (all_mask_ok != 0xffffffffu) {
          // ---- plain layout: the table rule with its scan over all objects (pw_fast.cuh) --------
          if (!mask_ok && go) {
            pushed = pushed_set_tab(env, (int)a);
            if (pushed) env.translate(pushed, dx, dy);
            if (pushed & ~1u) now = achieved_goals(env);
          }
          __syncwarp();
        }
        if (kCta) {
          __syncthreads();
          // ---- phase B: one posted environment per thread, resolved completely --------------------
          if ((uint32_t)tid < s_posted[parity]) {
            const int q = s_list[parity * kOccBlock + tid];
            const uint4 mb = s_mail[q];
            const uint4 cst = s_cst[q];
            const FastView fv{s_ft[q]};
            uint2* const sm = s_obj + q;
            const int qn = (int)cst.x, qg = (int)cst.y;
            const uint32_t qa = mb.y, qnn1 = 1u + (uint32_t)(qn * qn);
            const int qdx = action_dx((int)qa), qdy = action_dy((int)qa);
            uint32_t cand = mb.x, set = 1u, frontier = 0u;
            int i = 0, xi = (int)mb.z, yi = (int)mb.w;
            bool stop = false;
            while (true) {
              while (cand != 0u) {  // exact look-ups of pusher i's candidates (puzzle.py:567-593, 376-379)
                const int j = __ffs((int)cand) - 1;
                cand &= cand - 1;
                const uint2 v = sm[(j - 1) * kMaskPitch];
                const int xj = __ffs((int)v.y) - 1, yj = __ffs((int)v.x) - 1;
                if (tab_pushes(fv, 1u + (uint32_t)(i * qn + j), qa, xj - xi, yj - yi)) {
                  if (tab_wall(fv, qnn1, (uint32_t)j, qa, xj, yj)) {
                    stop = true;
                    break;
                  }
                  set |= 1u << j;
                  frontier |= 1u << j;
                }
              }
              if (stop || frontier == 0u) break;
              i = __ffs((int)frontier) - 1;
              frontier &= frontier - 1;
              const uint2 vi = sm[(i - 1) * kMaskPitch];
              xi = __ffs((int)vi.y) - 1;
              yi = __ffs((int)vi.x) - 1;
              const uint32_t prows = shift_mask(vi.x, qdy), pcols = shift_mask(vi.y, qdx);
              for (int j = 1; j < qn; ++j) {
                const uint2 v = sm[(j - 1) * kMaskPitch];
                const bool near = (v.x & prows) != 0u && (v.y & pcols) != 0u;
                cand = (cand >> 1) | (near ? 0x80000000u : 0u);
              }
              cand = (cand >> (32 - qn)) & ~set;
            }
            if (stop) set = 0u;
            int delta = 0;
            uint32_t rest = set & ~1u;
            while (rest != 0u) {  // displace the pushed objects; goal objects among them change the achieved count
              const int j = __ffs((int)rest) - 1;
              rest &= rest - 1;
              const uint2 v = sm[(j - 1) * kMaskPitch];
              const uint2 w = make_uint2(shift_mask(v.x, qdy), shift_mask(v.y, qdx));
              sm[(j - 1) * kMaskPitch] = w;
              if (j <= qg) {
                const uint32_t gw = __ldg(reinterpret_cast<const uint32_t*>(s_tb[q]) + cst.z + 4u * (uint32_t)j + 2u);
                const int gx = lo16(gw), gy = hi16(gw);
                const bool was = __ffs((int)v.y) - 1 == gx && __ffs((int)v.x) - 1 == gy;
                const bool is = __ffs((int)w.y) - 1 == gx && __ffs((int)w.x) - 1 == gy;
                delta += (is ? 1 : 0) - (was ? 1 : 0);
              }
            }
            s_mail[q] = make_uint4(set, (uint32_t)delta, 0u, 0u);
          }
          __syncthreads();
          // ---- phase C: the owners take the verdicts ------------------------------------------------
          if (tid == 0) s_posted[parity] = 0u;  // consumed; counted again two steps from now, after the next barrier
          parity ^= 1u;
          if (posted) {
            const uint4 res = s_mail[tid];
            pushed = res.x;
            if (pushed != 0u) {
              ax += dx;
              ay += dy;
            }
            now += (int)res.y;
          }
        }
        uint32_t code = code_idle;
        n_moved += pushed ? 1u : 0u;
        if (now != ach) {  // (a goal object moved onto / off its goal: rare)
          const int dg = now - ach;
          ach = now;
          const bool term = ach == g;  // puzzle.py:409-411
          code = record_byte(dg, term, false);
          code_idle = record_byte(0, term, false);
          n_gain += (uint32_t)max(dg, 0);
        }
        n_term += code >> 7;
        bool restart = false;
        if (kEpisodes) {
          ++ep_steps;
          const bool trunc = p.max_steps > 0 && ep_steps >= p.max_steps;  // gym_env.py:223
          restart = (code >> 7) != 0u || trunc;
          if (kPacked) {
            code |= trunc ? 0x40u : 0u;
          } else if (want_trunc) {
            p.truncated[out_off] = trunc ? 1 : 0;
          }
        }
        if (kPacked) {
          p.record[out_off] = (uint8_t)code;
        } else {
          p.reward[out_off] = s_lut[code & 0xbfu];
          p.terminated[out_off] = (uint8_t)(code >> 7);
        }
        out_off += p.N;
        if (kEpisodes && __any_sync(0xffffffffu, restart)) {
          if (restart) {  // gym_env.py:171-176: a new episode, on a freshly drawn puzzle if asked for
            if (p.flags & kFlagResample) {
              const uint32_t episode = p.episode[e] + 1u;
              p.episode[e] = episode;
              pid = draw_puzzle(p.seed, e, episode, p.num_puzzles);
              p.puzzle_id_mut[e] = pid;
              env.retarget(p.blob + p.puzzle_off[pid]);
              setup_puzzle();
              clear_unused_rows();
            }
            restart_episode();
            ep_steps = 0;
          }
          all_mask_ok = __ballot_sync(0xffffffffu, mask_ok);
          n_warp = __reduce_max_sync(0xffffffffu, n);
        }
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    c.terminated += n_term;
    c.moved += n_moved;
    c.dgoals += n_gain;
    c.invalid += n_invalid;
    env.store(p.x, p.y, p.stride, e);
    if (kEpisodes) p.steps[e] = ep_steps;
  }
  flush_counts(c, p, s_counters);
}

size_t mask_smem_bytes(int n_max, bool cta) {
  const int pos_rows = n_max > 1 ? n_max - 1 : 1;
  return (((size_t)pos_rows * kMaskPitch * sizeof(uint2) + 15) & ~(size_t)15) + (size_t)(kOccBlock / 32) * 2 * kTileBytes +
         256 * sizeof(float) +
         (cta ? (size_t)kOccBlock * (2 * sizeof(uint4) + 2 * sizeof(uint64_t*) + 2 * sizeof(uint16_t)) : 0);
}

// ------------------------------------------------------------------------------------------
// The layout BASELINE.json's north star sketches, built to be MEASURED against the kernel above
// (reachable with PW_WARP_PER_ENV=1): one WARP per environment, lane j <-> object j,
// positions in registers, the push closure (puzzle.py:359-382) as one pair-table look-up per lane
// and round with __ballot_sync / __shfl_sync.  Exact like every other path (same tables).
// ------------------------------------------------------------------------------------------
constexpr int kWarpEnvBlock = 128;  // 4 environments per CTA

__global__ void __launch_bounds__(kWarpEnvBlock) k_rollout_warp_per_env(const __grid_constant__ StepParams p) {
  __shared__ uint32_t s_counters[4];
  if (threadIdx.x < 4) s_counters[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u;
  const int64_t e = (int64_t)blockIdx.x * (kWarpEnvBlock / 32) + (threadIdx.x >> 5);
  StepCounts c;
  if (e < p.N) {
    const int pid = p.puzzle_id ? checked_puzzle_id(p.puzzle_id[e], p.num_puzzles, p.status) : p.puzzle_uniform;
    EnvBase<GlobalTable> env;
    env.init_header(GlobalTable{p.blob + p.puzzle_off[pid]});
    const int n = env.n(), g = env.g();
    const uint32_t nn1 = 1u + (uint32_t)(n * n);
    const bool live = (int)lane < n;
    int x = live ? p.x[lane * p.stride + e] : 0, y = live ? p.y[lane * p.stride + e] : 0;
    const bool goal_lane = lane >= 1u && (int)lane <= g;
    const uint32_t goal_word = goal_lane ? env.table_desc((int)lane, 2) : kNoGoal;
    const int gx = lo16(goal_word), gy = hi16(goal_word);
    auto achieved = [&]() { return __popc(__ballot_sync(0xffffffffu, goal_lane && x == gx && y == gy)); };
    int ach = achieved();
    const int T = p.T;
    uint32_t n_term = 0, n_moved = 0, n_gain = 0, n_invalid = 0;
    for (int t0 = 0; t0 < T; t0 += 32) {
      // lane l fetches the action of step t0 + l and keeps that step's result for one batched write
      const int tl = t0 + (int)lane;
      const uint32_t my_action = tl < T ? p.actions[(int64_t)tl * p.N + e] : 0u;
      uint32_t my_code = 0;
      const int steps = T - t0 < 32 ? T - t0 : 32;
      for (int k = 0; k < steps; ++k) {
        const uint32_t a = __shfl_sync(0xffffffffu, my_action, k);
        uint32_t pushed = 0;
        if (a <= 3u) {
          const int ax = __shfl_sync(0xffffffffu, x, 0), ay = __shfl_sync(0xffffffffu, y, 0);
          if (!tab_wall(env, nn1, 0u, a, ax, ay)) {  // puzzle.py:353-354
            pushed = 1u;
            uint32_t frontier = 1u;
            while (frontier) {
              const int i = __ffs((int)frontier) - 1;
              frontier &= frontier - 1;
              const int xi = __shfl_sync(0xffffffffu, x, i), yi = __shfl_sync(0xffffffffu, y, i);
              const bool candidate = live && lane >= 1u && !((pushed >> lane) & 1u);
              const bool push = candidate && tab_pushes(env, 1u + (uint32_t)(i * n) + lane, a, x - xi, y - yi);
              const uint32_t hits = __ballot_sync(0xffffffffu, push);
              const bool stop = push && tab_wall(env, nn1, lane, a, x, y);  // transitive stopping, :376-379
              if (__ballot_sync(0xffffffffu, stop)) {
                pushed = 0u;
                break;
              }
              pushed |= hits;
              frontier |= hits;
            }
          }
        } else {
          ++n_invalid;
        }
        if ((pushed >> lane) & 1u) {
          x += action_dx((int)a);
          y += action_dy((int)a);
        }
        const int now = (pushed & ~1u) ? achieved() : ach;
        const int dg = now - ach;
        ach = now;
        const bool term = ach == g;
        if ((int)lane == k) my_code = record_byte(dg, term, false);
        n_term += term ? 1u : 0u;
        n_moved += pushed ? 1u : 0u;
        n_gain += (uint32_t)max(dg, 0);
      }
      if (tl < T) {
        const int64_t idx = (int64_t)tl * p.N + e;
        if (p.record) {
          p.record[idx] = (uint8_t)my_code;
        } else {
          // gym_env.py:212-221, rounded to float32 once (IEEE: same value as the host-side table)
          p.reward[idx] = (my_code & 0x80u) ? 10.0f : (float)((double)((int)(my_code & 0x3fu) - kMaxObjects) - 0.01);
          p.terminated[idx] = (uint8_t)(my_code >> 7);
        }
      }
    }
    if (live) {
      p.x[lane * p.stride + e] = x;
      p.y[lane * p.stride + e] = y;
    }
    if (lane == 0) {  // every lane counted the same steps: report them once
      c.terminated = n_term;
      c.moved = n_moved;
      c.dgoals = n_gain;
      c.invalid = n_invalid;
    }
  }
  // flush_counts assumes one environment per thread for the env-step total: report it here
  if (p.counters && threadIdx.x == 0) {
    const int64_t first = (int64_t)blockIdx.x * (kWarpEnvBlock / 32);
    const int64_t rest = p.N - first;
    const int64_t live_envs = rest < (kWarpEnvBlock / 32) ? rest : (kWarpEnvBlock / 32);
    atomicAdd(&p.counters[0], (unsigned long long)(live_envs > 0 ? live_envs : 0) * (unsigned long long)p.T);
  }
  StepParams q = p;
  q.N = 0;  // (suppresses flush_counts' own env-step total)
  q.counters = p.counters;
  flush_counts(c, q, s_counters);
}

size_t occ_smem_bytes(int n_max, bool run_ahead) {
  const int pos_rows = n_max > 1 ? n_max - 1 : 1;
  return (size_t)kOccRows * kOccBlock * 4 + (size_t)2 * pos_rows * kPosPitch * 4 + 16 +
         (size_t)(kOccBlock / 32) * (run_ahead ? kActTiles + 2 : 2) * kTileBytes + 256 * sizeof(float);
}

}  // namespace

// Serves every plain / packed rollout of the general rule made of whole warps ...
bool occ_rollout_eligible(const StepParams& p) {
  const auto aligned16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  const bool plain = p.reward && p.terminated && !p.record;
  const bool packed = p.record && !p.reward && !p.terminated;
  return is_lean(p) && (plain || packed) && (p.N & 31) == 0 && aligned16(p.actions) &&
         (!packed || aligned16(p.record));
}

// ... and, on the mask kernel, the same output shapes with the episode handling of a vector
// environment (autoreset, with or without puzzle re-sampling; truncated flags optional).
bool occ_episodes_eligible(const StepParams& p) {
  const auto aligned16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  const bool plain = p.reward && p.terminated && !p.record;
  const bool packed = p.record && !p.reward && !p.terminated && !p.truncated;
  const bool resample = (p.flags & kFlagResample) != 0u;
  const char* mode = getenv("PW_OCC_MODE");
  return (p.flags & kFlagAutoreset) != 0u && (plain || packed) && !p.dgoals && !p.moved && !p.return_sum &&
         p.steps != nullptr && (!resample || (p.puzzle_id_mut && p.episode)) && (p.N & 31) == 0 &&
         aligned16(p.actions) && (mode == nullptr || mode[0] == 'm');
}

cudaError_t launch_rollout_occ(const StepParams& p_in, cudaStream_t stream) {
  StepParams p = p_in;
  // PW_OCC_WARP_CLOSURE=1: serve push closures one warp per environment instead of per thread (A/B)
  p.occ_warp_closure = getenv("PW_OCC_WARP_CLOSURE") != nullptr ? 1 : 0;
  if (getenv("PW_WARP_PER_ENV") != nullptr) {  // the measured alternative layout
    const unsigned warps = kWarpEnvBlock / 32;
    k_rollout_warp_per_env<<<(unsigned)((p.N + warps - 1) / warps), kWarpEnvBlock, 0, stream>>>(p);
    return cudaGetLastError();
  }
  const unsigned grid = (unsigned)((p.N + kOccBlock - 1) / kOccBlock);
  // PW_OCC_MODE=mask | lockstep | runahead: which general-rule kernel runs (A/B measurements)
  const char* mode = getenv("PW_OCC_MODE");
  if (mode == nullptr || mode[0] == 'm') {
    // PW_MASK_CTA=1 (whole CTAs only): the closure resolved by the CTA, two barriers per step -- measured,
    // 0.65x of the per-warp lock-step loop, so it is not the default
    const bool cta = p.N % kOccBlock == 0 && getenv("PW_MASK_CTA") != nullptr;
    const size_t smem = mask_smem_bytes(p.n_max, cta);
    const bool episodes = (p.flags & kFlagAutoreset) != 0u;
#define PW_MASK_GO(PACKED, EPISODES)                                                      \
  do {                                                                                    \
    if (cta) k_rollout_mask<PACKED, EPISODES, true><<<grid, kOccBlock, smem, stream>>>(p); \
    else k_rollout_mask<PACKED, EPISODES, false><<<grid, kOccBlock, smem, stream>>>(p);    \
  } while (0)
    if (p.record) {
      if (episodes) PW_MASK_GO(true, true);
      else PW_MASK_GO(true, false);
    } else {
      if (episodes) PW_MASK_GO(false, true);
      else PW_MASK_GO(false, false);
    }
#undef PW_MASK_GO
    return cudaGetLastError();
  }
  if (mode[0] == 'l') {
    const size_t smem = occ_smem_bytes(p.n_max, false);
    if (p.record) k_rollout_occ<true><<<grid, kOccBlock, smem, stream>>>(p);
    else k_rollout_occ<false><<<grid, kOccBlock, smem, stream>>>(p);
    return cudaGetLastError();
  }
  p.occ_batch = 8;  // parked lanes that trigger a closure pass (PW_OCC_BATCH overrides: tuning runs)
  if (const char* env = getenv("PW_OCC_BATCH")) {
    const int v = atoi(env);
    if (v >= 1 && v <= 32) p.occ_batch = v;
  }
  const size_t smem = occ_smem_bytes(p.n_max, true);
  if (p.record) k_rollout_occ_ra<true><<<grid, kOccBlock, smem, stream>>>(p);
  else k_rollout_occ_ra<false><<<grid, kOccBlock, smem, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t configure_occ_kernels() {
  const int cap = (int)occ_smem_bytes(kMaxObjects, true);
  cudaError_t err = cudaSuccess;
#define PW_OPT_IN(k) \
  if (err == cudaSuccess) err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, cap)
  PW_OPT_IN(k_rollout_occ<true>);
  PW_OPT_IN(k_rollout_occ<false>);
  PW_OPT_IN(k_rollout_occ_ra<true>);
  PW_OPT_IN(k_rollout_occ_ra<false>);
  PW_OPT_IN((k_rollout_mask<true, true, false>));
  PW_OPT_IN((k_rollout_mask<true, false, false>));
  PW_OPT_IN((k_rollout_mask<false, true, false>));
  PW_OPT_IN((k_rollout_mask<false, false, false>));
  PW_OPT_IN((k_rollout_mask<true, true, true>));
  PW_OPT_IN((k_rollout_mask<true, false, true>));
  PW_OPT_IN((k_rollout_mask<false, true, true>));
  PW_OPT_IN((k_rollout_mask<false, false, true>));
#undef PW_OPT_IN
  return err;
}

}  // namespace pw
