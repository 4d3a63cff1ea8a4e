// pw_occ.cu -- the general-rule rollout kernel (BASELINE config 3: mixed batch over all puzzles).
//
// What a random-policy step of the 223 benchmark puzzles looks like (measured with the C oracle,
// 7.3e6 steps): 84 % the agent walks into free space, 10 % it is stopped by a wall, 5 % it touches
// a movable (2 % end in transitive stopping, 3 % push 1.35 objects on average).  The table rule of
// pw_fast.cuh pays a scan over all n-1 other objects on EVERY step to find that out.  This kernel:
//
//  * keeps, per environment, the union occupancy of its movables as one uint32 per grid row in
//    shared memory (puzzles up to 32 x 32 cells incl. the border: 211 of the 223; puzzle.py's
//    rule evaluated on row masks like pw_core.cuh).  The common step is then ONE collision-table
//    look-up for the walls (tab_wall, exact incl. the "not already overlapping" clause of
//    puzzle.py:562) and ONE row test against the occupancy: if the moved agent touches no movable
//    cell, pushes(agent, j) is false for every j (puzzle.py:567-593 needs the moved pusher to
//    overlap j), so the agent just moves.  Only the agent's position lives in registers.
//  * lets the lanes of a warp RUN AHEAD of each other: every lane (one environment) has its own
//    step index inside a window of 64 steps.  A lane whose agent does touch something parks its
//    step; when enough lanes are parked (or nobody else can advance) the parked lanes run the
//    exact closure of pw_fast.cuh (pushed_set_tab: the reference's table look-ups) TOGETHER, so
//    the expensive part executes with many active lanes instead of the 1-2 that need it in any
//    one step.  Pushed objects move in shared memory and the occupancy rows are rebuilt.
//  * stages actions with cp.async (three [32 steps][32 envs] tiles per warp) and per-step results
//    as one record byte in a shared-memory ring that is flushed, fully coalesced, whenever all
//    lanes of the warp have left a 32-step tile: either expanded to reward f32 + terminated u8
//    (pw_rollout) or copied as is (pw_rollout_packed).
// Puzzles that do not fit the occupancy rows (wider or taller than 32) park every step that is
// not stopped by a wall, i.e. they run the table rule as before, inside the same batches.
//
// Results are independent of the schedule: an environment's steps are executed in order, each by
// exactly one of the two exact paths.
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
constexpr int kActTiles = 3;       // action tiles per warp: two readable + one in flight
constexpr int kTileBytes = 32 * 32;

// State policy of this kernel (see pw_env.cuh for the interface): the agent's position lives in
// registers (it changes in most steps), the other objects' positions in shared memory, column
// `threadIdx.x` of an [n_max - 1][kOccBlock] int32 array per coordinate (they are only touched when
// something is pushed).  Tables are read from global memory through the read-only path.
struct OccEnv : EnvBase<GlobalTable> {
  static constexpr int kStaticN = 0;
  static constexpr int kUnroll = 1;
  int ax, ay;
  int* sx;  // object j >= 1 at sx[(j - 1) * kOccBlock]
  int* sy;
  // masks: sy holds, instead of y, the ROW MASK of the object -- bit r set for every grid row
  // y .. y + h - 1 it spans (y = lowest set bit): the candidate filter of closure_occ.  Only for
  // environments whose objects all lie inside a grid of at most 32 rows (see build_occupancy).
  bool masks = false;
  __device__ __forceinline__ uint32_t rowmask(int j) const { return (uint32_t)sy[(j - 1) * kOccBlock]; }
  __device__ __forceinline__ void init(GlobalTable table, int* x_col, int* y_col) {
    this->init_header(table);
    sx = x_col;
    sy = y_col;
  }
  __device__ __forceinline__ void load(const int32_t* x, const int32_t* y, int64_t stride, int64_t e) {
    ax = x[e];
    ay = y[e];
    for (int j = 1; j < this->n_; ++j) {
      sx[(j - 1) * kOccBlock] = x[j * stride + e];
      sy[(j - 1) * kOccBlock] = y[j * stride + e];
    }
  }
  __device__ __forceinline__ void store(int32_t* x, int32_t* y, int64_t stride, int64_t e) const {
    x[e] = ax;
    y[e] = ay;
    for (int j = 1; j < this->n_; ++j) {
      x[j * stride + e] = sx[(j - 1) * kOccBlock];
      y[j * stride + e] = this->y(j);
    }
  }
  __device__ __forceinline__ uint32_t desc(int j) const { return this->table_desc(j, 0); }
  __device__ __forceinline__ uint32_t desc_dyn(int j) const { return this->table_desc(j, 0); }
  // j >= 1 in every loop over pushees / goal objects; index 0 only through x(0) / x_dyn(pusher)
  __device__ __forceinline__ int x(int j) const { return j == 0 ? ax : sx[(j - 1) * kOccBlock]; }
  __device__ __forceinline__ int y(int j) const {
    if (j == 0) return ay;
    const int v = sy[(j - 1) * kOccBlock];
    return masks ? __ffs(v) - 1 : v;
  }
  __device__ __forceinline__ int x_dyn(int j) const { return x(j); }
  __device__ __forceinline__ int y_dyn(int j) const { return y(j); }
  __device__ __forceinline__ int goal_x(int k) const { return lo16(this->table_desc(k + 1, 2)); }
  __device__ __forceinline__ int goal_y(int k) const { return hi16(this->table_desc(k + 1, 2)); }
  __device__ __forceinline__ void translate(uint32_t mask, int dx, int dy) {
    if (mask & 1u) {
      ax += dx;
      ay += dy;
    }
    mask &= ~1u;
    while (mask) {
      const int j = __ffs((int)mask) - 1;
      mask &= mask - 1;
      if (dx) sx[(j - 1) * kOccBlock] += dx;
      if (dy) sy[(j - 1) * kOccBlock] += dy;  // (plain positions: the generic closure never runs with masks)
    }
  }
};

// Builds the occupancy rows of this thread's environment from the positions of objects 1..n-1
// (occ: &s_occ[0][tid], row r at occ[r * kOccBlock]) and reports whether the environment can use
// them: every movable inside the grid, no two movables overlapping, none overlapping a wall.  Such a
// state stays that way: a movable that does not overlap a wall is stopped by the border walls
// (puzzle.py:160-168, 284-293), and a movable that would be moved onto another one pushes it
// (:359-382), so the occupancy can be maintained incrementally.  Every other environment runs the
// table rule on every step.
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

// The push closure of one step (puzzle.py:359-394) for an environment with occupancy rows:
//   * a pusher whose displaced cells touch no occupied cell pushes nothing (puzzle.py:567-593 needs
//     the moved pusher to overlap the pushee), so its scan over the other objects is skipped;
//   * otherwise only objects that span one of the rows with a hit are tested, with the exact
//     collision-table look-ups of pw_fast.cuh (tab_pushes, tab_wall);
//   * objects that join the pushed set leave the occupancy and re-enter it displaced (or where they
//     were, on transitive stopping, puzzle.py:376-379).
// Returns the moved-object mask (0 = nothing moves); positions and occupancy are updated.
__device__ __forceinline__ uint32_t closure_occ(OccEnv& env, uint32_t* occ, uint32_t a, int H, uint32_t nn1) {
  const int n = env.n();
  const int dx = action_dx((int)a), dy = action_dy((int)a);
  // (the caller has already found that the walls do not stop the agent, puzzle.py:353-354)
  uint32_t pushed = 1u, frontier = 1u;
  bool stop = false;
  while (frontier) {
    const int i = __ffs((int)frontier) - 1;
    frontier &= frontier - 1;
    const uint32_t di = env.desc(i);
    const int xi = env.x(i), yi = env.y(i);
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
    if (hitrows == 0u) continue;
    uint32_t cand = 0u;
#pragma unroll 4
    for (int j = 1; j < n; ++j) cand |= (env.rowmask(j) & hitrows) ? (1u << j) : 0u;
    cand &= ~pushed;
    while (cand) {
      const int j = __ffs((int)cand) - 1;
      cand &= cand - 1;
      const int xj = env.x(j), yj = env.y(j);
      if (!tab_pushes(env, 1u + (uint32_t)(i * n + j), a, xj - xi, yj - yi)) continue;
      if (tab_wall(env, nn1, (uint32_t)j, a, xj, yj)) {  // transitive stopping
        stop = true;
        break;
      }
      pushed |= 1u << j;
      frontier |= 1u << j;
      const uint32_t dj = env.desc(j);
      const int hj = desc_h(dj), oj = desc_off(dj);
      for (int r = 0; r < hj; ++r) occ[(yj + r) * kOccBlock] &= ~(uint32_t)(env.row(oj + r) << xj);
    }
    if (stop) break;
  }
  uint32_t rest = pushed & ~1u;
  while (rest) {
    const int j = __ffs((int)rest) - 1;
    rest &= rest - 1;
    const uint32_t dj = env.desc(j);
    const int hj = desc_h(dj), oj = desc_off(dj);
    const int nx = env.x(j) + (stop ? 0 : dx), ny = env.y(j) + (stop ? 0 : dy);
    for (int r = 0; r < hj; ++r) occ[(ny + r) * kOccBlock] |= (uint32_t)(env.row(oj + r) << nx);
    if (!stop) {
      env.sx[(j - 1) * kOccBlock] = nx;
      const uint32_t m = env.rowmask(j);
      env.sy[(j - 1) * kOccBlock] = (int)(dy > 0 ? m << 1 : (dy < 0 ? m >> 1 : m));
    }
  }
  if (stop) return 0u;
  env.ax += dx;
  env.ay += dy;
  return pushed;
}

template <bool kPacked>
__global__ void __launch_bounds__(kOccBlock)
    k_rollout_occ(const __grid_constant__ StepParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ uint32_t s_counters[4];
  // dynamic shared memory: occupancy | positions | action tiles | result ring | reward table
  uint32_t* s_occ = reinterpret_cast<uint32_t*>(smem_raw);
  int* s_pos = reinterpret_cast<int*>(s_occ + kOccRows * kOccBlock);
  const int pos_rows = p.n_max > 1 ? p.n_max - 1 : 1;
  uint8_t* s_act = reinterpret_cast<uint8_t*>(s_pos + 2 * pos_rows * kOccBlock);
  uint8_t* s_out = s_act + (kOccBlock / 32) * kActTiles * kTileBytes;
  float* s_lut = reinterpret_cast<float*>(s_out + (kOccBlock / 32) * 2 * kTileBytes);
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
    const int pid = p.puzzle_id ? p.puzzle_id[e] : p.puzzle_uniform;
    OccEnv env;
    env.init(GlobalTable{p.blob + p.puzzle_off[pid]}, s_pos + tid, s_pos + pos_rows * kOccBlock + tid);
    env.load(p.x, p.y, p.stride, e);
    const int n = env.n(), g = env.g(), W = env.width(), H = env.height();
    const uint32_t nn1 = 1u + (uint32_t)(n * n);
    // the agent's wall table (pw_fast.cuh: tab_wall), descriptor cached in registers
    const uint64_t wd = env.fast(nn1);
    const uint32_t aw_hi = (uint32_t)(wd >> 32);
    const uint32_t aw_xmax = aw_hi & 63u, aw_ymax = (aw_hi >> 6) & 0xfffu, aw_w = aw_hi >> 18;
    const uint32_t aw_word0 = (uint32_t)wd << 1;
    const uint32_t d0 = env.desc(0);
    const int h0 = desc_h(d0), o0 = desc_off(d0);
    const bool agent_1x1 = (d0 >> 16) == 0x0101u;
    uint32_t* const occ = s_occ + tid;
    // occupancy rows: puzzles of at most 32 x 32 cells whose state is regular (build_occupancy)
    const bool fast_ok = W <= 32 && H <= kOccRows && build_occupancy(env, occ, W, H);
    if (fast_ok) {  // rows instead of y for objects 1 .. n-1 (see OccEnv::masks)
      for (int j = 1; j < n; ++j) {
        const int h = desc_h(env.desc(j));
        env.sy[(j - 1) * kOccBlock] = (int)((0xffffffffu >> (32 - h)) << env.sy[(j - 1) * kOccBlock]);
      }
      env.masks = true;
    }
    int& ax = env.ax;
    int& ay = env.ay;
    int ach = achieved_goals(env);
    uint32_t code_idle = record_byte(0, ach == g, false);  // the record of a step that moves no movable

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
    // results of tile `tile` (steps 32*tile ...) leave the ring, coalesced
    auto flush_tile = [&](const int tile) {
      const int rows = T - tile * 32 < 32 ? T - tile * 32 : 32;
      const uint8_t* src = out_ring + (tile & 1) * kTileBytes;
      const size_t first = (size_t)tile * 32u * (size_t)p.N + (size_t)(e - (int64_t)lane);
      if (kPacked) {
        // 32 rows x 32 bytes = 64 chunks of 16 bytes, two per lane
#pragma unroll
        for (int k = 0; k < 2; ++k) {
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
    uint32_t n_term = 0, n_moved = 0, n_gain = 0, n_invalid = 0;
    // actions of a step as packed displacement bytes: LEFT, RIGHT, UP, DOWN (puzzle.py:43-50)
    constexpr uint32_t kDxBytes = 0x000001ffu, kDyBytes = 0x01ff0000u;
    const int batch = p.occ_batch;  // parked lanes that trigger a closure pass
    while (base < n_tiles) {
      const int lim = min(T, 32 * (base + 2));
      if (!parked && t < lim) {
        // ---- fast attempt (branch-free: the lanes of a warp stay together) ---------------------
        const uint32_t a = act_ring[(uint32_t)((t >> 5) % kActTiles) * kTileBytes + ((uint32_t)(t & 31) << 5) + lane];
        const bool valid = a <= 3u;
        const uint32_t sel = a * 0x1111u + 0x8880u;  // byte a, sign extended
        const int dx = (int)__byte_perm(kDxBytes, 0u, sel), dy = (int)__byte_perm(kDyBytes, 0u, sel);
        // walls (puzzle.py:353-354): one bit of the agent's collision table
        const bool here = (uint32_t)ax <= aw_xmax && (uint32_t)ay <= aw_ymax;
        const uint32_t idx = here ? ((((uint32_t)ay * aw_w + (uint32_t)ax) << 2) | (a & 3u)) : 0u;
        const bool stopped = here && tab_bit(env.fast32(aw_word0 + (idx >> 5)), idx);
        const int x = ax + dx, y = ay + dy;
        // does the moved agent (inside the grid: rows y .. y + h0 - 1 exist) touch a movable?
        const bool there = fast_ok && (uint32_t)x <= aw_xmax && (uint32_t)y <= aw_ymax;
        uint32_t hit;
        if (agent_1x1) {
          hit = (occ[(there ? y : 0) * kOccBlock] >> (x & 31)) & 1u;
        } else {
          hit = 0u;
          if (there)
            for (int r = 0; r < h0; ++r) hit |= (uint32_t)(env.row(o0 + r) << x) & occ[(y + r) * kOccBlock];
        }
        const bool go = valid && !stopped;
        const bool free_move = go && there && hit == 0u;
        n_invalid += valid ? 0u : 1u;
        n_moved += free_move ? 1u : 0u;
        ax = free_move ? x : ax;
        ay = free_move ? y : ay;
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
      const uint32_t m_ready = __ballot_sync(0xffffffffu, !parked && t < lim);
      if (m_parked != 0u && (__popc(m_parked) >= batch || m_ready == 0u)) {
        if (parked) {
          // ---- the exact closure (puzzle.py:359-394 on the collision tables), parked lanes only
          uint32_t pushed;
          if (fast_ok) {
            pushed = closure_occ(env, occ, parked_action, H, nn1);
          } else {
            pushed = pushed_set_tab(env, (int)parked_action);
            if (pushed) env.translate(pushed, action_dx((int)parked_action), action_dy((int)parked_action));
          }
          int now = ach;
          if (pushed) {
            ++n_moved;
            if (pushed & ~1u) now = achieved_goals(env);
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
    const int pid = p.puzzle_id ? p.puzzle_id[e] : p.puzzle_uniform;
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

size_t occ_smem_bytes(int n_max) {
  const int pos_rows = n_max > 1 ? n_max - 1 : 1;
  return (size_t)kOccRows * kOccBlock * 4 + (size_t)2 * pos_rows * kOccBlock * 4 +
         (size_t)(kOccBlock / 32) * (kActTiles + 2) * kTileBytes + 256 * sizeof(float);
}

}  // namespace

// Serves every plain / packed rollout of the general rule made of whole warps.
bool occ_rollout_eligible(const StepParams& p) {
  const auto aligned16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  const bool plain = p.reward && p.terminated && !p.record;
  const bool packed = p.record && !p.reward && !p.terminated;
  return is_lean(p) && (plain || packed) && (p.N & 31) == 0 && aligned16(p.actions) &&
         (!packed || aligned16(p.record));
}

cudaError_t launch_rollout_occ(const StepParams& p_in, cudaStream_t stream) {
  StepParams p = p_in;
  // how many parked lanes trigger a closure pass (PW_OCC_BATCH overrides: tuning runs)
  p.occ_batch = 12;
  if (const char* env = getenv("PW_OCC_BATCH")) {
    const int v = atoi(env);
    if (v >= 1 && v <= 32) p.occ_batch = v;
  }
  if (getenv("PW_WARP_PER_ENV") != nullptr) {  // the measured alternative layout
    const unsigned warps = kWarpEnvBlock / 32;
    k_rollout_warp_per_env<<<(unsigned)((p.N + warps - 1) / warps), kWarpEnvBlock, 0, stream>>>(p);
    return cudaGetLastError();
  }
  const unsigned grid = (unsigned)((p.N + kOccBlock - 1) / kOccBlock);
  const size_t smem = occ_smem_bytes(p.n_max);
  if (p.record) k_rollout_occ<true><<<grid, kOccBlock, smem, stream>>>(p);
  else k_rollout_occ<false><<<grid, kOccBlock, smem, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t configure_occ_kernels() {
  const int cap = (int)occ_smem_bytes(kMaxObjects);
  cudaError_t err = cudaFuncSetAttribute(k_rollout_occ<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
  if (err == cudaSuccess)
    err = cudaFuncSetAttribute(k_rollout_occ<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
  return err;
}

}  // namespace pw
