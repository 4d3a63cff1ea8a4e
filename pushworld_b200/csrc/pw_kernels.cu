// pw_kernels.cu -- sm_100a kernels for the PushWorld dynamics path (step / rollout / expand /
// reset) and their launchers.  One thread per environment; see pw_env.cuh for the two state
// policies and pw_core.cuh for the transition rule.  DESIGN.md section 3 has the rationale and
// the measured comparison against a warp-per-environment layout.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>
#include <type_traits>

#include "pw_board.cuh"
#include "pw_env.cuh"
#include "pw_internal.h"
#include "pw_steps.cuh"

namespace pw {

cudaError_t upload_reward_lut(const float* host_lut) {
  return cudaMemcpyToSymbol(c_reward_lut, host_lut, sizeof(float) * (2 * kMaxObjects + 1));
}

namespace {

#define COMMA ,
constexpr int kBlock = 128;
// most shared memory a uniform-puzzle launch may use for its staged tables (+ state arrays)
constexpr size_t kStageLimit = 160 * 1024;

// Copies one puzzle blob into shared memory (all threads of the block take part).
__device__ __forceinline__ void stage_table(uint64_t* dst, const uint64_t* src, int n_u64) {
  for (int i = threadIdx.x; i < n_u64; i += blockDim.x) dst[i] = __ldg(src + i);
  __syncthreads();
}

template <bool kEpisodes, bool kPacked, class Sim>
__device__ __forceinline__ void run_lean_sim(Sim& sim, const StepParams& p, const int64_t e,
                                             const float* s_rlut, uint8_t* tiles, StepCounts& c);
// dispatch on the output shape: reward + terminated, or one packed record byte per step
template <bool kEpisodes, class Sim>
__device__ __forceinline__ void run_lean_sim_any(Sim& sim, const StepParams& p, const int64_t e,
                                                 const float* s_rlut, uint8_t* tiles, StepCounts& c) {
  if (p.record) run_lean_sim<kEpisodes, true>(sim, p, e, s_rlut, tiles, c);
  else run_lean_sim<kEpisodes, false>(sim, p, e, s_rlut, tiles, c);
}
template <class Env>
struct GeneralLeanSim;

template <bool kLean, class Env>
__device__ __forceinline__ void run_steps(Env& env, const StepParams& p, int64_t e, bool active,
                                          uint32_t* s_counters) {
  StepCounts c;
  if constexpr (kLean) {
    // plain rollouts made of whole warps: action tiles + FMA-pipe addressing (run_lean_sim)
    __shared__ __align__(16) uint8_t s_tiles[kBlock / 32][2][32 * 32];
    __shared__ float s_rlut[2 * kMaxObjects + 1];
    if (p.lean_tiles) {
      for (int i = threadIdx.x; i < 2 * kMaxObjects + 1; i += blockDim.x) s_rlut[i] = c_reward_lut[i];
      __syncthreads();
      if (active) {
        env.load(p.x, p.y, p.stride, e);
        GeneralLeanSim<Env> sim{env};
        run_lean_sim_any<false>(sim, p, e, s_rlut, s_tiles[threadIdx.x >> 5][0], c);
      }
    } else if (active) {
      run_steps_thread<kLean>(env, p, e, c);
    }
  } else {
    if (active) run_steps_thread<kLean>(env, p, e, c);
  }
  flush_counts(c, p, s_counters);
}

// All environments of the launch share puzzle `p.puzzle_uniform`: its table is staged in
// shared memory once per block.
template <int NMAX, bool kLean>
__global__ void __launch_bounds__(kBlock)
    k_steps_uniform(const __grid_constant__ StepParams p) {
  extern __shared__ uint64_t smem[];
  __shared__ uint32_t s_counters[4];
  if (threadIdx.x < 4) s_counters[threadIdx.x] = 0;
  const uint64_t* src = p.blob + p.puzzle_off[p.puzzle_uniform];
  stage_table(smem, src, p.uniform_u64);
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = e < p.N;
  if constexpr (NMAX > 0) {
    RegEnv<NMAX, SharedTable> env;
    env.init(SharedTable{smem});
    run_steps<kLean>(env, p, e, active, s_counters);
  } else {
    SmemEnv<SharedTable> env;
    int* st = reinterpret_cast<int*>(smem + p.uniform_u64);
    env.init(SharedTable{smem}, st + threadIdx.x, st + p.n_max * kBlock + threadIdx.x, kBlock);
    run_steps<kLean>(env, p, e, active, s_counters);
  }
}

// ------------------------------------------------------------------------------------------
// Bitboard kernel: single puzzle with width*height <= 64 and n <= NMAX (pw_board.cuh).  The whole
// per-step state of an environment is NMAX uint64 registers and a step is branch-light shift /
// and / or work.  Environments whose loaded state is not regular take the general rule through
// a non-inlined slow path (positions in local memory), so results are exact for any input.
// ------------------------------------------------------------------------------------------
// per-block static boards, computed once by thread 0
template <typename Word>
struct BoardStatic {
  Word wall, awall, edge_up, edge_down;
  Word goal[8];
  Word init[8];
  BoardAction<Word> act[4];
  int P, n, g;
};

// Writes the positions back.  After a single step (pw_step) only the pushed objects changed, by
// the action's displacement: update those in place instead of re-deriving every position from its
// board (which costs an integer division per object).
template <int NMAX, typename Word>
__device__ __forceinline__ void board_store(const Board<NMAX, Word>& s, const StepParams& p, int64_t e,
                                            uint32_t last_moved, int last_action) {
  if (p.T == 1 && !(p.flags & kFlagAutoreset)) {
    const int dx = action_dx(last_action), dy = action_dy(last_action);
#pragma unroll
    for (int k = 0; k < NMAX; ++k)
      if ((last_moved >> k) & 1u) {
        if (dx) p.x[k * p.stride + e] += dx;
        if (dy) p.y[k * p.stride + e] += dy;
      }
    return;
  }
#pragma unroll
  for (int k = 0; k < NMAX; ++k)
    if (k < s.n) {
      int x, y;
      board_position<Word>(s.b[k], s.P, &x, &y);
      p.x[k * p.stride + e] = x;
      p.y[k * p.stride + e] = y;
    }
}

template <int NMAX, typename Word>
struct BoardSim {
  static constexpr bool kCheapAchieved = true;
  static constexpr int kPrefetch = 4;
  Board<NMAX, Word>& s;
  const BoardStatic<Word>& st;
  uint32_t last_moved = 0;
  int last_action = 0;
  __device__ __forceinline__ void store(const StepParams& p, int64_t e) {
    board_store(s, p, e, last_moved, last_action);
  }
  __device__ __forceinline__ void load_initial() {
#pragma unroll
    for (int k = 0; k < NMAX; ++k) s.b[k] = st.init[k];
  }
  __device__ __forceinline__ void resample(const StepParams&, int64_t) { load_initial(); }  // single puzzle
  __device__ __forceinline__ uint32_t do_step(int a) {
    last_action = a;
    return last_moved = board_step(s, a, st.act);
  }
  __device__ __forceinline__ int achieved() const { return board_achieved(s); }
  __device__ __forceinline__ int goals() const { return s.g; }
};

__device__ __noinline__ void general_slow_path(const uint64_t* table, const StepParams& p, int64_t e,
                                               StepCounts& c) {
  LocalEnv<SharedTable> env;
  env.init(SharedTable{table});
  run_steps_thread<false>(env, p, e, c);
}

template <int NMAX, typename Word>
__device__ __forceinline__ void run_board_lean(Board<NMAX, Word>& s, const StepParams& p, const int64_t e,
                                               const float* s_rlut, uint8_t* tiles,
                                               const BoardAction<Word>* acts, StepCounts& c);

template <int NMAX, typename Word, int MIN_BLOCKS, bool kLean>
__global__ void __launch_bounds__(kBlock, MIN_BLOCKS)
    k_steps_board(const __grid_constant__ StepParams p) {
  extern __shared__ uint64_t smem[];
  __shared__ uint32_t s_counters[4];
  __shared__ BoardStatic<Word> st;
  // lean launches only (see run_board_lean)
  __shared__ __align__(16) uint8_t s_tiles[kLean ? kBlock / 32 : 1][2][kLean ? 32 * 32 : 16];
  __shared__ BoardAction<Word> s_act[kLean ? 4 : 1][kLean ? kBlock : 1];
  __shared__ float s_rlut[kLean ? 2 * kMaxObjects + 1 : 1];
  if (kLean)
    for (int i = threadIdx.x; i < 2 * kMaxObjects + 1; i += blockDim.x) s_rlut[i] = c_reward_lut[i];
  if (threadIdx.x < 4) s_counters[threadIdx.x] = 0;
  const SharedTable tbl{smem};
  // static boards were computed on the host at engine creation (pw_abi.cu: make_board_static)
  const BoardStaticPod* __restrict__ pod = p.board_static;
  if (threadIdx.x < 8) {
    st.goal[threadIdx.x] = (Word)__ldg(&pod->goal[threadIdx.x]);
    st.init[threadIdx.x] = (Word)__ldg(&pod->init[threadIdx.x]);
  }
  if (threadIdx.x < 4) {
    st.act[threadIdx.x].edge = (Word)__ldg(&pod->act_edge[threadIdx.x]);
    st.act[threadIdx.x].sl = __ldg(&pod->act_sl[threadIdx.x]);
    st.act[threadIdx.x].sr = __ldg(&pod->act_sr[threadIdx.x]);
  }
  if (threadIdx.x == 0) {
    st.wall = (Word)__ldg(&pod->wall);
    st.awall = (Word)__ldg(&pod->awall);
    st.edge_up = (Word)__ldg(&pod->edge_up);
    st.edge_down = (Word)__ldg(&pod->edge_down);
    st.P = __ldg(&pod->P);
    st.n = __ldg(&pod->n);
    st.g = __ldg(&pod->g);
  }
  stage_table(smem, p.blob + p.puzzle_off[p.puzzle_uniform], p.uniform_u64);  // ends with a barrier
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  StepCounts c;
  if (e < p.N) {
    Board<NMAX, Word> s;
    s.wall = st.wall;
    s.awall = st.awall;
    s.edge_up = st.edge_up;
    s.edge_down = st.edge_down;
    s.P = st.P;
    s.n = st.n;
    s.g = st.g;
#pragma unroll
    for (int k = 0; k < NMAX; ++k) s.goal[k] = st.goal[k];
    const bool regular = board_load(s, tbl, [&](int k) { return p.x[k * p.stride + e]; },
                                    [&](int k) { return p.y[k * p.stride + e]; });
    if (kLean && p.lean_tiles) {
      // plain rollout with whole warps: the action tiles / addressing of the ring kernel (run_board_lean)
      if (__all_sync(0xffffffffu, regular)) {
#pragma unroll
        for (int a = 0; a < 4; ++a) s_act[a][threadIdx.x] = st.act[a];
        run_board_lean<NMAX, Word>(s, p, e, s_rlut, s_tiles[threadIdx.x >> 5][0], &s_act[0][threadIdx.x], c);
      } else {
        general_slow_path(smem, p, e, c);
      }
    } else if (regular) {
      BoardSim<NMAX, Word> sim{s, st, 0u, 0};
      run_steps_sim<kLean>(sim, p, e, c);
    } else {
      general_slow_path(smem, p, e, c);
    }
  }
  flush_counts(c, p, s_counters);
}

// ------------------------------------------------------------------------------------------
// Ring-board rollout (pw_board.cuh: ring_step): the plain-rollout shape (reward + terminated per
// step, nothing else) of a single puzzle whose board leaves P spare bits in a uint32.  The loop
// is written out here instead of going through run_steps_sim because the integer ALU pipe is the
// bound: the rule itself (one rotate per object, board tests, predicate logic,
// selects) has to run there, so the bookkeeping around it is kept off that pipe -- addresses, the
// terminated flag of the one-goal case (kG1) and everything derived from it are multiply-adds,
// and actions come out of shared-memory tiles filled by cp.async a block of steps ahead.
// ------------------------------------------------------------------------------------------
template <int NMAX, bool kG1, bool kPacked>
__device__ __forceinline__ void run_ring_lean(const Board<NMAX, uint32_t>& s, const StepParams& p,
                                              const int64_t e, const float* s_r4, uint8_t* tiles,
                                              StepCounts& c) {
  uint32_t b[NMAX];
#pragma unroll
  for (int k = 0; k < NMAX; ++k) b[k] = s.b[k];
  const uint32_t guard = ring_guard(p.board_bits);
  const uint32_t wall = s.wall | guard, awall = s.awall | guard;
  const uint32_t amounts = ring_amounts(s.P);
  const int g = s.g;
  // one-goal flag arithmetic below needs (board XOR goal) < 2^30: an unreachable goal (all ones)
  // becomes the guard bit 0, which no object ever owns
  const uint32_t goal0 = s.goal[0] == 0xffffffffu ? 1u : s.goal[0];
  const int T = p.T;
  // Addressing.  A lean launch has T * N < 2^32 and the ring launcher requires N even, so the
  // element offset of (t, e) is 2 * h + (e & 1) with the 32-bit h = (e >> 1) + t * (N / 2):
  // every address is one IMAD.WIDE (h times 2 or 8 bytes, plus a per-thread base) on the FMA pipe.
  // (A multiplier of 1 would be strength-reduced to an add-with-carry pair on the ALU pipe.)
  const uint32_t half_n = (uint32_t)(p.N >> 1);
  // kPacked (pw_rollout_packed): ONE record byte per step instead of reward + terminated
  uint8_t* term_base = (kPacked ? p.record : p.terminated) + (e & 1);
  float* const reward_base = p.reward + (e & 1);
  // keep the byte base an opaque per-thread register pair: otherwise the compiler splits it back
  // into (uniform pointer) + (offset) and adds the pointer with a carry pair per access
  // (and a PTX store below, because the opaque base has lost its address space)
  asm volatile("" : "+l"(term_base));
  uint32_t h_out = (uint32_t)(e >> 1);  // half offset of the step being written
  auto achieved = [&]() -> int {
    int count = 0;
#pragma unroll
    for (int k = 0; k + 1 < NMAX; ++k) count += (b[k + 1] == s.goal[k]) ? 1 : 0;  // goal[k >= g] never matches
    return count;
  };
  int prev = achieved();
  // kG1: on4 = 4 * (the goal object sat on its goal after the previous step); the reward is
  // s_r4[2 * terminated + previously_on_goal], read with the byte offset 8 * terminated + on4.
  // Everything derived from `terminated` below is a multiply-add, i.e. off the ALU pipe.
  // (packed records: the unit is 1 instead of 4, record = 32 + delta_goals + 128 * terminated)
  constexpr uint32_t kOn = kPacked ? 1u : 4u;
  uint32_t on4 = (kG1 && b[1] == goal0) ? kOn : 0u;
  uint32_t n_term = 0, n_both4 = 0, n_moved = 0, n_gain = 0, n_invalid = 0;
  auto one_step = [&](const uint32_t a, auto checked) {
    constexpr bool kChecked = decltype(checked)::value;
    const bool valid = kChecked ? a <= 3u : true;
    n_moved += ring_step<NMAX, false>(b, wall, awall, ring_amount_of(amounts, a), valid);
    uint32_t tbyte, code = 0;
    float r = 0.f;
    if (kG1) {
      // puzzle.py:409-411 with one goal object, as a multiply-add: d = board XOR goal < 2^30 (ring
      // boards have at most 30 bits), so the high word of d * -3 is -1 for d != 0 and 0 for d == 0.
      // (A compare would make the flag a predicate and everything derived from it below a select
      // on the ALU pipe; a multiplier of -1 would be strength-reduced to ALU instructions.)
      asm("mad.hi.s32 %0, %1, -3, 1;" : "=r"(tbyte) : "r"(b[1] ^ goal0));
      if (kPacked) {
        code = tbyte * 129u + ((uint32_t)kMaxObjects - on4);  // delta_goals = on now - on before
      } else {
        r = *reinterpret_cast<const float*>(reinterpret_cast<const char*>(s_r4) + (tbyte * 8u + on4));
      }
      n_both4 += tbyte * on4;
      on4 = tbyte * kOn;
    } else {
      const int ach = achieved();
      const bool term = ach == g;
      const int dg = ach - prev;
      prev = ach;
      // gym_env.py:212-221; the float64 value (dg - 0.01) was rounded to float32 once, on the host
      if (kPacked) code = record_byte(dg, term, false);
      else r = term ? 10.0f : c_reward_lut[dg + kMaxObjects];
      n_gain += (uint32_t)max(dg, 0);
      tbyte = term ? 1u : 0u;
    }
    n_term += tbyte;
    if (kChecked) n_invalid += valid ? 0u : 1u;
    if (!kPacked) reward_base[(size_t)h_out * 2u] = r;
    asm volatile("st.global.u8 [%0], %1;" ::"l"(term_base + (size_t)h_out * 2u), "r"(kPacked ? code : tbyte)
                 : "memory");
    h_out += half_n;
  };
  // Actions do not depend on the state, so they are staged a whole block of 32 steps ahead:
  // lane l copies the 32 action bytes of the warp's environments for step 32 * block + l
  // (one 32-byte sector of actions[T][N]) straight into shared memory with cp.async, into one of
  // two [32 steps][32 envs] tiles per warp.  A step then reads its action with one LDS at an
  // immediate offset: no address arithmetic, no registers held across the DRAM latency.
  // Each lane also tests the row it copied for invalid actions; the warp votes, and a block
  // without any runs with no per-step action checks.
  constexpr int kBlockSteps = 32, kGroup = 4;
  const uint32_t lane = threadIdx.x & 31u;
  const uint8_t* const warp_actions = p.actions + (e - (int64_t)lane);
  const uint32_t tiles_s = (uint32_t)__cvta_generic_to_shared(tiles);
  auto issue_block = [&](const int block) {
    const int row = block * kBlockSteps + (int)lane;
    const uint8_t* src = warp_actions + (size_t)(row < T ? row : T - 1) * (size_t)p.N;  // clamped: never used
    const uint32_t dst = tiles_s + (uint32_t)(block & 1) * 1024u + lane * 32u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u), "l"(src + 16) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const int blocks = (T + kBlockSteps - 1) / kBlockSteps;
  issue_block(0);
  for (int blk = 0; blk < blocks; ++blk) {
    __syncwarp();          // every lane is done with the tile the next block lands in
    issue_block(blk + 1);  // past the end: re-reads the last row, never used
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncwarp();  // block `blk` is visible to the whole warp
    const uint8_t* tile = tiles + (blk & 1) * 1024;
    const uint4 w0 = *reinterpret_cast<const uint4*>(tile + lane * 32u);
    const uint4 w1 = *reinterpret_cast<const uint4*>(tile + lane * 32u + 16u);
    const uint32_t seen = w0.x | w0.y | w0.z | w0.w | w1.x | w1.y | w1.z | w1.w;
    const bool any_invalid = __any_sync(0xffffffffu, (seen & 0xfcfcfcfcu) != 0u);
    const int steps = T - blk * kBlockSteps < kBlockSteps ? T - blk * kBlockSteps : kBlockSteps;
    const uint8_t* mine = tile + lane;
    if (!any_invalid && steps == kBlockSteps) {
#pragma unroll 1
      for (int gi = 0; gi < kBlockSteps / kGroup; ++gi, mine += kGroup * 32) {
#pragma unroll
        for (int k = 0; k < kGroup; ++k) one_step(mine[k * 32], std::false_type{});
      }
    } else {
#pragma unroll 1
      for (int k = 0; k < steps; ++k, mine += 32) one_step(mine[0], std::true_type{});
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  c.terminated += n_term;
  c.moved += n_moved;
  c.dgoals += kG1 ? n_term - n_both4 / kOn : n_gain;  // gained = on the goal now and not before
  c.invalid += n_invalid;
  // positions back from the boards
#pragma unroll
  for (int k = 0; k < NMAX; ++k)
    if (k < s.n) {
      int x, y;
      board_position<uint32_t>(b[k], s.P, &x, &y);
      p.x[k * p.stride + e] = x;
      p.y[k * p.stride + e] = y;
    }
}

template <int NMAX, bool kG1, int MIN_BLOCKS>
__global__ void __launch_bounds__(kBlock, MIN_BLOCKS)
    k_rollout_ring(const __grid_constant__ StepParams p) {
  extern __shared__ uint64_t smem[];
  __shared__ uint32_t s_counters[4];
  __shared__ BoardStatic<uint32_t> st;
  __shared__ float s_r4[4];  // gym_env.py:212-221 for one goal object: -0.01, -1.01, 10, 10
  __shared__ __align__(16) uint8_t s_tiles[kBlock / 32][2][32 * 32];  // per warp: two action tiles
  if (threadIdx.x < 4) {
    s_counters[threadIdx.x] = 0;
    s_r4[threadIdx.x] = threadIdx.x >= 2 ? 10.0f : c_reward_lut[kMaxObjects - (int)threadIdx.x];
  }
  const SharedTable tbl{smem};
  const BoardStaticPod* __restrict__ pod = p.board_static;
  if (threadIdx.x < 8) st.goal[threadIdx.x] = (uint32_t)__ldg(&pod->goal[threadIdx.x]);
  if (threadIdx.x == 0) {
    st.wall = (uint32_t)__ldg(&pod->wall);
    st.awall = (uint32_t)__ldg(&pod->awall);
    st.P = __ldg(&pod->P);
    st.n = __ldg(&pod->n);
    st.g = __ldg(&pod->g);
  }
  stage_table(smem, p.blob + p.puzzle_off[p.puzzle_uniform], p.uniform_u64);  // ends with a barrier
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  StepCounts c;
  if (e < p.N) {
    Board<NMAX, uint32_t> s;
    s.wall = st.wall;
    s.awall = st.awall;
    s.edge_up = s.edge_down = 0;
    s.P = st.P;
    s.n = st.n;
    s.g = st.g;
#pragma unroll
    for (int k = 0; k < NMAX; ++k) s.goal[k] = st.goal[k];
    const bool regular = board_load(s, tbl, [&](int k) { return p.x[k * p.stride + e]; },
                                    [&](int k) { return p.y[k * p.stride + e]; });
    // The action tiles are filled by whole warps (the launcher guarantees N % 32 == 0, so a warp
    // is full or empty): a warp with any irregular state takes the general rule for all its lanes.
    if (__all_sync(0xffffffffu, regular)) {
      if (p.record) run_ring_lean<NMAX, kG1, true>(s, p, e, s_r4, s_tiles[threadIdx.x >> 5][0], c);
      else run_ring_lean<NMAX, kG1, false>(s, p, e, s_r4, s_tiles[threadIdx.x >> 5][0], c);
    } else {
      general_slow_path(smem, p, e, c);
    }
  }
  flush_counts(c, p, s_counters);
}

// ------------------------------------------------------------------------------------------
// Bitboard kernel for per-environment puzzle ids (e.g. a batch over the 15,400 level-0 puzzles):
// every puzzle of the engine is board-eligible, so each thread loads its own puzzle's static
// boards (host-precomputed BoardStaticPod, L2 resident) into registers and computes the action
// constants from its pitch.  pw_rollout_episodes re-points a thread at a freshly drawn puzzle.
// ------------------------------------------------------------------------------------------
template <int NMAX, typename Word>
struct BoardSimMixed {
  static constexpr bool kCheapAchieved = true;
  static constexpr int kPrefetch = 4;
  Board<NMAX, Word>& s;
  const BoardStaticPod* pod;
  uint32_t last_moved = 0;
  int last_action = 0;
  __device__ __forceinline__ void load_static(const BoardStaticPod* q) {
    pod = q;
    s.wall = (Word)__ldg(&q->wall);
    s.awall = (Word)__ldg(&q->awall);
    s.edge_up = (Word)__ldg(&q->edge_up);
    s.edge_down = (Word)__ldg(&q->edge_down);
    s.P = __ldg(&q->P);
    s.n = __ldg(&q->n);
    s.g = __ldg(&q->g);
#pragma unroll
    for (int k = 0; k < NMAX; ++k) s.goal[k] = (Word)__ldg(&q->goal[k]);
  }
  __device__ __forceinline__ void store(const StepParams& p, int64_t e) {
    board_store(s, p, e, last_moved, last_action);
  }
  __device__ __forceinline__ void load_initial() {
#pragma unroll
    for (int k = 0; k < NMAX; ++k) s.b[k] = (Word)__ldg(&pod->init[k]);
  }
  __device__ __forceinline__ void resample(const StepParams& p, int64_t e) {
    const uint32_t episode = p.episode[e] + 1u;
    p.episode[e] = episode;
    const int pid = draw_puzzle(p.seed, e, episode, p.num_puzzles);
    p.puzzle_id_mut[e] = pid;
    load_static(p.board_static + pid);
    load_initial();
  }
  __device__ __forceinline__ uint32_t do_step(int a) {
    last_action = a;
    return last_moved = board_step(s, a, board_action_of(s, (uint32_t)a & 3u));
  }
  __device__ __forceinline__ int achieved() const { return board_achieved(s); }
  __device__ __forceinline__ int goals() const { return s.g; }
};

// ------------------------------------------------------------------------------------------
// The plain-rollout loop of the ring kernel for every other rule: same action tiles (cp.async, 32
// steps ahead), same half-offset addressing, same block-level action check.  `Sim` supplies the
// rule: do_step<kChecked>(a) -> moved mask, achieved(), goals(), store(), kCheapAchieved.
// Whole warps only (the launcher sets p.lean_tiles when N % 32 == 0 and actions are 16-byte aligned).
// ------------------------------------------------------------------------------------------
template <bool kEpisodes, bool kPacked, class Sim>
__device__ __forceinline__ void run_lean_sim(Sim& sim, const StepParams& p, const int64_t e,
                                             const float* s_rlut, uint8_t* tiles, StepCounts& c) {
  // kEpisodes (pw_rollout_episodes): step counter, truncation, and a terminated / truncated
  // environment restarts on a freshly drawn puzzle (gym_env.py:171-186,223 on the device)
  int g = sim.goals();
  int steps = kEpisodes ? p.steps[e] : 0;
  // kPacked (pw_rollout_packed): one record byte per step carries delta_goals, terminated, truncated
  uint8_t* trunc_base = (kEpisodes && !kPacked && p.truncated) ? p.truncated + (e & 1) : nullptr;
  asm volatile("" : "+l"(trunc_base));
  const int T = p.T;
  const uint32_t half_n = (uint32_t)(p.N >> 1);
  uint8_t* term_base = (kPacked ? p.record : p.terminated) + (e & 1);
  float* const reward_base = p.reward + (e & 1);
  asm volatile("" : "+l"(term_base));  // see run_ring_lean
  uint32_t h_out = (uint32_t)(e >> 1);
  int prev = sim.achieved();
  uint32_t n_term = 0, n_moved = 0, n_gain = 0, n_invalid = 0;
  auto one_step = [&](const uint32_t a, auto checked) {
    constexpr bool kChecked = decltype(checked)::value;
    const uint32_t moved = sim.template do_step<kChecked>(a);
    const int ach = Sim::kCheapAchieved ? sim.achieved() : (moved ? sim.achieved() : prev);
    const bool term = ach == g;  // puzzle.py:409-411
    const int dg = ach - prev;
    prev = ach;
    // gym_env.py:212-221; the float64 value (dg - 0.01) was rounded to float32 once, on the host
    const uint32_t tbyte = term ? 1u : 0u;
    n_term += tbyte;
    n_moved += moved ? 1u : 0u;
    n_gain += (uint32_t)max(dg, 0);
    if (kChecked) n_invalid += a > 3u ? 1u : 0u;
    bool trunc = false;
    if constexpr (kEpisodes) {
      ++steps;
      trunc = p.max_steps > 0 && steps >= p.max_steps;  // gym_env.py:223
    }
    if (kPacked) {
      asm volatile("st.global.u8 [%0], %1;" ::"l"(term_base + (size_t)h_out * 2u),
                   "r"(record_byte(dg, term, trunc))
                   : "memory");
    } else {
      reward_base[(size_t)h_out * 2u] = term ? 10.0f : s_rlut[dg + kMaxObjects];
      asm volatile("st.global.u8 [%0], %1;" ::"l"(term_base + (size_t)h_out * 2u), "r"(tbyte) : "memory");
    }
    if constexpr (kEpisodes) {
      if (trunc_base)
        asm volatile("st.global.u8 [%0], %1;" ::"l"(trunc_base + (size_t)h_out * 2u), "r"(trunc ? 1u : 0u)
                     : "memory");
      if (term || trunc) {
        sim.resample(p, e);
        g = sim.goals();
        prev = sim.achieved();
        steps = 0;
      }
    }
    h_out += half_n;
  };
  constexpr int kBlockSteps = 32, kGroup = Sim::kGroup;
  const uint32_t lane = threadIdx.x & 31u;
  const uint8_t* const warp_actions = p.actions + (e - (int64_t)lane);
  const uint32_t tiles_s = (uint32_t)__cvta_generic_to_shared(tiles);
  auto issue_block = [&](const int block) {
    const int row = block * kBlockSteps + (int)lane;
    const uint8_t* src = warp_actions + (size_t)(row < T ? row : T - 1) * (size_t)p.N;  // clamped: never used
    const uint32_t dst = tiles_s + (uint32_t)(block & 1) * 1024u + lane * 32u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u), "l"(src + 16) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const int blocks = (T + kBlockSteps - 1) / kBlockSteps;
  issue_block(0);
  for (int blk = 0; blk < blocks; ++blk) {
    __syncwarp();
    issue_block(blk + 1);
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncwarp();
    const uint8_t* tile = tiles + (blk & 1) * 1024;
    const uint4 w0 = *reinterpret_cast<const uint4*>(tile + lane * 32u);
    const uint4 w1 = *reinterpret_cast<const uint4*>(tile + lane * 32u + 16u);
    const uint32_t seen = w0.x | w0.y | w0.z | w0.w | w1.x | w1.y | w1.z | w1.w;
    const bool any_invalid = __any_sync(0xffffffffu, (seen & 0xfcfcfcfcu) != 0u);
    const int steps = T - blk * kBlockSteps < kBlockSteps ? T - blk * kBlockSteps : kBlockSteps;
    const uint8_t* mine = tile + lane;
    if (!any_invalid && steps == kBlockSteps) {
#pragma unroll 1
      for (int gi = 0; gi < kBlockSteps / kGroup; ++gi, mine += kGroup * 32) {
#pragma unroll
        for (int k = 0; k < kGroup; ++k) one_step(mine[k * 32], std::false_type{});
      }
    } else {
#pragma unroll 1
      for (int k = 0; k < steps; ++k, mine += 32) one_step(mine[0], std::true_type{});
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  c.terminated += n_term;
  c.moved += n_moved;
  c.dgoals += n_gain;
  c.invalid += n_invalid;
  sim.store(p, e);
  if (kEpisodes) p.steps[e] = steps;
}

// boards with the edge-mask rule; `acts` = this thread's per-action constants in shared memory,
// action a at acts[a * kBlock]
template <int NMAX, typename Word>
struct BoardLeanSim {
  static constexpr bool kCheapAchieved = true;
  static constexpr int kGroup = 4;
  Board<NMAX, Word>& s;
  const BoardAction<Word>* acts;
  template <bool kChecked>
  __device__ __forceinline__ uint32_t do_step(uint32_t a) {
    return board_step(s, (int)a, acts[(kChecked ? (a & 3u) : a) * kBlock]);
  }
  // branch-free (goal[k >= g] is all ones and never matches): the goal count differs per thread
  // in a mixed batch, so the early exit of board_achieved would diverge
  __device__ __forceinline__ int achieved() const {
    int count = 0;
#pragma unroll
    for (int k = 0; k + 1 < NMAX; ++k) count += (s.b[k + 1] == s.goal[k]) ? 1 : 0;
    return count;
  }
  __device__ __forceinline__ int goals() const { return s.g; }
  __device__ __forceinline__ void store(const StepParams& p, int64_t e) {
#pragma unroll
    for (int k = 0; k < NMAX; ++k)
      if (k < s.n) {
        int x, y;
        board_position<Word>(s.b[k], s.P, &x, &y);
        p.x[k * p.stride + e] = x;
        p.y[k * p.stride + e] = y;
      }
  }
};

// ... with per-thread puzzles that change at episode boundaries (pw_rollout_episodes)
template <int NMAX, typename Word>
struct BoardEpisodeSim : BoardLeanSim<NMAX, Word> {
  BoardAction<Word>* my_acts;  // the same shared-memory slots, writable
  __device__ __forceinline__ BoardEpisodeSim(Board<NMAX, Word>& board, BoardAction<Word>* a)
      : BoardLeanSim<NMAX, Word>{board, a}, my_acts(a) {}
  __device__ __forceinline__ void resample(const StepParams& p, int64_t e) {
    const uint32_t episode = p.episode[e] + 1u;
    p.episode[e] = episode;
    const int pid = draw_puzzle(p.seed, e, episode, p.num_puzzles);
    p.puzzle_id_mut[e] = pid;
    const BoardStaticPod* q = p.board_static + pid;
    Board<NMAX, Word>& s = this->s;
    s.wall = (Word)__ldg(&q->wall);
    s.awall = (Word)__ldg(&q->awall);
    s.edge_up = (Word)__ldg(&q->edge_up);
    s.edge_down = (Word)__ldg(&q->edge_down);
    s.P = __ldg(&q->P);
    s.n = __ldg(&q->n);
    s.g = __ldg(&q->g);
#pragma unroll
    for (int k = 0; k < NMAX; ++k) {
      s.goal[k] = (Word)__ldg(&q->goal[k]);
      s.b[k] = (Word)__ldg(&q->init[k]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) my_acts[a * kBlock] = board_action_of(s, (uint32_t)a);
  }
};

template <int NMAX, typename Word>
__device__ __forceinline__ void run_board_lean(Board<NMAX, Word>& s, const StepParams& p, const int64_t e,
                                               const float* s_rlut, uint8_t* tiles,
                                               const BoardAction<Word>* acts, StepCounts& c) {
  BoardLeanSim<NMAX, Word> sim{s, acts};
  run_lean_sim_any<false>(sim, p, e, s_rlut, tiles, c);
}

// the general rule (collision bit tables) on an Env policy; the (large) rule is not unrolled
template <class Env>
struct GeneralLeanSim {
  static constexpr bool kCheapAchieved = false;
  static constexpr int kGroup = 1;
  Env& env;
  template <bool kChecked>
  __device__ __forceinline__ uint32_t do_step(uint32_t a) { return step_tab(env, (int)a); }
  __device__ __forceinline__ int achieved() const { return achieved_goals(env); }
  __device__ __forceinline__ int goals() const { return env.g(); }
  __device__ __forceinline__ void store(const StepParams& p, int64_t e) { env.store(p.x, p.y, p.stride, e); }
};

__device__ __noinline__ void general_slow_path_global(const uint64_t* table, const StepParams& p,
                                                      int64_t e, StepCounts& c) {
  LocalEnv<GlobalTable> env;
  env.init(GlobalTable{table});
  run_steps_thread<false>(env, p, e, c);
}

template <int NMAX, typename Word, int MIN_BLOCKS, bool kLean>
__global__ void __launch_bounds__(kBlock, MIN_BLOCKS)
    k_steps_board_mixed(const __grid_constant__ StepParams p) {
  __shared__ uint32_t s_counters[4];
  // lean launches only (see run_board_lean): per-warp action tiles, per-thread action constants and a
  // shared-memory copy of the reward table
  __shared__ __align__(16) uint8_t s_tiles[kBlock / 32][2][32 * 32];
  __shared__ BoardAction<Word> s_act[4][kBlock];
  __shared__ float s_rlut[2 * kMaxObjects + 1];
  for (int i = threadIdx.x; i < 2 * kMaxObjects + 1; i += blockDim.x) s_rlut[i] = c_reward_lut[i];
  if (threadIdx.x < 4) s_counters[threadIdx.x] = 0;
  __syncthreads();
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  StepCounts c;
  if (e < p.N) {
    const int pid = p.puzzle_id ? checked_puzzle_id(p.puzzle_id[e], p.num_puzzles, p.status) : p.puzzle_uniform;
    const GlobalTable tbl{p.blob + p.puzzle_off[pid]};
    Board<NMAX, Word> s;
    BoardSimMixed<NMAX, Word> sim{s, nullptr, 0u, 0};
    sim.load_static(p.board_static + pid);
    const bool regular = board_load(s, tbl, [&](int k) { return p.x[k * p.stride + e]; },
                                    [&](int k) { return p.y[k * p.stride + e]; });
    // episode handling re-points the thread at other puzzles: stay on boards (all eligible)
    if (kLean && p.lean_tiles) {
      // plain rollout with whole warps (N % 32 == 0): action tiles + FMA-pipe addressing; a warp
      // with any irregular state takes the general rule for all its lanes
      if (__all_sync(0xffffffffu, regular)) {
#pragma unroll
        for (int a = 0; a < 4; ++a) s_act[a][threadIdx.x] = board_action_of(s, (uint32_t)a);
        run_board_lean<NMAX, Word>(s, p, e, s_rlut, s_tiles[threadIdx.x >> 5][0], &s_act[0][threadIdx.x], c);
      } else {
        general_slow_path_global(tbl.base, p, e, c);
      }
    } else if (!kLean && p.episode_tiles && __all_sync(0xffffffffu, regular)) {
      // pw_rollout_episodes made of whole warps: the plain-rollout loop with episode handling
#pragma unroll
      for (int a = 0; a < 4; ++a) s_act[a][threadIdx.x] = board_action_of(s, (uint32_t)a);
      BoardEpisodeSim<NMAX, Word> esim(s, &s_act[0][threadIdx.x]);
      run_lean_sim_any<true>(esim, p, e, s_rlut, s_tiles[threadIdx.x >> 5][0], c);
    } else if (regular) {
      run_steps_sim<kLean>(sim, p, e, c);
    } else if (!(p.flags & kFlagResample)) {
      general_slow_path_global(tbl.base, p, e, c);
    } else {
      // an irregular state under episode handling: general rule with re-pointable tables
      LocalEnv<GlobalTable> env;
      env.init(tbl);
      run_steps_thread<false>(env, p, e, c);
    }
  }
  flush_counts(c, p, s_counters);
}

// Per-environment puzzle ids: tables are read from global memory (L1/L2 resident, the whole
// benchmark set is < 1 MB).
template <int NMAX, bool kLean>
__global__ void __launch_bounds__(kBlock)
    k_steps_mixed(const __grid_constant__ StepParams p) {
  extern __shared__ uint64_t smem[];
  __shared__ uint32_t s_counters[4];
  if (threadIdx.x < 4) s_counters[threadIdx.x] = 0;
  __syncthreads();
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = e < p.N;
  const int pid = (active && p.puzzle_id) ? checked_puzzle_id(p.puzzle_id[e], p.num_puzzles, p.status) : p.puzzle_uniform;
  const GlobalTable table{p.blob + p.puzzle_off[pid]};
  if constexpr (NMAX > 0) {
    RegEnv<NMAX, GlobalTable> env;
    env.init(table);
    run_steps<kLean>(env, p, e, active, s_counters);
  } else {
    SmemEnv<GlobalTable> env;
    int* st = reinterpret_cast<int*>(smem);
    env.init(table, st + threadIdx.x, st + p.n_max * kBlock + threadIdx.x, kBlock);
    run_steps<kLean>(env, p, e, active, s_counters);
  }
}

// ------------------------------------------------------------------------------------------
// Successor expansion: 4 actions per state; a step is undone by translating the pushed set
// back, so the state is loaded once.
// ------------------------------------------------------------------------------------------
template <class Env>
__device__ __forceinline__ void run_expand(Env& env, const ExpandParams& p, int64_t e) {
  const unsigned live = __activemask();  // the warp's lanes that have a state (converged here)
  {
    const int64_t src = p.index ? p.index[e] : e;
    if (p.stride_in * (int64_t)p.n_max < ((int64_t)1 << 31)) env.load32(p.x_in, p.y_in, (uint32_t)p.stride_in, (uint32_t)src);
    else env.load(p.x_in, p.y_in, p.stride_in, src);
  }
  const int g = env.g(), n = env.n();
  const int jend = Env::kStaticN ? Env::kStaticN : n;
  // The state itself is never modified: successor a is the state with the pushed set displaced,
  // written straight to its column.  Offset = element index of (object j, successor s); with
  // 32-bit offsets (every launch whose output planes hold < 2^31 elements) an address is one
  // IMAD.WIDE on the FMA pipe instead of an add-with-carry pair on the ALU pipe.
  // what the agent pushes directly, for all four actions at once (one descriptor load and window
  // test per object instead of four): bit 4 * j + a <=> the agent, moved by a, pushes object j
  // (register policies: n <= 8, 4 bits per object in one 32-bit word)
  constexpr bool use_agent4 = Env::kStaticN > 0 && Env::kStaticN <= 8;
  uint32_t agent4 = 0;
  if constexpr (use_agent4) {
    env.pack_positions();  // the state is never modified here: dynamic-index reads become one byte permute
    const int x0 = env.x(0), y0 = env.y(0);
#pragma unroll
    for (int j = 1; j < Env::kStaticN; ++j) {
      if (j >= n) break;
      agent4 |= tab_pushes4(env, 1u + (uint32_t)j, env.x(j) - x0, env.y(j) - y0) << (4 * j);
    }
  }
  // bit 4 j + a of agent4 -> bit j: the nibbles' selected bits gathered by three shift-or-mask rounds
  auto agent_hits_of = [&](const int a) {
    uint32_t t = (agent4 >> a) & 0x11111111u;
    t = (t | (t >> 3)) & 0x03030303u;
    t = (t | (t >> 6)) & 0x000f000fu;
    return (t | (t >> 12)) & 0xffu;
  };
  // Register policies (n <= 8): the four pushed sets as the four bytes of one word.  An action whose agent
  // is stopped by a wall, or pushes nothing, is decided at once; the others (a few per cent of the lanes)
  // are collected and their closures run in ONE loop headed by a warp vote, so that the closure code is
  // executed max-over-lanes(pending actions) times per warp instead of once per action that any lane
  // needs it for.
  uint32_t moved4 = 0;
  if constexpr (use_agent4) {
    const uint32_t nn1 = 1u + (uint32_t)(n * n);
    uint32_t pending = 0;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      if (!tab_wall(env, nn1, 0u, (uint32_t)a, env.x(0), env.y(0))) {  // puzzle.py:353-354
        if (agent_hits_of(a) == 0u) moved4 |= 1u << (8 * a);
        else pending |= 1u << a;
      }
    }
    while (__any_sync(live, pending != 0u)) {
      if (pending != 0u) {
        const int a = __ffs((int)pending) - 1;
        pending &= pending - 1;
        moved4 |= pushed_set_tab(env, a, true, agent_hits_of(a)) << (8 * a);
      }
    }
  }
  auto expand = [&](auto first_offset, auto plane_stride) {
    using Offset = decltype(first_offset);
#pragma unroll 1
    for (int a = 0; a < 4; ++a) {
      uint32_t moved;
      if constexpr (use_agent4) {
        moved = (moved4 >> (8 * a)) & 0xffu;
      } else {
        moved = pushed_set_tab(env, a);
        // the closure's data-dependent loops leave the lanes apart: reconverge before the stores
        __syncwarp(live);
      }
      const int dx = moved ? action_dx(a) : 0, dy = moved ? action_dy(a) : 0;
      const Offset s = (Offset)a * (Offset)p.N + first_offset;
      Offset off = s;
#pragma unroll
      for (int j = 0; j < jend; ++j) {
        if (Env::kStaticN && j >= n) break;
        const bool m = (moved >> j) & 1u;
        p.x_out[off] = env.x(j) + (m ? dx : 0);
        p.y_out[off] = env.y(j) + (m ? dy : 0);
        off += plane_stride;
      }
      // goal objects are objects 1..g: a short loop with a uniform exit instead of a test on every object
      int at_goal = 0;
      auto goal_test = [&](const int j) {
        const bool m = (moved >> j) & 1u;
        const int x = env.x(j) + (m ? dx : 0), y = env.y(j) + (m ? dy : 0);
        at_goal += (x == env.goal_x(j - 1) && y == env.goal_y(j - 1)) ? 1 : 0;
      };
      if constexpr (Env::kStaticN > 0) {
#pragma unroll
        for (int j = 1; j < Env::kStaticN; ++j) {
          if (j > g) break;
          goal_test(j);
        }
      } else {
        for (int j = 1; j <= g; ++j) goal_test(j);
      }
      if (p.moved) p.moved[s] = moved;
      if (p.goal) p.goal[s] = at_goal == g ? 1 : 0;  // pushworld_puzzle.cc:462-469
    }
  };
  if (p.stride_out * (int64_t)p.n_max < ((int64_t)1 << 31) && 4 * p.N <= p.stride_out)
    expand((uint32_t)e, (uint32_t)p.stride_out);
  else
    expand((int64_t)e, (int64_t)p.stride_out);
}

template <int NMAX>
__global__ void __launch_bounds__(kBlock) k_expand_uniform(const ExpandParams p) {
  extern __shared__ uint64_t smem[];
  const uint64_t* src = p.blob + p.puzzle_off[p.puzzle_uniform];
  stage_table(smem, src, p.uniform_u64);
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= p.N) return;
  if constexpr (NMAX > 0) {
    RegEnv<NMAX, SharedTable> env;
    env.init(SharedTable{smem});
    run_expand(env, p, e);
  } else {
    SmemEnv<SharedTable> env;
    int* st = reinterpret_cast<int*>(smem + p.uniform_u64);
    env.init(SharedTable{smem}, st + threadIdx.x, st + p.n_max * kBlock + threadIdx.x, kBlock);
    run_expand(env, p, e);
  }
}

template <int NMAX>
__global__ void __launch_bounds__(kBlock) k_expand_mixed(const ExpandParams p) {
  extern __shared__ uint64_t smem[];
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= p.N) return;
  const GlobalTable table{
      p.blob + p.puzzle_off[p.puzzle_id ? checked_puzzle_id(p.puzzle_id[e], p.num_puzzles, p.status) : p.puzzle_uniform]};
  if constexpr (NMAX > 0) {
    RegEnv<NMAX, GlobalTable> env;
    env.init(table);
    run_expand(env, p, e);
  } else {
    SmemEnv<GlobalTable> env;
    int* st = reinterpret_cast<int*>(smem);
    env.init(table, st + threadIdx.x, st + p.n_max * kBlock + threadIdx.x, kBlock);
    run_expand(env, p, e);
  }
}

// ------------------------------------------------------------------------------------------
// reset: state <- initial_state (gym_env.py:171-176)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_reset(const ResetParams p) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= p.N) return;
  if (p.reset_mask && p.reset_mask[e] == 0) return;
  const int pid = p.puzzle_id ? checked_puzzle_id(p.puzzle_id[e], p.num_puzzles, p.status) : p.puzzle_uniform;
  const GlobalTable t{p.blob + p.puzzle_off[pid]};
  const int n = (int)(t.word(kHdrCounts) & 0xffffu);
  const int desc_word = (int)t.word(kHdrObjOff) * 2;
  for (int j = 0; j < p.n_max; ++j) {
    int vx = 0, vy = 0;
    if (j < n) {
      const uint32_t v = t.word(desc_word + 4 * j + 1);
      vx = lo16(v);
      vy = hi16(v);
    }
    p.x[j * p.stride + e] = vx;
    p.y[j * p.stride + e] = vy;
  }
  if (p.steps) p.steps[e] = 0;
}

// count_achieved_goals / is_goal_state of the current states (puzzle.py:396-411)
__global__ void __launch_bounds__(kBlock) k_goals(const GoalParams p) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= p.N) return;
  const int pid = p.puzzle_id ? checked_puzzle_id(p.puzzle_id[e], p.num_puzzles, p.status) : p.puzzle_uniform;
  const GlobalTable t{p.blob + p.puzzle_off[pid]};
  const int g = (int)(t.word(kHdrCounts) >> 16);
  const int desc_word = (int)t.word(kHdrObjOff) * 2;
  int count = 0;
  for (int k = 0; k < g; ++k) {
    const uint32_t goal = t.word(desc_word + 4 * (k + 1) + 2);
    count += (p.x[(k + 1) * p.stride + e] == lo16(goal) && p.y[(k + 1) * p.stride + e] == hi16(goal));
  }
  if (p.achieved) p.achieved[e] = count;
  if (p.is_goal) p.is_goal[e] = count == g ? 1 : 0;
}

// ------------------------------------------------------------------------------------------
// visited set: open addressing, entry = fingerprint24 << 40 | (state id + 1); exact compare
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t mix64(uint64_t h) {  // splitmix64 finaliser
  h ^= h >> 30;
  h *= 0xbf58476d1ce4e5b9ull;
  h ^= h >> 27;
  h *= 0x94d049bb133111ebull;
  h ^= h >> 31;
  return h;
}

__global__ void __launch_bounds__(kBlock) k_visited_insert(const VisitedParams p) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.count) return;
  if (p.valid && p.valid[i] == 0) {
    p.is_new[i] = 0;
    return;
  }
  const GlobalTable t{p.blob + p.puzzle_off[p.puzzle]};
  const int n = (int)(t.word(kHdrCounts) & 0xffffu);
  const int64_t id = p.first_id + i;
  uint64_t h = 0x9e3779b97f4a7c15ull;
  for (int j = 0; j < n; ++j) {
    const uint64_t v = ((uint64_t)(uint32_t)p.x[j * p.stride + id] << 32) | (uint32_t)p.y[j * p.stride + id];
    h = mix64(h ^ v) + 0x9e3779b97f4a7c15ull;
  }
  const uint64_t fp = (h >> 40) << 40;
  const unsigned long long mine = fp | (unsigned long long)(id + 1);
  const uint64_t mask = (uint64_t)p.capacity - 1;
  uint64_t slot = mix64(h) & mask;
  uint8_t is_new = 0;
  bool done = false;
  for (int64_t probes = 0; probes < p.capacity && !done; ++probes) {
    unsigned long long cur = p.table[slot];
    if (cur == 0ull) {
      cur = atomicCAS(&p.table[slot], 0ull, mine);
      if (cur == 0ull) {
        is_new = 1;
        done = true;
        break;
      }
    }
    if ((cur >> 40) == (mine >> 40)) {  // same fingerprint: compare the states themselves
      const int64_t other = (int64_t)(cur & ((1ull << 40) - 1ull)) - 1;
      bool same = true;
      for (int j = 0; j < n && same; ++j)
        same = p.x[j * p.stride + id] == p.x[j * p.stride + other] &&
               p.y[j * p.stride + id] == p.y[j * p.stride + other];
      if (same) {
        done = true;
        break;
      }
    }
    slot = (slot + 1) & mask;
  }
  if (!done && p.status) atomicAdd(&p.status[1], 1ull);  // table full
  p.is_new[i] = is_new;
}

// ------------------------------------------------------------------------------------------
// novelty (cpp/src/heuristics/novelty.cc:30-77), batch semantics: test pass, then insert pass
// bit layout: singles [i * C + pos_i], then pairs [base + (pair(i<j) * C + pos_i) * C + pos_j]
// ------------------------------------------------------------------------------------------
struct NoveltyIndex {
  int n, W, H;
  uint64_t C, pair_base;
  __device__ __forceinline__ NoveltyIndex(const NoveltyParams& p)
      : n(p.n), W(p.width), H(p.height), C((uint64_t)p.width * p.height), pair_base((uint64_t)p.n * C) {}
  __device__ __forceinline__ int64_t pos(int x, int y) const {  // -1 outside the grid
    return (x < 0 || y < 0 || x >= W || y >= H) ? -1 : (int64_t)y * W + x;
  }
  __device__ __forceinline__ uint64_t single(int i, int64_t pi) const { return (uint64_t)i * C + pi; }
  __device__ __forceinline__ uint64_t pair(int i, int j, int64_t pi, int64_t pj) const {  // i < j
    const uint64_t k = (uint64_t)i * (2 * n - i - 1) / 2 + (uint64_t)(j - i - 1);
    return pair_base + (k * C + (uint64_t)pi) * C + (uint64_t)pj;
  }
};

template <bool kInsert>
__global__ void __launch_bounds__(kBlock) k_novelty(const NoveltyParams p) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= p.count) return;
  if (p.valid && p.valid[s] == 0) {
    if (!kInsert) p.novelty[s] = 3;
    return;
  }
  const NoveltyIndex ix(p);
  const int64_t id = p.first_id + s;
  uint32_t moved = p.moved[s];
  int novelty = 3;
  while (moved) {
    const int i = __ffs((int)moved) - 1;
    moved &= moved - 1;
    if (i >= p.n) break;
    const int64_t pi = ix.pos(p.x[i * p.stride + id], p.y[i * p.stride + id]);
    if (pi < 0) continue;
    {
      const uint64_t bit = ix.single(i, pi);
      if (kInsert) {
        atomicOr(&p.seen[bit >> 5], 1u << (bit & 31));
      } else if (!((p.seen[bit >> 5] >> (bit & 31)) & 1u)) {
        novelty = 1;
      }
    }
    if (!kInsert && novelty == 1) continue;  // cannot get better; pairs are inserted by the other pass
    for (int j = 0; j < p.n; ++j) {
      if (j == i) continue;
      const int64_t pj = ix.pos(p.x[j * p.stride + id], p.y[j * p.stride + id]);
      if (pj < 0) continue;
      const uint64_t bit = i < j ? ix.pair(i, j, pi, pj) : ix.pair(j, i, pj, pi);
      if (kInsert) {
        atomicOr(&p.seen[bit >> 5], 1u << (bit & 31));
      } else if (novelty > 2 && !((p.seen[bit >> 5] >> (bit & 31)) & 1u)) {
        novelty = 2;
      }
    }
  }
  if (!kInsert) p.novelty[s] = (uint8_t)novelty;
}

inline unsigned grid_for(int64_t n) { return (unsigned)((n + kBlock - 1) / kBlock); }

}  // namespace

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
cudaError_t launch_steps(const StepParams& p_in, int n_limit, cudaStream_t stream) {
  if (p_in.N == 0 || p_in.T == 0) return cudaSuccess;
  StepParams p = p_in;
  // lean launches made of whole warps with 16-byte aligned actions can stage actions with cp.async
  const bool whole_warps = (p.N & 31) == 0 && (reinterpret_cast<uintptr_t>(p.actions) & 15u) == 0;
  p.lean_tiles = is_lean(p) && whole_warps;
  // pw_rollout_episodes in the shape an RL loop uses: reward + terminated (+ truncated) per step
  const bool plain_out = p.reward && p.terminated && !p.record;
  const bool packed_out = p.record && !p.reward && !p.terminated && !p.truncated;
  p.episode_tiles = whole_warps && p.flags == (kFlagAutoreset | kFlagResample) && (plain_out || packed_out) &&
                    !p.dgoals && !p.moved && !p.return_sum && p.steps && p.puzzle_id_mut && p.episode &&
                    (int64_t)p.T * p.N < ((int64_t)1 << 32);
  // a small puzzle with a reachable-state transition table: the rollout is table look-ups (pw_table.cu)
  if (table_rollout_eligible(p)) return launch_rollout_table(p, stream);
  const unsigned grid = grid_for(p.N);
  const size_t state_bytes = (size_t)p.n_max * 2 * kBlock * sizeof(int);
  const bool lean = is_lean(p);
  if (p.puzzle_id != nullptr && p.board_bits <= 64 && n_limit <= 8) {
    // every puzzle of the engine is board-eligible: per-thread boards (also under episode handling)
    const bool w32 = p.board_bits <= 32;
#define PW_BOARD_MIXED(NB, MINB)                                                                  \
  do {                                                                                            \
    if (w32 && lean)                                                                              \
      k_steps_board_mixed<NB, uint32_t, MINB, true><<<grid, kBlock, 0, stream>>>(p);              \
    else if (w32)                                                                                 \
      k_steps_board_mixed<NB, uint32_t, MINB, false><<<grid, kBlock, 0, stream>>>(p);             \
    else if (lean)                                                                                \
      k_steps_board_mixed<NB, uint64_t, MINB, true><<<grid, kBlock, 0, stream>>>(p);              \
    else                                                                                          \
      k_steps_board_mixed<NB, uint64_t, MINB, false><<<grid, kBlock, 0, stream>>>(p);             \
  } while (0)
    if (n_limit <= 3) PW_BOARD_MIXED(3, 8);
    else if (n_limit <= 5) PW_BOARD_MIXED(5, 8);
    else PW_BOARD_MIXED(8, 6);
#undef PW_BOARD_MIXED
    return cudaGetLastError();
  }
  // episode handling (autoreset / re-sampling) on the general rule: the mask kernel of pw_occ.cu
  // (a single board-eligible puzzle keeps its bitboard kernel)
  {
    const bool board_uniform = p.puzzle_id == nullptr && p.board_bits <= 64 && n_limit <= 8 &&
                               (size_t)p.uniform_u64 * 8 <= kStageLimit;
    if (getenv("PW_NO_OCC") == nullptr && !board_uniform && occ_episodes_eligible(p))
      return launch_rollout_occ(p, stream);
  }
  if (p.flags & kFlagResample) {  // puzzles change mid-rollout: dynamic-n policy, global tables
    k_steps_mixed<0, false><<<grid, kBlock, state_bytes, stream>>>(p);
    return cudaGetLastError();
  }
  // general rule, plain / packed rollouts made of whole warps: occupancy reject + batched closure
  // (pw_occ.cu); PW_NO_OCC=1 keeps the scan-every-step kernels below (A/B measurements)
  {
    const bool no_occ = getenv("PW_NO_OCC") != nullptr;
    const bool board_uniform = p.puzzle_id == nullptr && p.board_bits <= 64 && n_limit <= 8 &&
                               (size_t)p.uniform_u64 * 8 <= kStageLimit;
    if (!no_occ && !board_uniform && occ_rollout_eligible(p)) return launch_rollout_occ(p, stream);
  }
#define PW_GO(kernel, smem_bytes)                                         \
  do {                                                                    \
    if (lean)                                                             \
      kernel, true><<<grid, kBlock, smem_bytes, stream>>>(p);             \
    else                                                                  \
      kernel, false><<<grid, kBlock, smem_bytes, stream>>>(p);            \
  } while (0)
  // a single puzzle whose tables fit in shared memory is staged there; a larger one is read
  // through the global-table kernels like a mixed batch
  const size_t tbl = (size_t)p.uniform_u64 * 8;
  if (p.puzzle_id == nullptr && tbl + (n_limit > 8 ? state_bytes : 0) <= kStageLimit) {
    if (p.board_bits <= 64 && n_limit <= 8) {
      // bitboard kernels: uint32 / uint64 boards; lean rollouts get exact-n instantiations
      const bool w32 = p.board_bits <= 32;
#define PW_BOARD(NB, MINB)                                                                      \
  do {                                                                                          \
    if (w32)                                                                                    \
      k_steps_board<NB, uint32_t, MINB, true><<<grid, kBlock, tbl, stream>>>(p);                \
    else                                                                                        \
      k_steps_board<NB, uint64_t, MINB, true><<<grid, kBlock, tbl, stream>>>(p);                \
  } while (0)
      if (p.lean_tiles && p.ring_goals > 0 && n_limit >= 2) {
        // ring boards: one rotate per object, guard zone instead of edge rows
#define PW_RING(NB, MINB)                                                                       \
  do {                                                                                          \
    if (p.ring_goals == 1)                                                                      \
      k_rollout_ring<NB, true, MINB><<<grid, kBlock, tbl, stream>>>(p);                         \
    else                                                                                        \
      k_rollout_ring<NB, false, MINB><<<grid, kBlock, tbl, stream>>>(p);                        \
  } while (0)
        switch (n_limit) {
          case 2: PW_RING(2, 8); break;
          case 3: PW_RING(3, 8); break;
          case 4: PW_RING(4, 8); break;
          case 5:
          case 6: PW_RING(6, 8); break;
          default: PW_RING(8, 6); break;
        }
#undef PW_RING
      } else if (lean) {
        switch (n_limit) {
          case 1:
          case 2: PW_BOARD(2, 8); break;
          case 3: PW_BOARD(3, 8); break;
          case 4: PW_BOARD(4, 8); break;
          case 5: PW_BOARD(5, 8); break;
          case 6: PW_BOARD(6, 8); break;
          case 7: PW_BOARD(7, 6); break;
          default: PW_BOARD(8, 6); break;
        }
      } else if (n_limit <= 4) {
        if (w32) k_steps_board<4, uint32_t, 8, false><<<grid, kBlock, tbl, stream>>>(p);
        else k_steps_board<4, uint64_t, 8, false><<<grid, kBlock, tbl, stream>>>(p);
      } else {
        if (w32) k_steps_board<8, uint32_t, 6, false><<<grid, kBlock, tbl, stream>>>(p);
        else k_steps_board<8, uint64_t, 6, false><<<grid, kBlock, tbl, stream>>>(p);
      }
#undef PW_BOARD
    } else if (n_limit <= 4) {
      PW_GO(k_steps_uniform<4, tbl);
    } else if (n_limit <= 8) {
      PW_GO(k_steps_uniform<8, tbl);
    } else {
      PW_GO(k_steps_uniform<0, tbl + state_bytes);
    }
  } else {
    if (n_limit <= 4) {
      PW_GO(k_steps_mixed<4, 0);
    } else if (n_limit <= 8) {
      PW_GO(k_steps_mixed<8, 0);
    } else {
      PW_GO(k_steps_mixed<0, state_bytes);
    }
  }
#undef PW_GO
  return cudaGetLastError();
}

cudaError_t launch_expand(const ExpandParams& p, int n_limit, cudaStream_t stream) {
  if (p.N == 0) return cudaSuccess;
  const unsigned grid = grid_for(p.N);
  const size_t state_bytes = (size_t)p.n_max * 2 * kBlock * sizeof(int);
  const size_t tbl = (size_t)p.uniform_u64 * 8;
  if (p.puzzle_id == nullptr && tbl + (n_limit > 8 ? state_bytes : 0) <= kStageLimit) {
    if (n_limit <= 4) {
      k_expand_uniform<4><<<grid, kBlock, tbl, stream>>>(p);
    } else if (n_limit <= 8) {
      k_expand_uniform<8><<<grid, kBlock, tbl, stream>>>(p);
    } else {
      k_expand_uniform<0><<<grid, kBlock, tbl + state_bytes, stream>>>(p);
    }
  } else {
    if (n_limit <= 4) {
      k_expand_mixed<4><<<grid, kBlock, 0, stream>>>(p);
    } else if (n_limit <= 8) {
      k_expand_mixed<8><<<grid, kBlock, 0, stream>>>(p);
    } else {
      k_expand_mixed<0><<<grid, kBlock, state_bytes, stream>>>(p);
    }
  }
  return cudaGetLastError();
}

cudaError_t launch_reset(const ResetParams& p, cudaStream_t stream) {
  if (p.N == 0) return cudaSuccess;
  k_reset<<<grid_for(p.N), kBlock, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_goals(const GoalParams& p, cudaStream_t stream) {
  if (p.N == 0) return cudaSuccess;
  k_goals<<<grid_for(p.N), kBlock, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_visited(const VisitedParams& p, cudaStream_t stream) {
  if (p.count == 0) return cudaSuccess;
  k_visited_insert<<<grid_for(p.count), kBlock, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_novelty(const NoveltyParams& p, cudaStream_t stream) {
  if (p.count == 0) return cudaSuccess;
  k_novelty<false><<<grid_for(p.count), kBlock, 0, stream>>>(p);
  k_novelty<true><<<grid_for(p.count), kBlock, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t configure_kernels(size_t max_dynamic_smem) {
  // dynamic shared memory above 48 KB needs an explicit opt-in per kernel; a puzzle whose tables
  // do not fit the limit is read through the global-table kernels instead (launch_steps /
  // launch_expand), so the request is capped rather than refused
  if (max_dynamic_smem > kStageLimit) max_dynamic_smem = kStageLimit;
  cudaError_t err = cudaSuccess;
#define PW_OPT_IN(k)                                                                             \
  if (err == cudaSuccess)                                                                        \
  err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_dynamic_smem)
  PW_OPT_IN((k_steps_uniform<0, true>));
  PW_OPT_IN((k_steps_uniform<0, false>));
  PW_OPT_IN((k_steps_uniform<4, true>));
  PW_OPT_IN((k_steps_uniform<4, false>));
  PW_OPT_IN((k_steps_uniform<8, true>));
  PW_OPT_IN((k_steps_uniform<8, false>));
  PW_OPT_IN((k_steps_mixed<0, true>));
  PW_OPT_IN((k_steps_mixed<0, false>));
  PW_OPT_IN(k_expand_uniform<0>);
  PW_OPT_IN(k_expand_uniform<4>);
  PW_OPT_IN(k_expand_uniform<8>);
  PW_OPT_IN(k_expand_mixed<0>);
#undef PW_OPT_IN
  return err;
}

}  // namespace pw
