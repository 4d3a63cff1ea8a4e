// pw_steps.cuh -- T fused environment steps of ONE environment by ONE thread (device only).
// Shared by the dynamics kernels (pw_kernels.cu) and the fused step+render kernel
// (pw_render.cu).  T = 1 is pw_step.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pw_core.cuh"
#include "pw_fast.cuh"
#include "pw_internal.h"

namespace pw {

// (float)((double)dg - 0.01) for dg in [-32, 32]; one copy per translation unit that includes this
// header (uploaded by that file's upload function; pw_occ.cu computes the values on the device)
static __constant__ float c_reward_lut[2 * kMaxObjects + 1];

// The packed per-step record of pw_rollout_packed (include/pushworld_b200.h): everything
// PushWorldEnv.step returns besides the observation (gym_env.py:210-226) in one byte --
// bits 0..5 = delta_goals + 32, bit 6 = truncated, bit 7 = terminated.
__device__ __forceinline__ uint32_t record_byte(int dg, bool term, bool trunc) {
  return (uint32_t)(dg + kMaxObjects) | (trunc ? 0x40u : 0u) | (term ? 0x80u : 0u);
}

struct StepCounts {
  uint32_t terminated = 0, moved = 0, dgoals = 0, invalid = 0;
};

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------
// T fused steps of one environment per thread (pw_steps.cuh) + block-level counter flush.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void flush_counts(const StepCounts& c, const StepParams& p,
                                             uint32_t* s_counters, int envs_per_thread = 1) {
  // Launches without a counter vector (pw_step, pw_step_render) only report invalid actions, which
  // are rare: one block-wide vote decides whether anything has to be reduced at all.
  if (!p.counters && !__syncthreads_or(c.invalid != 0u)) return;
  if (p.counters || p.status) {
    const uint32_t c_term = warp_sum(c.terminated), c_moved = warp_sum(c.moved);
    const uint32_t c_dg = warp_sum(c.dgoals), c_invalid = warp_sum(c.invalid);
    if ((threadIdx.x & 31) == 0) {
      if (c_term) atomicAdd(&s_counters[0], c_term);
      if (c_moved) atomicAdd(&s_counters[1], c_moved);
      if (c_dg) atomicAdd(&s_counters[2], c_dg);
      if (c_invalid) atomicAdd(&s_counters[3], c_invalid);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      if (p.counters) {
        const int64_t per_cta = (int64_t)blockDim.x * envs_per_thread;
        const int64_t rest = p.N - (int64_t)blockIdx.x * per_cta;
        const int64_t live = rest < per_cta ? rest : per_cta;
        atomicAdd(&p.counters[0], (unsigned long long)(live > 0 ? live : 0) * (unsigned long long)p.T);
        if (s_counters[0]) atomicAdd(&p.counters[1], (unsigned long long)s_counters[0]);
        if (s_counters[1]) atomicAdd(&p.counters[2], (unsigned long long)s_counters[1]);
        if (s_counters[2]) atomicAdd(&p.counters[3], (unsigned long long)s_counters[2]);
      }
      if (p.status && s_counters[3]) atomicAdd(&p.status[0], (unsigned long long)s_counters[3]);
    }
  }
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t h) {
  h += 0x9e3779b97f4a7c15ull;
  h = (h ^ (h >> 30)) * 0xbf58476d1ce4e5b9ull;
  h = (h ^ (h >> 27)) * 0x94d049bb133111ebull;
  return h ^ (h >> 31);
}

// The puzzle an environment plays in its `episode`-th episode (pw_rollout_episodes).
__device__ __forceinline__ int draw_puzzle(uint64_t seed, int64_t env, uint32_t episode, int num_puzzles) {
  const uint64_t h = splitmix64(seed ^ ((uint64_t)env * 0x9e3779b97f4a7c15ull) ^
                                ((uint64_t)episode * 0xd1b54a32d192ed03ull));
  return (int)(((h >> 32) * (uint64_t)num_puzzles) >> 32);
}

// Adaptor: an Env policy (pw_env.cuh) driven by the general rule, read from the collision bit
// tables (pw_fast.cuh: step_tab == pw_core.cuh's step on every state).
template <class Env>
struct GeneralSim {
  Env& env;
  __device__ __forceinline__ void load(const StepParams& p, int64_t e) { env.load(p.x, p.y, p.stride, e); }
  __device__ __forceinline__ void store(const StepParams& p, int64_t e) { env.store(p.x, p.y, p.stride, e); }
  __device__ __forceinline__ void load_initial() { env.load_initial(); }
  __device__ __forceinline__ uint32_t do_step(int a) { return step_tab(env, a); }
  __device__ __forceinline__ int achieved() const { return achieved_goals(env); }
  __device__ __forceinline__ int goals() const { return env.g(); }
  // new episode on a freshly drawn puzzle (only table views that can be re-pointed support it)
  __device__ __forceinline__ void resample(const StepParams& p, int64_t e) {
    const uint32_t episode = p.episode[e] + 1u;
    p.episode[e] = episode;
    const int pid = draw_puzzle(p.seed, e, episode, p.num_puzzles);
    p.puzzle_id_mut[e] = pid;
    env.retarget(p.blob + p.puzzle_off[pid]);
    env.load_initial();
  }
  static constexpr bool kCheapAchieved = false;
  static constexpr int kPrefetch = 1;
};

// `Sim` provides load / store / load_initial / do_step / achieved / goals.
// kLean: the launch writes exactly reward[T][N] and terminated[T][N] (no step counter, no
// truncation, no autoreset, no dgoals / moved / return_sum) -- the shape of a plain rollout --
// so every optional output is compiled out of the loop.
template <bool kLean, class Sim>
__device__ __forceinline__ void run_steps_sim(Sim& sim, const StepParams& p, int64_t e,
                                              StepCounts& c) {
  int prev = sim.achieved();
  int steps = (!kLean && p.steps) ? p.steps[e] : 0;
  float ret = 0.f;
  int g = sim.goals();
  const int64_t N = p.N;
  const int T = p.T;
  // lean launches have T * N < 2^32 (is_lean), so one 32-bit element offset addresses all arrays
  const uint32_t n32 = (uint32_t)N;
  uint32_t off = (uint32_t)e;
  const uint8_t* __restrict__ actions = p.actions;
  float* __restrict__ reward = p.reward;
  uint8_t* __restrict__ terminated = p.terminated;
  uint8_t* __restrict__ record = p.record;
  // Actions are independent of the state, so a whole group of kPrefetch steps is fetched one
  // group ahead (double buffering in registers): the loads of group g+1 are issued before the
  // steps of group g run, which covers the DRAM latency even at ~8 warps per scheduler.
  constexpr int kPrefetch = Sim::kPrefetch;  // 1 for the (large) general rule: unrolling it thrashes the I-cache
  int ring[kPrefetch], nxt[kPrefetch];
  auto fetch = [&](int t) -> int {  // action of step t (clamped to the last step)
    const int tt = t < T ? t : T - 1;
    return kLean ? actions[(uint32_t)e + (uint32_t)tt * n32] : actions[(int64_t)tt * N + e];
  };
#pragma unroll
  for (int k = 0; k < kPrefetch; ++k) ring[k] = fetch(k);
  auto one_step = [&](const int t, const int k) {
    const int a = ring[k];
    const uint32_t moved = sim.do_step(a);
    const bool bad_action = (unsigned)a > 3u;
    const int ach = Sim::kCheapAchieved ? sim.achieved() : (moved ? sim.achieved() : prev);
    const bool term = ach == g;  // puzzle.py:409-411
    const int dg = ach - prev;
    prev = ach;
    // gym_env.py:212-221; the float64 value (dg - 0.01) was rounded to float32 once, on the host
    const float r = term ? 10.0f : c_reward_lut[dg + kMaxObjects];
    if (term) ++c.terminated;
    if (moved) ++c.moved;
    c.dgoals += (uint32_t)max(dg, 0);
    if (bad_action) ++c.invalid;
    if (kLean) {
      if (record) {  // packed variant of the plain rollout
        record[off] = (uint8_t)record_byte(dg, term, false);
      } else {
        reward[off] = r;
        terminated[off] = term ? 1 : 0;
      }
      off += n32;
    } else {
      const int64_t idx = (int64_t)t * N + e;
      ++steps;
      const bool trunc = p.max_steps > 0 && steps >= p.max_steps;  // gym_env.py:223
      if (p.reward) p.reward[idx] = r;
      if (p.dgoals) p.dgoals[idx] = (int8_t)dg;
      if (p.terminated) p.terminated[idx] = term ? 1 : 0;
      if (p.truncated) p.truncated[idx] = trunc ? 1 : 0;
      if (p.record) p.record[idx] = (uint8_t)record_byte(dg, term, trunc);
      if (p.moved) p.moved[idx] = moved;
      ret += r;
      if ((p.flags & kFlagAutoreset) && (term || trunc)) {
        if (p.flags & kFlagResample) {
          sim.resample(p, e);
          g = sim.goals();
        } else {
          sim.load_initial();
        }
        prev = sim.achieved();
        steps = 0;
      }
    }
  };
  // full groups of kPrefetch steps: straight-line, no exits, so the ring lives in fixed registers
  int t0 = 0;
  for (; t0 + kPrefetch <= T; t0 += kPrefetch) {
#pragma unroll
    for (int k = 0; k < kPrefetch; ++k) nxt[k] = fetch(t0 + kPrefetch + k);
#pragma unroll
    for (int k = 0; k < kPrefetch; ++k) one_step(t0 + k, k);
#pragma unroll
    for (int k = 0; k < kPrefetch; ++k) ring[k] = nxt[k];
  }
  // remainder (< kPrefetch steps): their actions are already in slots 0, 1, ...
#pragma unroll
  for (int k = 0; k + 1 < kPrefetch; ++k)
    if (t0 + k < T) one_step(t0 + k, k);
  sim.store(p, e);
  if (!kLean) {
    if (p.steps) p.steps[e] = steps;
    if (p.return_sum) p.return_sum[e] += ret;
  }
}

template <bool kLean, class Env>
__device__ __forceinline__ void run_steps_thread(Env& env, const StepParams& p, int64_t e,
                                                 StepCounts& c) {
  GeneralSim<Env> sim{env};
  sim.load(p, e);
  run_steps_sim<kLean>(sim, p, e, c);
}

// true when a launch has the "plain rollout" output shape (see run_steps_sim): reward + terminated
// per step, or -- the packed variant -- one record byte per step, and nothing else
inline bool is_lean(const StepParams& p) {
  const bool plain = p.reward && p.terminated && !p.record;
  const bool packed = p.record && !p.reward && !p.terminated;
  return (plain || packed) && !p.dgoals && !p.truncated && !p.moved && !p.steps &&
         !p.return_sum && p.flags == 0 && p.max_steps <= 0 &&
         (int64_t)p.T * p.N < ((int64_t)1 << 32);
}

}  // namespace pw
