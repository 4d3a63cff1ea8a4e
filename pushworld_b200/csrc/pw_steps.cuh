// pw_steps.cuh -- T fused environment steps of ONE environment by ONE thread (device only).
// Shared by the dynamics kernels (pw_kernels.cu) and the fused step+render kernel
// (pw_render.cu).  T = 1 is pw_step.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pw_core.cuh"
#include "pw_internal.h"

namespace pw {

// (float)((double)dg - 0.01) for dg in [-32, 32]; the library is built as ONE translation unit
// (pw_unity.cu), so this definition is unique.
__constant__ float c_reward_lut[2 * kMaxObjects + 1];

struct StepCounts {
  uint32_t terminated = 0, moved = 0, dgoals = 0, invalid = 0;
};

template <class Env>
__device__ __forceinline__ void run_steps_thread(Env& env, const StepParams& p, int64_t e,
                                                 StepCounts& c) {
  env.load(p.x, p.y, p.stride, e);
  int prev = achieved_goals(env);
  int steps = p.steps ? p.steps[e] : 0;
  float ret = 0.f;
  const int g = env.g();
  int a_next = p.actions[e];
  for (int t = 0; t < p.T; ++t) {
    const int64_t idx = (int64_t)t * p.N + e;
    const int a = a_next;
    if (t + 1 < p.T) a_next = p.actions[idx + p.N];  // prefetch: independent of the state
    const uint32_t moved = step(env, a);
    const int ach = moved ? achieved_goals(env) : prev;
    const bool term = ach == g;  // puzzle.py:409-411
    const int dg = ach - prev;
    prev = ach;
    ++steps;
    // gym_env.py:212-221; the float64 value (dg - 0.01) was rounded to float32 once, on the host
    const float r = term ? 10.0f : c_reward_lut[dg + kMaxObjects];
    const bool trunc = p.max_steps > 0 && steps >= p.max_steps;  // gym_env.py:223
    if (p.reward) p.reward[idx] = r;
    if (p.dgoals) p.dgoals[idx] = (int8_t)dg;
    if (p.terminated) p.terminated[idx] = term ? 1 : 0;
    if (p.truncated) p.truncated[idx] = trunc ? 1 : 0;
    if (p.moved) p.moved[idx] = moved;
    ret += r;
    c.terminated += term ? 1u : 0u;
    c.moved += moved ? 1u : 0u;
    c.dgoals += dg > 0 ? (uint32_t)dg : 0u;
    c.invalid += (unsigned)a > 3u ? 1u : 0u;
    if ((p.flags & kFlagAutoreset) && (term || trunc)) {
      env.load_initial();
      prev = achieved_goals(env);
      steps = 0;
    }
  }
  env.store(p.x, p.y, p.stride, e);
  if (p.steps) p.steps[e] = steps;
  if (p.return_sum) p.return_sum[e] += ret;
}

}  // namespace pw
