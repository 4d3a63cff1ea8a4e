// pw_table.cu -- reachable-state transition tables: the rollout of a SMALL puzzle as table look-ups.
//
// A puzzle such as BASELINE config 2's `Road Block` (8 x 5 cells, 3 objects) has only a few thousand
// states reachable from its initial state (4,648).  At engine creation the engine enumerates them ON
// THE DEVICE with the exact transition rule of pw_fast.cuh (breadth first, one CTA, a hash set of
// packed states) and records, for every (state, action), the successor's index, the record byte of
// the step (delta_goals, terminated: what PushWorldEnv.step derives from the two states,
// gym_env.py:210-221) and whether anything moved.  The table (16 bytes per state) is staged in
// shared memory by the rollout kernel, and a step of one environment becomes
//     entry = table[state * 4 + action];  state = entry & 0xffff;  record = entry >> 16
// -- one shared-memory load on the dependent chain instead of rotates, board tests and predicate
// logic -- which makes the rollout what its algorithmic bytes say it should be: an HBM stream of
// 1 action byte in and reward f32 + terminated u8 (or 1 record byte) out per environment step.
// Exact by construction: every entry was computed by the same rule every other kernel uses, and an
// environment whose loaded state is not in the table (not reachable from the initial state) makes
// its warp run the general rule instead.
//
// Kernels, fastest first (launch_rollout_table picks the first that fits the launch):
//   k_rollout_pairs4  two-step table (state, a1, a2), FOUR environments per thread, packed-byte bookkeeping,
//                     16-byte reward stores: one-goal puzzles with <= 6,144 states, N % 128 == 0, aligned outputs
//   k_rollout_pairs   the same table, one environment per thread (any N % 32 == 0, any alignment)
//   k_rollout_table   one-step table: compact entries (one goal, <= 8,192 states) or the 32-bit entries above
// DESIGN.md section 3.2c has the measurements that led here (tools/table_probe.py, tools/probes/).
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>
#include <mutex>
#include <type_traits>
#include <vector>

#include "pw_env.cuh"
#include "pw_fast.cuh"
#include "pw_internal.h"
#include "pw_steps.cuh"

namespace pw {
namespace {

constexpr int kBuildThreads = 1024;
constexpr int kTableMaxBlock = 1024;  // threads (= environments) per CTA of the rollout: chosen per launch
constexpr int kTileSteps = 16;               // steps per action tile (a warp's 32 actions per step)
constexpr int kTileBytes = kTileSteps * 32;
constexpr int kPairMaxStates = 6144;  // two-step table: 32 bytes per state next to the action tiles of a 1024-thread CTA
constexpr int kCompactBlock = 896;  // largest CTA of the compact kernels: two of them share an SM (32 registers per thread)

__device__ __forceinline__ uint64_t mix_key(uint64_t h) {  // splitmix64 finaliser
  h ^= h >> 30;
  h *= 0xbf58476d1ce4e5b9ull;
  h ^= h >> 27;
  h *= 0x94d049bb133111ebull;
  h ^= h >> 31;
  return h;
}

// table entry: successor index (16 bits) | record byte << 16 | moved << 24 | goals gained << 25
__device__ __forceinline__ uint32_t make_entry(uint32_t next, uint32_t record, bool moved, int gained) {
  return next | (record << 16) | (moved ? 1u << 24 : 0u) | ((uint32_t)gained << 25);
}

// ------------------------------------------------------------------------------------------
// Enumeration of the reachable states (one CTA).  A state is packed as one cell index
// (y * W + x, `bits` bits) per object; slot = key + 1 in an open-addressing set (0 = empty).
// Each round expands a chunk of states in three phases separated by barriers: (1) successors are
// computed and their keys claimed in the hash set, (2) the threads that claimed a new key number
// it, (3) everybody reads the numbers of its four successors and writes its table row.
// meta[0] = number of states, meta[1] != 0: not tabulated (too many states, or a state that leaves
// the grid / does not fit the key).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBuildThreads)
    k_build_state_table(const uint64_t* blob, uint32_t puzzle_off, StateTableView tv, int max_states, int* meta) {
  __shared__ int s_count, s_lo, s_fail;
  LocalEnv<GlobalTable> env;
  env.init(GlobalTable{blob + puzzle_off});
  const int n = env.n(), g = env.g(), W = env.width(), H = env.height();
  const int bits = tv.key_bits;
  const uint64_t hmask = (uint64_t)tv.hash_cap - 1;
  auto pack = [&](const int* px, const int* py, bool* ok) {
    uint64_t key = 0;
    for (int j = 0; j < n; ++j) {
      if ((unsigned)px[j] >= (unsigned)W || (unsigned)py[j] >= (unsigned)H) *ok = false;
      key |= (uint64_t)(uint32_t)(py[j] * W + px[j]) << (j * bits);
    }
    return key;
  };
  if (threadIdx.x == 0) {
    env.load_initial();
    bool ok = true;
    const uint64_t key = pack(env.px, env.py, &ok);
    tv.keys[0] = key;
    uint64_t slot = mix_key(key) & hmask;
    tv.hash_keys[slot] = key + 1;
    tv.hash_vals[slot] = 0u;
    s_count = 1;
    s_lo = 0;
    s_fail = ok ? 0 : 1;
  }
  __syncthreads();
  while (true) {
    const int lo = s_lo, hi = s_count;
    if (lo >= hi || s_fail) break;
    __syncthreads();  // everybody has read the round's bounds
    for (int base = lo; base < hi; base += kBuildThreads) {
      const int s = base + (int)threadIdx.x;
      const bool active = s < hi;
      uint64_t succ_key[4];
      uint32_t succ_slot[4], succ_meta[4];
      bool winner[4] = {false, false, false, false};
      if (active) {
        const uint64_t key = __ldcg(tv.keys + s);  // (written by other threads of this CTA: read through L2)
        for (int j = 0; j < n; ++j) {
          const uint32_t c = (uint32_t)(key >> (j * bits)) & ((1u << bits) - 1u);
          env.px[j] = (int)(c % (uint32_t)W);
          env.py[j] = (int)(c / (uint32_t)W);
        }
        const int ach0 = achieved_goals(env);
        for (int a = 0; a < 4; ++a) {
          const uint32_t pushed = pushed_set_tab(env, a);
          const int dx = pushed ? action_dx(a) : 0, dy = pushed ? action_dy(a) : 0;
          int nx[kMaxObjects], ny[kMaxObjects];
          int ach = 0;
          for (int j = 0; j < n; ++j) {
            const bool m = (pushed >> j) & 1u;
            nx[j] = env.px[j] + (m ? dx : 0);
            ny[j] = env.py[j] + (m ? dy : 0);
            if (j >= 1 && j <= g) ach += (nx[j] == env.goal_x(j - 1) && ny[j] == env.goal_y(j - 1)) ? 1 : 0;
          }
          bool ok = true;
          const uint64_t k2 = pack(nx, ny, &ok);
          if (!ok) atomicExch(&s_fail, 1);
          const int dg = ach - ach0;
          succ_key[a] = k2;
          succ_meta[a] = make_entry(0u, record_byte(dg, ach == g, false), pushed != 0u, dg > 0 ? dg : 0);
          // claim or find the key
          uint64_t slot = mix_key(k2) & hmask;
          for (uint64_t probes = 0; probes <= hmask; ++probes) {
            unsigned long long cur = __ldcg(tv.hash_keys + slot);
            if (cur == 0ull) {
              cur = atomicCAS(&tv.hash_keys[slot], 0ull, (unsigned long long)(k2 + 1));
              if (cur == 0ull) {
                winner[a] = true;
                break;
              }
            }
            if (cur == k2 + 1) break;
            slot = (slot + 1) & hmask;
          }
          succ_slot[a] = (uint32_t)slot;
        }
      }
      __syncthreads();
      if (active) {
        for (int a = 0; a < 4; ++a)
          if (winner[a]) {
            const int idx = atomicAdd(&s_count, 1);
            if (idx >= max_states) {
              atomicExch(&s_fail, 1);
            } else {
              tv.keys[idx] = succ_key[a];
              tv.hash_vals[succ_slot[a]] = (uint32_t)idx;
            }
          }
      }
      __syncthreads();
      if (active && !s_fail)
        for (int a = 0; a < 4; ++a) tv.table[s * 4 + a] = succ_meta[a] | __ldcg(tv.hash_vals + succ_slot[a]);
      __syncthreads();
    }
    if (threadIdx.x == 0) s_lo = hi;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    meta[0] = s_count;
    meta[1] = s_fail;
  }
}

// ------------------------------------------------------------------------------------------
// Renumbering (one CTA): states whose goal objects all sit on their goals -- the terminal ones,
// puzzle.py:409-411 -- get the highest indices, so that "terminal" is one comparison with n_off.
// Writes the permuted keys / table rows (successor indices remapped) to keys2 / table2, remaps the
// hash values in place, meta[0] = n_off.  Order inside the two groups is kept.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBuildThreads)
    k_renumber_state_table(const uint64_t* blob, uint32_t puzzle_off, StateTableView tv, int n_states, uint64_t* keys2,
                           uint32_t* table2, uint32_t* map, int* meta) {
  __shared__ int s_on[kBuildThreads];  // terminal states in the chunks before this thread's
  __shared__ int s_total_on;
  LocalEnv<GlobalTable> env;
  env.init(GlobalTable{blob + puzzle_off});
  const int n = env.n(), g = env.g(), W = env.width();
  const int bits = tv.key_bits;
  auto terminal = [&](const uint64_t key) {
    int ach = 0;
    for (int j = 1; j <= g && j < n; ++j) {
      const uint32_t c = (uint32_t)(key >> (j * bits)) & ((1u << bits) - 1u);
      ach += ((int)(c % (uint32_t)W) == env.goal_x(j - 1) && (int)(c / (uint32_t)W) == env.goal_y(j - 1)) ? 1 : 0;
    }
    return ach == g;
  };
  const int chunk = (n_states + kBuildThreads - 1) / kBuildThreads;
  const int lo = min(n_states, (int)threadIdx.x * chunk), hi = min(n_states, lo + chunk);
  int mine = 0;
  for (int s = lo; s < hi; ++s) mine += terminal(tv.keys[s]) ? 1 : 0;
  s_on[threadIdx.x] = mine;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int i = 0; i < kBuildThreads; ++i) {
      const int c = s_on[i];
      s_on[i] = run;
      run += c;
    }
    s_total_on = run;
  }
  __syncthreads();
  const int n_off = n_states - s_total_on;
  int on_before = s_on[threadIdx.x];
  for (int s = lo; s < hi; ++s) {
    const bool t = terminal(tv.keys[s]);
    map[s] = (uint32_t)(t ? n_off + on_before : s - on_before);
    on_before += t ? 1 : 0;
  }
  __threadfence_block();
  __syncthreads();
  for (int s = lo; s < hi; ++s) {
    const uint32_t to = map[s];
    keys2[to] = tv.keys[s];
    for (int a = 0; a < 4; ++a) {
      const uint32_t e = tv.table[s * 4 + a];
      table2[to * 4 + a] = (e & 0xffff0000u) | __ldcg(map + (e & 0xffffu));
    }
  }
  for (int slot = threadIdx.x; slot < tv.hash_cap; slot += kBuildThreads)
    if (tv.hash_keys[slot] != 0ull) tv.hash_vals[slot] = __ldcg(map + tv.hash_vals[slot]);
  if (threadIdx.x == 0) meta[0] = n_off;
}

// Two-step entries of a one-goal puzzle from the (renumbered) one-step table: one thread per (state, a1).
//   pairs[state * 16 + a1 * 4 + a2] = successor2 * 8 | terminal after a1 << 2 | (moved by a1) + (moved by a2)
// "terminal after a2" is successor2 >= n_off.
__global__ void k_build_pairs(StateTableView tv, int n_states) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_states * 4) return;
  const uint32_t e1 = tv.table[i];
  const uint32_t s1 = e1 & 0xffffu, on1 = (e1 >> 23) & 1u, m1 = (e1 >> 24) & 1u;
  for (int a2 = 0; a2 < 4; ++a2) {
    const uint32_t e2 = tv.table[s1 * 4 + a2];
    tv.pairs[i * 4 + a2] = (uint16_t)(((e2 & 0xffffu) << 3) | (on1 << 2) | (m1 + ((e2 >> 24) & 1u)));
  }
}

__device__ __noinline__ void table_slow_path(const uint64_t* table, const StepParams& p, int64_t e, StepCounts& c) {
  LocalEnv<GlobalTable> env;
  env.init(GlobalTable{table});
  run_steps_thread<false>(env, p, e, c);
}

// ------------------------------------------------------------------------------------------
// The rollout on the table.  Plain shape (reward + terminated) or packed records, whole warps.
// kCompact (one goal object, at most 8192 states): the staged entry is laid out so that a step is a
// handful of instructions --
//   bits 4..16  successor * 16 (= the byte offset of its table row: mask, add 4 * action, load)
//   bit  17     the goal object is on its goal after the step (= terminated, puzzle.py:409-411)
//   bit  22     the step moved the agent            bit 27  the step gained the goal
//   bit  31     copy of bit 17 (sign test -> reward select, shift -> terminated byte)
// The three counters are accumulated as 5 / 5 / 4-bit fields of ONE register (`acc += entry - state`)
// and unpacked once per 16-step action tile (a goal cannot be gained in two consecutive steps, so the
// fields cannot carry into each other).  The reward is selected from register constants: 10 when the
// step ends on the goal, else -1.01 / -0.01 by whether the PREVIOUS step ended there.  With 32
// registers per thread two 896-thread CTAs share an SM (56 warps) and cover the latency of the
// dependent shared-memory loads (~3.4-way bank conflicts for 32 random states).
// ------------------------------------------------------------------------------------------
template <bool kPacked, bool kCompact, int kCtasPerSm>
__global__ void __launch_bounds__(kCtasPerSm == 2 ? kCompactBlock : kTableMaxBlock, kCtasPerSm)
    k_rollout_table(const __grid_constant__ StepParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ uint32_t s_counters[4];
  const StateTableView& tv = p.state_table;
  uint32_t* s_tab = reinterpret_cast<uint32_t*>(smem_raw);
  uint8_t* s_act = smem_raw + (size_t)tv.n_states * 4 * sizeof(uint32_t);
  float* s_lut = reinterpret_cast<float*>(s_act + (blockDim.x / 32) * 2 * kTileBytes);
  if (threadIdx.x < 4) s_counters[threadIdx.x] = 0;
  for (int i = threadIdx.x; i < tv.n_states * 4; i += blockDim.x) {
    const uint32_t entry = __ldg(tv.table + i);
    if (kCompact) {
      const uint32_t on = (entry >> 23) & 1u, moved = (entry >> 24) & 1u, gain = (entry >> 25) & 1u;
      s_tab[i] = ((entry & 0x1fffu) << 4) | (on << 17) | (moved << 22) | (gain << 27) | (on << 31);
    } else {
      s_tab[i] = entry;
    }
  }
  // record byte -> reward (gym_env.py:212-221), float32 rounded once (IEEE: same as the host table)
  for (int c = threadIdx.x; c < 256; c += blockDim.x)
    s_lut[c] = (c & 0x80) ? 10.0f : (float)((double)((c & 0x3f) - kMaxObjects) - 0.01);
  __syncthreads();

  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  StepCounts c;
  if (e < p.N) {  // whole warps (the launcher requires N % 32 == 0)
    const uint64_t* blob = p.blob + p.puzzle_off[p.puzzle_uniform];
    EnvBase<GlobalTable> env;
    env.init_header(GlobalTable{blob});
    const int n = env.n(), g = env.g(), W = env.width(), H = env.height();
    // this environment's state -> its index in the table
    uint64_t key = 0;
    bool found = true;
    int ach = 0;
    for (int j = 0; j < n; ++j) {
      const int x = p.x[j * p.stride + e], y = p.y[j * p.stride + e];
      if ((unsigned)x >= (unsigned)W || (unsigned)y >= (unsigned)H) found = false;
      key |= (uint64_t)(uint32_t)(y * W + x) << (j * tv.key_bits);
      if (j >= 1 && j <= g) {
        const uint32_t gw = env.table_desc(j, 2);
        ach += (x == lo16(gw) && y == hi16(gw)) ? 1 : 0;
      }
    }
    uint32_t state = 0;
    if (found) {
      found = false;
      const uint64_t hmask = (uint64_t)tv.hash_cap - 1;
      uint64_t slot = mix_key(key) & hmask;
      for (uint64_t probes = 0; probes <= hmask; ++probes) {
        const unsigned long long cur = tv.hash_keys[slot];
        if (cur == key + 1) {
          state = tv.hash_vals[slot];
          found = true;
          break;
        }
        if (cur == 0ull) break;
        slot = (slot + 1) & hmask;
      }
    }
    if (kCompact) state <<= 4;
    if (!__all_sync(0xffffffffu, found)) {
      // a state the table does not hold (not reachable from the initial state): the general rule
      table_slow_path(blob, p, e, c);
    } else {
      const int T = p.T;
      uint8_t* const tiles = s_act + warp * (2 * kTileBytes);
      const uint32_t tiles_s = (uint32_t)__cvta_generic_to_shared(tiles);
      const uint8_t* const warp_actions = p.actions + (e - (int64_t)lane);
      auto issue_block = [&](const int block) {  // lane l copies 16 of the warp's 32 actions of step 16*block + l/2
        const int row = block * kTileSteps + (int)(lane >> 1);
        const uint8_t* src = warp_actions + (size_t)(row < T ? row : T - 1) * (size_t)p.N + (lane & 1u) * 16u;
        const uint32_t dst = tiles_s + (uint32_t)(block & 1) * kTileBytes + lane * 16u;  // clamped rows: never used
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
      };
      // Addressing as in the ring kernel (pw_kernels.cu): a lean launch has T * N < 2^32 and N even, so
      // the element offset of (t, e) is 2 * h + (e & 1) with the 32-bit h = (e >> 1) + t * (N / 2).
      const uint32_t half_n = (uint32_t)(p.N >> 1);
      uint8_t* byte_base = (kPacked ? p.record : p.terminated) + (e & 1);
      float* const reward_base = p.reward + (e & 1);
      asm volatile("" : "+l"(byte_base));
      uint32_t h_out = (uint32_t)(e >> 1);
      uint32_t idle = record_byte(0, ach == g, false);  // the record of a no-op in the current state
      // kCompact: what the previous step left -- on its goal or not -- as the value the next step needs
      uint32_t on = ach == g ? 1u : 0u;
      const float r_idle = s_lut[kMaxObjects], r_lost = s_lut[kMaxObjects - 1];  // -0.01, -1.01
      float r_off = on ? r_lost : r_idle;                                         // reward of a step that ends off the goal
      uint32_t c_off = (uint32_t)kMaxObjects - on;                                // its record byte (delta_goals + 32)
      uint32_t acc = 0;
      uint32_t n_term = 0, n_moved = 0, n_gain = 0, n_invalid = 0;
      auto one_step = [&](const uint32_t a, auto checked) {
        constexpr bool kChecked = decltype(checked)::value;
        if (kCompact) {
          // `state` is the byte offset of the state's table row (index * 16)
          uint32_t entry = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(s_tab) + state +
                                                              4u * (kChecked ? (a & 3u) : a));
          if (kChecked && a > 3u) {  // invalid action: a counted no-op
            entry = state | (on << 17) | (on << 31);
            ++n_invalid;
          }
          on = entry >> 31;
          state = entry & 0x1fff0u;
          acc += entry - state;  // the three counter fields (and bit 31, which falls off the top)
          const bool ends_on = (int32_t)entry < 0;
          if (kPacked) {
            const uint32_t code = c_off + (ends_on ? 0x81u : 0u);  // terminated | delta_goals one higher
            c_off = ends_on ? (uint32_t)kMaxObjects - 1u : (uint32_t)kMaxObjects;
            asm volatile("st.global.u8 [%0], %1;" ::"l"(byte_base + (size_t)h_out * 2u), "r"(code) : "memory");
          } else {
            reward_base[(size_t)h_out * 2u] = ends_on ? 10.0f : r_off;  // gym_env.py:212-221
            r_off = ends_on ? r_lost : r_idle;
            asm volatile("st.global.u8 [%0], %1;" ::"l"(byte_base + (size_t)h_out * 2u), "r"(on) : "memory");
          }
        } else {
          uint32_t entry = s_tab[state * 4u + (kChecked ? (a & 3u) : a)];
          if (kChecked && a > 3u) {  // invalid action: a counted no-op
            entry = state | (idle << 16);
            ++n_invalid;
          }
          state = entry & 0xffffu;
          const uint32_t code = (entry >> 16) & 0xffu;
          idle = (code & 0x80u) | (uint32_t)kMaxObjects;
          n_term += code >> 7;
          n_moved += (entry >> 24) & 1u;
          n_gain += entry >> 25;
          if (kPacked) {
            asm volatile("st.global.u8 [%0], %1;" ::"l"(byte_base + (size_t)h_out * 2u), "r"(code) : "memory");
          } else {
            reward_base[(size_t)h_out * 2u] = s_lut[code];
            asm volatile("st.global.u8 [%0], %1;" ::"l"(byte_base + (size_t)h_out * 2u), "r"(code >> 7) : "memory");
          }
        }
        h_out += half_n;
      };
      const int blocks = (T + kTileSteps - 1) / kTileSteps;
      issue_block(0);
      for (int blk = 0; blk < blocks; ++blk) {
        __syncwarp();          // every lane is done with the tile the next block lands in
        issue_block(blk + 1);  // past the end: re-reads the last row, never used
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();  // block `blk` is visible to the whole warp
        const uint8_t* tile = tiles + (blk & 1) * kTileBytes;
        const uint4 w0 = *reinterpret_cast<const uint4*>(tile + lane * 16u);
        const uint32_t seen = w0.x | w0.y | w0.z | w0.w;
        const bool any_invalid = __any_sync(0xffffffffu, (seen & 0xfcfcfcfcu) != 0u);
        const int steps = T - blk * kTileSteps < kTileSteps ? T - blk * kTileSteps : kTileSteps;
        const uint8_t* mine = tile + lane;
        if (!any_invalid && steps == kTileSteps) {
#pragma unroll
          for (int k = 0; k < kTileSteps; ++k) one_step(mine[k * 32], std::false_type{});
        } else {
#pragma unroll 1
          for (int k = 0; k < steps; ++k) one_step(mine[k * 32], std::true_type{});
          if (kCompact) {  // the unchecked path carries these through the selects
            r_off = on ? r_lost : r_idle;
            c_off = (uint32_t)kMaxObjects - on;
          }
        }
        if (kCompact) {  // at most 16 / 16 / 8 in the 5 / 5 / 4-bit fields
          n_term += (acc >> 17) & 31u;
          n_moved += (acc >> 22) & 31u;
          n_gain += (acc >> 27) & 15u;
          acc = 0;
        }
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      c.terminated += n_term;
      c.moved += n_moved;
      c.dgoals += n_gain;
      c.invalid += n_invalid;
      // positions back from the state index
      const uint64_t final_key = tv.keys[kCompact ? state >> 4 : state];
      for (int j = 0; j < n; ++j) {
        const uint32_t cell = (uint32_t)(final_key >> (j * tv.key_bits)) & ((1u << tv.key_bits) - 1u);
        p.x[j * p.stride + e] = (int)(cell % (uint32_t)W);
        p.y[j * p.stride + e] = (int)(cell / (uint32_t)W);
      }
    }
  }
  flush_counts(c, p, s_counters);
}

// ------------------------------------------------------------------------------------------
// The rollout on the TWO-STEP table of a one-goal puzzle (StateTableView::pairs).
// What bounds k_rollout_table is the shared-memory / LSU data pipe (ncu: 71-81 % of its peak): the
// 32 random states of a warp hit the 32 banks with ~3.7 wavefronts per table load, next to the action
// byte load and the stores.  Looking up (state, a1, a2) -> state'' halves the table loads per step.
// The 16-bit entry has no room for "terminal after a2", which is why the states are numbered
// terminal-last: it is `successor >= n_off`.  Per pair of steps: two action bytes, one table load, and
// selects for the two rewards / record bytes (gym_env.py:212-221: 10 when the step ends on the goal,
// else -1.01 / -0.01 by whether the previous step ended there).  Counters: the `moved` counts add up
// in a register; the terminal bits are shifted into a history word per action tile, and terminated
// steps / goals gained are pop-counts of it.
// Blocks with an out-of-range action, a last partial tile or an odd last step walk the one-step table
// in global memory (L2-resident, 16 bytes per state).
// ------------------------------------------------------------------------------------------
template <bool kPacked>
__global__ void __launch_bounds__(kTableMaxBlock, 1) k_rollout_pairs(const __grid_constant__ StepParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ uint32_t s_counters[4];
  const StateTableView& tv = p.state_table;
  uint8_t* s_act = smem_raw + (size_t)tv.n_states * 32;
  if (threadIdx.x < 4) s_counters[threadIdx.x] = 0;
  {
    const uint4* src = reinterpret_cast<const uint4*>(tv.pairs);
    uint4* dst = reinterpret_cast<uint4*>(smem_raw);
    for (int i = threadIdx.x; i < tv.n_states * 2; i += blockDim.x) dst[i] = __ldg(src + i);
  }
  __syncthreads();

  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  StepCounts c;
  if (e < p.N) {  // whole warps (the launcher requires N % 32 == 0)
    const uint64_t* blob = p.blob + p.puzzle_off[p.puzzle_uniform];
    EnvBase<GlobalTable> env;
    env.init_header(GlobalTable{blob});
    const int n = env.n(), W = env.width(), H = env.height();
    uint64_t key = 0;
    bool found = true;
    for (int j = 0; j < n; ++j) {
      const int x = p.x[j * p.stride + e], y = p.y[j * p.stride + e];
      if ((unsigned)x >= (unsigned)W || (unsigned)y >= (unsigned)H) found = false;
      key |= (uint64_t)(uint32_t)(y * W + x) << (j * tv.key_bits);
    }
    uint32_t state = 0;
    if (found) {
      found = false;
      const uint64_t hmask = (uint64_t)tv.hash_cap - 1;
      uint64_t slot = mix_key(key) & hmask;
      for (uint64_t probes = 0; probes <= hmask; ++probes) {
        const unsigned long long cur = tv.hash_keys[slot];
        if (cur == key + 1) {
          state = tv.hash_vals[slot];
          found = true;
          break;
        }
        if (cur == 0ull) break;
        slot = (slot + 1) & hmask;
      }
    }
    if (!__all_sync(0xffffffffu, found)) {
      table_slow_path(blob, p, e, c);  // a state the table does not hold: the general rule
    } else {
      const int T = p.T;
      uint8_t* const tiles = s_act + warp * (2 * kTileBytes);
      const uint32_t tiles_s = (uint32_t)__cvta_generic_to_shared(tiles);
      const uint8_t* const warp_actions = p.actions + (e - (int64_t)lane);
      auto issue_block = [&](const int block) {  // lane l copies 16 of the warp's 32 actions of step 16*block + l/2
        const int row = block * kTileSteps + (int)(lane >> 1);
        const uint8_t* src = warp_actions + (size_t)(row < T ? row : T - 1) * (size_t)p.N + (lane & 1u) * 16u;
        const uint32_t dst = tiles_s + (uint32_t)(block & 1) * kTileBytes + lane * 16u;  // clamped rows: never used
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
      };
      // output addressing as in the ring kernel: element offset of (t, e) = 2 * h + (e & 1), 32-bit h
      const uint32_t half_n = (uint32_t)(p.N >> 1);
      uint8_t* byte_base = (kPacked ? p.record : p.terminated) + (e & 1);
      float* const reward_base = p.reward + (e & 1);
      asm volatile("" : "+l"(byte_base));
      uint32_t h_out = (uint32_t)(e >> 1);
      const uint32_t thr8 = (uint32_t)tv.n_off << 3;
      uint32_t st8 = state << 3;  // state index * 8, the form the entries hold it in
      uint32_t on = st8 >= thr8 ? 1u : 0u;
      const float r_idle = c_reward_lut[kMaxObjects], r_lost = c_reward_lut[kMaxObjects - 1];  // -0.01, -1.01
      float r_off = on ? r_lost : r_idle;           // reward of a step that ends off the goal
      uint32_t c_off = (uint32_t)kMaxObjects - on;  // its record byte (delta_goals + 32)
      uint32_t n_term = 0, n_moved = 0, n_gain = 0, n_invalid = 0;
      auto emit = [&](const bool ends_on) {  // one step's outputs, given whether it ends on the goal
        if (kPacked) {
          const uint32_t code = c_off + (ends_on ? 0x81u : 0u);  // terminated | delta_goals one higher
          c_off = ends_on ? (uint32_t)kMaxObjects - 1u : (uint32_t)kMaxObjects;
          asm volatile("st.global.u8 [%0], %1;" ::"l"(byte_base + (size_t)h_out * 2u), "r"(code) : "memory");
        } else {
          reward_base[(size_t)h_out * 2u] = ends_on ? 10.0f : r_off;
          r_off = ends_on ? r_lost : r_idle;
          asm volatile("st.global.u8 [%0], %1;" ::"l"(byte_base + (size_t)h_out * 2u), "r"(ends_on ? 1u : 0u)
                       : "memory");
        }
        h_out += half_n;
      };
      const int blocks = (T + kTileSteps - 1) / kTileSteps;
      issue_block(0);
      for (int blk = 0; blk < blocks; ++blk) {
        __syncwarp();          // every lane is done with the tile the next block lands in
        issue_block(blk + 1);  // past the end: re-reads the last row, never used
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();  // block `blk` is visible to the whole warp
        const uint8_t* tile = tiles + (blk & 1) * kTileBytes;
        const uint4 w0 = *reinterpret_cast<const uint4*>(tile + lane * 16u);
        const uint32_t seen = w0.x | w0.y | w0.z | w0.w;
        const bool any_invalid = __any_sync(0xffffffffu, (seen & 0xfcfcfcfcu) != 0u);
        const int steps = T - blk * kTileSteps < kTileSteps ? T - blk * kTileSteps : kTileSteps;
        const uint8_t* mine = tile + lane;
        if (!any_invalid && steps == kTileSteps) {
          uint32_t hist = 0, moved = 0;
#pragma unroll
          for (int k = 0; k < kTileSteps; k += 2) {
            const uint32_t a1 = mine[k * 32], a2 = mine[k * 32 + 32];
            const uint32_t entry = *reinterpret_cast<const uint16_t*>(smem_raw + (st8 << 2) + (a1 * 8u + a2 * 2u));
            st8 = entry & 0xfff8u;
            moved += entry & 3u;
            const bool on1 = (entry & 4u) != 0u, on2 = st8 >= thr8;
            hist = hist * 4u + (on1 ? 2u : 0u) + (on2 ? 1u : 0u);
            emit(on1);
            emit(on2);
          }
          // bit 15 of hist is the tile's first step; a goal is gained where a step ends on the goal and the one before did not
          n_term += __popc(hist);
          n_gain += __popc(hist & ~(((on << 16) | hist) >> 1));
          n_moved += moved;
          on = hist & 1u;
        } else {
#pragma unroll 1
          for (int k = 0; k < steps; ++k) {  // one-step table in global memory: successor | record << 16 | moved << 24
            const uint32_t a = mine[k * 32];
            uint32_t succ = st8 >> 3, moved1 = 0;
            if (a > 3u) {
              ++n_invalid;
            } else {
              const uint32_t entry = __ldg(tv.table + (st8 >> 1) + a);
              succ = entry & 0xffffu;
              moved1 = (entry >> 24) & 1u;
            }
            st8 = succ << 3;
            const uint32_t now_on = st8 >= thr8 ? 1u : 0u;
            n_term += now_on;
            n_gain += now_on & ~on;
            n_moved += moved1;
            on = now_on;
            emit(now_on != 0u);
          }
        }
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      c.terminated += n_term;
      c.moved += n_moved;
      c.dgoals += n_gain;
      c.invalid += n_invalid;
      const uint64_t final_key = tv.keys[st8 >> 3];  // positions back from the state index
      for (int j = 0; j < n; ++j) {
        const uint32_t cell = (uint32_t)(final_key >> (j * tv.key_bits)) & ((1u << tv.key_bits) - 1u);
        p.x[j * p.stride + e] = (int)(cell % (uint32_t)W);
        p.y[j * p.stride + e] = (int)(cell / (uint32_t)W);
      }
    }
  }
  flush_counts(c, p, s_counters);
}

// ------------------------------------------------------------------------------------------
// The two-step table with FOUR environments per thread (4 e .. 4 e + 3) -- the kernel bench.py times.
// Why four: (a) the outputs of a step leave as ONE 16-byte reward store and ONE 4-byte terminated /
// record store per thread -- 512 B and 128 B per warp instruction instead of 128 B and 32 B; the
// synthetic probe tools/probes/write_pattern2.cu writes the [T][N] pattern at 5.4 TB/s with one
// environment per thread and at 7.0 TB/s with four; (b) four independent table chains per thread
// hide the shared-memory latency with 14 warps per SM, which is one wave for 2^18 environments;
// (c) per-step bookkeeping is done on four packed bytes at a time (the action pair's table column,
// the terminated bytes, the record bytes `32 + delta_goals | terminated << 7`, the counters).
// Staged entries have the table's own shared-memory address folded in (state index + base / 32), so
// that `(entry & 0xfff8) * 4 + column` IS the shared address of the next entry.
// ------------------------------------------------------------------------------------------
constexpr int kQuadBlock = 512;                   // largest CTA: 16 warps x 128 environments
constexpr int kQuadTileBytes = kTileSteps * 128;  // a warp's 128 actions of 16 steps

template <bool kPacked>
__global__ void __launch_bounds__(kQuadBlock, 1) k_rollout_pairs4(const __grid_constant__ StepParams p) {
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  __shared__ uint32_t s_counters[4];
  const StateTableView& tv = p.state_table;
  uint8_t* const smem_raw = smem_dyn + ((0u - (uint32_t)__cvta_generic_to_shared(smem_dyn)) & 31u);  // 32-byte rows
  uint8_t* s_act = smem_raw + (size_t)tv.n_states * 32;
  const uint32_t tab_s = (uint32_t)__cvta_generic_to_shared(smem_raw);
  const uint32_t base8 = (tab_s >> 5) << 3;                             // (index offset of the table) * 8
  if (threadIdx.x < 4) s_counters[threadIdx.x] = 0;
  {
    const uint4* src = reinterpret_cast<const uint4*>(tv.pairs);
    uint4* dst = reinterpret_cast<uint4*>(smem_raw);
    const uint32_t add = base8 * 0x00010001u;  // to both 16-bit entries of a word; (n_states + offset) * 8 < 2^16
    for (int i = threadIdx.x; i < tv.n_states * 2; i += blockDim.x) {
      uint4 v = __ldg(src + i);
      v.x += add, v.y += add, v.z += add, v.w += add;
      dst[i] = v;
    }
  }
  __syncthreads();

  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  StepCounts c;
  if (e < p.N) {  // whole warps (the launcher requires N % 128 == 0)
    const uint64_t* blob = p.blob + p.puzzle_off[p.puzzle_uniform];
    EnvBase<GlobalTable> env;
    env.init_header(GlobalTable{blob});
    const int n = env.n(), W = env.width(), H = env.height();
    uint32_t st[4];  // (state index + table offset) * 8
    bool found = true;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint64_t key = 0;
      bool ok = true;
      for (int j = 0; j < n; ++j) {
        const int x = p.x[j * p.stride + e + i], y = p.y[j * p.stride + e + i];
        if ((unsigned)x >= (unsigned)W || (unsigned)y >= (unsigned)H) ok = false;
        key |= (uint64_t)(uint32_t)(y * W + x) << (j * tv.key_bits);
      }
      uint32_t state = 0;
      bool hit = false;
      if (ok) {
        const uint64_t hmask = (uint64_t)tv.hash_cap - 1;
        uint64_t slot = mix_key(key) & hmask;
        for (uint64_t probes = 0; probes <= hmask; ++probes) {
          const unsigned long long cur = tv.hash_keys[slot];
          if (cur == key + 1) {
            state = tv.hash_vals[slot];
            hit = true;
            break;
          }
          if (cur == 0ull) break;
          slot = (slot + 1) & hmask;
        }
      }
      found = found && hit;
      st[i] = (state << 3) + base8;
    }
    if (!__all_sync(0xffffffffu, found)) {
      for (int i = 0; i < 4; ++i) table_slow_path(blob, p, e + i, c);  // a state outside the table: the general rule
    } else {
      const int T = p.T;
      uint8_t* const tiles = s_act + warp * (2 * kQuadTileBytes);
      const uint32_t tiles_s = (uint32_t)__cvta_generic_to_shared(tiles);
      const uint8_t* const warp_actions = p.actions + (e - (int64_t)lane * 4);
      auto issue_block = [&](const int block) {  // 16 rows of 128 bytes; lane l copies the 16-byte chunks l + 32 i
        const uint32_t dst = tiles_s + (uint32_t)(block & 1) * kQuadTileBytes + lane * 16u;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = block * kTileSteps + (int)(lane >> 3) + 4 * i;
          const uint8_t* src = warp_actions + (size_t)(row < T ? row : T - 1) * (size_t)p.N + (lane & 7u) * 16u;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)i * 512u), "l"(src) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");  // rows past the end are clamped and never used
      };
      // outputs: element offset of (t, e) = 4 * h with the 32-bit h = e / 4 + t * (N / 4)  (T * N < 2^32)
      const uint32_t quarter_n = (uint32_t)(p.N >> 2);
      uint8_t* byte_base = kPacked ? p.record : p.terminated;
      float* const reward_base = p.reward;
      asm volatile("" : "+l"(byte_base));
      uint32_t h_out = (uint32_t)(e >> 2);
      const uint32_t thr = ((uint32_t)tv.n_off << 3) + base8;  // terminal <=> st >= thr
      const float r_idle = c_reward_lut[kMaxObjects], r_lost = c_reward_lut[kMaxObjects - 1];  // -0.01, -1.01
      uint32_t t_prev = 0;  // byte i = 1: environment i's previous step ended on the goal
      float r_off[4];       // reward of a step that ends off the goal (-1.01 if it just left it)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool on = st[i] >= thr;
        t_prev |= on ? 1u << (8 * i) : 0u;
        r_off[i] = on ? r_lost : r_idle;
      }
      uint32_t n_term = 0, n_moved = 0, n_gain = 0, n_invalid = 0;
      // one step's outputs for the 4 environments: `t` = their packed "ends on the goal" bytes
      auto emit = [&](const uint32_t t) {
        if (kPacked) {
          // record = 32 + delta_goals | terminated << 7 with delta_goals = t - t_prev per byte (no carries)
          const uint32_t code = t * 0x81u + (0x20202020u - t_prev);
          asm volatile("st.global.u32 [%0], %1;" ::"l"(byte_base + (size_t)h_out * 4u), "r"(code) : "memory");
        } else {
          float4 r;
          r.x = (t & 0x00000001u) ? 10.0f : r_off[0];
          r.y = (t & 0x00000100u) ? 10.0f : r_off[1];
          r.z = (t & 0x00010000u) ? 10.0f : r_off[2];
          r.w = (t & 0x01000000u) ? 10.0f : r_off[3];
          r_off[0] = (t & 0x00000001u) ? r_lost : r_idle;
          r_off[1] = (t & 0x00000100u) ? r_lost : r_idle;
          r_off[2] = (t & 0x00010000u) ? r_lost : r_idle;
          r_off[3] = (t & 0x01000000u) ? r_lost : r_idle;
          *reinterpret_cast<float4*>(reward_base + (size_t)h_out * 4u) = r;
          asm volatile("st.global.u32 [%0], %1;" ::"l"(byte_base + (size_t)h_out * 4u), "r"(t) : "memory");
        }
        t_prev = t;
        h_out += quarter_n;
      };
      const int blocks = (T + kTileSteps - 1) / kTileSteps;
      issue_block(0);
      for (int blk = 0; blk < blocks; ++blk) {
        __syncwarp();          // every lane is done with the tile the next block lands in
        issue_block(blk + 1);  // past the end: re-reads the last row, never used
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();  // block `blk` is visible to the whole warp
        const uint8_t* mine = tiles + (blk & 1) * kQuadTileBytes + lane * 4u;
        uint32_t w[kTileSteps], seen = 0;  // this thread's 4 actions of each step of the tile
#pragma unroll
        for (int k = 0; k < kTileSteps; ++k) {
          w[k] = *reinterpret_cast<const uint32_t*>(mine + k * 128);
          seen |= w[k];
        }
        const bool any_invalid = __any_sync(0xffffffffu, (seen & 0xfcfcfcfcu) != 0u);
        const int steps = T - blk * kTileSteps < kTileSteps ? T - blk * kTileSteps : kTileSteps;
        if (!any_invalid && steps == kTileSteps) {
          uint32_t acc_t = 0, acc_g = 0, acc_m = 0;  // per-byte sums over the tile (at most 16 / 8 / 16)
#pragma unroll
          for (int k = 0; k < kTileSteps; k += 2) {
            const uint32_t col = (w[k] << 3) + (w[k + 1] << 1);  // per byte: byte offset of (a1, a2) in a table row
            uint32_t en[4], d[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t addr = (st[i] << 2) + __byte_perm(col, 0u, 0x4440u + (uint32_t)i);
              uint16_t v;
              asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
              en[i] = v;
              st[i] = en[i] & 0xfff8u;
              d[i] = st[i] - thr;  // negative <=> the pair ends off the goal
            }
            const uint32_t low = __byte_perm(__byte_perm(en[0], en[1], 0x0040u), __byte_perm(en[2], en[3], 0x0040u), 0x5410u);
            const uint32_t top = __byte_perm(__byte_perm(d[0], d[1], 0x0073u), __byte_perm(d[2], d[3], 0x0073u), 0x5410u);
            const uint32_t t1 = (low >> 2) & 0x01010101u, t2 = ~top & 0x01010101u;
            acc_m += low & 0x03030303u;
            acc_t += t1 + t2;
            acc_g += (t1 & ~t_prev) + (t2 & ~t1);
            emit(t1);
            emit(t2);
          }
          n_term += (acc_t * 0x01010101u) >> 24;
          n_gain += (acc_g * 0x01010101u) >> 24;
          n_moved += (acc_m * 0x01010101u) >> 24;
        } else {
#pragma unroll 1
          for (int k = 0; k < steps; ++k) {  // one-step table in global memory: successor | record << 16 | moved << 24
            const uint32_t wk = *reinterpret_cast<const uint32_t*>(mine + k * 128);
            uint32_t t = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t a = (wk >> (8 * i)) & 0xffu;
              if (a > 3u) {
                ++n_invalid;
              } else {
                const uint32_t entry = __ldg(tv.table + ((st[i] - base8) >> 1) + a);
                st[i] = ((entry & 0xffffu) << 3) + base8;
                n_moved += (entry >> 24) & 1u;
              }
              t |= st[i] >= thr ? 1u << (8 * i) : 0u;
            }
            n_term += __popc(t);
            n_gain += __popc(t & ~t_prev);
            emit(t);
          }
        }
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      c.terminated += n_term;
      c.moved += n_moved;
      c.dgoals += n_gain;
      c.invalid += n_invalid;
#pragma unroll
      for (int i = 0; i < 4; ++i) {  // positions back from the state index
        const uint64_t final_key = tv.keys[(st[i] - base8) >> 3];
        for (int j = 0; j < n; ++j) {
          const uint32_t cell = (uint32_t)(final_key >> (j * tv.key_bits)) & ((1u << tv.key_bits) - 1u);
          p.x[j * p.stride + e + i] = (int)(cell % (uint32_t)W);
          p.y[j * p.stride + e + i] = (int)(cell / (uint32_t)W);
        }
      }
    }
  }
  flush_counts(c, p, s_counters, 4);
}

// kind 0: one-step table (16 B per state), 1: two-step table, 2: two-step table, four environments per thread
size_t table_smem_bytes(int n_states, int block, int kind = 0) {
  if (kind == 2) return (size_t)n_states * 32 + 32 + (size_t)(block / 32) * 2 * kQuadTileBytes;
  return (size_t)n_states * (kind ? 32 : 16) + (size_t)(block / 32) * 2 * kTileBytes + (kind ? 0 : 256 * sizeof(float));
}

// Threads per CTA for a launch over n_envs environments.  The table is staged once per CTA and an
// environment is busy for the whole rollout, so what matters is (a) how many warps are resident per
// SM and (b) that the CTAs fill whole waves: 2^18 environments are 55.4 warps per SM, which 48
// resident warps process in 1.15 waves -- the second one nearly empty.  Picks the block size whose
// resident CTAs (occupancy API: registers + shared memory) waste the least of the last wave while
// keeping enough warps to cover the shared-memory latency.
template <class Kernel>
int pick_table_block_uncached(Kernel kernel, int64_t n_envs, int n_states, int max_block, int kind) {
  int device = 0, sms = 148;
  cudaGetDevice(&device);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  int best_block = 256;
  double best_score = -1.0;
  for (int block = 128; block <= max_block; block += 32) {
    const size_t smem = table_smem_bytes(n_states, block, kind);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem) != cudaSuccess || per_sm < 1) {
      cudaGetLastError();
      continue;
    }
    const int64_t n_threads = kind == 2 ? n_envs / 4 : n_envs;
    const double ctas = (double)((n_threads + block - 1) / block);
    const double waves = ctas / ((double)sms * per_sm);
    const double whole = waves <= 1.0 ? 1.0 : (double)(int64_t)(waves + 0.999999);
    const double fill = waves <= 1.0 ? 1.0 : waves / whole;
    const double warps = (double)per_sm * block / 32.0;
    const double enough = kind == 2 ? 12.0 : 28.0;  // warps per SM that cover the shared-memory latency
    const double cover = warps >= enough ? 1.0 : warps / enough;
    const double score = fill * cover;
    if (score >= best_score - 1e-9) {  // ties: the larger block (the table is staged once per CTA)
      best_score = score;
      best_block = block;
    }
  }
  return best_block;
}


// The choice depends on (kernel, batch size, table size) only: remembered, so that the chunked launches of the
// host pipeline do not pay ~30 occupancy queries each.
template <class Kernel>
int pick_table_block(Kernel kernel, int64_t n_envs, int n_states, int max_block, int kind = 0) {
  struct Entry {
    const void* kernel;
    int64_t n_envs;
    int n_states, block;
  };
  static std::mutex mutex;
  static std::vector<Entry> cache;
  const void* key = reinterpret_cast<const void*>(kernel);
  {
    std::lock_guard<std::mutex> lock(mutex);
    for (const Entry& e : cache)
      if (e.kernel == key && e.n_envs == n_envs && e.n_states == n_states) return e.block;
  }
  const int block = pick_table_block_uncached(kernel, n_envs, n_states, max_block, kind);
  std::lock_guard<std::mutex> lock(mutex);
  if (cache.size() < 256) cache.push_back(Entry{key, n_envs, n_states, block});
  return block;
}

}  // namespace

// (this file's copy of the reward table of pw_steps.cuh, used by the general-rule fall-back)
cudaError_t upload_reward_lut_table(const float* host_lut) {
  return cudaMemcpyToSymbol(c_reward_lut, host_lut, sizeof(float) * (2 * kMaxObjects + 1));
}

// Largest table a rollout CTA can stage next to its action tiles (227 KB of shared memory per SM).
int state_table_max_states() { return 12288; }  // 192 KB of 32-bit entries + action tiles of a 256-thread CTA

// Enumerates the reachable states of the puzzle at `puzzle_off` into `*tv` (device arrays allocated by
// the caller: keys[max_states], table[4 * max_states], hash_keys / hash_vals[hash_cap], zeroed hash),
// renumbers them terminal-last (tv->n_off) and, for a one-goal puzzle that is small enough, allocates
// and fills the two-step table tv->pairs.  Synchronises.  Returns the number of states, or 0 if the
// puzzle is not tabulated.
int build_state_table(const uint64_t* d_blob, uint32_t puzzle_off, StateTableView* tv, int max_states,
                      cudaError_t* err) {
  int* d_meta = nullptr;
  *err = cudaMalloc(&d_meta, 2 * sizeof(int));
  if (*err != cudaSuccess) return 0;
  k_build_state_table<<<1, kBuildThreads>>>(d_blob, puzzle_off, *tv, max_states, d_meta);
  int meta[2] = {0, 1};
  *err = cudaMemcpy(meta, d_meta, sizeof(meta), cudaMemcpyDeviceToHost);
  const int n_states = (*err == cudaSuccess && !meta[1]) ? meta[0] : 0;
  if (n_states > 0) {
    uint64_t* keys2 = nullptr;
    uint32_t *table2 = nullptr, *map = nullptr;
    *err = cudaMalloc(&keys2, sizeof(uint64_t) * (size_t)n_states);
    if (*err == cudaSuccess) *err = cudaMalloc(&table2, sizeof(uint32_t) * 4 * (size_t)n_states);
    if (*err == cudaSuccess) *err = cudaMalloc(&map, sizeof(uint32_t) * (size_t)n_states);
    if (*err == cudaSuccess) {
      k_renumber_state_table<<<1, kBuildThreads>>>(d_blob, puzzle_off, *tv, n_states, keys2, table2, map, d_meta);
      *err = cudaMemcpy(meta, d_meta, sizeof(int), cudaMemcpyDeviceToHost);
    }
    if (*err == cudaSuccess)
      *err = cudaMemcpy(tv->keys, keys2, sizeof(uint64_t) * (size_t)n_states, cudaMemcpyDeviceToDevice);
    if (*err == cudaSuccess)
      *err = cudaMemcpy(tv->table, table2, sizeof(uint32_t) * 4 * (size_t)n_states, cudaMemcpyDeviceToDevice);
    tv->n_off = meta[0];
    cudaFree(keys2);
    cudaFree(table2);
    cudaFree(map);
    if (*err == cudaSuccess && tv->goals == 1 && n_states <= kPairMaxStates) {
      *err = cudaMalloc(&tv->pairs, sizeof(uint16_t) * 16 * (size_t)n_states);
      if (*err == cudaSuccess) {
        k_build_pairs<<<(n_states * 4 + 255) / 256, 256>>>(*tv, n_states);
        *err = cudaDeviceSynchronize();
      }
    }
  }
  cudaFree(d_meta);
  if (*err != cudaSuccess) return 0;
  return n_states;
}

bool table_rollout_eligible(const StepParams& p) {
  const auto aligned16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  const bool plain = p.reward && p.terminated && !p.record;
  const bool packed = p.record && !p.reward && !p.terminated;
  return p.state_table.n_states > 0 && p.puzzle_id == nullptr && is_lean(p) && (plain || packed) &&
         (p.N & 31) == 0 && aligned16(p.actions) && getenv("PW_NO_TABLE") == nullptr;
}

cudaError_t launch_rollout_table(const StepParams& p, cudaStream_t stream) {
  // one goal object and at most 8192 states: the compact entries (see k_rollout_table); PW_TABLE_WIDE=1: A/B
  const bool compact = p.state_table.goals == 1 && p.state_table.n_states <= 8192 && getenv("PW_TABLE_WIDE") == nullptr;
  const int n_states = p.state_table.n_states;
  int block = 0;
  if (const char* env = getenv("PW_TABLE_BLOCK")) block = atoi(env);  // tuning runs
#define PW_TABLE_GO(PACKED, COMPACT, CTAS)                                                                     \
  do {                                                                                                    \
    const int max_block = CTAS == 2 ? kCompactBlock : kTableMaxBlock;                                     \
    if (block < 32 || block > max_block || (block & 31))                                                  \
      block = pick_table_block(k_rollout_table<PACKED, COMPACT, CTAS>, p.N, n_states, max_block);         \
    k_rollout_table<PACKED, COMPACT, CTAS><<<(unsigned)((p.N + block - 1) / block), block,                \
                                             table_smem_bytes(n_states, block), stream>>>(p);             \
  } while (0)
  const bool use_pairs = p.state_table.pairs && getenv("PW_TABLE_WIDE") == nullptr && getenv("PW_NO_PAIRS") == nullptr;
  const auto aligned = [](const void* q, uintptr_t a) { return (reinterpret_cast<uintptr_t>(q) & (a - 1)) == 0; };
  if (use_pairs && (p.N & 127) == 0 && getenv("PW_NO_QUAD") == nullptr &&
      (p.record ? aligned(p.record, 4) : aligned(p.reward, 16) && aligned(p.terminated, 4))) {
    const int64_t n_threads = p.N / 4;
    if (p.record) {
      if (block < 32 || block > kQuadBlock || (block & 31))
        block = pick_table_block(k_rollout_pairs4<true>, p.N, n_states, kQuadBlock, 2);
      k_rollout_pairs4<true><<<(unsigned)((n_threads + block - 1) / block), block, table_smem_bytes(n_states, block, 2),
                               stream>>>(p);
    } else {
      if (block < 32 || block > kQuadBlock || (block & 31))
        block = pick_table_block(k_rollout_pairs4<false>, p.N, n_states, kQuadBlock, 2);
      k_rollout_pairs4<false><<<(unsigned)((n_threads + block - 1) / block), block,
                                table_smem_bytes(n_states, block, 2), stream>>>(p);
    }
    return cudaGetLastError();
  }
  if (use_pairs) {
    if (p.record) {
      if (block < 32 || block > kTableMaxBlock || (block & 31))
        block = pick_table_block(k_rollout_pairs<true>, p.N, n_states, kTableMaxBlock, 1);
      k_rollout_pairs<true><<<(unsigned)((p.N + block - 1) / block), block, table_smem_bytes(n_states, block, 1),
                              stream>>>(p);
    } else {
      if (block < 32 || block > kTableMaxBlock || (block & 31))
        block = pick_table_block(k_rollout_pairs<false>, p.N, n_states, kTableMaxBlock, 1);
      k_rollout_pairs<false><<<(unsigned)((p.N + block - 1) / block), block, table_smem_bytes(n_states, block, 1),
                               stream>>>(p);
    }
    return cudaGetLastError();
  }
  bool two = p.record != nullptr;  // 32 registers, two CTAs per SM
  if (const char* env = getenv("PW_TABLE_CTAS")) two = atoi(env) == 2;
  if (compact) {
    if (p.record) { if (two) PW_TABLE_GO(true, true, 2); else PW_TABLE_GO(true, true, 1); }
    else { if (two) PW_TABLE_GO(false, true, 2); else PW_TABLE_GO(false, true, 1); }
  } else {
    if (p.record) PW_TABLE_GO(true, false, 1);
    else PW_TABLE_GO(false, false, 1);
  }
#undef PW_TABLE_GO
  return cudaGetLastError();
}

cudaError_t configure_table_kernels() {
  const int cap = 226 * 1024;  // the staged table + action tiles of the largest block
  cudaError_t err = cudaSuccess;
#define PW_OPT_IN(k) \
  if (err == cudaSuccess) err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, cap)
  PW_OPT_IN((k_rollout_pairs4<true>));
  PW_OPT_IN((k_rollout_pairs4<false>));
  PW_OPT_IN((k_rollout_pairs<true>));
  PW_OPT_IN((k_rollout_pairs<false>));
  PW_OPT_IN((k_rollout_table<true, true, 1>));
  PW_OPT_IN((k_rollout_table<false, true, 1>));
  PW_OPT_IN((k_rollout_table<true, true, 2>));
  PW_OPT_IN((k_rollout_table<false, true, 2>));
  PW_OPT_IN((k_rollout_table<true, false, 1>));
  PW_OPT_IN((k_rollout_table<false, false, 1>));
#undef PW_OPT_IN
  return err;
}

}  // namespace pw
