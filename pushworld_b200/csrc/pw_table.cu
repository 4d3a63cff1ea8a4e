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
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>
#include <type_traits>

#include "pw_env.cuh"
#include "pw_fast.cuh"
#include "pw_internal.h"
#include "pw_steps.cuh"

namespace pw {
namespace {

constexpr int kBuildThreads = 1024;
constexpr int kTableMaxBlock = 1024;  // threads (= environments) per CTA of the rollout: chosen per launch
constexpr int kTileBytes = 32 * 32;

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

__device__ __noinline__ void table_slow_path(const uint64_t* table, const StepParams& p, int64_t e, StepCounts& c) {
  LocalEnv<GlobalTable> env;
  env.init(GlobalTable{table});
  run_steps_thread<false>(env, p, e, c);
}

// ------------------------------------------------------------------------------------------
// The rollout on the table.  Plain shape (reward + terminated) or packed records, whole warps.
// kCompact (one goal object, at most 8192 states): the staged table holds 16-bit entries --
// successor * 8 (= the byte offset of its table row) | moved | (delta_goals + 1) << 1 -- half the
// shared memory, two instructions between dependent table loads (mask, add the action's offset), and
// the reward is selected from three register constants instead of a shared-memory table: the
// shared-memory pipe (action byte + table entry, the latter with ~3.4-way bank conflicts for 32
// random states) is what bounds this kernel once the arithmetic is gone.
// ------------------------------------------------------------------------------------------
template <bool kPacked, bool kCompact>
__global__ void __launch_bounds__(kTableMaxBlock, 1)
    k_rollout_table(const __grid_constant__ StepParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ uint32_t s_counters[4];
  const StateTableView& tv = p.state_table;
  using Entry = typename std::conditional<kCompact, uint16_t, uint32_t>::type;
  Entry* s_tab = reinterpret_cast<Entry*>(smem_raw);
  uint8_t* s_act = smem_raw + ((((size_t)tv.n_states * 4 * sizeof(Entry)) + 15) & ~(size_t)15);
  float* s_lut = reinterpret_cast<float*>(s_act + (blockDim.x / 32) * 2 * kTileBytes);
  if (threadIdx.x < 4) s_counters[threadIdx.x] = 0;
  for (int i = threadIdx.x; i < tv.n_states * 4; i += blockDim.x) {
    const uint32_t entry = __ldg(tv.table + i);
    if (kCompact) {
      const int dg = (int)((entry >> 16) & 0x3fu) - kMaxObjects;  // -1, 0 or +1 with one goal object
      s_tab[i] = (Entry)(((entry & 0x1fffu) << 3) | ((entry >> 24) & 1u) | ((uint32_t)(dg + 1) << 1));
    } else {
      s_tab[i] = (Entry)entry;
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
    if (kCompact) state <<= 3;
    if (!__all_sync(0xffffffffu, found)) {
      // a state the table does not hold (not reachable from the initial state): the general rule
      table_slow_path(blob, p, e, c);
    } else {
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
      // Addressing as in the ring kernel (pw_kernels.cu): a lean launch has T * N < 2^32 and N even, so
      // the element offset of (t, e) is 2 * h + (e & 1) with the 32-bit h = (e >> 1) + t * (N / 2).
      const uint32_t half_n = (uint32_t)(p.N >> 1);
      uint8_t* byte_base = (kPacked ? p.record : p.terminated) + (e & 1);
      float* const reward_base = p.reward + (e & 1);
      asm volatile("" : "+l"(byte_base));
      uint32_t h_out = (uint32_t)(e >> 1);
      uint32_t idle = record_byte(0, ach == g, false);  // the record of a no-op in the current state
      uint32_t on = ach == g ? 1u : 0u;                 // kCompact: the goal object sits on its goal
      const float r_idle = s_lut[kMaxObjects], r_lost = s_lut[kMaxObjects - 1];  // -0.01, -1.01
      uint32_t n_term = 0, n_moved = 0, n_gain = 0, n_invalid = 0;
      auto one_step = [&](const uint32_t a, auto checked) {
        constexpr bool kChecked = decltype(checked)::value;
        if (kCompact) {
          // `state` is the byte offset of the state's table row (index * 8)
          uint32_t entry = *reinterpret_cast<const uint16_t*>(reinterpret_cast<const char*>(s_tab) + state +
                                                              2u * (kChecked ? (a & 3u) : a));
          if (kChecked && a > 3u) {  // invalid action: a counted no-op
            entry = state | (1u << 1);
            ++n_invalid;
          }
          state = entry & 0xfff8u;
          const uint32_t dg1 = (entry >> 1) & 3u;  // delta_goals + 1
          n_moved += entry & 1u;
          n_gain += dg1 >> 1;
          on = on + dg1 - 1u;  // one goal object: terminated <=> it is on its goal (puzzle.py:409-411)
          n_term += on;
          if (kPacked) {
            const uint32_t code = (on << 7) + (uint32_t)(kMaxObjects - 1) + dg1;
            asm volatile("st.global.u8 [%0], %1;" ::"l"(byte_base + (size_t)h_out * 2u), "r"(code) : "memory");
          } else {
            reward_base[(size_t)h_out * 2u] = on ? 10.0f : (dg1 == 0u ? r_lost : r_idle);  // gym_env.py:212-221
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
      const int blocks = (T + 31) / 32;
      issue_block(0);
      for (int blk = 0; blk < blocks; ++blk) {
        __syncwarp();          // every lane is done with the tile the next block lands in
        issue_block(blk + 1);  // past the end: re-reads the last row, never used
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();  // block `blk` is visible to the whole warp
        const uint8_t* tile = tiles + (blk & 1) * kTileBytes;
        const uint4 w0 = *reinterpret_cast<const uint4*>(tile + lane * 32u);
        const uint4 w1 = *reinterpret_cast<const uint4*>(tile + lane * 32u + 16u);
        const uint32_t seen = w0.x | w0.y | w0.z | w0.w | w1.x | w1.y | w1.z | w1.w;
        const bool any_invalid = __any_sync(0xffffffffu, (seen & 0xfcfcfcfcu) != 0u);
        const int steps = T - blk * 32 < 32 ? T - blk * 32 : 32;
        const uint8_t* mine = tile + lane;
        if (!any_invalid && steps == 32) {
#pragma unroll 8
          for (int k = 0; k < 32; ++k) one_step(mine[k * 32], std::false_type{});
        } else {
#pragma unroll 1
          for (int k = 0; k < steps; ++k) one_step(mine[k * 32], std::true_type{});
        }
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      c.terminated += n_term;
      c.moved += n_moved;
      c.dgoals += n_gain;
      c.invalid += n_invalid;
      // positions back from the state index
      const uint64_t final_key = tv.keys[kCompact ? state >> 3 : state];
      for (int j = 0; j < n; ++j) {
        const uint32_t cell = (uint32_t)(final_key >> (j * tv.key_bits)) & ((1u << tv.key_bits) - 1u);
        p.x[j * p.stride + e] = (int)(cell % (uint32_t)W);
        p.y[j * p.stride + e] = (int)(cell / (uint32_t)W);
      }
    }
  }
  flush_counts(c, p, s_counters);
}

size_t table_smem_bytes(int n_states, bool compact, int block) {
  return ((((size_t)n_states * 4 * (compact ? 2 : 4)) + 15) & ~(size_t)15) + (size_t)(block / 32) * 2 * kTileBytes +
         256 * sizeof(float);
}

// Threads per CTA for a launch over n_envs environments.  The table is staged once per CTA and an
// environment is busy for the whole rollout, so what matters is (a) how many warps are resident per
// SM and (b) that the CTAs fill whole waves: 2^18 environments are 55.4 warps per SM, which 48
// resident warps process in 1.15 waves -- the second one nearly empty.  Picks the block size whose
// resident CTAs (occupancy API: registers + shared memory) waste the least of the last wave while
// keeping enough warps to cover the shared-memory latency.
template <class Kernel>
int pick_table_block(Kernel kernel, int64_t n_envs, int n_states, bool compact) {
  int device = 0, sms = 148;
  cudaGetDevice(&device);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  int best_block = 256;
  double best_score = -1.0;
  for (int block = 128; block <= kTableMaxBlock; block += 32) {
    const size_t smem = table_smem_bytes(n_states, compact, block);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem) != cudaSuccess || per_sm < 1) {
      cudaGetLastError();
      continue;
    }
    const double ctas = (double)((n_envs + block - 1) / block);
    const double waves = ctas / ((double)sms * per_sm);
    const double whole = waves <= 1.0 ? 1.0 : (double)(int64_t)(waves + 0.999999);
    const double fill = waves <= 1.0 ? 1.0 : waves / whole;
    const double warps = (double)per_sm * block / 32.0;
    const double cover = warps >= 28.0 ? 1.0 : warps / 28.0;
    const double score = fill * cover;
    if (score >= best_score - 1e-9) {  // ties: the larger block (the table is staged once per CTA)
      best_score = score;
      best_block = block;
    }
  }
  return best_block;
}

}  // namespace

// (this file's copy of the reward table of pw_steps.cuh, used by the general-rule fall-back)
cudaError_t upload_reward_lut_table(const float* host_lut) {
  return cudaMemcpyToSymbol(c_reward_lut, host_lut, sizeof(float) * (2 * kMaxObjects + 1));
}

// Largest table a rollout CTA can stage next to its action tiles (227 KB of shared memory per SM).
int state_table_max_states() { return 12288; }  // 192 KB of 32-bit entries + action tiles of a 256-thread CTA

// Enumerates the reachable states of the puzzle at `puzzle_off` into `tv` (device arrays allocated by
// the caller: keys[max_states], table[4 * max_states], hash_keys / hash_vals[hash_cap], zeroed hash).
// Synchronises.  Returns the number of states, or 0 if the puzzle is not tabulated.
int build_state_table(const uint64_t* d_blob, uint32_t puzzle_off, const StateTableView& tv, int max_states,
                      cudaError_t* err) {
  int* d_meta = nullptr;
  *err = cudaMalloc(&d_meta, 2 * sizeof(int));
  if (*err != cudaSuccess) return 0;
  k_build_state_table<<<1, kBuildThreads>>>(d_blob, puzzle_off, tv, max_states, d_meta);
  int meta[2] = {0, 1};
  *err = cudaMemcpy(meta, d_meta, sizeof(meta), cudaMemcpyDeviceToHost);
  cudaFree(d_meta);
  if (*err != cudaSuccess) return 0;
  return meta[1] ? 0 : meta[0];
}

bool table_rollout_eligible(const StepParams& p) {
  const auto aligned16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  const bool plain = p.reward && p.terminated && !p.record;
  const bool packed = p.record && !p.reward && !p.terminated;
  return p.state_table.n_states > 0 && p.puzzle_id == nullptr && is_lean(p) && (plain || packed) &&
         (p.N & 31) == 0 && aligned16(p.actions) && getenv("PW_NO_TABLE") == nullptr;
}

cudaError_t launch_rollout_table(const StepParams& p, cudaStream_t stream) {
  // one goal object and at most 8192 states: 16-bit entries (see k_rollout_table); PW_TABLE_WIDE=1: A/B
  const bool compact = p.state_table.goals == 1 && p.state_table.n_states <= 8192 && getenv("PW_TABLE_WIDE") == nullptr;
  const int n_states = p.state_table.n_states;
  int block = 0;
  if (const char* env = getenv("PW_TABLE_BLOCK")) block = atoi(env);  // tuning runs
#define PW_TABLE_GO(PACKED, COMPACT)                                                                      \
  do {                                                                                                    \
    if (block < 32 || block > kTableMaxBlock || (block & 31))                                             \
      block = pick_table_block(k_rollout_table<PACKED, COMPACT>, p.N, n_states, COMPACT);                 \
    k_rollout_table<PACKED, COMPACT><<<(unsigned)((p.N + block - 1) / block), block,                      \
                                       table_smem_bytes(n_states, COMPACT, block), stream>>>(p);          \
  } while (0)
  if (compact) {
    if (p.record) PW_TABLE_GO(true, true);
    else PW_TABLE_GO(false, true);
  } else {
    if (p.record) PW_TABLE_GO(true, false);
    else PW_TABLE_GO(false, false);
  }
#undef PW_TABLE_GO
  return cudaGetLastError();
}

cudaError_t configure_table_kernels() {
  const int cap = 226 * 1024;  // the staged table + action tiles of the largest block
  cudaError_t err = cudaSuccess;
#define PW_OPT_IN(k) \
  if (err == cudaSuccess) err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, cap)
  PW_OPT_IN((k_rollout_table<true, true>));
  PW_OPT_IN((k_rollout_table<false, true>));
  PW_OPT_IN((k_rollout_table<true, false>));
  PW_OPT_IN((k_rollout_table<false, false>));
#undef PW_OPT_IN
  return err;
}

}  // namespace pw
