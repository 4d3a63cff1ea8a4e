// pw_internal.h -- parameter blocks shared by the kernels and the C-ABI layer (not public).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pw {

#if defined(__CUDACC__)
// A per-environment puzzle id as the kernels use it: ids outside the engine's puzzles select puzzle
// 0 and are counted in status word 2 (pw_engine_poll) instead of indexing the tables out of bounds.
__device__ __forceinline__ int checked_puzzle_id(int pid, int num_puzzles, unsigned long long* status) {
  if ((unsigned)pid >= (unsigned)num_puzzles) {
    if (status) atomicAdd(&status[2], 1ull);
    return 0;
  }
  return pid;
}
#endif

constexpr uint32_t kFlagAutoreset = 1u;
constexpr uint32_t kFlagResample = 2u;  // autoreset draws a new puzzle (pw_rollout_episodes)

// Host-precomputed static boards of a bitboard-eligible puzzle (pw_board.cuh), one per puzzle.
struct BoardStaticPod {
  uint64_t wall, awall, edge_up, edge_down;
  uint64_t goal[8];
  uint64_t init[8];
  uint64_t act_edge[4];
  uint32_t act_sl[4], act_sr[4];
  int32_t P, n, g, bits;  // bits = 1 << 30 when the puzzle is not eligible
};

// Reachable-state transition table of one puzzle (pw_table.cu), device arrays owned by the engine.
struct StateTableView {
  uint64_t* keys;                 // [n_states] packed state of every index
  uint32_t* table;                // [n_states * 4] successor | record << 16 | moved << 24 | gained << 25
  unsigned long long* hash_keys;  // [hash_cap] open addressing, key + 1 (0 = empty)
  uint32_t* hash_vals;            // [hash_cap] index of the key
  int32_t n_states;               // 0 = the puzzle is not tabulated
  int32_t hash_cap;               // power of two
  int32_t key_bits;               // bits per object in a packed state (cell index y * W + x)
  int32_t goals;                  // the puzzle's number of goal objects
  // States are numbered so that the ones with every goal achieved come last: state s is terminal
  // (PushWorldPuzzle.is_goal_state) <=> s >= n_off.
  int32_t n_off;
  // [n_states * 16] two-step entries (state, a1, a2) -> successor * 8 | after-a1-terminal << 2 | steps that moved
  // (0..2); nullptr when the puzzle has several goal objects or too many states for a CTA to stage them
  uint16_t* pairs;
};

struct StepParams {
  const uint64_t* blob;        // all puzzle tables, device
  const uint32_t* puzzle_off;  // u64 offset of each puzzle inside blob, device
  int64_t N;
  int32_t T;
  const int32_t* puzzle_id;  // device int32[N] or nullptr
  int32_t puzzle_uniform;
  int32_t uniform_u64;  // blob size of puzzle_uniform (u64 units)
  int32_t n_max;        // planes in x / y
  int32_t board_bits;   // bits puzzle_uniform needs in board form if it is regular, else 1 << 30
  int32_t ring_goals;   // > 0: puzzle_uniform fits ring boards (pw_board.cuh), value = its goal count
  int32_t lean_tiles;   // set by launch_steps: lean shape, N % 32 == 0, 16-byte aligned actions
  int32_t occ_batch;         // set by launch_rollout_occ: parked lanes that trigger a closure pass (run-ahead kernel)
  int32_t occ_warp_closure;  // set by launch_rollout_occ (pw_occ.cu): closure per warp instead of per thread
  int32_t episode_tiles;  // set by launch_steps: pw_rollout_episodes with reward + terminated (+ truncated), whole warps
  const BoardStaticPod* board_static;  // device; entry of puzzle_uniform (valid if board_bits <= 64)
  StateTableView state_table;          // of puzzle_uniform (n_states = 0: none)
  int32_t* x;
  int32_t* y;
  int64_t stride;
  const uint8_t* actions;  // [T][N]
  float* reward;           // [T][N] or nullptr
  int8_t* dgoals;
  uint8_t* terminated;
  uint8_t* truncated;
  uint8_t* record;  // [T][N] or nullptr: packed per-step record (pw_rollout_packed), see record_byte()
  uint32_t* moved;
  int32_t* steps;  // [N] in/out or nullptr
  int32_t max_steps;
  float* return_sum;             // [N] or nullptr
  unsigned long long* counters;  // [4] or nullptr
  unsigned long long* status;    // engine status block
  uint32_t flags;
  int32_t num_puzzles;
  // pw_rollout_episodes only
  int32_t* puzzle_id_mut;
  uint32_t* episode;
  uint64_t seed;
};

struct ExpandParams {
  const uint64_t* blob;
  const uint32_t* puzzle_off;
  int32_t num_puzzles;
  unsigned long long* status;
  int64_t N;
  const int32_t* puzzle_id;
  int32_t puzzle_uniform;
  int32_t uniform_u64;
  int32_t n_max;
  const int64_t* index;  // optional: state e is column index[e] of x_in / y_in
  const int32_t* x_in;
  const int32_t* y_in;
  int64_t stride_in;
  int32_t* x_out;
  int32_t* y_out;
  int64_t stride_out;
  uint32_t* moved;
  uint8_t* goal;
};

struct ResetParams {
  const uint64_t* blob;
  const uint32_t* puzzle_off;
  int32_t num_puzzles;
  unsigned long long* status;
  int64_t N;
  const int32_t* puzzle_id;
  int32_t puzzle_uniform;
  int32_t n_max;
  int32_t* x;
  int32_t* y;
  int64_t stride;
  int32_t* steps;
  const uint8_t* reset_mask;
};

struct GoalParams {
  const uint64_t* blob;
  const uint32_t* puzzle_off;
  int32_t num_puzzles;
  unsigned long long* status;
  int64_t N;
  const int32_t* puzzle_id;
  int32_t puzzle_uniform;
  const int32_t* x;
  const int32_t* y;
  int64_t stride;
  int32_t* achieved;
  uint8_t* is_goal;
};

struct VisitedParams {
  const uint64_t* blob;
  const uint32_t* puzzle_off;
  int32_t num_puzzles;
  int32_t puzzle;
  const int32_t* x;
  const int32_t* y;
  int64_t stride;
  int64_t first_id, count;
  const uint8_t* valid;
  unsigned long long* table;
  int64_t capacity;
  uint8_t* is_new;
  unsigned long long* status;
};

struct NoveltyParams {
  const int32_t* x;
  const int32_t* y;
  int64_t stride;
  int64_t first_id, count;
  const uint32_t* moved;
  const uint8_t* valid;
  uint32_t* seen;
  uint8_t* novelty;
  int32_t n, width, height;
};

struct RenderParams {
  const uint64_t* blob;
  const uint32_t* puzzle_off;
  int32_t num_puzzles;
  unsigned long long* status;
  int64_t N;
  const int32_t* puzzle_id;
  int32_t puzzle_uniform;
  int32_t n_max;
  const int32_t* x;
  const int32_t* y;
  int64_t stride;
  void* out;
  int32_t dtype;
  int32_t ppc, bw;
  int32_t out_h, out_w;  // pixels
  int32_t max_cells;     // max width*height over the engine's puzzles
  int32_t max_blob_u64;
};

cudaError_t upload_reward_lut(const float* host_lut);
cudaError_t upload_reward_lut_table(const float* host_lut);  // pw_table.cu's copy
cudaError_t configure_kernels(size_t max_dynamic_smem);
cudaError_t configure_render_kernels();
// n_limit: max objects over the puzzles this launch can touch (selects the register policy)
cudaError_t launch_steps(const StepParams& p, int n_limit, cudaStream_t stream);
// pw_occ.cu: plain / packed rollouts of the general rule (occupancy reject + batched closure)
bool occ_rollout_eligible(const StepParams& p);
bool occ_episodes_eligible(const StepParams& p);
cudaError_t launch_rollout_occ(const StepParams& p, cudaStream_t stream);
cudaError_t configure_occ_kernels();
// pw_table.cu: rollouts of a small puzzle on its reachable-state transition table
int state_table_max_states();
int build_state_table(const uint64_t* d_blob, uint32_t puzzle_off, StateTableView* tv, int max_states,
                      cudaError_t* err);
bool table_rollout_eligible(const StepParams& p);
cudaError_t launch_rollout_table(const StepParams& p, cudaStream_t stream);
cudaError_t configure_table_kernels();
cudaError_t launch_expand(const ExpandParams& p, int n_limit, cudaStream_t stream);
cudaError_t launch_reset(const ResetParams& p, cudaStream_t stream);
cudaError_t launch_goals(const GoalParams& p, cudaStream_t stream);
cudaError_t launch_visited(const VisitedParams& p, cudaStream_t stream);
cudaError_t launch_novelty(const NoveltyParams& p, cudaStream_t stream);
cudaError_t launch_render(const RenderParams& p, const StepParams* fused_step, int n_limit,
                          cudaStream_t stream);
cudaError_t upload_palette(const uint8_t* rgb, const float* rgb_f32, int n_colors);

}  // namespace pw
