/*
 * pushworld_b200 -- C ABI of the B200-native batched PushWorld dynamics engine.
 *
 * This header is the drop-in boundary for ONE hot path of google-deepmind/pushworld:
 * PushWorldPuzzle.get_next_state + goal check / reward (+ RGB observation), batched
 * over N independent environments.  The reference has no FFI layer of its own; its
 * "operator API" for this path is the Python method set cited next to every entry
 * point below (paths relative to the reference repo root).  A maintainer binds these
 * functions with ctypes (see INTEGRATION.md for the stub).
 *
 * Conventions
 *  - plain C, no exceptions cross the ABI: every function returns 0 on success or a
 *    negative PW_E* code; pw_last_error() returns a thread-local message.
 *  - every `x`, `y`, `action`, `reward`, ... pointer is a DEVICE pointer owned by the
 *    caller (e.g. tensor.data_ptr() of a contiguous PyTorch CUDA tensor) on the
 *    engine's device.  The engine owns only the immutable puzzle tables it uploaded
 *    in pw_engine_create.  No entry point allocates or synchronises: kernels are
 *    enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default).
 *  - state layout (structure of arrays): object j of environment e lives at
 *    x[j * stride + e], y[j * stride + e] (int32), j < engine n_max, stride >= N.
 *    Index 0 is the agent, 1..g the goal objects, then the remaining movables, in the
 *    order PushWorldPuzzle.__init__ produces (python3/src/pushworld/puzzle.py:170-257).
 *  - actions: uint8, LEFT=0 RIGHT=1 UP=2 DOWN=3 (puzzle.py:35-50).  Any other value is
 *    a no-op for that environment and is counted in the engine's invalid-action
 *    counter (pw_engine_poll), which the Python shim turns into the reference's
 *    ValueError (gym_env.py:195-196).
 *  - puzzle selection: `puzzle_id` is a device int32[N] (per environment) or NULL, in
 *    which case every environment uses puzzle `puzzle_uniform`.
 */
#ifndef PUSHWORLD_B200_H_
#define PUSHWORLD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PW_ABI_VERSION 1

#if defined(__GNUC__)
#define PW_API __attribute__((visibility("default")))
#else
#define PW_API
#endif

/* limits of the table format */
#define PW_MAX_OBJECTS 32 /* movables incl. the agent (benchmark max: 19) */
#define PW_MAX_WIDTH 64   /* grid columns incl. the wall border (benchmark max: 47) */
#define PW_MAX_HEIGHT 4096

/* error codes */
#define PW_OK 0
#define PW_E_INVALID (-1)  /* bad argument */
#define PW_E_CUDA (-2)     /* CUDA runtime error, text in pw_last_error() */
#define PW_E_FORMAT (-3)   /* malformed table blob / puzzle text */
#define PW_E_LIMIT (-4)    /* puzzle exceeds PW_MAX_* */
#define PW_E_NOMEM (-5)

/* pw_render / pw_step_render output element types */
#define PW_DTYPE_U8 0  /* PushWorldPuzzle.render: uint8 RGB (puzzle.py:426-469) */
#define PW_DTYPE_F32 1 /* render_observation_padded: uint8/255 as float32 (env_utils.py:44-91) */

/* pw_rollout flags */
#define PW_FLAG_AUTORESET 1u /* re-initialise an environment after a terminated/truncated step */
#define PW_FLAG_RESAMPLE 2u  /* pw_rollout_packed: ... on a freshly drawn puzzle (as pw_rollout_episodes) */

/* the packed per-step record of pw_rollout_packed */
#define PW_RECORD_DGOALS_MASK 0x3fu /* bits 0..5: delta_goals + PW_RECORD_DGOALS_BIAS */
#define PW_RECORD_DGOALS_BIAS 32
#define PW_RECORD_TRUNCATED 0x40u
#define PW_RECORD_TERMINATED 0x80u

typedef struct PwEngine PwEngine;

/* ------------------------------------------------------------------------------------
 * Puzzle tables.  A table blob is the compiled, immutable form of one .pwp file: the
 * geometry (row bitmasks).  The reference's precomputed collision sets (puzzle.py:259-308)
 * are not part of the blob: pw_engine_create derives their bitmap equivalent from the
 * geometry on the host (milliseconds per puzzle) and keeps it on the device next to the
 * blob (pushworld_b200/csrc/pw_fast.cuh).  Layout (little endian, u64
 * aligned; all offsets in u64 units from the start of the puzzle's blob):
 *
 *   u32 word  0  magic 'PWT1' (0x31545750)
 *             1  width | height << 16          (incl. the 1-cell wall border, puzzle.py:160-168)
 *             2  n_objects | n_goals << 16
 *             3  flags: bit0 = file has agent walls ('aw'), bit1 = "regular" initial state
 *             4  offset of wall rows        u64[height], bit x of row y = wall at (x, y)
 *             5  offset of agent-wall rows  u64[height]  (aw UNION w, puzzle.py:273)
 *             6  offset of the object descriptors
 *             7  blob size in u64
 *             8..15 reserved (0)
 *   object descriptors: (n_objects + n_goals) x 4 u32; entries [n_objects, n_objects+n_goals)
 *   describe the goal outlines (shape of g<k>, drawn by render, puzzle.py:217-225,458)
 *             0  rows_off | bbox_w << 16 | bbox_h << 24      shape rows: u64[bbox_h]
 *             1  (uint16)init_x | (uint16)init_y << 16       initial_state (goal outlines: goal xy)
 *             2  (uint16)goal_x | (uint16)goal_y << 16       0x80008000 when the object has no goal
 *             3  fill palette index | border palette index << 8
 * ------------------------------------------------------------------------------------ */
#define PW_TABLE_MAGIC 0x31545750u
#define PW_TABLE_HEADER_WORDS 16
#define PW_TABLE_FLAG_HAS_AW 1u
#define PW_TABLE_FLAG_REGULAR 2u
#define PW_NO_GOAL 0x80008000u

/* Compiles the text of a .pwp file into a table blob, following PushWorldPuzzle.__init__
 * (puzzle.py:130-257): whitespace tokens, '+' joins ids, '.' empty, 1-cell border, ids
 * sorted descending, goals pull their movables to the front.  Writes at most `cap` bytes
 * to `out` and the needed size to `*size` (call with out=NULL to query).  Replaces the
 * 86 s Python table build of the reference. */
PW_API int pw_compile_puzzle(const char* text, size_t text_len, void* out, size_t cap, size_t* size);

/* Creates an engine on CUDA device `device` from `num_puzzles` concatenated table blobs.
 * offsets_u64[p] = start of puzzle p inside `blob` (u64 units).  `blob` is HOST memory
 * and is copied; the collision bit tables the kernels read are derived from it here.
 * Replaces: constructing PushWorldPuzzle objects (gym_env.py:66-69, puzzle.py:259-308). */
PW_API int pw_engine_create(const void* blob, size_t blob_bytes, const uint32_t* offsets_u64,
                     int num_puzzles, int device, PwEngine** out);
PW_API int pw_engine_destroy(PwEngine* engine);

/* Static facts: number of puzzles, max objects over all puzzles (= number of x/y planes
 * a state tensor must have), max width / height in cells. */
PW_API int pw_engine_info(const PwEngine* engine, int* num_puzzles, int* n_max, int* max_width,
                   int* max_height);

/* Copies the engine's device-side status block to host (then clears it): status[0] = invalid
 * actions seen since the last poll, status[1] = visited-set overflows (pw_visited_insert),
 * status[2] = per-environment puzzle ids outside [0, num_puzzles) (such environments are run on
 * puzzle 0; the tables are never indexed out of bounds).  SYNCHRONISES `stream`. */
PW_API int pw_engine_poll(PwEngine* engine, uint64_t status[4], void* stream);
/* Zeroes the status block on `stream` WITHOUT synchronising (drops counts of earlier launches). */
PW_API int pw_engine_clear_status(PwEngine* engine, void* stream);

/* Thread-local description of the last error on this thread ("" if none). */
PW_API const char* pw_last_error(void);
PW_API int pw_abi_version(void);

/* x/y <- initial_state of each environment's puzzle; steps <- 0 (nullable).
 * Replaces: PushWorldEnv.reset's state assignment (gym_env.py:171-176).  Planes
 * j >= n_objects(puzzle) are zero-filled.  If `reset_mask` (uint8[N]) is non-NULL only
 * environments with a non-zero mask are reset. */
PW_API int pw_reset(PwEngine* engine, int64_t n_envs, const int32_t* puzzle_id, int puzzle_uniform,
             int32_t* x, int32_t* y, int64_t stride, int32_t* steps, const uint8_t* reset_mask,
             void* stream);

/* One environment step for every environment (T = 1).
 * Replaces: PushWorldPuzzle.get_next_state (puzzle.py:348-394), is_goal_state (:409-411),
 * count_achieved_goals (:396-407) and the reward / terminated / truncated logic of
 * PushWorldEnv.step (gym_env.py:201-223, dm_env.py:204-230).
 *   reward[N]      float32: 10 on goal, else (float)(delta_goals - 0.01)   (nullable)
 *   dgoals[N]      int8: achieved goals after - before (exact integer part of the reward) (nullable)
 *   terminated[N]  uint8 (nullable)      truncated[N] uint8: steps >= max_steps (nullable)
 *   steps[N]       int32 in/out step counter (nullable);  max_steps <= 0 means "None"
 *   moved[N]       uint32 bit j = object j changed position, 0 if the move was blocked;
 *                  mirrors RelativeState.moved_object_indices (cpp/src/pushworld_puzzle.cc:439-452) (nullable)
 */
PW_API int pw_step(PwEngine* engine, int64_t n_envs, const int32_t* puzzle_id, int puzzle_uniform,
            int32_t* x, int32_t* y, int64_t stride, const uint8_t* action, float* reward,
            int8_t* dgoals, uint8_t* terminated, uint8_t* truncated, int32_t* steps,
            int32_t max_steps, uint32_t* moved, void* stream);

/* T fused steps per environment in ONE launch, state resident in registers / shared
 * memory between steps.  Equivalent to T successive pw_step calls with action[t*N + e]
 * (no reset on termination unless PW_FLAG_AUTORESET, matching the reference where
 * stepping after `terminated` keeps working, test_gym_env.py:89-102).
 *   actions     uint8[T][N]
 *   reward      float32[T][N] (nullable)   terminated uint8[T][N] (nullable)
 *   dgoals      int8[T][N]   (nullable)    truncated  uint8[T][N] (nullable)
 *   return_sum  float32[N]: += sum of the T rewards of the environment (nullable)
 *   counters    device uint64[4]: += {env-steps executed, terminated steps,
 *               steps in which anything moved, sum of max(delta_goals,0)}; this is the
 *               vector the multi-GPU driver all-reduces (nullable)
 */
PW_API int pw_rollout(PwEngine* engine, int64_t n_envs, int32_t n_steps, const int32_t* puzzle_id,
               int puzzle_uniform, int32_t* x, int32_t* y, int64_t stride,
               const uint8_t* actions, float* reward, int8_t* dgoals, uint8_t* terminated,
               uint8_t* truncated, int32_t* steps, int32_t max_steps, float* return_sum,
               uint64_t* counters, uint32_t flags, void* stream);

/* pw_rollout with the episode handling of a vector environment done entirely on the device
 * (SURVEY.md 8f-2): after a terminated or truncated step the environment draws a NEW puzzle
 * uniformly from the engine's puzzles -- what PushWorldEnv.reset does with random.choice
 * (gym_env.py:171-176) -- and restarts from that puzzle's initial state with steps = 0.
 * The draw is a counter-based hash of (seed, environment index, episode number), so a rollout is
 * reproducible and independent of launch geometry:
 *     puzzle = mulhi32(splitmix64(seed ^ env * 0x9E3779B97F4A7C15 ^ episode * 0xD1B54A32D192ED03) >> 32,
 *                      num_puzzles)
 * puzzle_id[N] (mandatory here), steps[N] and episode[N] are read and updated in place. */
PW_API int pw_rollout_episodes(PwEngine* engine, int64_t n_envs, int32_t n_steps, int32_t* puzzle_id,
                        int32_t* x, int32_t* y, int64_t stride, const uint8_t* actions,
                        float* reward, int8_t* dgoals, uint8_t* terminated, uint8_t* truncated,
                        int32_t* steps, int32_t max_steps, float* return_sum, uint64_t* counters,
                        uint64_t seed, uint32_t* episode, void* stream);

/* pw_rollout / pw_rollout_episodes writing ONE BYTE per environment step instead of the separate
 * reward / dgoals / terminated / truncated arrays: everything PushWorldEnv.step returns besides
 * the observation (gym_env.py:210-226), in the form a host-side consumer gets over PCIe at a fifth
 * of the bytes.
 *   record[T][N] uint8: bits 0..5 = delta_goals + 32, bit 6 = truncated, bit 7 = terminated
 * from which the reference's values follow exactly:
 *   reward = terminated ? 10.0 : (double)delta_goals - 0.01          (gym_env.py:212-221)
 * (pushworld_b200.vector_env.record_to_outputs is the host-side expansion).  flags:
 * PW_FLAG_AUTORESET as in pw_rollout; PW_FLAG_RESAMPLE = the episode handling of
 * pw_rollout_episodes (needs puzzle_id, steps, episode; `seed` selects the puzzle draws).
 * steps / return_sum / counters / episode are nullable as in pw_rollout. */
PW_API int pw_rollout_packed(PwEngine* engine, int64_t n_envs, int32_t n_steps, int32_t* puzzle_id,
                      int puzzle_uniform, int32_t* x, int32_t* y, int64_t stride,
                      const uint8_t* actions, uint8_t* record, int32_t* steps, int32_t max_steps,
                      float* return_sum, uint64_t* counters, uint32_t flags, uint64_t seed,
                      uint32_t* episode, void* stream);

/* Goal predicates of every environment's CURRENT state.
 * Replaces: PushWorldPuzzle.count_achieved_goals (puzzle.py:396-407) and is_goal_state
 * (puzzle.py:409-411) / satisfiesGoal (cpp/src/pushworld_puzzle.cc:462-469).
 *   achieved[N] int32 (nullable)      is_goal[N] uint8 (nullable) */
PW_API int pw_goals(PwEngine* engine, int64_t n_envs, const int32_t* puzzle_id, int puzzle_uniform,
                    const int32_t* x, const int32_t* y, int64_t stride, int32_t* achieved,
                    uint8_t* is_goal, void* stream);

/* Successor expansion: for each of N states, the 4 successors (LEFT, RIGHT, UP, DOWN).
 * Successor s = a * N + e: x_out[j * stride_out + s] etc. (stride_out >= 4N).
 * Replaces: the 4 x getNextState + satisfiesGoal calls per node of best_first_search
 * (cpp/include/search/best_first_search.h:76-92; pushworld_puzzle.cc:386-469).
 *   moved[4N] uint32 moved-object bitmask (RelativeState.moved_object_indices)
 *   goal[4N]  uint8  satisfiesGoal(successor)                           (both nullable)
 */
PW_API int pw_expand(PwEngine* engine, int64_t n_states, const int32_t* puzzle_id, int puzzle_uniform,
              const int32_t* x_in, const int32_t* y_in, int64_t stride_in, int32_t* x_out,
              int32_t* y_out, int64_t stride_out, uint32_t* moved, uint8_t* goal, void* stream);

/* pw_expand reading its inputs through an index list: state e is column index[e] of x_in / y_in
 * (a frontier scattered over a larger state arena).  index: device int64[n_states]. */
PW_API int pw_expand_indexed(PwEngine* engine, int64_t n_states, int puzzle, const int64_t* index,
                      const int32_t* x_in, const int32_t* y_in, int64_t stride_in, int32_t* x_out,
                      int32_t* y_out, int64_t stride_out, uint32_t* moved, uint8_t* goal,
                      void* stream);

/* Exact visited-set insertion for expanded frontiers (the `visited.find / insert` of
 * cpp/include/search/best_first_search.h:81-91 with StateHash, cpp/include/pushworld_puzzle.h:74-78,
 * batched).  All states of a search live in ONE caller-owned arena (SoA planes x / y, `stride`
 * columns); a state's id is its column.  `table` is a caller-owned open-addressing hash set:
 * uint64[capacity], capacity a power of two, zero-initialised = empty; an entry packs a 24-bit
 * fingerprint and the id of the first state inserted with that content.  For the `count` states
 * with ids first_id .. first_id+count-1 (only those with valid[i] != 0 if `valid` is non-NULL):
 *   is_new[i] = 1 if no equal state was in the set (the state is then inserted), else 0.
 * Equality is checked on the full positions, never on the hash alone, so the result is exact;
 * among equal states of one call exactly one is reported new.  If the table fills up the
 * engine's status word 1 (pw_engine_poll) is incremented and the state is reported not new. */
PW_API int pw_visited_insert(PwEngine* engine, int puzzle, const int32_t* x, const int32_t* y,
                      int64_t stride, int64_t first_id, int64_t count, const uint8_t* valid,
                      uint64_t* table, int64_t capacity, uint8_t* is_new, void* stream);

/* Batched novelty measure of width-based search (Lipovetzky & Geffner 2017), the heuristic the
 * reference's "N+RGD" planner mode combines with RGD (cpp/src/heuristics/novelty.cc:30-77,
 * cpp/src/run_planner.cc:47-57).  `seen` is a caller-owned, zero-initialised bit table of
 * pw_novelty_table_words() uint32 words holding, for the puzzle's n objects and C = width*height
 * grid positions, one bit per (object, position) and one bit per (object i < object j,
 * position_i, position_j).  For the `count` states with ids first_id .. first_id+count-1 of a
 * state arena (see pw_visited_insert) and their moved-object masks (as written by pw_expand):
 *   novelty[i] = 1  if a moved object is at a position never seen for that object,
 *                2  else if a pair (moved object, any other object) is in an unseen combination,
 *                3  otherwise (also when nothing moved),
 * judged against the table as it was BEFORE the call; afterwards the positions / pairs of all
 * `count` states are inserted.  With count = 1 per call this is exactly the reference's
 * sequential algorithm; larger batches evaluate a whole search layer against the previous
 * layers.  Objects outside the grid are ignored.  valid[i] == 0 (optional) skips a state. */
PW_API int pw_novelty_table_words(const PwEngine* engine, int puzzle, uint64_t* n_words);
PW_API int pw_novelty(PwEngine* engine, int puzzle, const int32_t* x, const int32_t* y, int64_t stride,
               int64_t first_id, int64_t count, const uint32_t* moved, const uint8_t* valid,
               uint32_t* seen, uint8_t* novelty, void* stream);

/* RGB image of every environment's state.
 * Replaces: PushWorldPuzzle.render + _draw_object (puzzle.py:426-469,596-638) and, with
 * dtype F32 and padding, render_observation_padded (env_utils.py:44-91).
 *   out: [N][out_h][out_w][3], out_h = pad_h_cells*ppc, out_w = pad_w_cells*ppc; the
 *   puzzle image (height*ppc x width*ppc) is centred with floor(pad/2) pixels on the
 *   top/left and zeros elsewhere.  pad_*_cells must be >= every rendered puzzle's dims.
 *   pixels_per_cell >= 1 + 2*border_width, border_width >= 1 (puzzle.py:444-448).
 */
PW_API int pw_render(PwEngine* engine, int64_t n_envs, const int32_t* puzzle_id, int puzzle_uniform,
              const int32_t* x, const int32_t* y, int64_t stride, void* out, int dtype,
              int pixels_per_cell, int border_width, int pad_h_cells, int pad_w_cells,
              void* stream);

/* pw_step and pw_render of the new state in one call (BASELINE config 5; the step kernel and
 * the render kernel are launched back to back on `stream`):
 * what PushWorldEnv.step returns -- observation, reward, terminated, truncated
 * (gym_env.py:188-226). */
PW_API int pw_step_render(PwEngine* engine, int64_t n_envs, const int32_t* puzzle_id,
                   int puzzle_uniform, int32_t* x, int32_t* y, int64_t stride,
                   const uint8_t* action, float* reward, int8_t* dgoals, uint8_t* terminated,
                   uint8_t* truncated, int32_t* steps, int32_t max_steps, void* out, int dtype,
                   int pixels_per_cell, int border_width, int pad_h_cells, int pad_w_cells,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PUSHWORLD_B200_H_ */
