// pw_abi.cu -- the extern "C" boundary declared in include/pushworld_b200.h: argument
// validation, engine lifetime, and translation into kernel parameter blocks.  No function here
// allocates device memory after pw_engine_create or synchronises (except pw_engine_poll).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/pushworld_b200.h"
#include "pw_board.cuh"
#include "pw_fast.cuh"
#include "pw_core.cuh"
#include "pw_host.h"
#include "pw_internal.h"

namespace pw {

static thread_local std::string t_error;

int set_error(int code, const std::string& message) {
  t_error = message;
  return code;
}
void clear_error() { t_error.clear(); }

static int cuda_error(cudaError_t err, const char* what) {
  return set_error(PW_E_CUDA, std::string(what) + ": " + cudaGetErrorString(err));
}

struct PuzzleInfo {
  int width, height, n, g;
  int size_u64;   // geometry blob + collision bit tables, as stored on the device
  int blob_u64;   // geometry blob alone (what the caller passed in)
  uint32_t flags;
};

class DeviceGuard {
 public:
  explicit DeviceGuard(int device) : prev_(-1), changed_(false) {
    if (cudaGetDevice(&prev_) == cudaSuccess && prev_ != device) {
      changed_ = cudaSetDevice(device) == cudaSuccess;
    }
  }
  ~DeviceGuard() {
    if (changed_) cudaSetDevice(prev_);
  }

 private:
  int prev_;
  bool changed_;
};

}  // namespace pw

struct PwEngine {
  int device = 0;
  int n_max = 0, max_w = 0, max_h = 0, max_cells = 0, max_blob_u64 = 0;
  int max_packed_u64 = 0;  // largest geometry blob + collision tables (shared-memory staging)
  std::vector<pw::PuzzleInfo> puzzles;
  uint64_t* d_blob = nullptr;
  uint32_t* d_off = nullptr;
  unsigned long long* d_status = nullptr;  // [4]
  pw::BoardStaticPod* d_board_static = nullptr;  // [num_puzzles]
  std::vector<pw::StateTableView> state_tables;   // per puzzle; n_states = 0: not tabulated
  std::vector<int> board_bits;                    // per puzzle, 1 << 30 = not eligible
  int board_bits_all = 1 << 30;                   // max over all puzzles (mixed board kernel)
};

using namespace pw;

namespace {

int check_common(const PwEngine* e, int64_t n, const int32_t* puzzle_id, int puzzle_uniform,
                 int64_t stride, int64_t min_stride) {
  if (e == nullptr) return set_error(PW_E_INVALID, "engine is NULL");
  if (n < 0) return set_error(PW_E_INVALID, "negative environment count");
  if (n > (int64_t)0x7fffffff * 128) return set_error(PW_E_INVALID, "too many environments");
  if (puzzle_id == nullptr && (puzzle_uniform < 0 || puzzle_uniform >= (int)e->puzzles.size()))
    return set_error(PW_E_INVALID, "puzzle_uniform out of range");
  if (stride < min_stride) return set_error(PW_E_INVALID, "stride smaller than the batch");
  return PW_OK;
}

// An empty batch is a successful no-op whatever the pointers are (empty tensors have NULL data).
#define PW_RETURN_IF_EMPTY(n) \
  if ((n) == 0) return PW_OK

// Host view of a table blob for the host-side use of pw_board.cuh.
struct HostTable {
  const uint64_t* base;
  PW_HD uint64_t row(int off) const { return base[off]; }
  PW_HD uint32_t word(int i) const { return reinterpret_cast<const uint32_t*>(base)[i]; }
};

// initial x / y of object k, read from the descriptors
struct InitCoord {
  HostTable t;
  int desc_word;
  bool want_y;
  PW_HD int operator()(int k) const {
    const uint32_t v = t.word(desc_word + 4 * k + 1);
    return want_y ? hi16(v) : lo16(v);
  }
};

// bitboard kernels (pw_board.cuh): the interior fits one machine word, n <= 8 and the puzzle
// starts in a regular state.  Fills `pod` (bits = 1 << 30 when not eligible).
void make_board_static(const PuzzleInfo& info, const uint64_t* blob, BoardStaticPod* pod) {
  memset(pod, 0, sizeof(*pod));
  pod->bits = 1 << 30;
  if (!(info.flags & PW_TABLE_FLAG_REGULAR) || info.n > 8) return;
  const int bits = board_bits(info.width, info.height);
  if (bits > 64) return;
  HostTable t{blob};
  Board<8, uint64_t> s;
  board_init(s, t);
  const int desc_word = (int)t.word(kHdrObjOff) * 2;
  const bool ok = board_load(s, t, InitCoord{t, desc_word, false}, InitCoord{t, desc_word, true});
  if (!ok) return;
  BoardAction<uint64_t> act[4];
  board_action_table(s, act);
  pod->wall = s.wall;
  pod->awall = s.awall;
  pod->edge_up = s.edge_up;
  pod->edge_down = s.edge_down;
  for (int k = 0; k < 8; ++k) {
    pod->goal[k] = s.goal[k];
    pod->init[k] = s.b[k];
  }
  for (int a = 0; a < 4; ++a) {
    pod->act_edge[a] = act[a].edge;
    pod->act_sl[a] = act[a].sl;
    pod->act_sr[a] = act[a].sr;
  }
  pod->P = s.P;
  pod->n = s.n;
  pod->g = s.g;
  pod->bits = bits;
}

int n_limit_of(const PwEngine* e, const int32_t* puzzle_id, int puzzle_uniform) {
  return puzzle_id ? e->n_max : e->puzzles[(size_t)puzzle_uniform].n;
}

}  // namespace

extern "C" {

int pw_abi_version(void) { return PW_ABI_VERSION; }

const char* pw_last_error(void) { return t_error.c_str(); }

int pw_compile_puzzle(const char* text, size_t text_len, void* out, size_t cap, size_t* size) {
  clear_error();
  if (text == nullptr || size == nullptr) return set_error(PW_E_INVALID, "NULL argument");
  std::vector<uint8_t> blob;
  const int rc = compile_puzzle(text, text_len, &blob);
  if (rc != PW_OK) return rc;
  *size = blob.size();
  if (out != nullptr) {
    if (cap < blob.size()) return set_error(PW_E_INVALID, "output buffer too small");
    memcpy(out, blob.data(), blob.size());
  }
  return PW_OK;
}

int pw_engine_create(const void* blob, size_t blob_bytes, const uint32_t* offsets_u64,
                     int num_puzzles, int device, PwEngine** out) {
  clear_error();
  if (blob == nullptr || offsets_u64 == nullptr || out == nullptr || num_puzzles <= 0)
    return set_error(PW_E_INVALID, "NULL / empty argument");
  if (blob_bytes % 8 != 0) return set_error(PW_E_FORMAT, "blob size is not a multiple of 8");
  PwEngine* e = new (std::nothrow) PwEngine();
  if (e == nullptr) return set_error(PW_E_NOMEM, "out of host memory");
  e->device = device;
  const size_t total_u64 = blob_bytes / 8;
  const uint8_t* bytes = static_cast<const uint8_t*>(blob);
  for (int p = 0; p < num_puzzles; ++p) {
    const size_t off = offsets_u64[p];
    if (off + PW_TABLE_HEADER_WORDS / 2 > total_u64) {
      delete e;
      return set_error(PW_E_FORMAT, "puzzle offset outside the blob");
    }
    uint32_t h[PW_TABLE_HEADER_WORDS];
    memcpy(h, bytes + off * 8, sizeof(h));
    PuzzleInfo info;
    info.width = (int)(h[1] & 0xffffu);
    info.height = (int)(h[1] >> 16);
    info.n = (int)(h[2] & 0xffffu);
    info.g = (int)(h[2] >> 16);
    info.flags = h[3];
    info.size_u64 = (int)h[7];
    info.blob_u64 = (int)h[7];
    const size_t n_desc = (size_t)info.n + (size_t)info.g;
    bool ok = h[0] == PW_TABLE_MAGIC && info.n >= 1 && info.n <= PW_MAX_OBJECTS &&
              info.g < info.n &&  /* goal objects are movables 1..g: g <= n - 1 */ info.width >= 3 && info.width <= PW_MAX_WIDTH &&
              info.height >= 3 && info.height <= PW_MAX_HEIGHT &&
              off + (size_t)info.size_u64 <= total_u64 && h[6] == PW_TABLE_HEADER_WORDS / 2 &&
              (size_t)h[6] + n_desc * 2 <= (size_t)info.size_u64 &&
              (size_t)h[4] + (size_t)info.height <= (size_t)info.size_u64 &&
              (size_t)h[5] + (size_t)info.height <= (size_t)info.size_u64;
    if (ok) {  // every shape's rows must lie inside the blob and inside 64 columns
      const uint8_t* d = bytes + (off + h[6]) * 8;
      for (size_t k = 0; k < n_desc && ok; ++k) {
        uint32_t w0;
        memcpy(&w0, d + k * 16, 4);
        const size_t ro = w0 & 0xffffu, bw = (w0 >> 16) & 0xffu, bh = w0 >> 24;
        ok = bw >= 1 && bw <= PW_MAX_WIDTH && bh >= 1 && ro + bh <= (size_t)info.size_u64;
      }
    }
    if (!ok) {
      delete e;
      return set_error(PW_E_FORMAT, "malformed table for puzzle " + std::to_string(p));
    }
    e->puzzles.push_back(info);
    e->n_max = info.n > e->n_max ? info.n : e->n_max;
    e->max_w = info.width > e->max_w ? info.width : e->max_w;
    e->max_h = info.height > e->max_h ? info.height : e->max_h;
    const int cells = info.width * info.height;
    e->max_cells = cells > e->max_cells ? cells : e->max_cells;
    e->max_blob_u64 = info.size_u64 > e->max_blob_u64 ? info.size_u64 : e->max_blob_u64;
  }

  std::vector<BoardStaticPod> pods((size_t)num_puzzles);
  for (int p = 0; p < num_puzzles; ++p) {
    make_board_static(e->puzzles[(size_t)p],
                      reinterpret_cast<const uint64_t*>(bytes) + offsets_u64[p], &pods[(size_t)p]);
    e->board_bits.push_back(pods[(size_t)p].bits);
  }
  e->board_bits_all = 0;
  for (int b : e->board_bits) e->board_bits_all = b > e->board_bits_all ? b : e->board_bits_all;

  // Device layout of a puzzle: its geometry blob, immediately followed by the collision bit tables
  // derived from it (pw_fast.cuh).  Kernels find the tables at base + header size.
  std::vector<uint64_t> packed;
  std::vector<uint32_t> packed_off((size_t)num_puzzles);
  e->max_packed_u64 = 0;
  for (int p = 0; p < num_puzzles; ++p) {
    PuzzleInfo& info = e->puzzles[(size_t)p];
    const uint64_t* src = reinterpret_cast<const uint64_t*>(bytes) + offsets_u64[p];
    packed_off[(size_t)p] = (uint32_t)packed.size();
    packed.insert(packed.end(), src, src + info.blob_u64);
    build_fast_table(src, &packed);
    info.size_u64 = (int)(packed.size() - packed_off[(size_t)p]);
    e->max_packed_u64 = info.size_u64 > e->max_packed_u64 ? info.size_u64 : e->max_packed_u64;
  }
  if (packed.size() > 0xffffffffull) {
    delete e;
    return set_error(PW_E_LIMIT, "puzzle tables exceed 32 GiB");
  }
  blob = packed.data();
  blob_bytes = packed.size() * 8;
  offsets_u64 = packed_off.data();

  DeviceGuard guard(device);
  cudaError_t err = cudaSetDevice(device);
  if (err == cudaSuccess) err = cudaMalloc(&e->d_blob, blob_bytes);
  if (err == cudaSuccess) err = cudaMalloc(&e->d_board_static, sizeof(BoardStaticPod) * pods.size());
  if (err == cudaSuccess)
    err = cudaMemcpy(e->d_board_static, pods.data(), sizeof(BoardStaticPod) * pods.size(),
                     cudaMemcpyHostToDevice);
  if (err == cudaSuccess) err = cudaMalloc(&e->d_off, sizeof(uint32_t) * (size_t)num_puzzles);
  if (err == cudaSuccess) err = cudaMalloc(&e->d_status, sizeof(unsigned long long) * 4);
  if (err == cudaSuccess) err = cudaMemcpy(e->d_blob, blob, blob_bytes, cudaMemcpyHostToDevice);
  if (err == cudaSuccess)
    err = cudaMemcpy(e->d_off, offsets_u64, sizeof(uint32_t) * (size_t)num_puzzles,
                     cudaMemcpyHostToDevice);
  if (err == cudaSuccess) err = cudaMemset(e->d_status, 0, sizeof(unsigned long long) * 4);
  if (err == cudaSuccess) {
    // gym_env.py:221 computes (int - int - 0.01) in float64; round to float32 once, here.
    float lut[2 * kMaxObjects + 1];
    for (int dg = -kMaxObjects; dg <= kMaxObjects; ++dg)
      lut[dg + kMaxObjects] = (float)((double)dg - 0.01);
    err = upload_reward_lut(lut);
    if (err == cudaSuccess) err = upload_reward_lut_table(lut);
  }
  if (err == cudaSuccess) {
    // puzzle.py:65-79; index 0 = background 255, 12 = zero padding (env_utils.py:84-91)
    static const uint8_t pal[13][3] = {
        {255, 255, 255},    {0x00, 0xDC, 0x00}, {0x00, 0x6E, 0x00}, {0xFA, 0xC7, 0x1E},
        {0x7D, 0x64, 0x0F}, {0xB9, 0x00, 0x00}, {0xDC, 0x00, 0x00}, {0x6E, 0x00, 0x00},
        {0x46, 0x9B, 0xFF}, {0x23, 0x48, 0x7F}, {0x0A, 0x0A, 0x0A}, {0x05, 0x05, 0x05},
        {0, 0, 0}};
    float palf[13][3];
    // env_utils.py:66-73: uint8 -> float32, then a true float32 division by 255
    for (int i = 0; i < 13; ++i)
      for (int c = 0; c < 3; ++c) palf[i][c] = (float)pal[i][c] / 255.0f;
    err = upload_palette(&pal[0][0], &palf[0][0], 13);
  }
  if (err == cudaSuccess) {
    // the opt-in is per process and device, not per engine: always ask for the staging limit, so a
    // later, smaller engine never lowers what an earlier one needs
    err = configure_kernels((size_t)1 << 30);
  }
  if (err == cudaSuccess) err = configure_occ_kernels();
  if (err == cudaSuccess) err = configure_table_kernels();
  // Reachable-state transition tables (pw_table.cu) for engines over a few small puzzles -- the
  // single-puzzle engines of the drop-in environments and of BASELINE config 2.  Enumerated on the
  // device with the exact rule; a puzzle with more states than a CTA can stage is simply not tabulated.
  e->state_tables.assign((size_t)num_puzzles, StateTableView{});
  if (err == cudaSuccess && num_puzzles <= 4 && getenv("PW_NO_TABLE") == nullptr) {
    const int max_states = state_table_max_states();
    for (int p = 0; p < num_puzzles && err == cudaSuccess; ++p) {
      const PuzzleInfo& info = e->puzzles[(size_t)p];
      int bits = 1;
      while ((1 << bits) < info.width * info.height) ++bits;
      if (info.n * bits > 63 || info.width * info.height > 4096) continue;
      StateTableView tv{};
      tv.key_bits = bits;
      tv.goals = info.g;
      tv.hash_cap = 1 << 16;  // >= 4 x the largest table
      if (err == cudaSuccess) err = cudaMalloc(&tv.keys, sizeof(uint64_t) * (size_t)max_states);
      if (err == cudaSuccess) err = cudaMalloc(&tv.table, sizeof(uint32_t) * 4 * (size_t)max_states);
      if (err == cudaSuccess) err = cudaMalloc(&tv.hash_keys, sizeof(unsigned long long) * (size_t)tv.hash_cap);
      if (err == cudaSuccess) err = cudaMalloc(&tv.hash_vals, sizeof(uint32_t) * (size_t)tv.hash_cap);
      if (err == cudaSuccess) err = cudaMemset(tv.hash_keys, 0, sizeof(unsigned long long) * (size_t)tv.hash_cap);
      if (err == cudaSuccess) tv.n_states = build_state_table(e->d_blob, packed_off[(size_t)p], &tv, max_states, &err);
      if (err != cudaSuccess || tv.n_states == 0) {
        cudaFree(tv.keys);
        cudaFree(tv.table);
        cudaFree(tv.hash_keys);
        cudaFree(tv.hash_vals);
        if (tv.pairs) cudaFree(tv.pairs);
        tv = StateTableView{};
      }
      e->state_tables[(size_t)p] = tv;
    }
  }
  if (err == cudaSuccess) err = configure_render_kernels();
  if (err != cudaSuccess) {
    const int rc = cuda_error(err, "pw_engine_create");
    pw_engine_destroy(e);
    return rc;
  }
  *out = e;
  return PW_OK;
}

int pw_engine_destroy(PwEngine* e) {
  if (e == nullptr) return PW_OK;
  DeviceGuard guard(e->device);
  if (e->d_blob) cudaFree(e->d_blob);
  if (e->d_off) cudaFree(e->d_off);
  if (e->d_status) cudaFree(e->d_status);
  if (e->d_board_static) cudaFree(e->d_board_static);
  for (const StateTableView& tv : e->state_tables) {
    if (tv.keys) cudaFree(tv.keys);
    if (tv.table) cudaFree(tv.table);
    if (tv.hash_keys) cudaFree(tv.hash_keys);
    if (tv.hash_vals) cudaFree(tv.hash_vals);
    if (tv.pairs) cudaFree(tv.pairs);
  }
  delete e;
  return PW_OK;
}

int pw_engine_info(const PwEngine* e, int* num_puzzles, int* n_max, int* max_width,
                   int* max_height) {
  if (e == nullptr) return set_error(PW_E_INVALID, "engine is NULL");
  if (num_puzzles) *num_puzzles = (int)e->puzzles.size();
  if (n_max) *n_max = e->n_max;
  if (max_width) *max_width = e->max_w;
  if (max_height) *max_height = e->max_h;
  return PW_OK;
}

int pw_engine_poll(PwEngine* e, uint64_t status[4], void* stream) {
  clear_error();
  if (e == nullptr || status == nullptr) return set_error(PW_E_INVALID, "NULL argument");
  DeviceGuard guard(e->device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  unsigned long long host[4];
  cudaError_t err = cudaMemcpyAsync(host, e->d_status, sizeof(host), cudaMemcpyDeviceToHost, s);
  if (err == cudaSuccess) err = cudaMemsetAsync(e->d_status, 0, sizeof(host), s);
  if (err == cudaSuccess) err = cudaStreamSynchronize(s);
  if (err != cudaSuccess) return cuda_error(err, "pw_engine_poll");
  for (int i = 0; i < 4; ++i) status[i] = host[i];
  return PW_OK;
}

int pw_engine_clear_status(PwEngine* e, void* stream) {
  clear_error();
  if (e == nullptr) return set_error(PW_E_INVALID, "engine is NULL");
  DeviceGuard guard(e->device);
  const cudaError_t err = cudaMemsetAsync(e->d_status, 0, sizeof(unsigned long long) * 4,
                                          static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? PW_OK : cuda_error(err, "pw_engine_clear_status");
}

int pw_reset(PwEngine* e, int64_t n_envs, const int32_t* puzzle_id, int puzzle_uniform, int32_t* x,
             int32_t* y, int64_t stride, int32_t* steps, const uint8_t* reset_mask, void* stream) {
  clear_error();
  const int rc = check_common(e, n_envs, puzzle_id, puzzle_uniform, stride, n_envs);
  if (rc != PW_OK) return rc;
  PW_RETURN_IF_EMPTY(n_envs);
  if (x == nullptr || y == nullptr) return set_error(PW_E_INVALID, "x / y is NULL");
  DeviceGuard guard(e->device);
  ResetParams p{};
  p.blob = e->d_blob;
  p.puzzle_off = e->d_off;
  p.num_puzzles = (int32_t)e->puzzles.size();
  p.status = e->d_status;
  p.N = n_envs;
  p.puzzle_id = puzzle_id;
  p.puzzle_uniform = puzzle_uniform;
  p.n_max = e->n_max;
  p.x = x;
  p.y = y;
  p.stride = stride;
  p.steps = steps;
  p.reset_mask = reset_mask;
  const cudaError_t err = launch_reset(p, static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? PW_OK : cuda_error(err, "pw_reset");
}

int pw_rollout(PwEngine* e, int64_t n_envs, int32_t n_steps, const int32_t* puzzle_id,
               int puzzle_uniform, int32_t* x, int32_t* y, int64_t stride, const uint8_t* actions,
               float* reward, int8_t* dgoals, uint8_t* terminated, uint8_t* truncated,
               int32_t* steps, int32_t max_steps, float* return_sum, uint64_t* counters,
               uint32_t flags, void* stream) {
  clear_error();
  const int rc = check_common(e, n_envs, puzzle_id, puzzle_uniform, stride, n_envs);
  if (rc != PW_OK) return rc;
  PW_RETURN_IF_EMPTY(n_envs);
  if (n_steps < 0) return set_error(PW_E_INVALID, "negative step count");
  if (x == nullptr || y == nullptr || actions == nullptr)
    return set_error(PW_E_INVALID, "x / y / actions is NULL");
  if (flags & ~PW_FLAG_AUTORESET) return set_error(PW_E_INVALID, "unknown flag");
  DeviceGuard guard(e->device);
  StepParams p{};
  p.blob = e->d_blob;
  p.puzzle_off = e->d_off;
  p.num_puzzles = (int32_t)e->puzzles.size();
  p.status = e->d_status;
  p.N = n_envs;
  p.T = n_steps;
  p.puzzle_id = puzzle_id;
  p.puzzle_uniform = puzzle_id ? 0 : puzzle_uniform;
  p.uniform_u64 = puzzle_id ? 0 : e->puzzles[(size_t)puzzle_uniform].size_u64;
  p.n_max = e->n_max;
  p.board_bits = puzzle_id ? e->board_bits_all : e->board_bits[(size_t)puzzle_uniform];
  if (!puzzle_id) {
    const PuzzleInfo& info = e->puzzles[(size_t)puzzle_uniform];
    if (p.board_bits <= 32 && ring_fits(p.board_bits, info.width - 1)) p.ring_goals = info.g;
  }
  p.board_static = puzzle_id ? e->d_board_static : e->d_board_static + puzzle_uniform;
  if (!puzzle_id) p.state_table = e->state_tables[(size_t)puzzle_uniform];
  p.x = x;
  p.y = y;
  p.stride = stride;
  p.actions = actions;
  p.reward = reward;
  p.dgoals = dgoals;
  p.terminated = terminated;
  p.truncated = truncated;
  p.moved = nullptr;
  p.steps = steps;
  p.max_steps = max_steps;
  p.return_sum = return_sum;
  p.counters = reinterpret_cast<unsigned long long*>(counters);
  p.flags = flags;
  const cudaError_t err = launch_steps(p, n_limit_of(e, puzzle_id, puzzle_uniform),
                                       static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? PW_OK : cuda_error(err, "pw_rollout");
}

int pw_rollout_episodes(PwEngine* e, int64_t n_envs, int32_t n_steps, int32_t* puzzle_id, int32_t* x,
                        int32_t* y, int64_t stride, const uint8_t* actions, float* reward,
                        int8_t* dgoals, uint8_t* terminated, uint8_t* truncated, int32_t* steps,
                        int32_t max_steps, float* return_sum, uint64_t* counters, uint64_t seed,
                        uint32_t* episode, void* stream) {
  clear_error();
  if (puzzle_id == nullptr) return set_error(PW_E_INVALID, "puzzle_id is mandatory");
  const int rc = check_common(e, n_envs, puzzle_id, 0, stride, n_envs);
  if (rc != PW_OK) return rc;
  PW_RETURN_IF_EMPTY(n_envs);
  if (n_steps < 0) return set_error(PW_E_INVALID, "negative step count");
  if (x == nullptr || y == nullptr || actions == nullptr || episode == nullptr || steps == nullptr)
    return set_error(PW_E_INVALID, "x / y / actions / steps / episode is NULL");
  DeviceGuard guard(e->device);
  StepParams p{};
  p.blob = e->d_blob;
  p.puzzle_off = e->d_off;
  p.num_puzzles = (int32_t)e->puzzles.size();
  p.status = e->d_status;
  p.N = n_envs;
  p.T = n_steps;
  p.puzzle_id = puzzle_id;
  p.n_max = e->n_max;
  p.board_bits = e->board_bits_all;
  p.board_static = e->d_board_static;
  p.x = x;
  p.y = y;
  p.stride = stride;
  p.actions = actions;
  p.reward = reward;
  p.dgoals = dgoals;
  p.terminated = terminated;
  p.truncated = truncated;
  p.steps = steps;
  p.max_steps = max_steps;
  p.return_sum = return_sum;
  p.counters = reinterpret_cast<unsigned long long*>(counters);
  p.status = e->d_status;
  p.flags = kFlagAutoreset | kFlagResample;
  p.puzzle_id_mut = puzzle_id;
  p.episode = episode;
  p.seed = seed;
  p.num_puzzles = (int32_t)e->puzzles.size();
  const cudaError_t err = launch_steps(p, e->n_max, static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? PW_OK : cuda_error(err, "pw_rollout_episodes");
}

int pw_rollout_packed(PwEngine* e, int64_t n_envs, int32_t n_steps, int32_t* puzzle_id,
                      int puzzle_uniform, int32_t* x, int32_t* y, int64_t stride, const uint8_t* actions,
                      uint8_t* record, int32_t* steps, int32_t max_steps, float* return_sum,
                      uint64_t* counters, uint32_t flags, uint64_t seed, uint32_t* episode, void* stream) {
  clear_error();
  const int rc = check_common(e, n_envs, puzzle_id, puzzle_uniform, stride, n_envs);
  if (rc != PW_OK) return rc;
  PW_RETURN_IF_EMPTY(n_envs);
  if (n_steps < 0) return set_error(PW_E_INVALID, "negative step count");
  if (x == nullptr || y == nullptr || actions == nullptr || record == nullptr)
    return set_error(PW_E_INVALID, "x / y / actions / record is NULL");
  if (flags & ~(PW_FLAG_AUTORESET | PW_FLAG_RESAMPLE)) return set_error(PW_E_INVALID, "unknown flag");
  const bool resample = (flags & PW_FLAG_RESAMPLE) != 0;
  if (resample && (puzzle_id == nullptr || steps == nullptr || episode == nullptr))
    return set_error(PW_E_INVALID, "PW_FLAG_RESAMPLE needs puzzle_id, steps and episode");
  DeviceGuard guard(e->device);
  StepParams p{};
  p.blob = e->d_blob;
  p.puzzle_off = e->d_off;
  p.num_puzzles = (int32_t)e->puzzles.size();
  p.status = e->d_status;
  p.N = n_envs;
  p.T = n_steps;
  p.puzzle_id = puzzle_id;
  p.puzzle_uniform = puzzle_id ? 0 : puzzle_uniform;
  p.uniform_u64 = puzzle_id ? 0 : e->puzzles[(size_t)puzzle_uniform].size_u64;
  p.n_max = e->n_max;
  p.board_bits = puzzle_id ? e->board_bits_all : e->board_bits[(size_t)puzzle_uniform];
  if (!puzzle_id) {
    const PuzzleInfo& info = e->puzzles[(size_t)puzzle_uniform];
    if (p.board_bits <= 32 && ring_fits(p.board_bits, info.width - 1)) p.ring_goals = info.g;
  }
  p.board_static = puzzle_id ? e->d_board_static : e->d_board_static + puzzle_uniform;
  if (!puzzle_id) p.state_table = e->state_tables[(size_t)puzzle_uniform];
  p.x = x;
  p.y = y;
  p.stride = stride;
  p.actions = actions;
  p.record = record;
  p.steps = steps;
  p.max_steps = max_steps;
  p.return_sum = return_sum;
  p.counters = reinterpret_cast<unsigned long long*>(counters);
  p.status = e->d_status;
  p.flags = resample ? (kFlagAutoreset | kFlagResample) : (flags & PW_FLAG_AUTORESET ? kFlagAutoreset : 0u);
  if (resample) {
    p.puzzle_id_mut = puzzle_id;
    p.episode = episode;
    p.seed = seed;
  }
  const cudaError_t err = launch_steps(p, resample ? e->n_max : n_limit_of(e, puzzle_id, puzzle_uniform),
                                       static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? PW_OK : cuda_error(err, "pw_rollout_packed");
}

int pw_step(PwEngine* e, int64_t n_envs, const int32_t* puzzle_id, int puzzle_uniform, int32_t* x,
            int32_t* y, int64_t stride, const uint8_t* action, float* reward, int8_t* dgoals,
            uint8_t* terminated, uint8_t* truncated, int32_t* steps, int32_t max_steps,
            uint32_t* moved, void* stream) {
  clear_error();
  const int rc = check_common(e, n_envs, puzzle_id, puzzle_uniform, stride, n_envs);
  if (rc != PW_OK) return rc;
  PW_RETURN_IF_EMPTY(n_envs);
  if (x == nullptr || y == nullptr || action == nullptr)
    return set_error(PW_E_INVALID, "x / y / action is NULL");
  DeviceGuard guard(e->device);
  StepParams p{};
  p.blob = e->d_blob;
  p.puzzle_off = e->d_off;
  p.num_puzzles = (int32_t)e->puzzles.size();
  p.status = e->d_status;
  p.N = n_envs;
  p.T = 1;
  p.puzzle_id = puzzle_id;
  p.puzzle_uniform = puzzle_id ? 0 : puzzle_uniform;
  p.uniform_u64 = puzzle_id ? 0 : e->puzzles[(size_t)puzzle_uniform].size_u64;
  p.n_max = e->n_max;
  p.board_bits = puzzle_id ? e->board_bits_all : e->board_bits[(size_t)puzzle_uniform];
  p.board_static = puzzle_id ? e->d_board_static : e->d_board_static + puzzle_uniform;
  p.x = x;
  p.y = y;
  p.stride = stride;
  p.actions = action;
  p.reward = reward;
  p.dgoals = dgoals;
  p.terminated = terminated;
  p.truncated = truncated;
  p.moved = moved;
  p.steps = steps;
  p.max_steps = max_steps;
  p.return_sum = nullptr;
  p.counters = nullptr;
  p.status = e->d_status;
  p.flags = 0;
  const cudaError_t err = launch_steps(p, n_limit_of(e, puzzle_id, puzzle_uniform),
                                       static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? PW_OK : cuda_error(err, "pw_step");
}

int pw_goals(PwEngine* e, int64_t n_envs, const int32_t* puzzle_id, int puzzle_uniform,
             const int32_t* x, const int32_t* y, int64_t stride, int32_t* achieved, uint8_t* is_goal,
             void* stream) {
  clear_error();
  const int rc = check_common(e, n_envs, puzzle_id, puzzle_uniform, stride, n_envs);
  if (rc != PW_OK) return rc;
  PW_RETURN_IF_EMPTY(n_envs);
  if (x == nullptr || y == nullptr) return set_error(PW_E_INVALID, "x / y is NULL");
  DeviceGuard guard(e->device);
  GoalParams p{};
  p.blob = e->d_blob;
  p.puzzle_off = e->d_off;
  p.num_puzzles = (int32_t)e->puzzles.size();
  p.status = e->d_status;
  p.N = n_envs;
  p.puzzle_id = puzzle_id;
  p.puzzle_uniform = puzzle_id ? 0 : puzzle_uniform;
  p.x = x;
  p.y = y;
  p.stride = stride;
  p.achieved = achieved;
  p.is_goal = is_goal;
  const cudaError_t err = launch_goals(p, static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? PW_OK : cuda_error(err, "pw_goals");
}

int pw_expand(PwEngine* e, int64_t n_states, const int32_t* puzzle_id, int puzzle_uniform,
              const int32_t* x_in, const int32_t* y_in, int64_t stride_in, int32_t* x_out,
              int32_t* y_out, int64_t stride_out, uint32_t* moved, uint8_t* goal, void* stream) {
  clear_error();
  int rc = check_common(e, n_states, puzzle_id, puzzle_uniform, stride_in, n_states);
  if (rc != PW_OK) return rc;
  PW_RETURN_IF_EMPTY(n_states);
  if (stride_out < 4 * n_states) return set_error(PW_E_INVALID, "stride_out smaller than 4*N");
  if (x_in == nullptr || y_in == nullptr || x_out == nullptr || y_out == nullptr)
    return set_error(PW_E_INVALID, "state pointer is NULL");
  DeviceGuard guard(e->device);
  ExpandParams p{};
  p.blob = e->d_blob;
  p.puzzle_off = e->d_off;
  p.num_puzzles = (int32_t)e->puzzles.size();
  p.status = e->d_status;
  p.N = n_states;
  p.puzzle_id = puzzle_id;
  p.puzzle_uniform = puzzle_id ? 0 : puzzle_uniform;
  p.uniform_u64 = puzzle_id ? 0 : e->puzzles[(size_t)puzzle_uniform].size_u64;
  p.n_max = e->n_max;
  p.x_in = x_in;
  p.y_in = y_in;
  p.stride_in = stride_in;
  p.x_out = x_out;
  p.y_out = y_out;
  p.stride_out = stride_out;
  p.moved = moved;
  p.goal = goal;
  const cudaError_t err = launch_expand(p, n_limit_of(e, puzzle_id, puzzle_uniform),
                                        static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? PW_OK : cuda_error(err, "pw_expand");
}

int pw_expand_indexed(PwEngine* e, int64_t n_states, int puzzle, const int64_t* index,
                      const int32_t* x_in, const int32_t* y_in, int64_t stride_in, int32_t* x_out,
                      int32_t* y_out, int64_t stride_out, uint32_t* moved, uint8_t* goal,
                      void* stream) {
  clear_error();
  const int rc = check_common(e, n_states, nullptr, puzzle, stride_in, 0);
  if (rc != PW_OK) return rc;
  PW_RETURN_IF_EMPTY(n_states);
  if (stride_out < 4 * n_states) return set_error(PW_E_INVALID, "stride_out smaller than 4*N");
  if (index == nullptr || x_in == nullptr || y_in == nullptr || x_out == nullptr || y_out == nullptr)
    return set_error(PW_E_INVALID, "index / state pointer is NULL");
  DeviceGuard guard(e->device);
  ExpandParams p{};
  p.blob = e->d_blob;
  p.puzzle_off = e->d_off;
  p.num_puzzles = (int32_t)e->puzzles.size();
  p.status = e->d_status;
  p.N = n_states;
  p.puzzle_id = nullptr;
  p.puzzle_uniform = puzzle;
  p.uniform_u64 = e->puzzles[(size_t)puzzle].size_u64;
  p.n_max = e->n_max;
  p.index = index;
  p.x_in = x_in;
  p.y_in = y_in;
  p.stride_in = stride_in;
  p.x_out = x_out;
  p.y_out = y_out;
  p.stride_out = stride_out;
  p.moved = moved;
  p.goal = goal;
  const cudaError_t err = launch_expand(p, e->puzzles[(size_t)puzzle].n, static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? PW_OK : cuda_error(err, "pw_expand_indexed");
}

int pw_visited_insert(PwEngine* e, int puzzle, const int32_t* x, const int32_t* y, int64_t stride,
                      int64_t first_id, int64_t count, const uint8_t* valid, uint64_t* table,
                      int64_t capacity, uint8_t* is_new, void* stream) {
  clear_error();
  const int rc = check_common(e, count, nullptr, puzzle, stride, 0);
  if (rc != PW_OK) return rc;
  PW_RETURN_IF_EMPTY(count);
  if (x == nullptr || y == nullptr || table == nullptr || is_new == nullptr)
    return set_error(PW_E_INVALID, "NULL argument");
  if (capacity <= 0 || (capacity & (capacity - 1)) != 0)
    return set_error(PW_E_INVALID, "capacity must be a power of two");
  if (first_id < 0 || first_id + count > stride || first_id + count >= ((int64_t)1 << 40))
    return set_error(PW_E_INVALID, "state ids outside the arena");
  DeviceGuard guard(e->device);
  VisitedParams p{};
  p.blob = e->d_blob;
  p.puzzle_off = e->d_off;
  p.num_puzzles = (int32_t)e->puzzles.size();
  p.status = e->d_status;
  p.puzzle = puzzle;
  p.x = x;
  p.y = y;
  p.stride = stride;
  p.first_id = first_id;
  p.count = count;
  p.valid = valid;
  p.table = reinterpret_cast<unsigned long long*>(table);
  p.capacity = capacity;
  p.is_new = is_new;
  p.status = e->d_status;
  const cudaError_t err = launch_visited(p, static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? PW_OK : cuda_error(err, "pw_visited_insert");
}

static uint64_t novelty_bits(const PuzzleInfo& info) {
  const uint64_t n = (uint64_t)info.n, c = (uint64_t)info.width * (uint64_t)info.height;
  return n * c + n * (n - 1) / 2 * c * c;
}

int pw_novelty_table_words(const PwEngine* e, int puzzle, uint64_t* n_words) {
  clear_error();
  if (e == nullptr || n_words == nullptr) return set_error(PW_E_INVALID, "NULL argument");
  if (puzzle < 0 || puzzle >= (int)e->puzzles.size())
    return set_error(PW_E_INVALID, "puzzle out of range");
  *n_words = (novelty_bits(e->puzzles[(size_t)puzzle]) + 31) / 32;
  return PW_OK;
}

int pw_novelty(PwEngine* e, int puzzle, const int32_t* x, const int32_t* y, int64_t stride,
               int64_t first_id, int64_t count, const uint32_t* moved, const uint8_t* valid,
               uint32_t* seen, uint8_t* novelty, void* stream) {
  clear_error();
  const int rc = check_common(e, count, nullptr, puzzle, stride, 0);
  if (rc != PW_OK) return rc;
  PW_RETURN_IF_EMPTY(count);
  if (x == nullptr || y == nullptr || moved == nullptr || seen == nullptr || novelty == nullptr)
    return set_error(PW_E_INVALID, "NULL argument");
  if (first_id < 0 || first_id + count > stride) return set_error(PW_E_INVALID, "state ids outside the arena");
  DeviceGuard guard(e->device);
  const PuzzleInfo& info = e->puzzles[(size_t)puzzle];
  NoveltyParams p{};
  p.x = x;
  p.y = y;
  p.stride = stride;
  p.first_id = first_id;
  p.count = count;
  p.moved = moved;
  p.valid = valid;
  p.seen = seen;
  p.novelty = novelty;
  p.n = info.n;
  p.width = info.width;
  p.height = info.height;
  const cudaError_t err = launch_novelty(p, static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? PW_OK : cuda_error(err, "pw_novelty");
}

static int render_common(PwEngine* e, int64_t n_envs, const int32_t* puzzle_id, int puzzle_uniform,
                         const int32_t* x, const int32_t* y, int64_t stride, void* out, int dtype,
                         int ppc, int bw, int pad_h, int pad_w, const StepParams* fused,
                         void* stream, const char* what) {
  if (x == nullptr || y == nullptr || out == nullptr)
    return set_error(PW_E_INVALID, "x / y / out is NULL");
  if (dtype != PW_DTYPE_U8 && dtype != PW_DTYPE_F32) return set_error(PW_E_INVALID, "bad dtype");
  if (bw < 1) return set_error(PW_E_INVALID, "border_width must be >= 1");  // puzzle.py:444-445
  if (ppc < 1 + 2 * bw)
    return set_error(PW_E_INVALID, "pixels_per_cell must be >= 1 + 2*border_width");  // :447-448
  const int need_w = puzzle_id ? e->max_w : e->puzzles[(size_t)puzzle_uniform].width;
  const int need_h = puzzle_id ? e->max_h : e->puzzles[(size_t)puzzle_uniform].height;
  if (pad_h < need_h || pad_w < need_w)
    return set_error(PW_E_INVALID, "padded size is smaller than the puzzle");
  if ((int64_t)pad_h * ppc > 0x7fffffff / 4 || (int64_t)pad_w * ppc * 12 > 0x7fffffff)
    return set_error(PW_E_INVALID, "image too large");
  DeviceGuard guard(e->device);
  RenderParams p{};
  p.blob = e->d_blob;
  p.puzzle_off = e->d_off;
  p.num_puzzles = (int32_t)e->puzzles.size();
  p.status = e->d_status;
  p.N = n_envs;
  p.puzzle_id = puzzle_id;
  p.puzzle_uniform = puzzle_id ? 0 : puzzle_uniform;
  p.n_max = e->n_max;
  p.x = x;
  p.y = y;
  p.stride = stride;
  p.out = out;
  p.dtype = dtype;
  p.ppc = ppc;
  p.bw = bw;
  p.out_h = pad_h * ppc;
  p.out_w = pad_w * ppc;
  p.max_cells = e->max_cells;
  p.max_blob_u64 = e->max_blob_u64;
  const cudaError_t err = launch_render(p, fused, n_limit_of(e, puzzle_id, puzzle_uniform),
                                        static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? PW_OK : cuda_error(err, what);
}

int pw_render(PwEngine* e, int64_t n_envs, const int32_t* puzzle_id, int puzzle_uniform,
              const int32_t* x, const int32_t* y, int64_t stride, void* out, int dtype,
              int pixels_per_cell, int border_width, int pad_h_cells, int pad_w_cells,
              void* stream) {
  clear_error();
  const int rc = check_common(e, n_envs, puzzle_id, puzzle_uniform, stride, n_envs);
  if (rc != PW_OK) return rc;
  PW_RETURN_IF_EMPTY(n_envs);
  return render_common(e, n_envs, puzzle_id, puzzle_uniform, x, y, stride, out, dtype,
                       pixels_per_cell, border_width, pad_h_cells, pad_w_cells, nullptr, stream,
                       "pw_render");
}

int pw_step_render(PwEngine* e, int64_t n_envs, const int32_t* puzzle_id, int puzzle_uniform,
                   int32_t* x, int32_t* y, int64_t stride, const uint8_t* action, float* reward,
                   int8_t* dgoals, uint8_t* terminated, uint8_t* truncated, int32_t* steps,
                   int32_t max_steps, void* out, int dtype, int pixels_per_cell, int border_width,
                   int pad_h_cells, int pad_w_cells, void* stream) {
  clear_error();
  const int rc = check_common(e, n_envs, puzzle_id, puzzle_uniform, stride, n_envs);
  if (rc != PW_OK) return rc;
  PW_RETURN_IF_EMPTY(n_envs);
  if (action == nullptr) return set_error(PW_E_INVALID, "action is NULL");
  StepParams p{};
  p.blob = e->d_blob;
  p.puzzle_off = e->d_off;
  p.num_puzzles = (int32_t)e->puzzles.size();
  p.status = e->d_status;
  p.N = n_envs;
  p.T = 1;
  p.puzzle_id = puzzle_id;
  p.puzzle_uniform = puzzle_id ? 0 : puzzle_uniform;
  p.uniform_u64 = puzzle_id ? 0 : e->puzzles[(size_t)puzzle_uniform].size_u64;
  p.n_max = e->n_max;
  p.board_bits = puzzle_id ? e->board_bits_all : e->board_bits[(size_t)puzzle_uniform];
  p.board_static = puzzle_id ? e->d_board_static : e->d_board_static + puzzle_uniform;
  p.x = x;
  p.y = y;
  p.stride = stride;
  p.actions = action;
  p.reward = reward;
  p.dgoals = dgoals;
  p.terminated = terminated;
  p.truncated = truncated;
  p.steps = steps;
  p.max_steps = max_steps;
  p.status = e->d_status;
  return render_common(e, n_envs, puzzle_id, puzzle_uniform, x, y, stride, out, dtype,
                       pixels_per_cell, border_width, pad_h_cells, pad_w_cells, &p, stream,
                       "pw_step_render");
}

}  // extern "C"
