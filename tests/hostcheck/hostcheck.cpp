// CPU-side check harness (TEST INFRASTRUCTURE): compiles the very transition rule the CUDA
// kernels use (pushworld_b200/csrc/pw_core.cuh is host/device code) with g++ so that the
// bit logic can be checked against the oracle and the golden vectors without a GPU.
// Never linked into, or reachable from, the product library.
#include <stdint.h>
#include <string.h>

#include "pw_board.cuh"
#include "pw_core.cuh"
#include "pw_fast.cuh"
#include "pw_render_core.cuh"

namespace {

struct HostEnv {
  static constexpr int kStaticN = 0;
  static constexpr int kUnroll = 1;
  const uint64_t* base;
  const uint64_t* ftab = nullptr;  // collision bit tables (pw_fast.cuh), built by hc_walk_tab
  int n_, g_, w_, h_, wall_off_, awall_off_, desc_word_;
  int* px;
  int* py;
  uint64_t fast(uint32_t k) const { return ftab[k]; }
  uint32_t fast32(uint32_t k) const { return reinterpret_cast<const uint32_t*>(ftab)[k]; }
  uint32_t word(int i) const { return reinterpret_cast<const uint32_t*>(base)[i]; }
  void init(const uint64_t* b, int* x, int* y) {
    base = b;
    px = x;
    py = y;
    w_ = word(pw::kHdrDims) & 0xffff;
    h_ = word(pw::kHdrDims) >> 16;
    n_ = word(pw::kHdrCounts) & 0xffff;
    g_ = word(pw::kHdrCounts) >> 16;
    wall_off_ = word(pw::kHdrWallOff);
    awall_off_ = word(pw::kHdrAwallOff);
    desc_word_ = word(pw::kHdrObjOff) * 2;
  }
  int n() const { return n_; }
  int g() const { return g_; }
  int width() const { return w_; }
  int height() const { return h_; }
  int wall_off() const { return wall_off_; }
  int awall_off() const { return awall_off_; }
  uint64_t row(int off) const { return base[off]; }
  uint32_t desc(int j) const { return word(desc_word_ + 4 * j); }
  uint32_t desc_dyn(int j) const { return desc(j); }
  int x(int j) const { return px[j]; }
  int y(int j) const { return py[j]; }
  int x_dyn(int j) const { return px[j]; }
  int y_dyn(int j) const { return py[j]; }
  int goal_x(int k) const { return pw::lo16(word(desc_word_ + 4 * (k + 1) + 2)); }
  int goal_y(int k) const { return pw::hi16(word(desc_word_ + 4 * (k + 1) + 2)); }
  void translate(uint32_t mask, int dx, int dy) {
    for (int j = 0; j < n_; ++j)
      if ((mask >> j) & 1u) {
        px[j] += dx;
        py[j] += dy;
      }
  }
};

}  // namespace

struct HostTable {
  const uint64_t* base;
  uint64_t row(int off) const { return base[off]; }
  uint32_t word(int i) const { return reinterpret_cast<const uint32_t*>(base)[i]; }
};
struct HostPos {
  const int* px;
  const int* py;
  int x(int j) const { return px[j]; }
  int y(int j) const { return py[j]; }
};

template <typename Word, int NMAX>
static int board_walk_n(const uint64_t* blob, int* x, int* y, const uint8_t* actions, int n_actions,
                        int* xs, int* ys, uint32_t* moved, int* achieved) {
  HostTable t{blob};
  const int W = t.word(pw::kHdrDims) & 0xffff, H = t.word(pw::kHdrDims) >> 16;
  const int n = t.word(pw::kHdrCounts) & 0xffff;
  if (pw::board_bits(W, H) > (int)sizeof(Word) * 8 || n > NMAX) return -1;
  pw::Board<NMAX, Word> s;
  pw::board_init(s, t);
  if (!pw::board_load(s, t, [&](int k) { return x[k]; }, [&](int k) { return y[k]; })) return -2;
  pw::BoardAction<Word> lut[4];
  pw::board_action_table(s, lut);
  for (int i = 0; i < n_actions; ++i) {
    moved[i] = pw::board_step(s, actions[i], lut);
    achieved[i] = pw::board_achieved(s);
    for (int k = 0; k < n; ++k)
      pw::board_position<Word>(s.b[k], s.P, &xs[(size_t)i * n + k], &ys[(size_t)i * n + k]);
  }
  return n;
}

// exact-n instantiations like the kernels' (2 and 3 objects use closed-form closures)
template <typename Word>
static int board_walk(const uint64_t* blob, int* x, int* y, const uint8_t* actions, int n_actions,
                      int* xs, int* ys, uint32_t* moved, int* achieved) {
  const int n = reinterpret_cast<const uint32_t*>(blob)[pw::kHdrCounts] & 0xffff;
  if (n <= 2) return board_walk_n<Word, 2>(blob, x, y, actions, n_actions, xs, ys, moved, achieved);
  if (n == 3) return board_walk_n<Word, 3>(blob, x, y, actions, n_actions, xs, ys, moved, achieved);
  return board_walk_n<Word, 8>(blob, x, y, actions, n_actions, xs, ys, moved, achieved);
}

// ring boards (pw_board.cuh: one rotate per object, guard zone instead of edge rows)
template <int NMAX>
static int ring_walk_n(const uint64_t* blob, int* x, int* y, const uint8_t* actions, int n_actions,
                       int* xs, int* ys, uint32_t* moved, int* achieved) {
  HostTable t{blob};
  const int W = t.word(pw::kHdrDims) & 0xffff, H = t.word(pw::kHdrDims) >> 16;
  const int n = t.word(pw::kHdrCounts) & 0xffff;
  const int bits = pw::board_bits(W, H);
  if (!pw::ring_fits(bits, W - 1) || n > NMAX) return -1;
  pw::Board<NMAX, uint32_t> s;
  pw::board_init(s, t);
  if (!pw::board_load(s, t, [&](int k) { return x[k]; }, [&](int k) { return y[k]; })) return -2;
  const uint32_t guard = pw::ring_guard(bits), amounts = pw::ring_amounts(s.P);
  uint32_t b[NMAX];
  for (int k = 0; k < NMAX; ++k) b[k] = s.b[k];
  for (int i = 0; i < n_actions; ++i) {
    const uint32_t a = actions[i];
    moved[i] = pw::ring_step<NMAX, true>(b, s.wall | guard, s.awall | guard,
                                         pw::ring_amount_of(amounts, a), a <= 3u);
    for (int k = 0; k < NMAX; ++k) s.b[k] = b[k];
    achieved[i] = pw::board_achieved(s);
    for (int k = 0; k < n; ++k)
      pw::board_position<uint32_t>(b[k], s.P, &xs[(size_t)i * n + k], &ys[(size_t)i * n + k]);
  }
  return n;
}

static int ring_walk(const uint64_t* blob, int* x, int* y, const uint8_t* actions, int n_actions,
                     int* xs, int* ys, uint32_t* moved, int* achieved) {
  const int n = reinterpret_cast<const uint32_t*>(blob)[pw::kHdrCounts] & 0xffff;
  if (n <= 2) return ring_walk_n<2>(blob, x, y, actions, n_actions, xs, ys, moved, achieved);
  if (n == 3) return ring_walk_n<3>(blob, x, y, actions, n_actions, xs, ys, moved, achieved);
  return ring_walk_n<8>(blob, x, y, actions, n_actions, xs, ys, moved, achieved);
}

extern "C" {

// Same walk through the compact bitboard rule (pw_board.cuh, NMAX = 8).  word_bits selects
// uint32 / uint64 boards.  Returns -1 if the puzzle does not fit, -2 if the start state is not
// regular.
int hc_board_walk(const uint64_t* blob, int word_bits, int* x, int* y, const uint8_t* actions,
                  int n_actions, int* xs, int* ys, uint32_t* moved, int* achieved) {
  if (word_bits == 31)  // ring boards
    return ring_walk(blob, x, y, actions, n_actions, xs, ys, moved, achieved);
  if (word_bits == 32)
    return board_walk<uint32_t>(blob, x, y, actions, n_actions, xs, ys, moved, achieved);
  return board_walk<uint64_t>(blob, x, y, actions, n_actions, xs, ys, moved, achieved);
}

// Expands pw::cell_code() into a uint8 image [H*ppc][W*ppc][3] (the CUDA kernel does the same
// expansion with shared-memory row patterns; this loop is deliberately the naive one).
void hc_render_u8(const uint64_t* blob, const int* x, const int* y, int ppc, int bw,
                  const uint8_t* palette /*[16][3]*/, uint8_t* out) {
  HostTable t{blob};
  HostPos pos{x, y};
  const int W = t.word(pw::kHdrDims) & 0xffff, H = t.word(pw::kHdrDims) >> 16;
  for (int cy = 0; cy < H; ++cy)
    for (int cx = 0; cx < W; ++cx) {
      const uint64_t code = pw::cell_code(t, pos, cx, cy);
      for (int oy = 0; oy < ppc; ++oy)
        for (int ox = 0; ox < ppc; ++ox) {
          const int k = pw::region_of(oy, ppc, bw) * 3 + pw::region_of(ox, ppc, bw);
          const int pal = (int)((code >> (4 * k)) & 15);
          uint8_t* px = out + ((size_t)(cy * ppc + oy) * (size_t)(W * ppc) + (size_t)(cx * ppc + ox)) * 3;
          px[0] = palette[pal * 3 + 0];
          px[1] = palette[pal * 3 + 1];
          px[2] = palette[pal * 3 + 2];
        }
    }
}

// The same walk through the table-driven rule (pw_fast.cuh: tables built from the blob here, then
// step_tab); also returns the table size in u64 units through *table_u64.
int hc_walk_tab(const uint64_t* blob, int* x, int* y, const uint8_t* actions, int n_actions,
                int* xs, int* ys, uint32_t* moved, int* achieved, int* table_u64) {
  std::vector<uint64_t> table;
  pw::build_fast_table(blob, &table);
  if (table_u64) *table_u64 = (int)table.size();
  HostEnv env;
  env.init(blob, x, y);
  env.ftab = table.data();
  for (int t = 0; t < n_actions; ++t) {
    moved[t] = pw::step_tab(env, actions[t]);
    achieved[t] = pw::achieved_goals(env);
    memcpy(xs + (size_t)t * env.n(), x, sizeof(int) * env.n());
    memcpy(ys + (size_t)t * env.n(), y, sizeof(int) * env.n());
  }
  return env.n();
}

// Applies `n_actions` actions in sequence to one environment, recording after every action the
// state (xs/ys: [n_actions][n]), the moved mask and the achieved-goal count.
int hc_walk(const uint64_t* blob, int* x, int* y, const uint8_t* actions, int n_actions,
            int* xs, int* ys, uint32_t* moved, int* achieved) {
  HostEnv env;
  env.init(blob, x, y);
  for (int t = 0; t < n_actions; ++t) {
    moved[t] = pw::step(env, actions[t]);
    achieved[t] = pw::achieved_goals(env);
    memcpy(xs + (size_t)t * env.n(), x, sizeof(int) * env.n());
    memcpy(ys + (size_t)t * env.n(), y, sizeof(int) * env.n());
  }
  return env.n();
}
}
