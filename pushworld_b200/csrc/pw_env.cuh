// pw_env.cuh -- per-thread environment state policies for pw_core.cuh (device only).
//
// One CUDA thread owns one environment.  Two storage policies:
//   RegEnv<NMAX>  positions, shape descriptors and goal positions live in registers;
//                 every loop over objects is fully unrolled so no array is ever indexed
//                 dynamically (the pusher index, the only runtime index, goes through a
//                 select chain).  Used when the launch's puzzles have n <= NMAX (4 or 8).
//   SmemEnv       positions live in shared memory, column `threadIdx.x` of an
//                 [n_max][blockDim] int32 array per coordinate: a dynamic object index is
//                 a different row of the same bank, so accesses are conflict free however
//                 the lanes of a warp diverge.  Descriptors are read from the table.
// The table view is either the block's shared-memory copy of a puzzle blob (all threads of
// the launch share one puzzle) or the engine's global blob read through the read-only path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pw_core.cuh"
#include "pw_fast.cuh"

namespace pw {

struct SharedTable {
  const uint64_t* base;
  // the table stored `offset_u64` after `t` (shared memory: plain pointer arithmetic, LDS)
  static __device__ __forceinline__ SharedTable after(const SharedTable& t, uint32_t offset_u64) {
    return SharedTable{t.base + offset_u64};
  }
  __device__ __forceinline__ uint64_t row(int off) const { return base[off]; }
  __device__ __forceinline__ uint64_t qword(uint32_t off) const { return base[off]; }  // unsigned: one IMAD.WIDE
  __device__ __forceinline__ uint32_t dword(uint32_t i) const {
    return reinterpret_cast<const uint32_t*>(base)[i];
  }
  __device__ __forceinline__ uint32_t word(int i) const {
    return reinterpret_cast<const uint32_t*>(base)[i];
  }
};

struct GlobalTable {
  const uint64_t* base;
  // the table stored `offset_u64` after `t`, as an OPAQUE per-thread pointer: otherwise every
  // address is rebuilt from the uniform blob pointer plus the puzzle offset (four instructions
  // instead of one IMAD.WIDE); reads go through __ldg, so the lost address space does not matter
  static __device__ __forceinline__ GlobalTable after(const GlobalTable& t, uint32_t offset_u64) {
    const uint64_t* p = t.base + offset_u64;
    asm volatile("" : "+l"(p));
    return GlobalTable{p};
  }
  __device__ __forceinline__ uint64_t row(int off) const { return __ldg(base + off); }
  __device__ __forceinline__ uint64_t qword(uint32_t off) const { return __ldg(base + off); }
  __device__ __forceinline__ uint32_t dword(uint32_t i) const {
    return __ldg(reinterpret_cast<const uint32_t*>(base) + i);
  }
  __device__ __forceinline__ uint32_t word(int i) const {
    return __ldg(reinterpret_cast<const uint32_t*>(base) + i);
  }
};

template <class Table>
struct EnvBase {
  Table t;
  Table ft;  // collision bit tables of the same puzzle (pw_fast.cuh): stored right after the blob
  int n_, g_, w_, h_, wall_off_, awall_off_, desc_word_;
  __device__ __forceinline__ uint64_t fast(uint32_t k) const { return ft.qword(k); }
  __device__ __forceinline__ uint32_t fast32(uint32_t k) const { return ft.dword(k); }
  __device__ __forceinline__ void init_header(Table table) {
    t = table;
    const uint32_t dims = t.word(kHdrDims), counts = t.word(kHdrCounts);
    w_ = (int)(dims & 0xffffu);
    h_ = (int)(dims >> 16);
    n_ = (int)(counts & 0xffffu);
    g_ = (int)(counts >> 16);
    wall_off_ = (int)t.word(kHdrWallOff);
    awall_off_ = (int)t.word(kHdrAwallOff);
    desc_word_ = (int)t.word(kHdrObjOff) * 2;
    ft = Table::after(t, t.word(kHdrSizeU64));
  }
  // points the view at another puzzle's tables (same address space)
  __device__ __forceinline__ void retarget(const uint64_t* base) { init_header(Table{base}); }
  __device__ __forceinline__ int n() const { return n_; }
  __device__ __forceinline__ int g() const { return g_; }
  __device__ __forceinline__ int width() const { return w_; }
  __device__ __forceinline__ int height() const { return h_; }
  __device__ __forceinline__ int wall_off() const { return wall_off_; }
  __device__ __forceinline__ int awall_off() const { return awall_off_; }
  __device__ __forceinline__ uint64_t row(int off) const { return t.row(off); }
  __device__ __forceinline__ uint32_t table_desc(int j, int k) const {
    return t.word(desc_word_ + 4 * j + k);
  }
};

template <int NMAX, class Table>
struct RegEnv : EnvBase<Table> {
  static constexpr int kStaticN = NMAX;
  static constexpr int kUnroll = NMAX;
  int px[NMAX], py[NMAX];
  uint32_t d[NMAX];
  int gx[NMAX], gy[NMAX];  // goal of object k+1 at index k

  __device__ __forceinline__ void init(Table table) {
    this->init_header(table);
#pragma unroll
    for (int j = 0; j < NMAX; ++j) {
      const bool live = j < this->n_;
      d[j] = live ? this->table_desc(j, 0) : 0u;
      const uint32_t goal = (live && j >= 1) ? this->table_desc(j, 2) : kNoGoal;
      if (j >= 1) {
        gx[j - 1] = lo16(goal);
        gy[j - 1] = hi16(goal);
      }
      px[j] = 0;
      py[j] = 0;
    }
    gx[NMAX - 1] = gy[NMAX - 1] = 0;
  }
  __device__ __forceinline__ void load(const int32_t* x, const int32_t* y, int64_t stride,
                                       int64_t e) {
#pragma unroll
    for (int j = 0; j < NMAX; ++j)
      if (j < this->n_) {
        px[j] = x[j * stride + e];
        py[j] = y[j * stride + e];
      }
  }
  // the same with 32-bit element offsets (planes of fewer than 2^31 elements): one IMAD.WIDE per address
  __device__ __forceinline__ void load32(const int32_t* x, const int32_t* y, uint32_t stride, uint32_t e) {
    uint32_t off = e;
#pragma unroll
    for (int j = 0; j < NMAX; ++j) {
      if (j < this->n_) {
        px[j] = x[off];
        py[j] = y[off];
      }
      off += stride;
    }
  }
  __device__ __forceinline__ void load_initial() {
#pragma unroll
    for (int j = 0; j < NMAX; ++j)
      if (j < this->n_) {
        const uint32_t v = this->table_desc(j, 1);
        px[j] = lo16(v);
        py[j] = hi16(v);
      }
  }
  __device__ __forceinline__ void store(int32_t* x, int32_t* y, int64_t stride, int64_t e) const {
#pragma unroll
    for (int j = 0; j < NMAX; ++j)
      if (j < this->n_) {
        x[j * stride + e] = px[j];
        y[j * stride + e] = py[j];
      }
  }
  __device__ __forceinline__ uint32_t desc(int j) const { return d[j]; }
  __device__ __forceinline__ int x(int j) const { return px[j]; }
  __device__ __forceinline__ int y(int j) const { return py[j]; }
  __device__ __forceinline__ int goal_x(int k) const { return gx[k]; }
  __device__ __forceinline__ int goal_y(int k) const { return gy[k]; }
  __device__ __forceinline__ uint32_t desc_dyn(int i) const {
    uint32_t v = d[0];
#pragma unroll
    for (int k = 1; k < NMAX; ++k) v = (i == k) ? d[k] : v;
    return v;
  }
  // Dynamic index: a chain of selects over the register array -- or, after pack_positions() (callers that
  // never move the objects: successor expansion), ONE byte permute on a packed copy of the positions.
  uint32_t xp_[2] = {0u, 0u}, yp_[2] = {0u, 0u};
  bool packed_ = false;
  __device__ __forceinline__ void pack_positions() {
    static_assert(NMAX <= 8, "two words of byte positions");
    bool ok = true;
    xp_[0] = xp_[1] = yp_[0] = yp_[1] = 0u;
#pragma unroll
    for (int j = 0; j < NMAX; ++j) {
      ok = ok && (uint32_t)px[j] <= 255u && (uint32_t)py[j] <= 255u;
      xp_[j >> 2] |= ((uint32_t)px[j] & 0xffu) << (8 * (j & 3));
      yp_[j >> 2] |= ((uint32_t)py[j] & 0xffu) << (8 * (j & 3));
    }
    packed_ = ok;
  }
  __device__ __forceinline__ int x_dyn(int i) const {
    if (packed_) return (int)(__byte_perm(xp_[0], xp_[1], (uint32_t)i) & 0xffu);
    int v = px[0];
#pragma unroll
    for (int k = 1; k < NMAX; ++k) v = (i == k) ? px[k] : v;
    return v;
  }
  __device__ __forceinline__ int y_dyn(int i) const {
    if (packed_) return (int)(__byte_perm(yp_[0], yp_[1], (uint32_t)i) & 0xffu);
    int v = py[0];
#pragma unroll
    for (int k = 1; k < NMAX; ++k) v = (i == k) ? py[k] : v;
    return v;
  }
  __device__ __forceinline__ void translate(uint32_t mask, int dx, int dy) {
#pragma unroll
    for (int j = 0; j < NMAX; ++j) {
      const bool m = (mask >> j) & 1u;
      px[j] += m ? dx : 0;
      py[j] += m ? dy : 0;
    }
  }
};

template <class Table>
struct SmemEnv : EnvBase<Table> {
  static constexpr int kStaticN = 0;
  static constexpr int kUnroll = 1;
  int* sx;  // &s_x[threadIdx.x]; object j at sx[j * pitch]
  int* sy;
  int pitch;

  // x_col / y_col: this thread's column; object j lives at x_col[j * col_pitch]
  __device__ __forceinline__ void init(Table table, int* x_col, int* y_col, int col_pitch) {
    this->init_header(table);
    pitch = col_pitch;
    sx = x_col;
    sy = y_col;
  }
  __device__ __forceinline__ void load(const int32_t* x, const int32_t* y, int64_t stride,
                                       int64_t e) {
    for (int j = 0; j < this->n_; ++j) {
      sx[j * pitch] = x[j * stride + e];
      sy[j * pitch] = y[j * stride + e];
    }
  }
  __device__ __forceinline__ void load32(const int32_t* x, const int32_t* y, uint32_t stride, uint32_t e) {
    uint32_t off = e;
    for (int j = 0; j < this->n_; ++j, off += stride) {
      sx[j * pitch] = x[off];
      sy[j * pitch] = y[off];
    }
  }
  __device__ __forceinline__ void load_initial() {
    for (int j = 0; j < this->n_; ++j) {
      const uint32_t v = this->table_desc(j, 1);
      sx[j * pitch] = lo16(v);
      sy[j * pitch] = hi16(v);
    }
  }
  __device__ __forceinline__ void store(int32_t* x, int32_t* y, int64_t stride, int64_t e) const {
    for (int j = 0; j < this->n_; ++j) {
      x[j * stride + e] = sx[j * pitch];
      y[j * stride + e] = sy[j * pitch];
    }
  }
  __device__ __forceinline__ uint32_t desc(int j) const { return this->table_desc(j, 0); }
  __device__ __forceinline__ uint32_t desc_dyn(int j) const { return this->table_desc(j, 0); }
  __device__ __forceinline__ int x(int j) const { return sx[j * pitch]; }
  __device__ __forceinline__ int y(int j) const { return sy[j * pitch]; }
  __device__ __forceinline__ int x_dyn(int j) const { return sx[j * pitch]; }
  __device__ __forceinline__ int y_dyn(int j) const { return sy[j * pitch]; }
  __device__ __forceinline__ int goal_x(int k) const { return lo16(this->table_desc(k + 1, 2)); }
  __device__ __forceinline__ int goal_y(int k) const { return hi16(this->table_desc(k + 1, 2)); }
  __device__ __forceinline__ void translate(uint32_t mask, int dx, int dy) {
    while (mask) {
      const int j = __ffs((int)mask) - 1;
      mask &= mask - 1;
      if (dx) sx[j * pitch] += dx;
      if (dy) sy[j * pitch] += dy;
    }
  }
};

// Positions in (per-thread) local memory: the slow-path policy for the rare environments a
// specialised kernel cannot take (e.g. an irregular state handed to the bitboard kernel).
template <class Table>
struct LocalEnv : EnvBase<Table> {
  static constexpr int kStaticN = 0;
  static constexpr int kUnroll = 1;
  int px[kMaxObjects], py[kMaxObjects];

  __device__ __forceinline__ void init(Table table) { this->init_header(table); }
  __device__ __forceinline__ void load(const int32_t* x, const int32_t* y, int64_t stride,
                                       int64_t e) {
    for (int j = 0; j < this->n_; ++j) {
      px[j] = x[j * stride + e];
      py[j] = y[j * stride + e];
    }
  }
  __device__ __forceinline__ void load_initial() {
    for (int j = 0; j < this->n_; ++j) {
      const uint32_t v = this->table_desc(j, 1);
      px[j] = lo16(v);
      py[j] = hi16(v);
    }
  }
  __device__ __forceinline__ void store(int32_t* x, int32_t* y, int64_t stride, int64_t e) const {
    for (int j = 0; j < this->n_; ++j) {
      x[j * stride + e] = px[j];
      y[j * stride + e] = py[j];
    }
  }
  __device__ __forceinline__ uint32_t desc(int j) const { return this->table_desc(j, 0); }
  __device__ __forceinline__ uint32_t desc_dyn(int j) const { return this->table_desc(j, 0); }
  __device__ __forceinline__ int x(int j) const { return px[j]; }
  __device__ __forceinline__ int y(int j) const { return py[j]; }
  __device__ __forceinline__ int x_dyn(int j) const { return px[j]; }
  __device__ __forceinline__ int y_dyn(int j) const { return py[j]; }
  __device__ __forceinline__ int goal_x(int k) const { return lo16(this->table_desc(k + 1, 2)); }
  __device__ __forceinline__ int goal_y(int k) const { return hi16(this->table_desc(k + 1, 2)); }
  __device__ __forceinline__ void translate(uint32_t mask, int dx, int dy) {
    while (mask) {
      const int j = __ffs((int)mask) - 1;
      mask &= mask - 1;
      px[j] += dx;
      py[j] += dy;
    }
  }
};

}  // namespace pw
