// pw_render.cu -- RGB observation kernel.  pw_step_render (BASELINE config 5) is the step launch of
// pw_kernels.cu followed by this kernel on the same stream.
//
// One CTA renders a few environments, one after the other:
//   1. every thread computes the 9-region colour code of some grid cells (pw_render_core.cuh)
//      into shared memory -- W*H codes, 8 bytes each;
//   2. per cell row, the block builds the 3 distinct pixel-row patterns (top border rows,
//      middle rows, bottom border rows; all `pixels_per_cell` rows of a cell row are one of
//      these) INCLUDING the left/right zero padding, in the output element type, in shared
//      memory; then streams them to the `pixels_per_cell` output rows with 128-bit coalesced
//      stores.  Compute is therefore amortised pixels_per_cell/3 times and the kernel is a pure
//      HBM write stream (algorithmic bytes = out_h*out_w*3*sizeof(elem) per frame).
// Rows above/below the centred image are zero (env_utils.py:75-91).
#include <cuda_runtime.h>
#include <stdint.h>

#include "pw_env.cuh"
#include "pw_internal.h"
#include "pw_render_core.cuh"

namespace pw {

__constant__ uint8_t c_pal_u8[16][4];
__constant__ float c_pal_f32[16][4];

cudaError_t upload_palette(const uint8_t* rgb, const float* rgb_f32, int n_colors) {
  uint8_t pu[16][4] = {};
  float pf[16][4] = {};
  for (int i = 0; i < n_colors && i < 16; ++i)
    for (int c = 0; c < 3; ++c) {
      pu[i][c] = rgb[i * 3 + c];
      pf[i][c] = rgb_f32[i * 3 + c];
    }
  cudaError_t err = cudaMemcpyToSymbol(c_pal_u8, pu, sizeof(pu));
  if (err == cudaSuccess) err = cudaMemcpyToSymbol(c_pal_f32, pf, sizeof(pf));
  return err;
}

namespace {

constexpr int kRenderBlock = 256;
constexpr int kBulkChunks = 4;  // 16-byte chunks per thread in the bulk cell-row emit (<= 16 KB per cell row)
constexpr size_t kRenderMaxDynamicSmem = 227 * 1024 - 2048;

struct SmemPos {
  const int* sx;
  const int* sy;
  __device__ __forceinline__ int x(int j) const { return sx[j]; }
  __device__ __forceinline__ int y(int j) const { return sy[j]; }
};

template <typename T>
struct Elem;
template <>
struct Elem<uint8_t> {
  static __device__ __forceinline__ uint8_t get(uint32_t pal, int c) { return c_pal_u8[pal][c]; }
};
template <>
struct Elem<float> {
  static __device__ __forceinline__ float get(uint32_t pal, int c) { return c_pal_f32[pal][c]; }
};

// One output row: `bytes` from shared `src` to global `dst`, by ONE warp (lanes over chunks).
template <int VEC>
__device__ __forceinline__ void copy_row(uint8_t* dst, const uint8_t* src, int bytes, int lane) {
  if (VEC == 16) {
    const int n = bytes >> 4;
    for (int i = lane; i < n; i += 32)
      reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
  } else if (VEC == 4) {
    const int n = bytes >> 2;
    for (int i = lane; i < n; i += 32)
      reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(src)[i];
  } else {
    for (int i = lane; i < bytes; i += 32) dst[i] = src[i];
  }
}

template <typename T>
__device__ __forceinline__ void put_pixel(T* dst, uint32_t pal, const float4* s_pal_f, const uint32_t* s_pal_u);
template <>
__device__ __forceinline__ void put_pixel<float>(float* dst, uint32_t pal, const float4* s_pal_f,
                                                 const uint32_t*) {
  const float4 c = s_pal_f[pal];
  dst[0] = c.x;
  dst[1] = c.y;
  dst[2] = c.z;
}
template <>
__device__ __forceinline__ void put_pixel<uint8_t>(uint8_t* dst, uint32_t pal, const float4*,
                                                   const uint32_t* s_pal_u) {
  const uint32_t c = s_pal_u[pal];
  dst[0] = (uint8_t)(c & 0xffu);
  dst[1] = (uint8_t)((c >> 8) & 0xffu);
  dst[2] = (uint8_t)((c >> 16) & 0xffu);
}

// Persistent CTAs: CTA b owns environments b, b + G, b + 2G, ... (G = grid size), so that the CTAs
// running at the same time write ADJACENT frames and the whole grid sweeps the output buffer
// like one stream (contiguous chunks per CTA cost ~25 % of the HBM write bandwidth).  The puzzle
// table, the pixel-column map and the palette are staged once per CTA and only re-staged when
// the puzzle changes.
template <typename T>
__global__ void __launch_bounds__(kRenderBlock) k_render(const RenderParams p) {
  extern __shared__ __align__(16) uint64_t smem[];
  __shared__ float4 s_pal_f[16];
  __shared__ uint32_t s_pal_u[16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const int64_t G = gridDim.x;

  // ---- shared memory carve-up -----------------------------------------------------------
  uint64_t* s_tbl = smem;                                   // max_blob_u64
  uint64_t* s_code = s_tbl + p.max_blob_u64;                // max_cells
  int* s_x = reinterpret_cast<int*>(s_code + p.max_cells);  // 32 ints
  int* s_y = s_x + 32;
  uint16_t* s_xmap = reinterpret_cast<uint16_t*>(s_y + 32);  // out_w entries: cx << 2 | region
  const int row_bytes = p.out_w * 3 * (int)sizeof(T);
  const int row_pitch = (row_bytes + 15) & ~15;
  // 2 pattern sets (double buffered) x 3 region rows, then a zero row; 16-byte aligned
  uint8_t* s_pat = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(s_xmap + p.out_w) + 15) & ~static_cast<uintptr_t>(15));
  uint8_t* s_zero = s_pat + 6 * row_pitch;

  if (threadIdx.x < 16) {
    s_pal_f[threadIdx.x] = make_float4(c_pal_f32[threadIdx.x][0], c_pal_f32[threadIdx.x][1],
                                       c_pal_f32[threadIdx.x][2], 0.f);
    s_pal_u[threadIdx.x] = (uint32_t)c_pal_u8[threadIdx.x][0] | ((uint32_t)c_pal_u8[threadIdx.x][1] << 8) |
                           ((uint32_t)c_pal_u8[threadIdx.x][2] << 16);
  }
  for (int i = threadIdx.x; i < (row_pitch >> 2); i += blockDim.x)
    reinterpret_cast<uint32_t*>(s_zero)[i] = 0u;

  __syncthreads();

  const int ppc = p.ppc, bw = p.bw;
  // bulk emit (see emit_cell_row_bulk): chunk c of a cell-row block lies in output row
  // c * 16 / row_bytes; its source is that row's region pattern at the same column offset
  const int bulk_chunks = (ppc * row_bytes) >> 4;
  const bool use_bulk = (row_bytes & 15) == 0 && bulk_chunks <= kBulkChunks * kRenderBlock;
  int bulk_src[kBulkChunks];
#pragma unroll
  for (int k = 0; k < kBulkChunks; ++k) {
    const int byte = ((int)threadIdx.x + k * kRenderBlock) << 4;
    const int oy = use_bulk ? byte / row_bytes : 0;
    bulk_src[k] = region_of(oy < ppc ? oy : 0, ppc, bw) * row_pitch + (byte - oy * row_bytes);
  }
  int cur_pid = -1, W = 0, H = 0, n = 0, img_h = 0, img_w = 0, pad_top = 0, pad_left = 0;
  for (int64_t e = blockIdx.x; e < p.N; e += G) {
    // (every thread of the CTA reads the id; thread 0 reports a bad one)
    const int pid = p.puzzle_id ? checked_puzzle_id(p.puzzle_id[e], p.num_puzzles, threadIdx.x == 0 ? p.status : nullptr)
                                : p.puzzle_uniform;
    if (pid != cur_pid) {  // block-uniform branch
      cur_pid = pid;
      const uint64_t* g_tbl = p.blob + p.puzzle_off[pid];
      const int n_u64 = (int)__ldg(reinterpret_cast<const uint32_t*>(g_tbl) + kHdrSizeU64);
      for (int i = threadIdx.x; i < n_u64; i += blockDim.x) s_tbl[i] = __ldg(g_tbl + i);
      const uint32_t dims = __ldg(reinterpret_cast<const uint32_t*>(g_tbl) + kHdrDims);
      W = (int)(dims & 0xffffu);
      H = (int)(dims >> 16);
      n = (int)(__ldg(reinterpret_cast<const uint32_t*>(g_tbl) + kHdrCounts) & 0xffffu);
      img_h = H * ppc;
      img_w = W * ppc;
      pad_top = (p.out_h - img_h) / 2;   // env_utils.py:81-82
      pad_left = (p.out_w - img_w) / 2;
      // pixel column -> (cell column, region) map, shared by every cell row of the frame
      for (int X = threadIdx.x; X < p.out_w; X += blockDim.x) {
        const int xi = X - pad_left;
        uint16_t m = 0xffffu;  // zero padding
        if (xi >= 0 && xi < img_w) {
          const int cx = xi / ppc;
          m = (uint16_t)((cx << 2) | region_of(xi - cx * ppc, ppc, bw));
        }
        s_xmap[X] = m;
      }
    }
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      s_x[j] = p.x[j * p.stride + e];
      s_y[j] = p.y[j * p.stride + e];
    }
    __syncthreads();
    const SharedTable tbl{s_tbl};

    // ---- 1. cell codes ------------------------------------------------------------------
    const SmemPos pos{s_x, s_y};
    for (int c = threadIdx.x; c < W * H; c += blockDim.x) {
      const int cy = c / W, cx = c - cy * W;
      s_code[c] = cell_code(tbl, pos, cx, cy);
    }

    // ---- 2. pixel rows ------------------------------------------------------------------
    uint8_t* frame = static_cast<uint8_t*>(p.out) + (size_t)e * (size_t)p.out_h * (size_t)row_bytes;
    const bool vec16 = (row_bytes & 15) == 0 && (reinterpret_cast<uintptr_t>(frame) & 15) == 0;
    const bool vec4 = (row_bytes & 3) == 0 && (reinterpret_cast<uintptr_t>(frame) & 3) == 0;
    auto emit = [&](int Y, const uint8_t* src) {  // one warp writes one output row
      uint8_t* dst = frame + (size_t)Y * (size_t)row_bytes;
      if (vec16) copy_row<16>(dst, src, row_bytes, lane);
      else if (vec4) copy_row<4>(dst, src, row_bytes, lane);
      else copy_row<1>(dst, src, row_bytes, lane);
    };
    // builds the 3 region-row patterns of cell row cy into pattern set `buf`
    auto build = [&](int cy, int buf) {
      uint8_t* set = s_pat + buf * 3 * row_pitch;
      const uint64_t* codes = s_code + cy * W;
      for (int X = threadIdx.x; X < p.out_w; X += blockDim.x) {
        const uint32_t m = s_xmap[X];
        const bool inside = m != 0xffffu;
        const uint64_t code = inside ? codes[m >> 2] >> (4u * (m & 3u)) : 0ull;
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) {
          const uint32_t pal = inside ? (uint32_t)(code >> (12 * rr)) & 15u : kPalPad;
          put_pixel<T>(reinterpret_cast<T*>(set + rr * row_pitch) + 3 * X, pal, s_pal_f, s_pal_u);
        }
      }
    };
    // Small rows (uint8 frames): the pixels_per_cell output rows of a cell row are one contiguous
    // block of the frame, written here as 16-byte chunks with per-thread source offsets that were
    // computed once per kernel (bulk_src), instead of one warp and ~40 instructions per row.
    auto emit_cell_row_bulk = [&](int cy, const uint8_t* set) {
      uint4* dst = reinterpret_cast<uint4*>(frame + (size_t)(pad_top + cy * ppc) * (size_t)row_bytes);
#pragma unroll
      for (int k = 0; k < kBulkChunks; ++k) {
        const int c = (int)threadIdx.x + k * kRenderBlock;
        if (c < bulk_chunks) dst[c] = *reinterpret_cast<const uint4*>(set + bulk_src[k]);
      }
    };

    for (int Y = warp; Y < pad_top; Y += n_warps) emit(Y, s_zero);
    for (int Y = pad_top + img_h + warp; Y < p.out_h; Y += n_warps) emit(Y, s_zero);
    __syncthreads();  // codes ready
    build(0, 0);
    __syncthreads();
    for (int cy = 0; cy < H; ++cy) {
      if (cy + 1 < H) build(cy + 1, (cy + 1) & 1);  // next cell row while this one streams out
      const uint8_t* set = s_pat + (cy & 1) * 3 * row_pitch;
      if (use_bulk && vec16) {
        emit_cell_row_bulk(cy, set);
      } else {
        const int Y0 = pad_top + cy * ppc;
        for (int oy = warp; oy < ppc; oy += n_warps)
          emit(Y0 + oy, set + region_of(oy, ppc, bw) * row_pitch);
      }
      __syncthreads();
    }
  }
}

}  // namespace

cudaError_t launch_render(const RenderParams& p, const StepParams* fused_step, int n_limit,
                          cudaStream_t stream) {
  if (p.N == 0) return cudaSuccess;
  const int elem = p.dtype == 1 ? 4 : 1;
  const size_t row_pitch = ((size_t)p.out_w * 3 * elem + 15) & ~(size_t)15;
  const size_t smem = (size_t)p.max_blob_u64 * 8 + (size_t)p.max_cells * 8 + 64 * sizeof(int) +
                      (size_t)p.out_w * 2 + 16 + 7 * row_pitch;
  if (smem > kRenderMaxDynamicSmem) return cudaErrorInvalidValue;
  // pw_step_render: the step runs first, as its own launch on the same stream (the dynamics kernels
  // of pw_kernels.cu: thousands of environments per CTA-wave), then the frames are rendered from the
  // new state.  Stepping inside the render kernel (a few environments per CTA, one thread each)
  // cost 0.3 ms per 2^16 environments; the separate launch costs ~10 us.
  if (fused_step) {
    const cudaError_t err = launch_steps(*fused_step, n_limit, stream);
    if (err != cudaSuccess) return err;
  }
  // persistent grid: enough CTAs to fill the machine several times over, never more than N
  int sm_count = 148;
  {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess)
      cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
  }
  // at most ~4 frames per CTA: amortises the per-CTA staging without turning the output into
  // thousands of far-apart write streams
  int64_t want = (int64_t)sm_count * 16;
  if ((p.N + 3) / 4 > want) want = (p.N + 3) / 4;
  const unsigned grid = (unsigned)(p.N < want ? p.N : want);
  if (p.dtype == 1) k_render<float><<<grid, kRenderBlock, smem, stream>>>(p);
  else k_render<uint8_t><<<grid, kRenderBlock, smem, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t configure_render_kernels() {
  // opt in to (almost) the full 227 KB of shared memory once per device; the kernels also use a
  // little static shared memory, which counts against the same limit
  const int cap = (int)kRenderMaxDynamicSmem;
  cudaError_t err =
      cudaFuncSetAttribute(k_render<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
  if (err == cudaSuccess)
    err = cudaFuncSetAttribute(k_render<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
  return err;
}

}  // namespace pw
