// pw_render.cu -- RGB observation kernel (optionally fused with one environment step).
//
// One CTA per environment:
//   0. (fused only) thread 0 performs the step on shared-memory state (BASELINE config 5);
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
#include "pw_steps.cuh"

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

// Copies `bytes` from shared `src` to global `dst` (same alignment class decided by caller).
template <int VEC>
__device__ __forceinline__ void copy_row(uint8_t* dst, const uint8_t* src, int bytes) {
  if (VEC == 16) {
    const int n = bytes >> 4;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
      reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
  } else if (VEC == 4) {
    const int n = bytes >> 2;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
      reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(src)[i];
  } else {
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) dst[i] = src[i];
  }
}

template <typename T, bool kFused>
__global__ void __launch_bounds__(kRenderBlock) k_render(const RenderParams p, const StepParams sp) {
  extern __shared__ __align__(16) uint64_t smem[];
  const int64_t e = blockIdx.x;
  const int pid = p.puzzle_id ? p.puzzle_id[e] : p.puzzle_uniform;
  const uint64_t* g_tbl = p.blob + p.puzzle_off[pid];

  // ---- shared memory carve-up -----------------------------------------------------------
  uint64_t* s_tbl = smem;                          // max_blob_u64
  uint64_t* s_code = s_tbl + p.max_blob_u64;       // max_cells
  int* s_x = reinterpret_cast<int*>(s_code + p.max_cells);  // n_max (padded to 32)
  int* s_y = s_x + 32;
  const int row_bytes = p.out_w * 3 * (int)sizeof(T);
  const int row_pitch = (row_bytes + 15) & ~15;
  // 4 patterns (3 region rows + a zero row), 16-byte aligned for the 128-bit copies
  uint8_t* s_pat = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(s_y + 32) + 15) & ~static_cast<uintptr_t>(15));

  const int n_u64 = (int)__ldg(reinterpret_cast<const uint32_t*>(g_tbl) + kHdrSizeU64);
  for (int i = threadIdx.x; i < n_u64; i += blockDim.x) s_tbl[i] = __ldg(g_tbl + i);
  __syncthreads();
  const SharedTable tbl{s_tbl};
  const uint32_t dims = tbl.word(kHdrDims);
  const int W = (int)(dims & 0xffffu), H = (int)(dims >> 16);
  const int n = (int)(tbl.word(kHdrCounts) & 0xffffu);

  // ---- state (and the fused step) -----------------------------------------------------
  if (kFused) {
    if (threadIdx.x == 0) {
      SmemEnv<SharedTable> env;
      env.init(tbl, s_x, s_y, 1);
      StepCounts counts;
      run_steps_thread<false>(env, sp, e, counts);
      if (counts.invalid && sp.status) atomicAdd(&sp.status[0], (unsigned long long)counts.invalid);
    }
  } else {
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      s_x[j] = p.x[j * p.stride + e];
      s_y[j] = p.y[j * p.stride + e];
    }
  }
  __syncthreads();

  // ---- 1. cell codes --------------------------------------------------------------------
  const SmemPos pos{s_x, s_y};
  for (int c = threadIdx.x; c < W * H; c += blockDim.x) {
    const int cy = c / W, cx = c - cy * W;
    s_code[c] = cell_code(tbl, pos, cx, cy);
  }
  // zero pattern (index 3)
  for (int i = threadIdx.x; i < (row_pitch >> 2); i += blockDim.x)
    reinterpret_cast<uint32_t*>(s_pat + 3 * row_pitch)[i] = 0u;
  __syncthreads();

  // ---- 2. pixel rows --------------------------------------------------------------------
  const int ppc = p.ppc, bw = p.bw;
  const int img_h = H * ppc, img_w = W * ppc;
  const int pad_top = (p.out_h - img_h) / 2, pad_left = (p.out_w - img_w) / 2;  // env_utils.py:81-82
  uint8_t* frame = static_cast<uint8_t*>(p.out) + (size_t)e * (size_t)p.out_h * (size_t)row_bytes;
  const bool vec16 = (row_bytes & 15) == 0 && (reinterpret_cast<uintptr_t>(frame) & 15) == 0;
  const bool vec4 = (row_bytes & 3) == 0 && (reinterpret_cast<uintptr_t>(frame) & 3) == 0;

  auto emit = [&](int Y, const uint8_t* src) {
    uint8_t* dst = frame + (size_t)Y * (size_t)row_bytes;
    if (vec16) copy_row<16>(dst, src, row_bytes);
    else if (vec4) copy_row<4>(dst, src, row_bytes);
    else copy_row<1>(dst, src, row_bytes);
  };

  for (int Y = 0; Y < pad_top; ++Y) emit(Y, s_pat + 3 * row_pitch);
  for (int Y = pad_top + img_h; Y < p.out_h; ++Y) emit(Y, s_pat + 3 * row_pitch);

  for (int cy = 0; cy < H; ++cy) {
    // build the three patterns of this cell row: element (rr, X, channel)
    for (int i = threadIdx.x; i < 3 * p.out_w; i += blockDim.x) {
      const int rr = i / p.out_w, X = i - rr * p.out_w;
      const int xi = X - pad_left;
      uint32_t pal = kPalPad;
      if (xi >= 0 && xi < img_w) {
        const int cx = xi / ppc, ox = xi - cx * ppc;
        const int rc = region_of(ox, ppc, bw);
        pal = (uint32_t)(s_code[cy * W + cx] >> (4 * (rr * 3 + rc))) & 15u;
      }
      T* dst = reinterpret_cast<T*>(s_pat + rr * row_pitch) + 3 * X;
      dst[0] = Elem<T>::get(pal, 0);
      dst[1] = Elem<T>::get(pal, 1);
      dst[2] = Elem<T>::get(pal, 2);
    }
    __syncthreads();
    const int Y0 = pad_top + cy * ppc;
    for (int oy = 0; oy < ppc; ++oy) emit(Y0 + oy, s_pat + region_of(oy, ppc, bw) * row_pitch);
    __syncthreads();
  }
}

}  // namespace

cudaError_t launch_render(const RenderParams& p, const StepParams* fused_step, int n_limit,
                          cudaStream_t stream) {
  (void)n_limit;
  if (p.N == 0) return cudaSuccess;
  const int elem = p.dtype == 1 ? 4 : 1;
  const size_t row_pitch = ((size_t)p.out_w * 3 * elem + 15) & ~(size_t)15;
  const size_t smem = (size_t)p.max_blob_u64 * 8 + (size_t)p.max_cells * 8 + 64 * sizeof(int) +
                      16 + 4 * row_pitch;
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  StepParams sp{};
  if (fused_step) sp = *fused_step;
#define PW_LAUNCH(T, F) k_render<T, F><<<(unsigned)p.N, kRenderBlock, smem, stream>>>(p, sp)
  if (p.dtype == 1) {
    if (fused_step) PW_LAUNCH(float, true);
    else PW_LAUNCH(float, false);
  } else {
    if (fused_step) PW_LAUNCH(uint8_t, true);
    else PW_LAUNCH(uint8_t, false);
  }
#undef PW_LAUNCH
  return cudaGetLastError();
}

cudaError_t configure_render_kernels() {
  // opt in to the full 227 KB of dynamic shared memory once per device
  const int cap = 227 * 1024;
  cudaError_t err = cudaFuncSetAttribute(k_render<float, false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
  if (err == cudaSuccess)
    err = cudaFuncSetAttribute(k_render<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
  if (err == cudaSuccess)
    err = cudaFuncSetAttribute(k_render<uint8_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
  if (err == cudaSuccess)
    err = cudaFuncSetAttribute(k_render<uint8_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
  return err;
}

}  // namespace pw
