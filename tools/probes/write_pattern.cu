// Synthetic probe: how fast can the [T][N] output pattern of a fused rollout be written at all?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/write_pattern tools/probes/write_pattern.cu && /tmp/write_pattern
// Variants: stores only (reward f32 + terminated u8, one row per step), stores + action reads, and
// different CTA sizes; same 2^18 x 1024 shape as BASELINE config 2.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <bool kReadActions, bool kPacked>
__global__ void k_pattern(const uint8_t* actions, float* reward, uint8_t* term, int N, int T) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N) return;
  uint32_t s = e;
  for (int t = 0; t < T; ++t) {
    if (kReadActions) s += __ldg(actions + (size_t)t * N + e);
    const size_t o = (size_t)t * N + e;
    if (!kPacked) reward[o] = (s & 1) ? 10.0f : -0.01f;
    term[o] = (uint8_t)(s & 1);
    s = s * 3 + 1;
  }
}

int main() {
  const int N = 1 << 18, T = 1024;
  uint8_t *actions, *term;
  float* reward;
  cudaMalloc(&actions, (size_t)N * T);
  cudaMalloc(&term, (size_t)N * T);
  cudaMalloc(&reward, (size_t)N * T * 4);
  cudaMemset(actions, 1, (size_t)N * T);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  for (int variant = 0; variant < 4; ++variant)
    for (int block : {128, 256, 448, 512, 896, 1024}) {
      float best = 1e9f;
      for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(a);
        const int grid = (N + block - 1) / block;
        if (variant == 0) k_pattern<false, false><<<grid, block>>>(actions, reward, term, N, T);
        if (variant == 1) k_pattern<true, false><<<grid, block>>>(actions, reward, term, N, T);
        if (variant == 2) k_pattern<false, true><<<grid, block>>>(actions, reward, term, N, T);
        if (variant == 3) k_pattern<true, true><<<grid, block>>>(actions, reward, term, N, T);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (rep >= 2 && ms < best) best = ms;
      }
      const double bytes = (double)N * T * ((variant & 1 ? 1 : 0) + (variant & 2 ? 1 : 5));
      printf("{\"read_actions\": %d, \"packed\": %d, \"block\": %d, \"ms\": %.4f, \"GBps\": %.0f}\n", variant & 1, variant >> 1,
             block, best, bytes / best / 1e6);
    }
  return 0;
}
