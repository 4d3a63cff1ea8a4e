// Synthetic probe 5: does the power-of-two row pitch matter?  Pattern of probe 3 (four environments per
// thread, reads 16 rows ahead) with the rows of all three arrays padded to `pitch` elements.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k_pattern(const uint8_t* actions, float* reward, uint8_t* term, int N, int T, size_t pitch) {
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (e >= N) return;
  uint32_t s = e;
  uint32_t w[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) w[k] = __ldcs(reinterpret_cast<const uint32_t*>(actions + (size_t)k * pitch + e));
  for (int t0 = 0; t0 < T; t0 += 16) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int t = t0 + k;
      s += w[k];
      const int tn = t + 16 < T ? t + 16 : T - 1;
      w[k] = __ldcs(reinterpret_cast<const uint32_t*>(actions + (size_t)tn * pitch + e));
      const size_t o = (size_t)t * pitch + e;
      *reinterpret_cast<float4*>(reward + o) = make_float4((s & 1) ? 10.f : -0.01f, (s & 2) ? 10.f : -0.01f, (s & 4) ? 10.f : -0.01f, (s & 8) ? 10.f : -0.01f);
      *reinterpret_cast<uint32_t*>(term + o) = s & 0x01010101u;
      s = s * 3 + 1;
    }
  }
}

int main() {
  const int N = 1 << 18, T = 1024;
  const size_t max_pitch = N + 65536;
  uint8_t *term, *actions;
  float* reward;
  cudaMalloc(&term, max_pitch * T);
  cudaMalloc(&actions, max_pitch * T);
  cudaMalloc(&reward, max_pitch * T * 4);
  cudaMemset(actions, 1, max_pitch * T);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  for (size_t pad : {0, 128, 512, 2048, 8192, 4096 + 128, 32768 + 2048}) {
    for (int block : {256, 512}) {
      float best = 1e9f;
      for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(a);
        k_pattern<<<(N / 4 + block - 1) / block, block>>>(actions, reward, term, N, T, N + pad);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (rep >= 2 && ms < best) best = ms;
      }
      printf("{\"pad\": %zu, \"block\": %d, \"ms\": %.4f, \"GBps\": %.0f}\n", pad, block, best, (double)N * T * 6 / best / 1e6);
    }
  }
  return 0;
}
