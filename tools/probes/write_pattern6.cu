// Synthetic probe 6: prefetch distance of the action reads (rows ahead) on the full memory pattern
// (four environments per thread, no arithmetic), 128 CTAs x 512 threads and 256 x 256.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int kAhead>
__global__ void __launch_bounds__(512) k_pattern(const uint8_t* actions, float* reward, uint8_t* term, int N, int T) {
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (e >= N) return;
  uint32_t s = e;
  uint32_t w[kAhead];
#pragma unroll
  for (int k = 0; k < kAhead; ++k) w[k] = __ldcs(reinterpret_cast<const uint32_t*>(actions + (size_t)k * N + e));
  for (int t0 = 0; t0 < T; t0 += kAhead) {
#pragma unroll
    for (int k = 0; k < kAhead; ++k) {
      const int t = t0 + k;
      s += w[k];
      const int tn = t + kAhead < T ? t + kAhead : T - 1;
      w[k] = __ldcs(reinterpret_cast<const uint32_t*>(actions + (size_t)tn * N + e));
      const size_t o = (size_t)t * N + e;
      *reinterpret_cast<float4*>(reward + o) = make_float4((s & 1) ? 10.f : -0.01f, (s & 2) ? 10.f : -0.01f, (s & 4) ? 10.f : -0.01f, (s & 8) ? 10.f : -0.01f);
      *reinterpret_cast<uint32_t*>(term + o) = s & 0x01010101u;
      s = s * 3 + 1;
    }
  }
}

template <int kAhead>
void run(const uint8_t* actions, float* reward, uint8_t* term, int N, int T) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  for (int block : {256, 512}) {
    float best = 1e9f;
    for (int rep = 0; rep < 6; ++rep) {
      cudaEventRecord(a);
      k_pattern<kAhead><<<(N / 4 + block - 1) / block, block>>>(actions, reward, term, N, T);
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms;
      cudaEventElapsedTime(&ms, a, b);
      if (rep >= 2 && ms < best) best = ms;
    }
    printf("{\"rows_ahead\": %d, \"block\": %d, \"ms\": %.4f, \"GBps\": %.0f}\n", kAhead, block, best, (double)N * T * 6 / best / 1e6);
  }
}

int main() {
  const int N = 1 << 18, T = 1024;
  uint8_t *term, *actions;
  float* reward;
  cudaMalloc(&term, (size_t)N * T);
  cudaMalloc(&actions, (size_t)N * T);
  cudaMalloc(&reward, (size_t)N * T * 4);
  cudaMemset(actions, 1, (size_t)N * T);
  run<4>(actions, reward, term, N, T);
  run<8>(actions, reward, term, N, T);
  run<16>(actions, reward, term, N, T);
  run<32>(actions, reward, term, N, T);
  run<64>(actions, reward, term, N, T);
  return 0;
}
