// Synthetic probe 3: the whole memory pattern of the fused plain rollout with four environments per thread and no
// arithmetic: per step a thread reads its 4 action bytes (one row of the [T][N] u8 action array, prefetched 16
// steps ahead) and writes a float4 of rewards and 4 terminated bytes.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <bool kRead, bool kReward, bool kTerm>
__global__ void k_pattern(const uint8_t* actions, float* reward, uint8_t* term, int N, int T) {
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (e >= N) return;
  uint32_t s = e;
  uint32_t w[16];
  if (kRead)
#pragma unroll
    for (int k = 0; k < 16; ++k) w[k] = __ldcs(reinterpret_cast<const uint32_t*>(actions + (size_t)k * N + e));
  for (int t0 = 0; t0 < T; t0 += 16) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int t = t0 + k;
      if (kRead) {
        s += w[k];
        const int tn = t + 16 < T ? t + 16 : T - 1;
        w[k] = __ldcs(reinterpret_cast<const uint32_t*>(actions + (size_t)tn * N + e));
      }
      const size_t o = (size_t)t * N + e;
      if (kReward)
        *reinterpret_cast<float4*>(reward + o) = make_float4((s & 1) ? 10.f : -0.01f, (s & 2) ? 10.f : -0.01f, (s & 4) ? 10.f : -0.01f, (s & 8) ? 10.f : -0.01f);
      if (kTerm) *reinterpret_cast<uint32_t*>(term + o) = s & 0x01010101u;
      s = s * 3 + 1;
    }
  }
}

template <bool kRead, bool kReward, bool kTerm>
void run(const char* name, const uint8_t* actions, float* reward, uint8_t* term, int N, int T) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  for (int block : {128, 256, 448, 512}) {
    float best = 1e9f;
    for (int rep = 0; rep < 6; ++rep) {
      cudaEventRecord(a);
      const int threads = N / 4;
      k_pattern<kRead, kReward, kTerm><<<(threads + block - 1) / block, block>>>(actions, reward, term, N, T);
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms;
      cudaEventElapsedTime(&ms, a, b);
      if (rep >= 2 && ms < best) best = ms;
    }
    const double bytes = (double)N * T * ((kRead ? 1 : 0) + (kReward ? 4 : 0) + (kTerm ? 1 : 0));
    printf("{\"variant\": \"%s\", \"block\": %d, \"ms\": %.4f, \"GBps\": %.0f}\n", name, block, best, bytes / best / 1e6);
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
  run<true, true, true>("read+reward+term", actions, reward, term, N, T);
  run<true, false, true>("read+term", actions, reward, term, N, T);
  run<true, false, false>("read", actions, reward, term, N, T);
  run<false, true, true>("reward+term", actions, reward, term, N, T);
  run<true, true, false>("read+reward", actions, reward, term, N, T);
  return 0;
}
