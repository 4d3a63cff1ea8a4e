// Synthetic probe 2: the [T][N] output pattern written by threads that own kEnvs consecutive environments
// (float4 / uchar4 stores per step), and reward-only / terminated-only variants.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int kEnvs, bool kReward, bool kTerm>
__global__ void k_pattern(float* reward, uint8_t* term, int N, int T) {
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) * kEnvs;
  if (e >= N) return;
  uint32_t s = e;
  for (int t = 0; t < T; ++t) {
    const size_t o = (size_t)t * N + e;
    if (kReward) {
      if (kEnvs == 4) *reinterpret_cast<float4*>(reward + o) = make_float4((s & 1) ? 10.f : -0.01f, (s & 2) ? 10.f : -0.01f, (s & 4) ? 10.f : -0.01f, (s & 8) ? 10.f : -0.01f);
      else if (kEnvs == 2) *reinterpret_cast<float2*>(reward + o) = make_float2((s & 1) ? 10.f : -0.01f, (s & 2) ? 10.f : -0.01f);
      else reward[o] = (s & 1) ? 10.f : -0.01f;
    }
    if (kTerm) {
      if (kEnvs == 4) *reinterpret_cast<uint32_t*>(term + o) = s & 0x01010101u;
      else if (kEnvs == 2) *reinterpret_cast<uint16_t*>(term + o) = (uint16_t)(s & 0x0101u);
      else term[o] = (uint8_t)(s & 1);
    }
    s = s * 3 + 1;
  }
}

template <int kEnvs, bool kReward, bool kTerm>
void run(const char* name, float* reward, uint8_t* term, int N, int T) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  for (int block : {128, 256, 448, 512, 896}) {
    float best = 1e9f;
    for (int rep = 0; rep < 6; ++rep) {
      cudaEventRecord(a);
      const int threads = N / kEnvs;
      k_pattern<kEnvs, kReward, kTerm><<<(threads + block - 1) / block, block>>>(reward, term, N, T);
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms;
      cudaEventElapsedTime(&ms, a, b);
      if (rep >= 2 && ms < best) best = ms;
    }
    const double bytes = (double)N * T * ((kReward ? 4 : 0) + (kTerm ? 1 : 0));
    printf("{\"variant\": \"%s\", \"envs_per_thread\": %d, \"block\": %d, \"ms\": %.4f, \"GBps\": %.0f}\n", name, kEnvs, block, best, bytes / best / 1e6);
  }
}

int main() {
  const int N = 1 << 18, T = 1024;
  uint8_t* term;
  float* reward;
  cudaMalloc(&term, (size_t)N * T);
  cudaMalloc(&reward, (size_t)N * T * 4);
  run<1, true, true>("reward+term", reward, term, N, T);
  run<2, true, true>("reward+term", reward, term, N, T);
  run<4, true, true>("reward+term", reward, term, N, T);
  run<1, true, false>("reward", reward, term, N, T);
  run<4, true, false>("reward", reward, term, N, T);
  run<1, false, true>("term", reward, term, N, T);
  run<2, false, true>("term", reward, term, N, T);
  run<4, false, true>("term", reward, term, N, T);
  return 0;
}
