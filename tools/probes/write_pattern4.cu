// Synthetic probe 4: cache hints for the stores and an L2 prefetch for the action rows, on the full memory
// pattern of probe 3 (four environments per thread, no arithmetic), 128 CTAs x 512 threads.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int kHint>
__device__ __forceinline__ void store4(float* p, float4 v) {
  if (kHint == 0) asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  if (kHint == 1) asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  if (kHint == 2) asm volatile("st.global.wt.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  if (kHint == 3) asm volatile("st.global.cg.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
template <int kHint>
__device__ __forceinline__ void store1(uint8_t* p, uint32_t v) {
  if (kHint == 0) asm volatile("st.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
  if (kHint == 1) asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
  if (kHint == 2) asm volatile("st.global.wt.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
  if (kHint == 3) asm volatile("st.global.cg.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int kHint, int kPrefetch>
__global__ void k_pattern(const uint8_t* actions, float* reward, uint8_t* term, int N, int T) {
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (e >= N) return;
  const int lane = threadIdx.x & 31;
  const uint8_t* warp_actions = actions + (e - lane * 4);
  uint32_t s = e;
  uint32_t w[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) w[k] = __ldcs(reinterpret_cast<const uint32_t*>(actions + (size_t)k * N + e));
  for (int t0 = 0; t0 < T; t0 += 16) {
    if (kPrefetch && (t0 & 31) == 0) {  // lane l: the warp's 128-byte line of row t0 + kPrefetch + l
      const int row = t0 + kPrefetch + lane;
      if (row < T) asm volatile("prefetch.global.L2 [%0];" ::"l"(warp_actions + (size_t)row * N));
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int t = t0 + k;
      s += w[k];
      const int tn = t + 16 < T ? t + 16 : T - 1;
      w[k] = __ldcs(reinterpret_cast<const uint32_t*>(actions + (size_t)tn * N + e));
      const size_t o = (size_t)t * N + e;
      store4<kHint>(reward + o, make_float4((s & 1) ? 10.f : -0.01f, (s & 2) ? 10.f : -0.01f, (s & 4) ? 10.f : -0.01f, (s & 8) ? 10.f : -0.01f));
      store1<kHint>(term + o, s & 0x01010101u);
      s = s * 3 + 1;
    }
  }
}

template <int kHint, int kPrefetch>
void run(const char* name, const uint8_t* actions, float* reward, uint8_t* term, int N, int T) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  for (int block : {256, 512}) {
    float best = 1e9f;
    for (int rep = 0; rep < 6; ++rep) {
      cudaEventRecord(a);
      const int threads = N / 4;
      k_pattern<kHint, kPrefetch><<<(threads + block - 1) / block, block>>>(actions, reward, term, N, T);
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms;
      cudaEventElapsedTime(&ms, a, b);
      if (rep >= 2 && ms < best) best = ms;
    }
    printf("{\"variant\": \"%s\", \"block\": %d, \"ms\": %.4f, \"GBps\": %.0f}\n", name, block, best, (double)N * T * 6 / best / 1e6);
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
  run<0, 0>("default", actions, reward, term, N, T);
  run<1, 0>("st.cs", actions, reward, term, N, T);
  run<2, 0>("st.wt", actions, reward, term, N, T);
  run<3, 0>("st.cg", actions, reward, term, N, T);
  run<0, 64>("prefetch64", actions, reward, term, N, T);
  run<0, 128>("prefetch128", actions, reward, term, N, T);
  run<1, 128>("st.cs+prefetch128", actions, reward, term, N, T);
  run<0, 512>("prefetch512", actions, reward, term, N, T);
  return 0;
}
