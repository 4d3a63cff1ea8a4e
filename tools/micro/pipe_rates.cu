// Micro-benchmark (experiment support, not product code): issue cost of the integer instruction
// kinds the bitboard kernels are made of, on sm_100a.  Each kernel runs ITER iterations of 8
// independent chains per thread; prints cycles per warp-instruction per SM sub-partition.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITER = 4096;

#define CHAINS8(OP) OP(0) OP(1) OP(2) OP(3) OP(4) OP(5) OP(6) OP(7)

template <int KIND>
__global__ void __launch_bounds__(1024) k(uint32_t* out, uint32_t seed, long long* cycles) {
  uint32_t x[8], y = seed * 3u + threadIdx.x, z = seed ^ 0x5555u;
  for (int i = 0; i < 8; ++i) x[i] = seed + threadIdx.x * 8u + i;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
    if (KIND == 0) {  // LOP3 word
#define OP(i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y), "r"(z));
      CHAINS8(OP)
#undef OP
    } else if (KIND == 1) {  // LOP3 -> predicate, consumed by SEL
#define OP(i) asm volatile("{.reg .pred p; .reg .b32 t; and.b32 t, %0, %1; setp.ne.u32 p, t, 0; selp.u32 %0, %2, %0, p;}" : "+r"(x[i]) : "r"(y), "r"(z));
      CHAINS8(OP)
#undef OP
    } else if (KIND == 2) {  // SEL with an invariant predicate + LOP3 word (same count as kind 1)
#define OP(i) asm volatile("{.reg .pred p; .reg .b32 t; setp.ne.u32 p, %2, 77; and.b32 t, %0, %1; selp.u32 %0, t, %0, p;}" : "+r"(x[i]) : "r"(y), "r"(z));
      CHAINS8(OP)
#undef OP
    } else if (KIND == 3) {  // mad.hi.s32 (nonzero -> mask)
#define OP(i) asm volatile("mad.hi.s32 %0, %0, -1000, %1;" : "+r"(x[i]) : "r"(y));
      CHAINS8(OP)
#undef OP
    } else if (KIND == 4) {  // funnel shift (rotate)
#define OP(i) asm volatile("shf.l.wrap.b32 %0, %0, %0, %1;" : "+r"(x[i]) : "r"(y));
      CHAINS8(OP)
#undef OP
    } else if (KIND == 5) {  // two predicates from two tests, combined (PLOP3), consumed by SEL
#define OP(i) asm volatile("{.reg .pred p, q; .reg .b32 t, u; and.b32 t, %0, %1; setp.ne.u32 p, t, 0; and.b32 u, %0, %2; setp.ne.u32 q, u, 0; and.pred p, p, q; selp.u32 %0, %2, %0, p;}" : "+r"(x[i]) : "r"(y), "r"(z));
      CHAINS8(OP)
#undef OP
    } else if (KIND == 6) {  // mad.lo (IMAD)
#define OP(i) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y), "r"(z));
      CHAINS8(OP)
#undef OP
    } else if (KIND == 7) {  // LOP3 word x4 + IMAD x4 interleaved (dual pipe)
#define OP(i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y), "r"(z));
      OP(0) OP(1) OP(2) OP(3)
#undef OP
#define OP(i) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y), "r"(z));
      OP(4) OP(5) OP(6) OP(7)
#undef OP
    } else if (KIND == 8) {  // SEL only: 1 ISETP + 8 SEL
      asm volatile("{.reg .pred p; setp.ne.u32 p, %9, 77;\n"
                   "selp.u32 %0, %8, %0, p; selp.u32 %1, %8, %1, p; selp.u32 %2, %8, %2, p; selp.u32 %3, %8, %3, p;\n"
                   "selp.u32 %4, %8, %4, p; selp.u32 %5, %8, %5, p; selp.u32 %6, %8, %6, p; selp.u32 %7, %8, %7, p;}"
                   : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7])
                   : "r"(y), "r"(z));
      y += 1;
    } else if (KIND == 9) {  // PLOP3 chain: 2 ISETP + 8 PLOP3 + 1 SEL
      asm volatile("{.reg .pred p, q; setp.ne.u32 p, %1, 77; setp.ne.u32 q, %2, 78;\n"
                   "xor.pred p, p, q; xor.pred q, p, q; xor.pred p, p, q; xor.pred q, p, q;\n"
                   "xor.pred p, p, q; xor.pred q, p, q; xor.pred p, p, q; xor.pred q, p, q;\n"
                   "selp.u32 %0, %1, %0, q;}"
                   : "+r"(x[0]) : "r"(y), "r"(z));
      y += 1;
    } else if (KIND == 10) {  // 6 LOP3->P + 6 SEL
      asm volatile("{.reg .pred p0, p1, p2, p3, p4, p5; .reg .b32 t;\n"
                   "and.b32 t, %0, %6; setp.ne.u32 p0, t, 0; and.b32 t, %1, %6; setp.ne.u32 p1, t, 0;\n"
                   "and.b32 t, %2, %6; setp.ne.u32 p2, t, 0; and.b32 t, %3, %6; setp.ne.u32 p3, t, 0;\n"
                   "and.b32 t, %4, %6; setp.ne.u32 p4, t, 0; and.b32 t, %5, %6; setp.ne.u32 p5, t, 0;\n"
                   "selp.u32 %0, %7, %0, p1; selp.u32 %1, %7, %1, p2; selp.u32 %2, %7, %2, p3;\n"
                   "selp.u32 %3, %7, %3, p4; selp.u32 %4, %7, %4, p5; selp.u32 %5, %7, %5, p0;}"
                   : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5])
                   : "r"(y), "r"(z));
      y += 1;
    } else if (KIND == 11) {  // 8 LOP3 (word) + 8 SEL with one invariant predicate, independent streams
      asm volatile("{.reg .pred p; setp.ne.u32 p, %9, 77;\n"
                   "selp.u32 %0, %8, %0, p; selp.u32 %1, %8, %1, p; selp.u32 %2, %8, %2, p; selp.u32 %3, %8, %3, p;\n"
                   "lop3.b32 %4, %4, %8, %9, 0x96; lop3.b32 %5, %5, %8, %9, 0x96; lop3.b32 %6, %6, %8, %9, 0x96; lop3.b32 %7, %7, %8, %9, 0x96;}"
                   : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7])
                   : "r"(y), "r"(z));
      y += 1;
    } else if (KIND == 12) {  // IMAD.WIDE.U32 (64-bit product of two 32-bit values), halves folded by a LOP3
      unsigned long long w;
#define OP(i) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(x[i]), "r"(y)); x[i] = (uint32_t)w ^ (uint32_t)(w >> 32);
      CHAINS8(OP)
#undef OP
    }
  }
  const long long t1 = clock64();
  uint32_t acc = 0;
  for (int i = 0; i < 8; ++i) acc ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int KIND>
void run(const char* name, int instr_per_iter) {
  uint32_t* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 148 * 8);
  k<KIND><<<148, 1024>>>(out, 1, cyc);
  k<KIND><<<148, 1024>>>(out, 2, cyc);  // one block per SM: 8 warps per sub-partition
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
  long long lo = h[0], hi = h[0];
  for (int i = 1; i < 148; ++i) { lo = h[i] < lo ? h[i] : lo; hi = h[i] > hi ? h[i] : hi; }
  const double per_iter = (double)lo / ITER;  // cycles for 8 warps x one iteration
  printf("%-44s %7.2f cycles / iteration of 8 warps (%2d instr each) -> IPC %.2f (max over SMs %+.1f%%)\n", name,
         per_iter, instr_per_iter, 8.0 * instr_per_iter / per_iter, 100.0 * (hi - lo) / lo);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0>("LOP3 x8", 8 + 3);
  run<4>("SHF x8", 8 + 3);
  run<6>("IMAD x8", 8 + 3);
  run<3>("IMAD.HI x8", 8 + 3);
  run<7>("LOP3 x4 + IMAD x4", 8 + 3);
  run<8>("SEL x8 (+ISETP, IADD)", 10 + 3);
  run<11>("SEL x4 + LOP3 x4 (+ISETP, IADD)", 10 + 3);
  run<9>("PLOP3 x8 (+2 ISETP, SEL, IADD)", 12 + 3);
  run<10>("LOP3->P x6 + SEL x6 (+IADD)", 13 + 3);
  run<12>("(IMAD.WIDE, LOP3) x8", 16 + 3);
  run<1>("(LOP3->P, SEL) x8", 16 + 3);
  run<5>("(2 LOP3->P, PLOP3, SEL) x8", 32 + 3);
  return 0;
}
