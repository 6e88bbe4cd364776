// Microbenchmark (B200, sm_100a): issue rate of the multiply instruction forms a big-integer
// Montgomery multiplier can be built from.  Every chain feeds its own result back as a multiplier
// operand, so nothing can be hoisted or strength-reduced (the first version of this benchmark,
// imad_forms.cu MODE 0, was folded by the compiler into one IMAD.WIDE + adds and over-reported).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipes pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

enum { WIDE = 0, WIDE_X = 1, LO = 2, HI = 3, DFMA = 4, WIDE_PLUS_DFMA = 5, WIDE_PLUS_ALU = 6, LO_PLUS_HI = 7, FFMA = 8 };

template <int MODE, int CH>
__global__ void k(unsigned* out, unsigned seed, int iters) {
  unsigned lo[CH], hi[CH], b[CH];
  double d[CH], e[CH];
  unsigned long long x[CH];
  for (int i = 0; i < CH; i++) {
    lo[i] = threadIdx.x * 7 + i + seed;
    hi[i] = threadIdx.x * 3 + i;
    b[i] = (threadIdx.x + i * 11 + seed) | 1;
    d[i] = 1.0 + 1e-9 * (threadIdx.x + i);
    e[i] = 1.0 - 1e-9 * (threadIdx.x + 2 * i);
    x[i] = ((unsigned long long)hi[i] << 32) | lo[i];
  }
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int i = 0; i < CH; i++) {
        if (MODE == WIDE || MODE == WIDE_PLUS_DFMA || MODE == WIDE_PLUS_ALU) {
          // (hi:lo) = lo * b + (hi:lo)   -> IMAD.WIDE.U32 with a 64-bit addend, multiplier depends on the chain
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[i]) : "r"((unsigned)x[(i + 1) % CH]), "r"(b[i]));
        }
        if (MODE == WIDE_X) {
          asm volatile("mad.lo.cc.u32 %0, %0, %2, %0;\n\tmadc.hi.u32 %1, %0, %2, %1;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(b[i]));
        }
        if (MODE == LO) asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(lo[i]) : "r"(b[i]));
        if (MODE == HI) asm volatile("mad.hi.u32 %0, %0, %1, %0;" : "+r"(lo[i]) : "r"(b[i]));
        if (MODE == LO_PLUS_HI) {
          asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(lo[i]) : "r"(b[i]));
          asm volatile("mad.hi.u32 %0, %0, %1, %0;" : "+r"(hi[i]) : "r"(b[i]));
        }
        if (MODE == DFMA || MODE == WIDE_PLUS_DFMA) asm volatile("fma.rz.f64 %0, %0, %1, %0;" : "+d"(d[i]) : "d"(e[i]));
        if (MODE == WIDE_PLUS_ALU) {
          asm volatile("add.u32 %0, %0, %1;" : "+r"(hi[i]) : "r"(lo[i]));
          asm volatile("shf.r.wrap.b32 %0, %0, %1, 30;" : "+r"(lo[i]) : "r"(hi[i]));
        }
        if (MODE == FFMA) {
          float f = __uint_as_float(lo[i]);
          asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(f) : "f"(__uint_as_float(b[i])));
          lo[i] = __float_as_uint(f);
        }
      }
    }
  }
  unsigned r = 0;
  for (int i = 0; i < CH; i++) r ^= lo[i] ^ hi[i] ^ b[i] ^ (unsigned)x[i] ^ (unsigned)(x[i] >> 32) ^ (unsigned)__double2uint_rz(d[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE, int CH>
void run(const char* name, int warps_per_sm) {
  int threads = 128, blocks = 148 * warps_per_sm / 4, iters = 2000;
  unsigned* out;
  cudaMalloc(&out, (size_t)blocks * threads * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e9;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    k<MODE, CH><<<blocks, threads>>>(out, 12345 + rep, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  double ops = (double)blocks * threads * iters * 8 * CH;  // "ops" = one unit of the mode (see names)
  double per_clk_sm = ops / (best * 1e-3) / 148 / 1.965e9;
  // per-warp latency view: clocks per unit per warp when a single warp per SMSP runs
  printf("%-44s chains=%2d warps/SM=%2d  %8.3f ms  %6.2f units/clk/SM  (%.2f clk per warp-instr-unit per SMSP)\n", name, CH,
         warps_per_sm, best, per_clk_sm, 128.0 / per_clk_sm);
  cudaFree(out);
}

#define SWEEP(MODE, NAME)                \
  run<MODE, 1>(NAME, 4);                 \
  run<MODE, 4>(NAME, 4);                 \
  run<MODE, 8>(NAME, 4);                 \
  run<MODE, 8>(NAME, 16);                \
  run<MODE, 8>(NAME, 32);

int main() {
  SWEEP(WIDE, "IMAD.WIDE.U32 (64-bit addend, no carry)");
  SWEEP(WIDE_X, "mad.lo.cc + madc.hi (1 unit = 1 product)");
  SWEEP(LO, "IMAD (lo 32)");
  SWEEP(HI, "IMAD.HI");
  SWEEP(LO_PLUS_HI, "IMAD lo + IMAD.HI (1 unit = both)");
  SWEEP(DFMA, "DFMA");
  SWEEP(FFMA, "FFMA");
  SWEEP(WIDE_PLUS_DFMA, "IMAD.WIDE + DFMA (1 unit = both)");
  SWEEP(WIDE_PLUS_ALU, "IMAD.WIDE + IADD3 + SHF (1 unit = all 3)");
  return 0;
}
