// Microbenchmark: throughput of the integer multiply-add instruction forms the Montgomery
// kernels can be built from (B200, sm_100a).  8 independent chains per thread, 2 x 1024 threads/SM.
#include <cstdio>
#include <cuda_runtime.h>
#define REP8(x) x x x x x x x x
template <int MODE>
__global__ void k(unsigned* out, unsigned a, unsigned b, int iters) {
  unsigned lo[8], hi[8];
  for (int i = 0; i < 8; i++) { lo[i] = threadIdx.x + i; hi[i] = threadIdx.x * 3 + i; }
  unsigned m = a + threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (MODE == 0) {  // plain mad.wide (IMAD.WIDE.U32), independent accumulators
#pragma unroll
        for (int i = 0; i < 8; i++) {
          unsigned long long x = ((unsigned long long)hi[i] << 32) | lo[i];
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x) : "r"(m), "r"(b));
          lo[i] = (unsigned)x; hi[i] = (unsigned)(x >> 32);
        }
      } else if (MODE == 1) {  // one carry chain through all 8 pairs: mad.lo.cc / madc.hi.cc (IMAD.WIDE.U32.X)
        asm volatile(
            "mad.lo.cc.u32 %0, %16, %17, %0;\n\tmadc.hi.cc.u32 %1, %16, %17, %1;\n\t"
            "madc.lo.cc.u32 %2, %16, %17, %2;\n\tmadc.hi.cc.u32 %3, %16, %17, %3;\n\t"
            "madc.lo.cc.u32 %4, %16, %17, %4;\n\tmadc.hi.cc.u32 %5, %16, %17, %5;\n\t"
            "madc.lo.cc.u32 %6, %16, %17, %6;\n\tmadc.hi.cc.u32 %7, %16, %17, %7;\n\t"
            "madc.lo.cc.u32 %8, %16, %17, %8;\n\tmadc.hi.cc.u32 %9, %16, %17, %9;\n\t"
            "madc.lo.cc.u32 %10, %16, %17, %10;\n\tmadc.hi.cc.u32 %11, %16, %17, %11;\n\t"
            "madc.lo.cc.u32 %12, %16, %17, %12;\n\tmadc.hi.cc.u32 %13, %16, %17, %13;\n\t"
            "madc.lo.cc.u32 %14, %16, %17, %14;\n\tmadc.hi.u32 %15, %16, %17, %15;"
            : "+r"(lo[0]), "+r"(hi[0]), "+r"(lo[1]), "+r"(hi[1]), "+r"(lo[2]), "+r"(hi[2]), "+r"(lo[3]), "+r"(hi[3]),
              "+r"(lo[4]), "+r"(hi[4]), "+r"(lo[5]), "+r"(hi[5]), "+r"(lo[6]), "+r"(hi[6]), "+r"(lo[7]), "+r"(hi[7])
            : "r"(m), "r"(b));
      } else if (MODE == 2) {  // 8 independent 1-link chains with carry-out only (mad.lo.cc + madc.hi)
#pragma unroll
        for (int i = 0; i < 8; i++)
          asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(m), "r"(b));
      } else if (MODE == 3) {  // separate IMAD lo + IMAD.HI (no carry)
#pragma unroll
        for (int i = 0; i < 8; i++)
          asm volatile("mad.lo.u32 %0, %2, %3, %0;\n\tmad.hi.u32 %1, %2, %3, %1;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(m), "r"(b));
      } else if (MODE == 4) {  // add.cc / addc chains (IADD3.X) 16 links
        asm volatile(
            "add.cc.u32 %0, %0, %16;\n\taddc.cc.u32 %1, %1, %17;\n\taddc.cc.u32 %2, %2, %16;\n\taddc.cc.u32 %3, %3, %17;\n\t"
            "addc.cc.u32 %4, %4, %16;\n\taddc.cc.u32 %5, %5, %17;\n\taddc.cc.u32 %6, %6, %16;\n\taddc.cc.u32 %7, %7, %17;\n\t"
            "addc.cc.u32 %8, %8, %16;\n\taddc.cc.u32 %9, %9, %17;\n\taddc.cc.u32 %10, %10, %16;\n\taddc.cc.u32 %11, %11, %17;\n\t"
            "addc.cc.u32 %12, %12, %16;\n\taddc.cc.u32 %13, %13, %17;\n\taddc.cc.u32 %14, %14, %16;\n\taddc.u32 %15, %15, %17;"
            : "+r"(lo[0]), "+r"(hi[0]), "+r"(lo[1]), "+r"(hi[1]), "+r"(lo[2]), "+r"(hi[2]), "+r"(lo[3]), "+r"(hi[3]),
              "+r"(lo[4]), "+r"(hi[4]), "+r"(lo[5]), "+r"(hi[5]), "+r"(lo[6]), "+r"(hi[6]), "+r"(lo[7]), "+r"(hi[7])
            : "r"(m), "r"(b));
      } else if (MODE == 5) {  // two independent 4-pair carry chains (closer to the even/odd structure)
        asm volatile(
            "mad.lo.cc.u32 %0, %16, %17, %0;\n\tmadc.hi.cc.u32 %1, %16, %17, %1;\n\t"
            "madc.lo.cc.u32 %2, %16, %17, %2;\n\tmadc.hi.cc.u32 %3, %16, %17, %3;\n\t"
            "madc.lo.cc.u32 %4, %16, %17, %4;\n\tmadc.hi.cc.u32 %5, %16, %17, %5;\n\t"
            "madc.lo.cc.u32 %6, %16, %17, %6;\n\tmadc.hi.u32 %7, %16, %17, %7;\n\t"
            "mad.lo.cc.u32 %8, %16, %17, %8;\n\tmadc.hi.cc.u32 %9, %16, %17, %9;\n\t"
            "madc.lo.cc.u32 %10, %16, %17, %10;\n\tmadc.hi.cc.u32 %11, %16, %17, %11;\n\t"
            "madc.lo.cc.u32 %12, %16, %17, %12;\n\tmadc.hi.cc.u32 %13, %16, %17, %13;\n\t"
            "madc.lo.cc.u32 %14, %16, %17, %14;\n\tmadc.hi.u32 %15, %16, %17, %15;"
            : "+r"(lo[0]), "+r"(hi[0]), "+r"(lo[1]), "+r"(hi[1]), "+r"(lo[2]), "+r"(hi[2]), "+r"(lo[3]), "+r"(hi[3]),
              "+r"(lo[4]), "+r"(hi[4]), "+r"(lo[5]), "+r"(hi[5]), "+r"(lo[6]), "+r"(hi[6]), "+r"(lo[7]), "+r"(hi[7])
            : "r"(m), "r"(b));
      }
    }
  }
  unsigned r = 0;
  for (int i = 0; i < 8; i++) r ^= lo[i] ^ hi[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE>
void run(const char* name, double ops_per_iter) {
  int dev_sms = 148, threads = 1024, blocks = dev_sms * 2, iters = 1000;
  unsigned* out;
  cudaMalloc(&out, blocks * threads * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, 12345, 678, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  double total = (double)blocks * threads * iters * 4 * ops_per_iter;
  printf("%-52s %8.3f ms  %.3e products/s  (%.1f per clk per SM @1.965GHz)\n", name, best, total / (best * 1e-3),
         total / (best * 1e-3) / 148 / 1.965e9);
  cudaFree(out);
}
int main() {
  run<0>("mad.wide.u32 (IMAD.WIDE.U32), 8 indep", 8);
  run<1>("mad.lo.cc/madc.hi.cc one 8-pair chain (.X)", 8);
  run<5>("two 4-pair chains", 8);
  run<2>("mad.lo.cc + madc.hi (carry-out only), 8 indep", 8);
  run<3>("mad.lo + mad.hi separate (no carry)", 8);
  run<4>("add.cc/addc 16-link chain (counts adds/2)", 8);
  return 0;
}
