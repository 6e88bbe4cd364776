// Throughput of the Fq multiplier forms (dependent chain per thread, MSM-like occupancy).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/fqmul_forms tools/ubench/fqmul_forms.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../typlonk_b200/csrc/field.cuh"
using namespace tp;

template <int FORM>
__global__ void kform(uint32_t* out, const uint32_t* in, int iters) {
  Fq x, y;
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = 0; i < 12; i++) { x.v[i] = in[(t * 7 + i) & 1023]; y.v[i] = in[(t * 13 + i + 5) & 1023]; }
  x.v[11] &= 0x0fffffff; y.v[11] &= 0x0fffffff;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    Fq r;
    if (FORM == 0) fq_mul_ptx(r.v, x.v, y.v);
    if (FORM == 1) fq_mul_sep_ptx(r.v, x.v, y.v);
    if (FORM == 2) fq_mul_kar_ptx(r.v, x.v, y.v);
    if (FORM == 3) fq_sqr_ptx(r.v, x.v);
    if (FORM == 4) fq_mul2_ptx(r.v, x.v, y.v, y.v, x.v);
    if (FORM == 5) fq_mul2_kar_ptx(r.v, x.v, y.v, y.v, x.v);
    y = x; x = r;
  }
  uint32_t acc = 0;
  for (int i = 0; i < 12; i++) acc ^= x.v[i] ^ y.v[i];
  out[t] = acc;
}
template <int FORM>
void run(const char* name, int threads, int blocks_per_sm, double products) {
  int blocks = 148 * blocks_per_sm, iters = 400;
  uint32_t *out, *in;
  cudaMalloc(&out, blocks * threads * 4); cudaMalloc(&in, 4096);
  cudaMemset(in, 0x5a, 4096);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0); kform<FORM><<<blocks, threads>>>(out, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
  }
  double ops = (double)blocks * threads * iters;
  printf("%-34s warps/SM=%2d %8.3f ms  %.3e op/s  %.3e products/s\n", name, threads / 32 * blocks_per_sm, best,
         ops / (best * 1e-3), ops * products / (best * 1e-3));
  cudaFree(out); cudaFree(in);
}
int main() {
  for (int bps : {2, 3, 4, 8}) {
    run<0>("mul CIOS interleaved (288)", 128, bps, 288);
    run<1>("mul wide+redc schoolbook (288)", 128, bps, 288);
    run<2>("mul wide+redc Karatsuba (252)", 128, bps, 252);
    run<3>("sqr wide+redc (222)", 128, bps, 222);
    run<4>("lazy pair CIOS (432)", 128, bps, 432);
    run<5>("lazy pair Karatsuba (360)", 128, bps, 360);
  }
  return 0;
}
