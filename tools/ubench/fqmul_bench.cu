// Throughput of the two Fq multipliers at MSM-like occupancies (dependent chain per thread).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../typlonk_b200/csrc/field.cuh"
#include "../../typlonk_b200/csrc/fq30.cuh"
using namespace tp;

template <int ILP>
__global__ void k32(uint32_t* out, const uint32_t* in, int iters) {
  Fq x[ILP], y[ILP];
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  for (int k = 0; k < ILP; k++) for (int i = 0; i < 12; i++) { x[k].v[i] = in[(t * 7 + i + k) & 1023]; y[k].v[i] = in[(t * 13 + i + 5 * k) & 1023]; }
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) { Fq r = fq_mul(x[k], y[k]); y[k] = x[k]; x[k] = r; }
  }
  uint32_t acc = 0;
  for (int k = 0; k < ILP; k++) for (int i = 0; i < 12; i++) acc ^= x[k].v[i];
  out[t] = acc;
}
template <int ILP>
__global__ void k30(uint32_t* out, const uint32_t* in, int iters) {
  Fq30 x[ILP], y[ILP];
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  for (int k = 0; k < ILP; k++) for (int i = 0; i < 13; i++) { x[k].l[i] = in[(t * 7 + i + k) & 1023] & 0x3fffffff; y[k].l[i] = in[(t * 13 + i + 5 * k) & 1023] & 0x3fffffff; }
  for (int k = 0; k < ILP; k++) { x[k].l[12] &= 0xfffff; y[k].l[12] &= 0xfffff; }
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) { Fq30 r = fq30_mul(x[k], y[k]); y[k] = x[k]; x[k] = r; }
  }
  uint32_t acc = 0;
  for (int k = 0; k < ILP; k++) for (int i = 0; i < 13; i++) acc ^= x[k].l[i];
  out[t] = acc;
}
template <typename F>
void run(const char* name, F launch, int threads, int blocks_per_sm, int ilp) {
  int blocks = 148 * blocks_per_sm, iters = 400;
  uint32_t *out, *in;
  cudaMalloc(&out, blocks * threads * 4); cudaMalloc(&in, 4096);
  cudaMemset(in, 0x5a, 4096);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0); launch(blocks, threads, out, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
  }
  double muls = (double)blocks * threads * iters * ilp;
  printf("%-28s warps/SM=%2d ilp=%d  %8.3f ms  %.3e mul/s\n", name, threads / 32 * blocks_per_sm, ilp, best, muls / (best * 1e-3));
  cudaFree(out); cudaFree(in);
}
int main() {
  for (int bps : {1, 2, 4}) {
    int th = 128;
    run("fq32 carry-chain PTX", [](int b, int t, uint32_t* o, const uint32_t* i, int it) { k32<1><<<b, t>>>(o, i, it); }, th, bps * 1, 1);
    run("fq32 carry-chain PTX", [](int b, int t, uint32_t* o, const uint32_t* i, int it) { k32<2><<<b, t>>>(o, i, it); }, th, bps * 1, 2);
    run("fq30 reduced radix", [](int b, int t, uint32_t* o, const uint32_t* i, int it) { k30<1><<<b, t>>>(o, i, it); }, th, bps * 1, 1);
    run("fq30 reduced radix", [](int b, int t, uint32_t* o, const uint32_t* i, int it) { k30<2><<<b, t>>>(o, i, it); }, th, bps * 1, 2);
  }
  for (int bps : {2, 4}) {
    run("fq32 (256 thr)", [](int b, int t, uint32_t* o, const uint32_t* i, int it) { k32<1><<<b, t>>>(o, i, it); }, 256, bps, 1);
    run("fq30 (256 thr)", [](int b, int t, uint32_t* o, const uint32_t* i, int it) { k30<1><<<b, t>>>(o, i, it); }, 256, bps, 1);
  }
  return 0;
}
