// Device self-test (PTX field/curve arithmetic vs the host implementation in host_field.h) and
// the integer-pipe microbenchmark that supplies the measured IMAD roofline.
#include "common.cuh"

namespace tp {

struct SelfIn {
  Fr fa, fb;
  Fq qa, qb;
};
struct SelfOut {
  Fr mul, add, sub, inv, frommont;
  Fq qmul, qadd, qsub, qinv, qsqr, qlazy, qkar, qsep;
  G1Xyzz madd, dbl, add2;
};

__global__ void k_selftest(const SelfIn* in, SelfOut* out, const G1Affine* pts, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  SelfIn x = in[i];
  SelfOut o;
  o.mul = fr_mul(x.fa, x.fb);
  o.add = fr_add(x.fa, x.fb);
  o.sub = fr_sub(x.fa, x.fb);
  o.inv = fr_inv(x.fa);
  o.frommont = fr_from_mont(x.fa);
  o.qmul = fq_mul(x.qa, x.qb);
  o.qadd = fq_add(x.qa, x.qb);
  o.qsub = fq_sub(x.qa, x.qb);
  o.qinv = fq_inv(x.qa);
  o.qsqr = fq_sqr(x.qa);
  o.qlazy = fq_mul_sub2(x.qa, x.qb, x.qb, x.qb);  // a*b - b*b
  fq_mul_kar_ptx(o.qkar.v, x.qa.v, x.qb.v);
  fq_mul_sep_ptx(o.qsep.v, x.qa.v, x.qb.v);
  // points: P = pts[i], Q = pts[(i+1)%n]
  G1Affine p = affine_load(pts + i), q = affine_load(pts + (i + 1) % n);
  G1Xyzz acc = xyzz_identity();
  xyzz_madd(acc, p, false);
  xyzz_madd(acc, q, (i & 1) != 0);
  o.madd = acc;  // P +- Q
  G1Xyzz d = xyzz_identity();
  xyzz_madd(d, p, false);
  xyzz_madd(d, p, false);  // exercises the doubling branch
  o.dbl = d;
  G1Xyzz s = acc;
  xyzz_add(s, d);  // (P +- Q) + 2P
  o.add2 = s;
  out[i] = o;
}

static uint64_t lcg(uint64_t& s) {
  s = s * 6364136223846793005ull + 1442695040888963407ull;
  return s;
}

static bool xyzz_matches(const G1Xyzz& d, const tph::HG1& expect) {
  tph::HFq x, y, zz, zzz;
  memcpy(x.v, d.x.v, 48);
  memcpy(y.v, d.y.v, 48);
  memcpy(zz.v, d.zz.v, 48);
  memcpy(zzz.v, d.zzz.v, 48);
  tph::HG1 got = tph::g1_from_xyzz(x, y, zz, zzz);
  tph::HFq gx, gy, ex, ey;
  bool gi = !tph::g1_to_affine(got, &gx, &gy), ei = !tph::g1_to_affine(expect, &ex, &ey);
  if (gi || ei) return gi == ei;
  return gx == ex && gy == ey;
}

tph::HG1 host_generator();

int selftest_dev(tp_ctx* ctx, int* failures) {
  const int n = 64;
  std::vector<SelfIn> in(n);
  std::vector<uint8_t> pts(n * 96);
  std::vector<tph::HG1> hp(n);
  uint64_t seed = 0x1234567;
  tph::HG1 g = host_generator();
  for (int i = 0; i < n; i++) {
    uint64_t a[4], b[4], c[6], d[6];
    for (int k = 0; k < 4; k++) {
      a[k] = lcg(seed);
      b[k] = lcg(seed);
    }
    for (int k = 0; k < 6; k++) {
      c[k] = lcg(seed);
      d[k] = lcg(seed);
    }
    a[3] &= 0x3fffffffffffffffull;
    b[3] &= 0x3fffffffffffffffull;
    c[5] &= 0x0fffffffffffffffull;
    d[5] &= 0x0fffffffffffffffull;
    if (i == 0) memset(a, 0, sizeof(a));
    if (i == 1) {
      memcpy(a, tph::FR_PARAMS.mod, 32);
      a[0] -= 1;  // r - 1
      memcpy(c, tph::FQ_PARAMS.mod, 48);
      c[0] -= 1;
    }
    memcpy(in[i].fa.v, a, 32);
    memcpy(in[i].fb.v, b, 32);
    memcpy(in[i].qa.v, c, 48);
    memcpy(in[i].qb.v, d, 48);
    uint64_t k[1] = {lcg(seed) | 1};
    hp[i] = tph::g1_mul_u64limbs(g, k, 1);
    tph::HFq x, y;
    tph::g1_to_affine(hp[i], &x, &y);
    memcpy(&pts[i * 96], x.v, 48);
    memcpy(&pts[i * 96 + 48], y.v, 48);
  }
  SelfIn* din;
  SelfOut* dout;
  G1Affine* dpts;
  TP_CUDA_OK(ctx, cudaMalloc(&din, n * sizeof(SelfIn)));
  TP_CUDA_OK(ctx, cudaMalloc(&dout, n * sizeof(SelfOut)));
  TP_CUDA_OK(ctx, cudaMalloc(&dpts, n * 96));
  TP_CUDA_OK(ctx, cudaMemcpyAsync(din, in.data(), n * sizeof(SelfIn), cudaMemcpyHostToDevice, ctx->stream));
  TP_CUDA_OK(ctx, cudaMemcpyAsync(dpts, pts.data(), n * 96, cudaMemcpyHostToDevice, ctx->stream));
  k_selftest<<<1, n, 0, ctx->stream>>>(din, dout, dpts, n);
  TP_LAUNCH(ctx, "k_selftest");
  std::vector<SelfOut> out(n);
  TP_CUDA_OK(ctx, cudaMemcpyAsync(out.data(), dout, n * sizeof(SelfOut), cudaMemcpyDeviceToHost, ctx->stream));
  TP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFree(din);
  cudaFree(dout);
  cudaFree(dpts);
  int bad = 0;
  for (int i = 0; i < n; i++) {
    tph::HFr a = to_host(in[i].fa), b = to_host(in[i].fb);
    tph::HFq c, d;
    memcpy(c.v, in[i].qa.v, 48);
    memcpy(d.v, in[i].qb.v, 48);
    auto eqr = [&](const Fr& x, const tph::HFr& y) { return memcmp(x.v, y.v, 32) == 0; };
    auto eqq = [&](const Fq& x, const tph::HFq& y) { return memcmp(x.v, y.v, 48) == 0; };
    bad += !eqr(out[i].mul, a * b);
    bad += !eqr(out[i].add, a + b);
    bad += !eqr(out[i].sub, a - b);
    bad += !eqr(out[i].inv, a.inv());
    bad += !eqr(out[i].frommont, a.from_mont());
    bad += !eqq(out[i].qmul, c * d);
    bad += !eqq(out[i].qadd, c + d);
    bad += !eqq(out[i].qsub, c - d);
    bad += !eqq(out[i].qinv, c.inv());
    bad += !eqq(out[i].qsqr, c * c);
    bad += !eqq(out[i].qlazy, c * d - d * d);
    bad += !eqq(out[i].qkar, c * d);
    bad += !eqq(out[i].qsep, c * d);
    tph::HG1 p = hp[i], q = hp[(i + 1) % n];
    if (i & 1) q.y = q.y.neg();
    tph::HG1 pq = tph::g1_add(p, q), p2 = tph::g1_dbl(p);
    bad += !xyzz_matches(out[i].madd, pq);
    bad += !xyzz_matches(out[i].dbl, p2);
    bad += !xyzz_matches(out[i].add2, tph::g1_add(pq, p2));
  }
  *failures = bad;
  return TP_OK;
}

// ---- integer pipe microbenchmark --------------------------------------------------------
__global__ void k_imad_bench(unsigned* out, unsigned a, unsigned b, int iters) {
  unsigned x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      asm volatile(
          "mad.lo.u32 %0, %0, %8, %9;\n\tmad.lo.u32 %1, %1, %8, %9;\n\tmad.lo.u32 %2, %2, %8, %9;\n\t"
          "mad.lo.u32 %3, %3, %8, %9;\n\tmad.lo.u32 %4, %4, %8, %9;\n\tmad.lo.u32 %5, %5, %8, %9;\n\t"
          "mad.lo.u32 %6, %6, %8, %9;\n\tmad.lo.u32 %7, %7, %8, %9;"
          : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7)
          : "r"(a), "r"(b));
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
}
// Wide multiply-add in the form the Montgomery multipliers use: a carry chain of
// mad.lo.cc / madc.hi.cc pairs (ptxas fuses each pair into one IMAD.WIDE.U32.X).  The multiplier
// of every row comes out of the previous row, so nothing is loop-invariant (an earlier version
// with constant operands was folded into 64-bit additions and over-reported by 2x).
__global__ void k_imad_wide_bench(unsigned long long* out, unsigned a, unsigned b, int iters) {
  unsigned r0 = threadIdx.x, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5, r6 = r0 + 6, r7 = r0 + 7;
  unsigned m = a + threadIdx.x;
  const unsigned b0 = b, b1 = b + 2, b2 = b + 4, b3 = b + 6;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      asm volatile(
          "mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\t"
          "madc.lo.cc.u32 %2, %8, %10, %2;\n\tmadc.hi.cc.u32 %3, %8, %10, %3;\n\t"
          "madc.lo.cc.u32 %4, %8, %11, %4;\n\tmadc.hi.cc.u32 %5, %8, %11, %5;\n\t"
          "madc.lo.cc.u32 %6, %8, %12, %6;\n\tmadc.hi.u32 %7, %8, %12, %7;"
          : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7)
          : "r"(m), "r"(b0), "r"(b1), "r"(b2), "r"(b3));
      m = r7;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((unsigned long long)(r0 ^ r2 ^ r4 ^ r6) << 32) | (r1 ^ r3 ^ r5 ^ r7);
}

int measure_imad_dev(tp_ctx* ctx, double* imad, double* wide) {
  const int threads = 1024, iters = 2000;
  int blocks = ctx->sm_count * 2;
  void* buf;
  TP_CUDA_OK(ctx, cudaMalloc(&buf, (size_t)blocks * threads * 8));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float ms = 0;
  double best[2] = {0, 0};
  for (int which = 0; which < 2; which++) {
    for (int rep = 0; rep < 4; rep++) {
      cudaEventRecord(e0, ctx->stream);
      if (which == 0)
        k_imad_bench<<<blocks, threads, 0, ctx->stream>>>((unsigned*)buf, 12345u, 678u, iters);
      else
        k_imad_wide_bench<<<blocks, threads, 0, ctx->stream>>>((unsigned long long*)buf, 12345u, 678u, iters);
      ctx->launches++;
      cudaEventRecord(e1, ctx->stream);
      TP_CUDA_OK(ctx, cudaEventSynchronize(e1));
      cudaEventElapsedTime(&ms, e0, e1);
      double ops = (double)blocks * threads * iters * 64.0;
      double rate = ops / (ms * 1e-3);
      if (rep > 0 && rate > best[which]) best[which] = rate;
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  *imad = best[0];
  *wide = best[1];
  return TP_OK;
}

}  // namespace tp
