// Reduced-radix BLS12-381 Fq for the MSM kernels: 13 limbs x 30 bits in u32 registers, Montgomery
// radix R' = 2^390, lazily reduced (values are kept below 16 q, limbs below 2^30).
//
// Why: on sm_100a `IMAD.WIDE.U32` retires 63 per clk per SM, but every carry predicate
// (mad.lo.cc / madc.hi.cc -> IMAD.WIDE.U32.X) halves that (profiles/r1_summary.md, section C).  With 30-bit limbs
// the 13 column sums of a product fit 64-bit accumulators without any carry, so the 351 limb
// products of a Montgomery multiplication are plain carry-free IMAD.WIDE and the (few) carries move to
// shifts/adds on the ALU pipe, which runs beside the FMA-heavy pipe.
//
// Plain C++ (host + device): the same code is unit-tested on the CPU (tests/test_fq30_host.py).
// Replaces ark-ff 0.3.0 Fp384 arithmetic inside kzg/src/lib.rs:46-53 (commit).
#pragma once
#include <stdint.h>

#include "fq30_consts.h"

#if defined(__CUDACC__)
#define TP_HD __host__ __device__ __forceinline__
#define TP_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define TP_HD inline
#define TP_HD_NOINLINE static inline
#endif

namespace tp {

#define FQ30_L 13
#define FQ30_MASK 0x3fffffffu

struct Fq30 {
  uint32_t l[FQ30_L];
};

// Compile-time tables: after full unrolling every use is an immediate operand.
struct Fq30Tables {
  uint32_t kq[17][FQ30_L];
};
TP_HD constexpr Fq30Tables fq30_tables() { return Fq30Tables{FQ30_KQ_TABLE}; }

// acc += a * b: compiles to one carry-free IMAD.WIDE.U32 with the accumulator as the addend.
TP_HD void fq30_madw(uint64_t& acc, uint32_t a, uint32_t b) { acc += (uint64_t)a * b; }

// m = (t0 * -q^-1) mod 2^30.  Opaque 32-bit on the device: otherwise NVVM widens m to 64 bits and the
// 13 products m * q_j become 64 x 64 multiplies (one stray IADD3 each).
TP_HD uint32_t fq30_mont_digit(uint32_t t0) {
  uint32_t m;
#if defined(__CUDA_ARCH__)
  asm("mul.lo.u32 %0, %1, %2;" : "=r"(m) : "r"(t0), "r"(FQ30_PINV));
#else
  m = t0 * FQ30_PINV;
#endif
  return m & FQ30_MASK;
}

TP_HD Fq30 fq30_zero() {
  Fq30 r;
#pragma unroll
  for (int i = 0; i < FQ30_L; i++) r.l[i] = 0;
  return r;
}
TP_HD Fq30 fq30_one() {
  constexpr uint32_t one[FQ30_L] = FQ30_ONE;
  Fq30 r;
#pragma unroll
  for (int i = 0; i < FQ30_L; i++) r.l[i] = one[i];
  return r;
}
TP_HD bool fq30_is_exact_zero(const Fq30& a) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < FQ30_L; i++) o |= a.l[i];
  return o == 0;
}

// Montgomery product a * b * 2^-390 mod q.  Requires limbs < 2^30 and a * b < 256 q^2 (any two
// values below 16 q); returns a value < 1.41 q with limbs < 2^30.
TP_HD Fq30 fq30_mul(const Fq30& a, const Fq30& b) {
  constexpr Fq30Tables T = fq30_tables();
  uint64_t t[FQ30_L];
#pragma unroll
  for (int j = 0; j < FQ30_L; j++) t[j] = 0;
#pragma unroll
  for (int i = 0; i < FQ30_L; i++) {
#pragma unroll
    for (int j = 0; j < FQ30_L; j++) fq30_madw(t[j], a.l[j], b.l[i]);
    uint32_t m = fq30_mont_digit((uint32_t)t[0]);
#pragma unroll
    for (int j = 0; j < FQ30_L; j++) fq30_madw(t[j], m, T.kq[1][j]);
    uint64_t carry = t[0] >> 30;
#pragma unroll
    for (int j = 0; j < FQ30_L - 1; j++) t[j] = t[j + 1];
    t[FQ30_L - 1] = 0;
    t[0] += carry;
    if (i == 6) {  // keep every column below 2^64: at most 14 products between normalisations
#pragma unroll
      for (int j = 0; j < FQ30_L - 1; j++) {
        t[j + 1] += t[j] >> 30;
        t[j] &= FQ30_MASK;
      }
    }
  }
  Fq30 r;
#pragma unroll
  for (int j = 0; j < FQ30_L - 1; j++) {
    t[j + 1] += t[j] >> 30;
    r.l[j] = (uint32_t)t[j] & FQ30_MASK;
  }
  r.l[FQ30_L - 1] = (uint32_t)t[FQ30_L - 1];
  return r;
}
TP_HD Fq30 fq30_sqr(const Fq30& a) { return fq30_mul(a, a); }

// a + b (no reduction; bound adds up)
TP_HD Fq30 fq30_add(const Fq30& a, const Fq30& b) {
  Fq30 r;
  uint32_t c = 0;
#pragma unroll
  for (int j = 0; j < FQ30_L - 1; j++) {
    uint32_t s = a.l[j] + b.l[j] + c;
    r.l[j] = s & FQ30_MASK;
    c = s >> 30;
  }
  r.l[FQ30_L - 1] = a.l[FQ30_L - 1] + b.l[FQ30_L - 1] + c;
  return r;
}
TP_HD Fq30 fq30_dbl(const Fq30& a) { return fq30_add(a, a); }

// a + K q - b, for b < K q  (result < a + K q)
template <int K>
TP_HD Fq30 fq30_sub(const Fq30& a, const Fq30& b) {
  constexpr Fq30Tables T = fq30_tables();
  Fq30 r;
  int32_t c = 0;
#pragma unroll
  for (int j = 0; j < FQ30_L - 1; j++) {
    int32_t s = (int32_t)(a.l[j] + T.kq[K][j]) - (int32_t)b.l[j] + c;
    r.l[j] = (uint32_t)s & FQ30_MASK;
    c = s >> 30;  // arithmetic shift: -1, 0 or 1
  }
  r.l[FQ30_L - 1] = (uint32_t)((int32_t)(a.l[FQ30_L - 1] + T.kq[K][FQ30_L - 1]) - (int32_t)b.l[FQ30_L - 1] + c);
  return r;
}
// K q - b
template <int K>
TP_HD Fq30 fq30_neg(const Fq30& b) {
  return fq30_sub<K>(fq30_zero(), b);
}

// a == k q for some 0 <= k <= 16 (i.e. a = 0 mod q for any lazily reduced a < 17 q).  Cheap low-limb
// filter first; the exact comparison runs with probability ~2^-26.
TP_HD_NOINLINE bool fq30_is_zero_mod_q_slow(const Fq30& a) {
  constexpr Fq30Tables T = fq30_tables();
  for (int k = 0; k <= 16; k++) {
    bool eq = true;
    for (int j = 0; j < FQ30_L; j++) eq = eq && (a.l[j] == T.kq[k][j]);
    if (eq) return true;
  }
  return false;
}
TP_HD bool fq30_is_zero_mod_q(const Fq30& a) {
  constexpr Fq30Tables T = fq30_tables();
  bool hit = false;
#pragma unroll
  for (int k = 0; k <= 16; k++) hit = hit || (a.l[0] == T.kq[k][0]);
  if (!hit) return false;
  return fq30_is_zero_mod_q_slow(a);
}

// ---- G1 in XYZZ coordinates over Fq30 -------------------------------------------------------------
// Bounds maintained: x < 8q, y < 4q, zz < 2q, zzz < 2q (products < 1.41 q).  The identity is stored
// as exact zeros in zz (a computed zz is never 0 mod q).
struct G1Aff30 {
  Fq30 x, y;  // fully reduced (< q); (0, 0) = infinity
};
struct G1Xyzz30 {
  Fq30 x, y, zz, zzz;
};
TP_HD G1Xyzz30 xyzz30_identity() {
  G1Xyzz30 r;
  r.x = fq30_zero(); r.y = fq30_zero(); r.zz = fq30_zero(); r.zzz = fq30_zero();
  return r;
}
TP_HD bool xyzz30_is_identity(const G1Xyzz30& p) { return fq30_is_exact_zero(p.zz); }
TP_HD bool aff30_is_identity(const G1Aff30& p) { return fq30_is_exact_zero(p.x) && fq30_is_exact_zero(p.y); }

// 2 * P for affine P (x, y < q)
TP_HD_NOINLINE G1Xyzz30 xyzz30_mdbl(const G1Aff30 p) {
  Fq30 u = fq30_dbl(p.y);                         // < 2q
  Fq30 v = fq30_sqr(u);
  Fq30 w = fq30_mul(u, v);
  Fq30 s = fq30_mul(p.x, v);
  Fq30 xx = fq30_sqr(p.x);
  Fq30 m = fq30_add(fq30_dbl(xx), xx);            // < 6q
  G1Xyzz30 r;
  r.x = fq30_sub<4>(fq30_sqr(m), fq30_dbl(s));    // < 6q
  r.y = fq30_sub<2>(fq30_mul(m, fq30_sub<8>(s, r.x)), fq30_mul(w, p.y));  // < 4q
  r.zz = v;
  r.zzz = w;
  return r;
}
TP_HD_NOINLINE void xyzz30_dbl(G1Xyzz30& r) {
  if (xyzz30_is_identity(r)) return;
  Fq30 u = fq30_dbl(r.y);                         // < 8q
  Fq30 v = fq30_sqr(u);
  Fq30 w = fq30_mul(u, v);
  Fq30 s = fq30_mul(r.x, v);
  Fq30 xx = fq30_sqr(r.x);
  Fq30 m = fq30_add(fq30_dbl(xx), xx);            // < 6q
  Fq30 x3 = fq30_sub<4>(fq30_sqr(m), fq30_dbl(s));                       // < 6q
  Fq30 y3 = fq30_sub<2>(fq30_mul(m, fq30_sub<8>(s, x3)), fq30_mul(w, r.y));  // < 4q
  r.x = x3;
  r.y = y3;
  r.zz = fq30_mul(v, r.zz);
  r.zzz = fq30_mul(w, r.zzz);
}

// acc += (+-) P, P affine and not the identity.
TP_HD void xyzz30_madd(G1Xyzz30& acc, const G1Aff30& p_in, bool neg) {
  G1Aff30 p = p_in;
  if (neg) p.y = fq30_neg<1>(p.y);                // q - y  (< q... exactly q when y = 0, which no G1 point has)
  if (xyzz30_is_identity(acc)) {
    acc.x = p.x; acc.y = p.y; acc.zz = fq30_one(); acc.zzz = fq30_one();
    return;
  }
  Fq30 u2 = fq30_mul(p.x, acc.zz);
  Fq30 s2 = fq30_mul(p.y, acc.zzz);
  Fq30 pp_ = fq30_sub<8>(u2, acc.x);              // < 10q
  Fq30 rr = fq30_sub<4>(s2, acc.y);               // < 6q
  if (fq30_is_zero_mod_q(pp_)) {
    if (fq30_is_zero_mod_q(rr)) {
      acc = xyzz30_mdbl(p);
    } else {
      acc = xyzz30_identity();
    }
    return;
  }
  Fq30 pp = fq30_sqr(pp_);
  Fq30 ppp = fq30_mul(pp_, pp);
  Fq30 q = fq30_mul(acc.x, pp);
  Fq30 x3 = fq30_sub<2>(fq30_sub<2>(fq30_sub<2>(fq30_sqr(rr), ppp), q), q);   // < 8q
  Fq30 y3 = fq30_sub<2>(fq30_mul(rr, fq30_sub<8>(q, x3)), fq30_mul(acc.y, ppp));  // < 4q
  acc.x = x3;
  acc.y = y3;
  acc.zz = fq30_mul(acc.zz, pp);
  acc.zzz = fq30_mul(acc.zzz, ppp);
}

// acc += b
TP_HD_NOINLINE void xyzz30_add(G1Xyzz30& acc, const G1Xyzz30& b) {
  if (xyzz30_is_identity(b)) return;
  if (xyzz30_is_identity(acc)) {
    acc = b;
    return;
  }
  Fq30 u1 = fq30_mul(acc.x, b.zz);
  Fq30 u2 = fq30_mul(b.x, acc.zz);
  Fq30 s1 = fq30_mul(acc.y, b.zzz);
  Fq30 s2 = fq30_mul(b.y, acc.zzz);
  Fq30 pp_ = fq30_sub<2>(u2, u1);                 // < 4q
  Fq30 rr = fq30_sub<2>(s2, s1);                  // < 4q
  if (fq30_is_zero_mod_q(pp_)) {
    if (fq30_is_zero_mod_q(rr)) {
      xyzz30_dbl(acc);
    } else {
      acc = xyzz30_identity();
    }
    return;
  }
  Fq30 pp = fq30_sqr(pp_);
  Fq30 ppp = fq30_mul(pp_, pp);
  Fq30 q = fq30_mul(u1, pp);
  Fq30 x3 = fq30_sub<2>(fq30_sub<2>(fq30_sub<2>(fq30_sqr(rr), ppp), q), q);   // < 8q
  Fq30 y3 = fq30_sub<2>(fq30_mul(rr, fq30_sub<8>(q, x3)), fq30_mul(s1, ppp));  // < 4q
  acc.x = x3;
  acc.y = y3;
  acc.zz = fq30_mul(fq30_mul(acc.zz, b.zz), pp);
  acc.zzz = fq30_mul(fq30_mul(acc.zzz, b.zzz), ppp);
}

// acc = k * acc, k >= 0 small
TP_HD_NOINLINE void xyzz30_mul_small(G1Xyzz30& acc, uint32_t k) {
  if (k == 0 || xyzz30_is_identity(acc)) {
    acc = xyzz30_identity();
    return;
  }
  G1Xyzz30 base = acc;
  int top = 31;
  while (!((k >> top) & 1)) top--;
  for (int b = top - 1; b >= 0; b--) {
    xyzz30_dbl(acc);
    if ((k >> b) & 1) xyzz30_add(acc, base);
  }
}

}  // namespace tp
