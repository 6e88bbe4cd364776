// Fq inversion by constant-time "safegcd" division steps (Bernstein-Yang 2019, half-delta
// variant), used for the one inversion per thread of the batch-affine MSM rounds (msm.cu).
// Replaces what ark-ff 0.3.0 does inside `GroupAffine` additions / `into_affine`
// (reached from kzg/src/lib.rs:46-53); the value computed is the unique inverse, so results are
// bit-identical to any other inversion.
//
// Why not Fermat: a^(q-2) costs ~570 Fq multiplications = 1.6e5 wide multiplies on the FMA-heavy
// pipe (the MSM's binding resource, profiles/r1_summary.md F); 30 rounds of 30 division steps cost
// ~3.9e3 wide multiplies plus ~3e4 ALU-pipe instructions, which the multiplier-bound kernels have
// issue slots to spare for.  All lanes of a warp run the same instruction sequence (no
// data-dependent branches).
//
// Representation: 13 signed limbs of 30 bits (value = sum v[i] 2^(30 i)).  Plain C++ so that the
// same code is unit-tested on the host (tests/test_host_logic.py builds tools/host_tests/fq_inv_test.cpp).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define TP_HD __host__ __device__ __forceinline__
#else
#define TP_HD inline
#endif

namespace tp {

#ifndef FQINV_UNROLL
#define FQINV_UNROLL 30   // division steps unrolled per loop trip (instruction footprint vs loop overhead)
#endif
#define FQINV_L 13
#define FQINV_ROUNDS 30   // 900 division steps >= the 879 needed for a 381-bit modulus
constexpr int kFqinvUnroll = FQINV_UNROLL;
struct S30 {
  int32_t v[FQINV_L];
};

// q = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
// in 30-bit limbs, q^-1 mod 2^30 and R^3 mod q (checked against Python big integers by the host unit test).
#define FQINV_MOD {0x3fffaaab, 0x27fbffff, 0x153ffffb, 0x2affffac, 0x30f6241e, 0x034a83da, 0x112bf673, 0x12e13ce1, 0x2cd76477, 0x1ed90d2e, 0x29a4b1ba, 0x3a8e5ff9, 0x001a0111}
#define FQINV_MODINV30 0x00030003u   // q^-1 mod 2^30
// R^3 mod q (Montgomery, 32-bit limbs): montmul(x^-1, R^3) = x^-1 R^2 = Montgomery form of the inverse of x / R
#define FQINV_R3 {0xd94ca1e0u, 0xed48ac6bu, 0x03a7adf8u, 0x315f831eu, 0x615e29ddu, 0x9a53352au, 0x921e1761u, 0x34c04e5eu, 0x65724728u, 0x2512d435u, 0x91755d4du, 0x0aa63460u}

// 30 division steps on the low words; returns the transition matrix t = (u, v; q, r) scaled by 2^30.
TP_HD int32_t fqinv_divsteps_30(int32_t zeta, uint32_t f0, uint32_t g0, int32_t t[4]) {
  uint32_t u = 1, v = 0, q = 0, r = 1;
  uint32_t f = f0, g = g0;
#pragma unroll kFqinvUnroll
  for (int i = 0; i < 30; i++) {
    uint32_t c1 = (uint32_t)(zeta >> 31);      // all ones if zeta < 0
    uint32_t c2 = (uint32_t)0 - (g & 1);       // all ones if g is odd
    uint32_t x = (f ^ c1) - c1, y = (u ^ c1) - c1, z = (v ^ c1) - c1;  // conditionally negate f, u, v
    g += x & c2;
    q += y & c2;
    r += z & c2;
    c1 &= c2;                                   // swap only if both hold
    zeta = (int32_t)((uint32_t)zeta ^ c1) - 1;
    f += g & c1;
    u += q & c1;
    v += r & c1;
    g >>= 1;
    u <<= 1;
    v <<= 1;
  }
  t[0] = (int32_t)u;
  t[1] = (int32_t)v;
  t[2] = (int32_t)q;
  t[3] = (int32_t)r;
  return zeta;
}

// (f, g) <- t * (f, g) / 2^30 (exact)
TP_HD void fqinv_update_fg(S30& f, S30& g, const int32_t t[4]) {
  const int32_t M30 = 0x3fffffff;
  const int64_t u = t[0], v = t[1], q = t[2], r = t[3];
  int64_t cf = u * f.v[0] + v * g.v[0];
  int64_t cg = q * f.v[0] + r * g.v[0];
  cf >>= 30;
  cg >>= 30;
#pragma unroll
  for (int i = 1; i < FQINV_L; i++) {
    int32_t fi = f.v[i], gi = g.v[i];
    cf += u * fi + v * gi;
    cg += q * fi + r * gi;
    f.v[i - 1] = (int32_t)cf & M30;
    g.v[i - 1] = (int32_t)cg & M30;
    cf >>= 30;
    cg >>= 30;
  }
  f.v[FQINV_L - 1] = (int32_t)cf;
  g.v[FQINV_L - 1] = (int32_t)cg;
}

// (d, e) <- t * (d, e) / 2^30 mod q, kept in (-2q, q)
TP_HD void fqinv_update_de(S30& d, S30& e, const int32_t t[4]) {
  const int32_t M30 = 0x3fffffff;
  const int32_t mod[FQINV_L] = FQINV_MOD;
  const int64_t u = t[0], v = t[1], q = t[2], r = t[3];
  const int32_t sd = d.v[FQINV_L - 1] >> 31, se = e.v[FQINV_L - 1] >> 31;
  int32_t md = (t[0] & sd) + (t[1] & se);
  int32_t me = (t[2] & sd) + (t[3] & se);
  int64_t cd = u * d.v[0] + v * e.v[0];
  int64_t ce = q * d.v[0] + r * e.v[0];
  md -= (int32_t)((FQINV_MODINV30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
  me -= (int32_t)((FQINV_MODINV30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
  cd += (int64_t)mod[0] * md;
  ce += (int64_t)mod[0] * me;
  cd >>= 30;
  ce >>= 30;
#pragma unroll
  for (int i = 1; i < FQINV_L; i++) {
    int32_t di = d.v[i], ei = e.v[i];
    cd += u * di + v * ei + (int64_t)mod[i] * md;
    ce += q * di + r * ei + (int64_t)mod[i] * me;
    d.v[i - 1] = (int32_t)cd & M30;
    e.v[i - 1] = (int32_t)ce & M30;
    cd >>= 30;
    ce >>= 30;
  }
  d.v[FQINV_L - 1] = (int32_t)cd;
  e.v[FQINV_L - 1] = (int32_t)ce;
}

// r in (-2q, q) -> [0, q), negated first if sign < 0
TP_HD void fqinv_normalize(S30& r, int32_t sign) {
  const int32_t M30 = 0x3fffffff;
  const int32_t mod[FQINV_L] = FQINV_MOD;
  int32_t cond_add = r.v[FQINV_L - 1] >> 31;
  const int32_t cond_neg = sign >> 31;
#pragma unroll
  for (int i = 0; i < FQINV_L; i++) {
    int32_t x = r.v[i] + (mod[i] & cond_add);
    r.v[i] = (x ^ cond_neg) - cond_neg;
  }
#pragma unroll
  for (int i = 0; i < FQINV_L - 1; i++) {
    r.v[i + 1] += r.v[i] >> 30;
    r.v[i] &= M30;
  }
  cond_add = r.v[FQINV_L - 1] >> 31;
#pragma unroll
  for (int i = 0; i < FQINV_L; i++) r.v[i] += mod[i] & cond_add;
#pragma unroll
  for (int i = 0; i < FQINV_L - 1; i++) {
    r.v[i + 1] += r.v[i] >> 30;
    r.v[i] &= M30;
  }
}

// x (12 x u32, little-endian, 0 <= x < q) -> x^-1 mod q as 12 x u32 (0 for 0).  Returns false if
// the division steps did not terminate (g != 0) -- impossible within the proven bound; callers
// fall back to Fermat rather than trust it blindly.
TP_HD bool fqinv_plain(const uint32_t x[12], uint32_t out[12]) {
  const int32_t mod[FQINV_L] = FQINV_MOD;
  S30 d, e, f, g;
#pragma unroll
  for (int i = 0; i < FQINV_L; i++) {
    d.v[i] = 0;
    e.v[i] = 0;
    f.v[i] = mod[i];
    const int bit = 30 * i;
    const int w = bit >> 5, s = bit & 31;
    uint64_t two = w < 12 ? x[w] : 0;
    if (w + 1 < 12) two |= (uint64_t)x[w + 1] << 32;
    g.v[i] = (int32_t)((uint32_t)(two >> s) & 0x3fffffffu);
  }
  e.v[0] = 1;
  int32_t zeta = -1;
#pragma unroll 1
  for (int it = 0; it < FQINV_ROUNDS; it++) {
    int32_t t[4];
    zeta = fqinv_divsteps_30(zeta, (uint32_t)f.v[0], (uint32_t)g.v[0], t);
    fqinv_update_de(d, e, t);
    fqinv_update_fg(f, g, t);
  }
  int32_t gz = 0;
#pragma unroll
  for (int i = 0; i < FQINV_L; i++) gz |= g.v[i];
  fqinv_normalize(d, f.v[FQINV_L - 1]);
  // 13 x 30 bits -> 12 x 32 bits
#pragma unroll
  for (int k = 0; k < 12; k++) {
    const int bit = 32 * k;
    const int i = bit / 30, s = bit % 30;
    uint64_t acc = (uint64_t)(uint32_t)d.v[i] >> s;
    if (i + 1 < FQINV_L) acc |= (uint64_t)(uint32_t)d.v[i + 1] << (30 - s);
    if (i + 2 < FQINV_L && 60 - s < 32) acc |= (uint64_t)(uint32_t)d.v[i + 2] << (60 - s);
    out[k] = (uint32_t)acc;
  }
  return gz == 0;
}

}  // namespace tp
