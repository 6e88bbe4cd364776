// BLS12-381 G1 device arithmetic in XYZZ coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2;
// identity <=> ZZ == 0), a = 0, b = 4.  Replaces ark-ec 0.3.0 `GroupAffine::mul` /
// `GroupProjective` additions behind kzg/src/lib.rs:46-53 and kzg/src/srs.rs:15-24.
// Mixed addition = 8M + 2S, general addition = 12M + 2S, doubling = 6M + 3S (Fq).
#pragma once
#include "field.cuh"
#include "fq_inv.cuh"

namespace tp {

// Packed affine point as stored in the device SRS: (x, y) Montgomery, 96 bytes.
// (0, 0) encodes the point at infinity (it is not on y^2 = x^3 + 4).
struct alignas(16) G1Affine {
  Fq x, y;
};
struct alignas(16) G1Xyzz {
  Fq x, y, zz, zzz;
};

__device__ __forceinline__ G1Xyzz xyzz_identity() {
  G1Xyzz r;
  r.x = fq_zero(); r.y = fq_zero(); r.zz = fq_zero(); r.zzz = fq_zero();
  return r;
}
__device__ __forceinline__ bool xyzz_is_identity(const G1Xyzz& p) { return fq_is_zero(p.zz); }
__device__ __forceinline__ bool affine_is_identity(const G1Affine& p) { return fq_is_zero(p.x) && fq_is_zero(p.y); }

__device__ __forceinline__ G1Affine affine_load(const G1Affine* p) {
  G1Affine r;
  r.x = fq_load(&p->x);
  r.y = fq_load(&p->y);
  return r;
}
__device__ __forceinline__ void xyzz_store(G1Xyzz* p, const G1Xyzz& a) {
  fq_store(&p->x, a.x); fq_store(&p->y, a.y); fq_store(&p->zz, a.zz); fq_store(&p->zzz, a.zzz);
}
__device__ __forceinline__ G1Xyzz xyzz_load(const G1Xyzz* p) {
  G1Xyzz r;
  r.x = fq_load(&p->x); r.y = fq_load(&p->y); r.zz = fq_load(&p->zz); r.zzz = fq_load(&p->zzz);
  return r;
}

// Montgomery-form inverse by division steps (fq_inv.cuh); Fermat only if those did not terminate.
static __device__ __noinline__ Fq fq_inverse(const Fq& a) {
  Fq plain, r;
  if (!fqinv_plain(a.v, plain.v)) return fq_inv(a);
  const uint32_t r3[12] = FQINV_R3;
  Fq k;
#pragma unroll
  for (int i = 0; i < 12; i++) k.v[i] = r3[i];
  r = fq_mul(plain, k);   // (a R)^-1 * R^3 / R = a^-1 R
  return r;
}

// ---- affine pair addition with a shared (batched) inversion ------------------------------
// kind of a pair P1 + P2: 0 = chord (denominator x2 - x1), 1 = tangent (denominator 2 y1),
// 2 = result is the identity, 3 = result P1 (P2 is the identity), 4 = result P2.
// For kinds >= 2 the denominator is 1 so that it can sit in the running product.
static __device__ __noinline__ int affine_pair_special(const G1Affine& p1, const G1Affine& p2, Fq& d) {
  d = fq_one();
  if (affine_is_identity(p1)) return 4;
  if (affine_is_identity(p2)) return 3;
  if (fq_eq(p1.x, p2.x)) {
    if (fq_eq(p1.y, p2.y) && !fq_is_zero(p1.y)) {
      d = fq_dbl(p1.y);
      return 1;
    }
    return 2;
  }
  d = fq_sub(p2.x, p1.x);  // x = 0 on a non-identity point: an ordinary chord after all
  return 0;
}
// P1 + P2 given dinv = 1 / (denominator of `kind`); all cases of affine_pair_special.
__device__ __forceinline__ G1Affine affine_pair_finish(const G1Affine& p1, const G1Affine& p2, int kind, const Fq& dinv) {
  if (kind >= 2) {
    G1Affine r;
    if (kind == 3) return p1;
    if (kind == 4) return p2;
    r.x = fq_zero();
    r.y = fq_zero();
    return r;
  }
  Fq num;
  if (kind == 0) {
    num = fq_sub(p2.y, p1.y);
  } else {
    Fq xx = fq_sqr(p1.x);
    num = fq_add(fq_dbl(xx), xx);
  }
  Fq lam = fq_mul(num, dinv);
  G1Affine r;
  r.x = fq_sub(fq_sub(fq_sqr(lam), p1.x), p2.x);
  r.y = fq_sub(fq_mul(lam, fq_sub(p1.x, r.x)), p1.y);
  return r;
}

// 2 * (affine P), P != identity.
static __device__ __noinline__ void xyzz_mdbl(G1Xyzz& r, const G1Affine p) {
  Fq u = fq_dbl(p.y);
  Fq v = fq_sqr(u);
  Fq w = fq_mul(u, v);
  Fq s = fq_mul(p.x, v);
  Fq xx = fq_sqr(p.x);
  Fq m = fq_add(fq_dbl(xx), xx);
  Fq x3 = fq_sub(fq_sqr(m), fq_dbl(s));
  r.y = fq_mul_sub2(m, fq_sub(s, x3), w, p.y);
  r.x = x3;
  r.zz = v;
  r.zzz = w;
}

static __device__ __noinline__ void xyzz_dbl(G1Xyzz& r) {
  if (xyzz_is_identity(r)) return;
  Fq u = fq_dbl(r.y);
  Fq v = fq_sqr(u);
  Fq w = fq_mul(u, v);
  Fq s = fq_mul(r.x, v);
  Fq xx = fq_sqr(r.x);
  Fq m = fq_add(fq_dbl(xx), xx);
  Fq x3 = fq_sub(fq_sqr(m), fq_dbl(s));
  Fq y3 = fq_mul_sub2(m, fq_sub(s, x3), w, r.y);
  r.x = x3;
  r.y = y3;
  r.zz = fq_mul(v, r.zz);
  r.zzz = fq_mul(w, r.zzz);
}

// acc += P (affine, not identity); `neg` adds -P.  Handles acc == identity, acc == P, acc == -P.
__device__ __forceinline__ void xyzz_madd(G1Xyzz& acc, const G1Affine& p_in, bool neg) {
  G1Affine p = p_in;
  if (neg) p.y = fq_neg(p.y);
  if (xyzz_is_identity(acc)) {
    acc.x = p.x; acc.y = p.y; acc.zz = fq_one(); acc.zzz = fq_one();
    return;
  }
  Fq u2 = fq_mul(p.x, acc.zz);
  Fq s2 = fq_mul(p.y, acc.zzz);
  Fq pp_ = fq_sub(u2, acc.x);
  Fq rr = fq_sub(s2, acc.y);
  if (fq_is_zero(pp_)) {
    if (fq_is_zero(rr)) {
      G1Xyzz d;  // separate object so `acc` itself never has its address taken
      xyzz_mdbl(d, p);
      acc = d;
    } else {
      acc = xyzz_identity();
    }
    return;
  }
  Fq pp = fq_sqr(pp_);
  Fq ppp = fq_mul(pp_, pp);
  Fq q = fq_mul(acc.x, pp);
  Fq x3 = fq_sub(fq_sub(fq_sqr(rr), ppp), fq_dbl(q));
  Fq y3 = fq_mul_sub2(rr, fq_sub(q, x3), acc.y, ppp);
  acc.x = x3;
  acc.y = y3;
  acc.zz = fq_mul(acc.zz, pp);
  acc.zzz = fq_mul(acc.zzz, ppp);
}

// acc += b (both XYZZ), all special cases handled.
static __device__ __noinline__ void xyzz_add(G1Xyzz& acc, const G1Xyzz& b) {
  if (xyzz_is_identity(b)) return;
  if (xyzz_is_identity(acc)) {
    acc = b;
    return;
  }
  Fq u1 = fq_mul(acc.x, b.zz);
  Fq u2 = fq_mul(b.x, acc.zz);
  Fq s1 = fq_mul(acc.y, b.zzz);
  Fq s2 = fq_mul(b.y, acc.zzz);
  Fq pp_ = fq_sub(u2, u1);
  Fq rr = fq_sub(s2, s1);
  if (fq_is_zero(pp_)) {
    if (fq_is_zero(rr)) {
      xyzz_dbl(acc);
    } else {
      acc = xyzz_identity();
    }
    return;
  }
  Fq pp = fq_sqr(pp_);
  Fq ppp = fq_mul(pp_, pp);
  Fq q = fq_mul(u1, pp);
  Fq x3 = fq_sub(fq_sub(fq_sqr(rr), ppp), fq_dbl(q));
  Fq y3 = fq_mul_sub2(rr, fq_sub(q, x3), s1, ppp);
  acc.x = x3;
  acc.y = y3;
  acc.zz = fq_mul(fq_mul(acc.zz, b.zz), pp);
  acc.zzz = fq_mul(fq_mul(acc.zzz, b.zzz), ppp);
}

// acc = k * acc for a small non-negative integer k (double-and-add, MSB first).
static __device__ __noinline__ void xyzz_mul_small(G1Xyzz& acc, uint32_t k) {
  if (k == 0 || xyzz_is_identity(acc)) {
    acc = xyzz_identity();
    return;
  }
  G1Xyzz base = acc;
  int top = 31 - __clz(k);
  for (int b = top - 1; b >= 0; b--) {
    xyzz_dbl(acc);
    if ((k >> b) & 1) xyzz_add(acc, base);
  }
}

}  // namespace tp
