// Host-side BLS12-381 Fr / Fq Montgomery arithmetic (64-bit limbs, same memory layout as
// the device types and as ark_ff::Fp256 / Fp384) and a small Jacobian G1.  Used by the
// prover driver for the O(1) scalar work between kernels (challenge algebra for
// plonk/src/proof.rs:376-439) and for the serial tail of each MSM (window Horner + affine
// conversion), where one CPU thread beats one GPU thread by an order of magnitude.
// Product code -- independent of oracle/.
#pragma once
#include <stdint.h>
#include <string.h>

namespace tph {

typedef unsigned __int128 u128;

template <int N>
struct Mont {
  uint64_t v[N];
};

template <int N>
struct FieldParams {
  uint64_t mod[N];
  uint64_t one[N];  // R mod p
  uint64_t r2[N];   // R^2 mod p
  uint64_t inv;     // -p^-1 mod 2^64
};

template <int N>
static inline bool ge(const uint64_t* a, const uint64_t* b) {
  for (int i = N - 1; i >= 0; i--) {
    if (a[i] != b[i]) return a[i] > b[i];
  }
  return true;
}
template <int N>
static inline uint64_t sub_n(uint64_t* r, const uint64_t* a, const uint64_t* b) {
  uint64_t borrow = 0;
  for (int i = 0; i < N; i++) {
    u128 t = (u128)a[i] - b[i] - borrow;
    r[i] = (uint64_t)t;
    borrow = (uint64_t)(t >> 64) & 1;
  }
  return borrow;
}
template <int N>
static inline uint64_t add_n(uint64_t* r, const uint64_t* a, const uint64_t* b) {
  uint64_t carry = 0;
  for (int i = 0; i < N; i++) {
    u128 t = (u128)a[i] + b[i] + carry;
    r[i] = (uint64_t)t;
    carry = (uint64_t)(t >> 64);
  }
  return carry;
}

template <int N, const FieldParams<N>& P>
struct Fp {
  uint64_t v[N];

  static Fp zero() {
    Fp r;
    memset(r.v, 0, sizeof(r.v));
    return r;
  }
  static Fp one() {
    Fp r;
    memcpy(r.v, P.one, sizeof(r.v));
    return r;
  }
  bool is_zero() const {
    uint64_t o = 0;
    for (int i = 0; i < N; i++) o |= v[i];
    return o == 0;
  }
  bool operator==(const Fp& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
  bool operator!=(const Fp& o) const { return !(*this == o); }

  Fp operator+(const Fp& o) const {
    Fp r;
    uint64_t c = add_n<N>(r.v, v, o.v);
    if (c || ge<N>(r.v, P.mod)) sub_n<N>(r.v, r.v, P.mod);
    return r;
  }
  Fp operator-(const Fp& o) const {
    Fp r;
    if (sub_n<N>(r.v, v, o.v)) add_n<N>(r.v, r.v, P.mod);
    return r;
  }
  Fp neg() const { return zero() - *this; }
  Fp dbl() const { return *this + *this; }

  // CIOS Montgomery product.  Both moduli leave the top bit of their highest limb clear, so the running value
  // never needs a second overflow word (the "no-carry" form); the loops are unrolled because a single thread's
  // latency is what the MSM tails and the verifier wait for.
  Fp operator*(const Fp& o) const {
    static_assert(N <= 8, "unroll pragmas below assume at most 8 limbs");
    uint64_t t[N + 1];
#pragma GCC unroll 8
    for (int k = 0; k <= N; k++) t[k] = 0;
#pragma GCC unroll 8
    for (int i = 0; i < N; i++) {
      uint64_t c = 0;
#pragma GCC unroll 8
      for (int j = 0; j < N; j++) {
        u128 x = (u128)v[j] * o.v[i] + t[j] + c;
        t[j] = (uint64_t)x;
        c = (uint64_t)(x >> 64);
      }
      const uint64_t tn = t[N] + c;
      const uint64_t m = t[0] * P.inv;
      u128 x = (u128)m * P.mod[0] + t[0];
      c = (uint64_t)(x >> 64);
#pragma GCC unroll 8
      for (int j = 1; j < N; j++) {
        x = (u128)m * P.mod[j] + t[j] + c;
        t[j - 1] = (uint64_t)x;
        c = (uint64_t)(x >> 64);
      }
      x = (u128)tn + c;
      t[N - 1] = (uint64_t)x;
      t[N] = (uint64_t)(x >> 64);
    }
    Fp r;
#pragma GCC unroll 8
    for (int k = 0; k < N; k++) r.v[k] = t[k];
    if (t[N] || ge<N>(r.v, P.mod)) sub_n<N>(r.v, r.v, P.mod);
    return r;
  }
  Fp sqr() const { return *this * *this; }

  // exponent as little-endian u64 limbs
  Fp pow(const uint64_t* e, int n) const {
    Fp r = one();
    bool started = false;
    for (int i = n - 1; i >= 0; i--)
      for (int b = 63; b >= 0; b--) {
        if (started) r = r.sqr();
        if ((e[i] >> b) & 1) {
          r = started ? r * *this : *this;
          started = true;
        }
      }
    return r;
  }
  Fp pow_u64(uint64_t e) const { return pow(&e, 1); }
  Fp inv() const {  // Fermat; inv(0) = 0
    uint64_t e[N];
    uint64_t two[N];
    memset(two, 0, sizeof(two));
    two[0] = 2;
    sub_n<N>(e, P.mod, two);
    return pow(e, N);
  }
  static Fp from_u64(uint64_t x) {  // canonical small integer -> Montgomery
    Fp a = zero();
    a.v[0] = x;
    Fp r2;
    memcpy(r2.v, P.r2, sizeof(r2.v));
    return a * r2;
  }
  Fp from_mont() const {  // Montgomery -> canonical integer limbs
    Fp o = zero();
    o.v[0] = 1;
    return *this * o;
  }
  static Fp to_mont(const uint64_t* canonical) {
    Fp a;
    memcpy(a.v, canonical, sizeof(a.v));
    Fp r2;
    memcpy(r2.v, P.r2, sizeof(r2.v));
    return a * r2;
  }
};

// BLS12-381 parameters (ark-bls12-381 0.3.0 FrParameters / FqParameters).
inline constexpr FieldParams<4> FR_PARAMS = {
    {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull},
    {0x00000001fffffffeull, 0x5884b7fa00034802ull, 0x998c4fefecbc4ff5ull, 0x1824b159acc5056full},
    {0xc999e990f3f29c6dull, 0x2b6cedcb87925c23ull, 0x05d314967254398full, 0x0748d9d99f59ff11ull},
    0xfffffffeffffffffull};
inline constexpr FieldParams<6> FQ_PARAMS = {
    {0xb9feffffffffaaabull, 0x1eabfffeb153ffffull, 0x6730d2a0f6b0f624ull, 0x64774b84f38512bfull, 0x4b1ba7b6434bacd7ull,
     0x1a0111ea397fe69aull},
    {0x760900000002fffdull, 0xebf4000bc40c0002ull, 0x5f48985753c758baull, 0x77ce585370525745ull, 0x5c071a97a256ec6dull,
     0x15f65ec3fa80e493ull},
    {0xf4df1f341c341746ull, 0x0a76e6a609d104f1ull, 0x8de5476c4c95b6d5ull, 0x67eb88a9939d83c0ull, 0x9a793e85b519952dull,
     0x11988fe592cae3aaull},
    0x89f3fffcfffcfffdull};

typedef Fp<4, FR_PARAMS> HFr;
typedef Fp<6, FQ_PARAMS> HFq;

// ---- Jacobian G1 on the host (y^2 = x^3 + 4) ------------------------------------------
struct HG1 {
  HFq x, y, z;  // z == 0 -> identity
  static HG1 identity() { return {HFq::one(), HFq::one(), HFq::zero()}; }
  bool is_identity() const { return z.is_zero(); }
};

static inline HG1 g1_dbl(const HG1& p) {
  if (p.is_identity() || p.y.is_zero()) return HG1::identity();
  HFq a = p.x.sqr(), b = p.y.sqr(), c = b.sqr();
  HFq d = ((p.x + b).sqr() - a - c).dbl();
  HFq e = a.dbl() + a;
  HFq f = e.sqr();
  HG1 r;
  r.x = f - d.dbl();
  r.y = e * (d - r.x) - c.dbl().dbl().dbl();
  r.z = (p.y * p.z).dbl();
  return r;
}
static inline HG1 g1_add(const HG1& p, const HG1& q) {
  if (p.is_identity()) return q;
  if (q.is_identity()) return p;
  HFq z1z1 = p.z.sqr(), z2z2 = q.z.sqr();
  HFq u1 = p.x * z2z2, u2 = q.x * z1z1;
  HFq s1 = p.y * q.z * z2z2, s2 = q.y * p.z * z1z1;
  if (u1 == u2) {
    if (s1 == s2) return g1_dbl(p);
    return HG1::identity();
  }
  HFq h = u2 - u1;
  HFq i = h.dbl().sqr();
  HFq j = h * i;
  HFq rr = (s2 - s1).dbl();
  HFq v = u1 * i;
  HG1 r;
  r.x = rr.sqr() - j - v.dbl();
  r.y = rr * (v - r.x) - (s1 * j).dbl();
  r.z = ((p.z + q.z).sqr() - z1z1 - z2z2) * h;
  return r;
}
// XYZZ (x, y, zz, zzz) -> Jacobian: x/zz = X/Z^2, y/zzz = Y/Z^3 with Z = zzz/zz requires a
// division; instead use the identity (x*zz^2... ) cheap map: X = x*zzz^2*zz... see msm.cu.
// (Implemented there as Jacobian with Z = zz*zzz: X = x * zz * zzz^2, Y = y * zz^3 * zzz^2.)
static inline HG1 g1_from_xyzz(const HFq& x, const HFq& y, const HFq& zz, const HFq& zzz) {
  if (zz.is_zero()) return HG1::identity();
  // Z = zz*zzz  =>  Z^2 = zz^2 zzz^2, Z^3 = zz^3 zzz^3; affine x_a = x/zz, y_a = y/zzz
  // X = x_a Z^2 = x * zz * zzz^2 ; Y = y_a Z^3 = y * zz^3 * zzz^2
  HFq zzz2 = zzz.sqr();
  HFq zz2 = zz.sqr();
  HG1 r;
  r.x = x * zz * zzz2;
  r.y = y * zz2 * zz * zzz2;
  r.z = zz * zzz;
  return r;
}
// -> affine (x, y); returns false for identity.
static inline bool g1_to_affine(const HG1& p, HFq* ax, HFq* ay) {
  if (p.is_identity()) return false;
  HFq zi = p.z.inv();
  HFq zi2 = zi.sqr();
  *ax = p.x * zi2;
  *ay = p.y * zi2 * zi;
  return true;
}
static inline HG1 g1_from_affine(const HFq& x, const HFq& y) { return {x, y, HFq::one()}; }

static inline HG1 g1_mul_u64limbs(const HG1& p, const uint64_t* k, int n) {
  HG1 r = HG1::identity();
  for (int i = n - 1; i >= 0; i--)
    for (int b = 63; b >= 0; b--) {
      r = g1_dbl(r);
      if ((k[i] >> b) & 1) r = g1_add(r, p);
    }
  return r;
}

}  // namespace tph
