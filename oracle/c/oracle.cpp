// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into or called from the product library
// (typlonk_b200/); used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs as the checker and the CPU baseline.
//
// C++ restatement of fabrizio-m/TyPLONK's prover for the CPU, multi-threaded (OpenMP), with
// "arkworks-grade" algorithms where the reference is quadratic:
//   kzg/src/lib.rs:37-64        commit / open            -> Pippenger MSM (ark-ec 0.3 window rule)
//   kzg/src/srs.rs:15-34        Srs::from_secret         -> fixed-base windowed multiplication
//   permutation/src/lib.rs:101-154  Permutation::compile, cosets
//   permutation/src/proving.rs:7-31 grand product        -> batch inversion + running product
//   plonk/src/proof.rs:96-194   prove                    -> same round structure
//   plonk/src/proof.rs:292-375  quotient_polynomial      -> products on a 4n domain instead of
//                                                           naive_mul; the division by X^n - 1 is
//                                                           done in coefficient space so the floor
//                                                           quotient matches the reference even for
//                                                           unsatisfied copy constraints
//   plonk/src/proof.rs:376-439  linearisation_poly
//   plonk/src/proof/challenges.rs:9-46  Fiat-Shamir (Blake2b-512 -> ChaCha12 StdRng -> Fr::rand)
// The arithmetic crates (ark-ff/ark-ec/ark-poly/ark-serialize 0.3.0, rand 0.8.4, blake2 0.9.2)
// are not vendored in the reference tree; their published algorithms are restated.
// Parity: cross-checked against the Python big-integer oracle (oracle/pyoracle) in
// tests/test_oracle_c.py; byte-level parity with real arkworks is UNPINNED (no Rust toolchain,
// the reference ships no golden vectors) -- see DESIGN.md.
//
// Build: g++ -O3 -fopenmp -shared -fPIC oracle.cpp -o ../_build/liboracle.so   (oracle/coracle.py)
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <vector>
#include <cstdio>
#include <cstdlib>
extern "C" double oracle_now_ms();

typedef unsigned __int128 u128;

// ============================================================================================
// Prime fields (Montgomery, little-endian u64 limbs, same layout as ark_ff::Fp256 / Fp384)
// ============================================================================================
template <int N>
struct Params {
  uint64_t p[N], one[N], r2[N], inv;
};
static const Params<4> PR = {{0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull},
                             {0x00000001fffffffeull, 0x5884b7fa00034802ull, 0x998c4fefecbc4ff5ull, 0x1824b159acc5056full},
                             {0xc999e990f3f29c6dull, 0x2b6cedcb87925c23ull, 0x05d314967254398full, 0x0748d9d99f59ff11ull},
                             0xfffffffeffffffffull};
static const Params<6> PQ = {{0xb9feffffffffaaabull, 0x1eabfffeb153ffffull, 0x6730d2a0f6b0f624ull, 0x64774b84f38512bfull,
                              0x4b1ba7b6434bacd7ull, 0x1a0111ea397fe69aull},
                             {0x760900000002fffdull, 0xebf4000bc40c0002ull, 0x5f48985753c758baull, 0x77ce585370525745ull,
                              0x5c071a97a256ec6dull, 0x15f65ec3fa80e493ull},
                             {0xf4df1f341c341746ull, 0x0a76e6a609d104f1ull, 0x8de5476c4c95b6d5ull, 0x67eb88a9939d83c0ull,
                              0x9a793e85b519952dull, 0x11988fe592cae3aaull},
                             0x89f3fffcfffcfffdull};

template <int N, const Params<N>& P>
struct Fp {
  uint64_t l[N];
  static Fp zero() { Fp r; memset(r.l, 0, sizeof r.l); return r; }
  static Fp one() { Fp r; memcpy(r.l, P.one, sizeof r.l); return r; }
  bool is_zero() const { uint64_t o = 0; for (int i = 0; i < N; i++) o |= l[i]; return o == 0; }
  bool operator==(const Fp& o) const { return memcmp(l, o.l, sizeof l) == 0; }
  bool operator!=(const Fp& o) const { return !(*this == o); }
  static bool geq_p(const uint64_t* a) {
    for (int i = N - 1; i >= 0; i--) if (a[i] != P.p[i]) return a[i] > P.p[i];
    return true;
  }
  static void sub_p(uint64_t* a) {
    uint64_t br = 0;
    for (int i = 0; i < N; i++) { u128 t = (u128)a[i] - P.p[i] - br; a[i] = (uint64_t)t; br = (uint64_t)(t >> 64) & 1; }
  }
  Fp operator+(const Fp& o) const {
    Fp r; uint64_t c = 0;
    for (int i = 0; i < N; i++) { u128 t = (u128)l[i] + o.l[i] + c; r.l[i] = (uint64_t)t; c = (uint64_t)(t >> 64); }
    if (c || geq_p(r.l)) sub_p(r.l);
    return r;
  }
  Fp operator-(const Fp& o) const {
    Fp r; uint64_t br = 0;
    for (int i = 0; i < N; i++) { u128 t = (u128)l[i] - o.l[i] - br; r.l[i] = (uint64_t)t; br = (uint64_t)(t >> 64) & 1; }
    if (br) { uint64_t c = 0; for (int i = 0; i < N; i++) { u128 t = (u128)r.l[i] + P.p[i] + c; r.l[i] = (uint64_t)t; c = (uint64_t)(t >> 64); } }
    return r;
  }
  Fp neg() const { return zero() - *this; }
  Fp dbl() const { return *this + *this; }
  // Montgomery product, CIOS with the "no-carry" shortcut ark-ff 0.3 uses for moduli whose top bit is clear
  // (both r and q here): the two inner chains never overflow a limb pair, so no extra carry words are kept.
  Fp operator*(const Fp& o) const {
    uint64_t t[N];
#pragma GCC unroll 8
    for (int j = 0; j < N; j++) t[j] = 0;
#pragma GCC unroll 8
    for (int i = 0; i < N; i++) {
      u128 x = (u128)l[0] * o.l[i] + t[0];
      uint64_t c = (uint64_t)(x >> 64);
      const uint64_t m = (uint64_t)x * P.inv;
      u128 y = (u128)m * P.p[0] + (uint64_t)x;
      uint64_t c2 = (uint64_t)(y >> 64);
#pragma GCC unroll 8
      for (int j = 1; j < N; j++) {
        x = (u128)l[j] * o.l[i] + t[j] + c;
        c = (uint64_t)(x >> 64);
        y = (u128)m * P.p[j] + (uint64_t)x + c2;
        c2 = (uint64_t)(y >> 64);
        t[j - 1] = (uint64_t)y;
      }
      t[N - 1] = c + c2;
    }
    Fp r;
#pragma GCC unroll 8
    for (int j = 0; j < N; j++) r.l[j] = t[j];
    if (geq_p(r.l)) sub_p(r.l);
    return r;
  }
  Fp sqr() const { return *this * *this; }
  Fp pow(const uint64_t* e, int n) const {
    Fp r = one(); bool st = false;
    for (int i = n - 1; i >= 0; i--) for (int b = 63; b >= 0; b--) {
      if (st) r = r.sqr();
      if ((e[i] >> b) & 1) { r = st ? r * *this : *this; st = true; }
    }
    return r;
  }
  Fp pow64(uint64_t e) const { return pow(&e, 1); }
  Fp inv() const { uint64_t e[N]; memcpy(e, P.p, sizeof e); e[0] -= 2; return pow(e, N); }
  static Fp from_u64(uint64_t x) { Fp a = zero(); a.l[0] = x; Fp r2; memcpy(r2.l, P.r2, sizeof r2.l); return a * r2; }
  Fp canonical() const { Fp o = zero(); o.l[0] = 1; return *this * o; }
  static Fp from_canonical(const uint64_t* c) { Fp a; memcpy(a.l, c, sizeof a.l); Fp r2; memcpy(r2.l, P.r2, sizeof r2.l); return a * r2; }
};
typedef Fp<4, PR> Fr;
typedef Fp<6, PQ> Fq;

// ============================================================================================
// G1: y^2 = x^3 + 4, Jacobian coordinates
// ============================================================================================
struct Aff { Fq x, y; bool inf; };
struct Jac {
  Fq x, y, z;
  static Jac id() { return {Fq::one(), Fq::one(), Fq::zero()}; }
  bool is_id() const { return z.is_zero(); }
};
static Jac jdbl(const Jac& p) {
  if (p.is_id() || p.y.is_zero()) return Jac::id();
  Fq a = p.x.sqr(), b = p.y.sqr(), c = b.sqr();
  Fq d = ((p.x + b).sqr() - a - c).dbl();
  Fq e = a.dbl() + a, f = e.sqr();
  Jac r; r.x = f - d.dbl(); r.y = e * (d - r.x) - c.dbl().dbl().dbl(); r.z = (p.y * p.z).dbl();
  return r;
}
static Jac jadd(const Jac& p, const Jac& q) {
  if (p.is_id()) return q;
  if (q.is_id()) return p;
  Fq z1z1 = p.z.sqr(), z2z2 = q.z.sqr();
  Fq u1 = p.x * z2z2, u2 = q.x * z1z1, s1 = p.y * q.z * z2z2, s2 = q.y * p.z * z1z1;
  if (u1 == u2) return s1 == s2 ? jdbl(p) : Jac::id();
  Fq h = u2 - u1, i = h.dbl().sqr(), j = h * i, rr = (s2 - s1).dbl(), v = u1 * i;
  Jac r; r.x = rr.sqr() - j - v.dbl(); r.y = rr * (v - r.x) - (s1 * j).dbl(); r.z = ((p.z + q.z).sqr() - z1z1 - z2z2) * h;
  return r;
}
static Jac jmadd(const Jac& p, const Aff& q) {  // mixed addition
  if (q.inf) return p;
  if (p.is_id()) return {q.x, q.y, Fq::one()};
  Fq z1z1 = p.z.sqr(), u2 = q.x * z1z1, s2 = q.y * p.z * z1z1;
  if (p.x == u2) return p.y == s2 ? jdbl(p) : Jac::id();
  Fq h = u2 - p.x, hh = h.sqr(), i = hh.dbl().dbl(), j = h * i, rr = (s2 - p.y).dbl(), v = p.x * i;
  Jac r; r.x = rr.sqr() - j - v.dbl(); r.y = rr * (v - r.x) - (p.y * j).dbl(); r.z = (p.z + h).sqr() - z1z1 - hh;
  return r;
}
static Aff to_aff(const Jac& p) {
  if (p.is_id()) return {Fq::zero(), Fq::one(), true};
  Fq zi = p.z.inv(), zi2 = zi.sqr();
  return {p.x * zi2, p.y * zi2 * zi, false};
}
static void batch_to_aff(const std::vector<Jac>& in, Aff* out) {
  size_t n = in.size();
  std::vector<Fq> pre(n);
  Fq acc = Fq::one();
  for (size_t i = 0; i < n; i++) { pre[i] = acc; if (!in[i].is_id()) acc = acc * in[i].z; }
  Fq inv = acc.inv();
  for (size_t i = n; i-- > 0;) {
    if (in[i].is_id()) { out[i] = {Fq::zero(), Fq::one(), true}; continue; }
    Fq zi = inv * pre[i]; inv = inv * in[i].z;
    Fq zi2 = zi.sqr();
    out[i] = {in[i].x * zi2, in[i].y * zi2 * zi, false};
  }
}
static const uint64_t GX[6] = {0xfb3af00adb22c6bbull, 0x6c55e83ff97a1aefull, 0xa14e3a3f171bac58ull, 0xc3688c4f9774b905ull, 0x2695638c4fa9ac0full, 0x17f1d3a73197d794ull};
static const uint64_t GY[6] = {0x0caa232946c5e7e1ull, 0xd03cc744a2888ae4ull, 0x00db18cb2c04b3edull, 0xfcf5e095d5d00af6ull, 0xa09e30ed741d8ae4ull, 0x08b3f481e3aaa0f1ull};
static Aff generator() { return {Fq::from_canonical(GX), Fq::from_canonical(GY), false}; }

// kzg/src/lib.rs:46-53 as written: per-point double-and-add (used only by the "as written" timing)
static Jac scalar_mul(const Aff& p, const Fr& k) {
  Fr c = k.canonical();
  Jac r = Jac::id();
  for (int i = 3; i >= 0; i--) for (int b = 63; b >= 0; b--) { r = jdbl(r); if ((c.l[i] >> b) & 1) r = jmadd(r, p); }
  return r;
}

// Pippenger with the ark-ec 0.3.0 window rule (c = 3 if n < 32 else ln(n) + 2); point range split
// over threads, each running the full bucket method on its slice.
static unsigned ln_without_floats(size_t a) { unsigned l = 0; while (a >>= 1) l++; return l * 69 / 100; }
static Jac msm_slice(const Aff* pts, const Fr* sc, size_t n) {
  if (n == 0) return Jac::id();
  unsigned c = n < 32 ? 3 : ln_without_floats(n) + 2;
  std::vector<Fr> can(n);
  for (size_t i = 0; i < n; i++) can[i] = sc[i].canonical();
  unsigned nwin = (255 + c - 1) / c;
  std::vector<Jac> buckets((size_t)1 << c);
  Jac total = Jac::id();
  for (int w = (int)nwin - 1; w >= 0; w--) {
    for (unsigned d = 0; d < c; d++) total = jdbl(total);
    for (auto& b : buckets) b = Jac::id();
    unsigned bit = w * c;
    for (size_t i = 0; i < n; i++) {
      unsigned limb = bit >> 6, off = bit & 63;
      uint64_t v = can[i].l[limb] >> off;
      if (off + c > 64 && limb + 1 < 4) v |= can[i].l[limb + 1] << (64 - off);
      v &= ((uint64_t)1 << c) - 1;
      if (v) buckets[v] = jmadd(buckets[v], pts[i]);
    }
    Jac run = Jac::id(), acc = Jac::id();
    for (size_t v = buckets.size() - 1; v >= 1; v--) { run = jadd(run, buckets[v]); acc = jadd(acc, run); }
    total = jadd(total, acc);
  }
  return total;
}
static Jac msm(const Aff* pts, const Fr* sc, size_t n) {
  int T = omp_get_max_threads();
  if (n < 256) T = 1;
  std::vector<Jac> parts(T, Jac::id());
#pragma omp parallel for schedule(static, 1)
  for (int t = 0; t < T; t++) {
    size_t lo = n * t / T, hi = n * (t + 1) / T;
    parts[t] = msm_slice(pts + lo, sc + lo, hi - lo);
  }
  Jac r = Jac::id();
  for (auto& p : parts) r = jadd(r, p);
  return r;
}

// Srs::g1 (kzg/src/srs.rs:15-24) via a fixed-base table of 32 x 255 multiples of G.
static void srs_g1(const Fr& tau, size_t len, Aff* out) {
  std::vector<Jac> tabj(32 * 255);
  Jac base = {generator().x, generator().y, Fq::one()};
  for (int j = 0; j < 32; j++) {
    Jac acc = base;
    for (int d = 0; d < 255; d++) { tabj[j * 255 + d] = acc; acc = jadd(acc, base); }
    base = acc;
  }
  std::vector<Aff> tab(32 * 255);
  batch_to_aff(tabj, tab.data());
  std::vector<Fr> pw(len);
  Fr cur = Fr::one();
  for (size_t i = 0; i < len; i++) { pw[i] = cur; cur = cur * tau; }
  const size_t CH = 1024;
#pragma omp parallel for schedule(dynamic, 1)
  for (size_t lo = 0; lo < len; lo += CH) {
    size_t hi = std::min(len, lo + CH);
    std::vector<Jac> tmp(hi - lo);
    for (size_t i = lo; i < hi; i++) {
      Fr c = pw[i].canonical();
      Jac acc = Jac::id();
      for (int j = 0; j < 32; j++) {
        unsigned d = (c.l[j >> 3] >> ((j & 7) * 8)) & 0xff;
        if (d) acc = jmadd(acc, tab[j * 255 + d - 1]);
      }
      tmp[i - lo] = acc;
    }
    batch_to_aff(tmp, out + lo);
  }
}

// ============================================================================================
// Radix-2 NTT (ark-poly 0.3 Radix2EvaluationDomain semantics: natural order in/out)
// ============================================================================================
static const uint64_t ROOT32[4] = {0x3829971f439f0d2bull, 0xb63683508c2280b9ull, 0xd09b681922c813b4ull, 0x16a2a19edfe81f20ull};
static Fr root_of_unity(unsigned log_n) { Fr w = Fr::from_canonical(ROOT32); for (unsigned i = log_n; i < 32; i++) w = w.sqr(); return w; }
static void ntt(Fr* a, unsigned log_n, bool inverse) {
  size_t n = (size_t)1 << log_n;
  for (size_t i = 0; i < n; i++) {
    size_t j = 0;
    for (unsigned b = 0; b < log_n; b++) if (i >> b & 1) j |= (size_t)1 << (log_n - 1 - b);
    if (i < j) std::swap(a[i], a[j]);
  }
  Fr w = root_of_unity(log_n);
  if (inverse) w = w.inv();
  std::vector<Fr> tw(n / 2 ? n / 2 : 1);
  tw[0] = Fr::one();
  for (size_t i = 1; i < n / 2; i++) tw[i] = tw[i - 1] * w;
  for (unsigned s = 1; s <= log_n; s++) {
    size_t half = (size_t)1 << (s - 1), step = n >> s;
#pragma omp parallel for schedule(static) if (n >= 4096)
    for (size_t b = 0; b < n / 2; b++) {
      size_t blk = b / half, k = b % half;
      size_t i0 = blk * 2 * half + k, i1 = i0 + half;
      Fr u = a[i0], v = a[i1] * tw[k * step];
      a[i0] = u + v;
      a[i1] = u - v;
    }
  }
  if (inverse) {
    Fr ni = Fr::from_u64(n).inv();
#pragma omp parallel for schedule(static) if (n >= 4096)
    for (size_t i = 0; i < n; i++) a[i] = a[i] * ni;
  }
}

// ============================================================================================
// Blake2b-512, ChaCha12 StdRng, Fr::rand  (plonk/src/proof/challenges.rs:31-45)
// ============================================================================================
static inline uint64_t rotr64(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
static void blake2b_512(const uint8_t* in, size_t len, uint8_t out[64]) {
  static const uint64_t IV[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                                 0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
  static const uint8_t S[12][16] = {{0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
                                    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
                                    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
                                    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
                                    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
                                    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
  uint64_t h[8];
  memcpy(h, IV, sizeof h);
  h[0] ^= 0x01010040ull;  // digest 64 bytes, no key, fanout = depth = 1
  uint64_t t = 0;
  auto compress = [&](const uint8_t* blk, bool last) {
    uint64_t m[16], v[16];
    for (int i = 0; i < 16; i++) { m[i] = 0; for (int j = 7; j >= 0; j--) m[i] = (m[i] << 8) | blk[8 * i + j]; }
    for (int i = 0; i < 8; i++) { v[i] = h[i]; v[i + 8] = IV[i]; }
    v[12] ^= t;
    if (last) v[14] = ~v[14];
#define G(a, b, c, d, x, y) v[a] += v[b] + x; v[d] = rotr64(v[d] ^ v[a], 32); v[c] += v[d]; v[b] = rotr64(v[b] ^ v[c], 24); \
                            v[a] += v[b] + y; v[d] = rotr64(v[d] ^ v[a], 16); v[c] += v[d]; v[b] = rotr64(v[b] ^ v[c], 63);
    for (int r = 0; r < 12; r++) {
      const uint8_t* s = S[r];
      G(0, 4, 8, 12, m[s[0]], m[s[1]]) G(1, 5, 9, 13, m[s[2]], m[s[3]]) G(2, 6, 10, 14, m[s[4]], m[s[5]]) G(3, 7, 11, 15, m[s[6]], m[s[7]])
      G(0, 5, 10, 15, m[s[8]], m[s[9]]) G(1, 6, 11, 12, m[s[10]], m[s[11]]) G(2, 7, 8, 13, m[s[12]], m[s[13]]) G(3, 4, 9, 14, m[s[14]], m[s[15]])
    }
#undef G
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
  };
  size_t off = 0;
  while (len - off > 128) { t += 128; compress(in + off, false); off += 128; }
  uint8_t blk[128];
  memset(blk, 0, 128);
  memcpy(blk, in + off, len - off);
  t += len - off;
  compress(blk, true);
  for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) out[8 * i + j] = (uint8_t)(h[i] >> (8 * j));
}
struct Rng {
  uint32_t key[8]; uint64_t ctr; uint32_t buf[16]; int pos;
  static uint32_t rl(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
  static Rng seed(uint64_t s) {
    Rng r;
    for (int i = 0; i < 8; i++) {
      s = s * 6364136223846793005ull + 11634580027462260723ull;
      uint32_t xs = (uint32_t)(((s >> 18) ^ s) >> 27), rot = (uint32_t)(s >> 59);
      r.key[i] = (xs >> rot) | (xs << ((32 - rot) & 31));
    }
    r.ctr = 0; r.pos = 16;
    return r;
  }
  void refill() {
    uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
    for (int i = 0; i < 8; i++) in[4 + i] = key[i];
    in[12] = (uint32_t)ctr; in[13] = (uint32_t)(ctr >> 32); in[14] = in[15] = 0;
    uint32_t x[16]; memcpy(x, in, sizeof x);
#define QR(a, b, c, d) x[a] += x[b]; x[d] = rl(x[d] ^ x[a], 16); x[c] += x[d]; x[b] = rl(x[b] ^ x[c], 12); \
                       x[a] += x[b]; x[d] = rl(x[d] ^ x[a], 8); x[c] += x[d]; x[b] = rl(x[b] ^ x[c], 7);
    for (int r = 0; r < 6; r++) { QR(0, 4, 8, 12) QR(1, 5, 9, 13) QR(2, 6, 10, 14) QR(3, 7, 11, 15) QR(0, 5, 10, 15) QR(1, 6, 11, 12) QR(2, 7, 8, 13) QR(3, 4, 9, 14) }
#undef QR
    for (int i = 0; i < 16; i++) buf[i] = x[i] + in[i];
    ctr++; pos = 0;
  }
  uint64_t next64() { if (pos >= 16) refill(); uint64_t lo = buf[pos++]; if (pos >= 16) refill(); uint64_t hi = buf[pos++]; return lo | (hi << 32); }
};
static Fr fr_rand(Rng& g) {
  for (;;) {
    Fr r; for (int i = 0; i < 4; i++) r.l[i] = g.next64();
    r.l[3] &= 0x7fffffffffffffffull;
    if (!Fr::geq_p(r.l)) return r;  // the limbs ARE the Montgomery representation (ark-ff 0.3)
  }
}
static void ser_g1(const Aff& p, uint8_t out[96]) {  // ark-serialize 0.3 serialize_unchecked
  if (p.inf) { memset(out, 0, 96); out[48] = 1; out[95] |= 1 << 6; return; }
  Fq x = p.x.canonical(), y = p.y.canonical();
  memcpy(out, x.l, 48); memcpy(out + 48, y.l, 48);
}
static void challenges2(const std::vector<Aff>& coms, Fr* a, Fr* b) {
  std::vector<uint8_t> tr(coms.size() * 96);
  for (size_t i = 0; i < coms.size(); i++) ser_g1(coms[i], &tr[i * 96]);
  uint8_t d[64];
  blake2b_512(tr.data(), tr.size(), d);
  uint64_t seed = 0; for (int i = 7; i >= 0; i--) seed = (seed << 8) | d[i];
  Rng g = Rng::seed(seed);
  *a = fr_rand(g); *b = fr_rand(g);
}

// ============================================================================================
// Circuit + prover
// ============================================================================================
struct Circuit {
  size_t n; unsigned log_n;
  std::vector<Aff> srs;
  std::vector<Fr> sel_eval[5], sel_coef[5], id[3], sig[3], sig_coef[3];
  Fr k[3];
};
static std::vector<Fr> horner_div(const std::vector<Fr>& p, const Fr& z, Fr* y) {  // kzg/src/lib.rs:57-61
  size_t n = p.size();
  std::vector<Fr> q(n ? n - 1 : 0);
  Fr acc = Fr::zero();
  for (size_t k = n; k-- > 0;) { acc = p[k] + z * acc; if (k >= 1) q[k - 1] = acc; }
  *y = acc;
  return q;
}
static Fr eval(const std::vector<Fr>& p, const Fr& z) { Fr acc = Fr::zero(); for (size_t k = p.size(); k-- > 0;) acc = p[k] + z * acc; return acc; }
static Aff commit(const Circuit& c, const std::vector<Fr>& p) { return to_aff(msm(c.srs.data(), p.data(), std::min(p.size(), c.srs.size()))); }

extern "C" {

int oracle_threads() { return omp_get_max_threads(); }

// count x Fr::rand of StdRng::seed_from_u64(seed) as Montgomery limbs
void oracle_fr_rand_stream(uint64_t seed, size_t count, uint64_t* out) {
  Rng g = Rng::seed(seed);
  for (size_t i = 0; i < count; i++) { Fr x = fr_rand(g); memcpy(out + 4 * i, x.l, 32); }
}
void oracle_ntt(uint64_t* data, unsigned log_n, int inverse) { ntt((Fr*)data, log_n, inverse != 0); }
// points: 96 B (x | y Montgomery), all-zero = infinity; out: 97 B ABI point
void oracle_msm(const uint8_t* pts, const uint64_t* scalars, size_t n, uint8_t out[97]) {
  std::vector<Aff> a(n);
  for (size_t i = 0; i < n; i++) {
    memcpy(a[i].x.l, pts + i * 96, 48); memcpy(a[i].y.l, pts + i * 96 + 48, 48);
    a[i].inf = a[i].x.is_zero() && a[i].y.is_zero();
  }
  Aff r = to_aff(msm(a.data(), (const Fr*)scalars, n));
  memcpy(out, r.x.l, 48); memcpy(out + 48, r.y.l, 48); out[96] = r.inf ? 1 : 0;
}
void oracle_srs(const uint64_t tau[4], size_t len, uint8_t* out) {
  Fr t; memcpy(t.l, tau, 32);
  std::vector<Aff> pts(len);
  srs_g1(t, len, pts.data());
  for (size_t i = 0; i < len; i++) {
    if (pts[i].inf) { memset(out + i * 96, 0, 96); continue; }
    memcpy(out + i * 96, pts[i].x.l, 48); memcpy(out + i * 96 + 48, pts[i].y.l, 48);
  }
}

// Setup numerics of CircuitBuilder::compile (plonk/src/builder.rs:70-88): SRS from tau, selector
// interpolation, sigma / id tables.  selector_evals: 5 x n Montgomery Fr; perm: 3n indices.
void* oracle_circuit_new(const uint64_t tau[4], const uint64_t* const selector_evals[5], const uint64_t* perm, size_t n) {
  Circuit* c = new Circuit();
  c->n = n; c->log_n = 0; while (((size_t)1 << c->log_n) < n) c->log_n++;
  Fr t; memcpy(t.l, tau, 32);
  c->srs.resize(n + 3);
  srs_g1(t, n + 3, c->srs.data());
  for (int i = 0; i < 5; i++) {
    c->sel_eval[i].assign((const Fr*)selector_evals[i], (const Fr*)selector_evals[i] + n);
    c->sel_coef[i] = c->sel_eval[i];
    ntt(c->sel_coef[i].data(), c->log_n, true);
  }
  uint64_t kv = 1;
  for (int i = 0; i < 3; i++) {  // permutation/src/lib.rs:141-154
    while (Fr::from_u64(kv).pow64(n) == Fr::one()) kv++;
    c->k[i] = Fr::from_u64(kv++);
  }
  std::vector<Fr> roots(n);
  Fr w = root_of_unity(c->log_n);
  roots[0] = Fr::one();
  for (size_t j = 1; j < n; j++) roots[j] = roots[j - 1] * w;
  for (int i = 0; i < 3; i++) {  // permutation/src/lib.rs:108-118
    c->id[i].resize(n); c->sig[i].resize(n);
    for (size_t j = 0; j < n; j++) {
      uint64_t p = perm[i * n + j];
      c->id[i][j] = c->k[i] * roots[j];
      c->sig[i][j] = c->k[p / n] * roots[p % n];
    }
    c->sig_coef[i] = c->sig[i];
    ntt(c->sig_coef[i].data(), c->log_n, true);
  }
  return c;
}
void oracle_circuit_free(void* c) { delete (Circuit*)c; }
void oracle_circuit_fixed_commitments(void* cv, uint8_t out[5 * 97]) {
  Circuit* c = (Circuit*)cv;
  for (int i = 0; i < 5; i++) { Aff a = commit(*c, c->sel_coef[i]); memcpy(out + 97 * i, a.x.l, 48); memcpy(out + 97 * i + 48, a.y.l, 48); out[97 * i + 96] = a.inf; }
}

// prove() (plonk/src/proof.rs:96-194).  advice: 3 x n evaluations (blinders included);
// public_inputs: n.  Writes the 1472-byte fixed proof block (same layout as tp_prove).
// Returns 0, or 6 where the reference's `vanishes(line1)` assert fires, 5 on a zero denominator.
int oracle_prove(void* cv, const uint64_t* const advice[3], const uint64_t* public_inputs, uint8_t* out) {
  Circuit& c = *(Circuit*)cv;
  const size_t n = c.n;
  const bool timing = getenv("ORACLE_TIMING") != nullptr;
  double t_last = oracle_now_ms();
  auto lap = [&](const char* what) { if (timing) { double t = oracle_now_ms(); fprintf(stderr, "[oracle] %-12s %9.1f ms\n", what, t - t_last); t_last = t; } };
  std::vector<Fr> ev[3], co[3], pi((const Fr*)public_inputs, (const Fr*)public_inputs + n), pic;
  for (int i = 0; i < 3; i++) { ev[i].assign((const Fr*)advice[i], (const Fr*)advice[i] + n); co[i] = ev[i]; ntt(co[i].data(), c.log_n, true); }
  pic = pi; ntt(pic.data(), c.log_n, true);
  lap("intt x4");
  std::vector<Aff> com(4);
  for (int i = 0; i < 3; i++) com[i] = commit(c, co[i]);
  lap("commit abc");
  Fr beta, gamma;
  challenges2({com[0], com[1], com[2]}, &beta, &gamma);
  for (size_t j = 0; j < n; j++) {  // vanishes(line1), proof.rs:317-321
    Fr g = c.sel_eval[0][j] * ev[0][j] + c.sel_eval[1][j] * ev[1][j] - c.sel_eval[2][j] * ev[2][j] + c.sel_eval[3][j] * ev[0][j] * ev[1][j] + c.sel_eval[4][j] + pi[j];
    if (!g.is_zero()) return 6;
  }
  // grand product (proving.rs:7-31) with batch inversion
  std::vector<Fr> num(n), den(n), z(n + 1);
  for (size_t j = 0; j < n; j++) {
    Fr nu = Fr::one(), de = Fr::one();
    for (int i = 0; i < 3; i++) {
      Fr v = ev[i][j] + gamma;
      Fr d = v + beta * c.sig[i][j];
      if (d.is_zero()) return 5;
      nu = nu * (v + beta * c.id[i][j]);
      de = de * d;
    }
    num[j] = nu; den[j] = de;
  }
  {
    std::vector<Fr> pre(n);
    Fr acc = Fr::one();
    for (size_t j = 0; j < n; j++) { pre[j] = acc; acc = acc * den[j]; }
    Fr inv = acc.inv();
    for (size_t j = n; j-- > 0;) { Fr di = inv * pre[j]; inv = inv * den[j]; num[j] = num[j] * di; }
    z[0] = Fr::one();
    for (size_t j = 0; j < n; j++) z[j + 1] = z[j] * num[j];
  }
  lap("grand prod");
  std::vector<Fr> zc(z.begin(), z.begin() + n);
  ntt(zc.data(), c.log_n, true);
  com[3] = commit(c, zc);
  lap("commit z");
  Fr alpha, zeta;
  challenges2(com, &alpha, &zeta);

  // quotient on the 4n domain, division by X^n - 1 in coefficient space
  const size_t n4 = 4 * n;
  auto to4 = [&](const std::vector<Fr>& p) { std::vector<Fr> r(n4, Fr::zero()); std::copy(p.begin(), p.end(), r.begin()); ntt(r.data(), c.log_n + 2, false); return r; };
  std::vector<Fr> a4 = to4(co[0]), b4 = to4(co[1]), c4 = to4(co[2]), z4 = to4(zc), pi4 = to4(pic);
  std::vector<Fr> s4[5], g4[3];
  for (int i = 0; i < 5; i++) s4[i] = to4(c.sel_coef[i]);
  for (int i = 0; i < 3; i++) g4[i] = to4(c.sig_coef[i]);
  std::vector<Fr> l0c(n, Fr::from_u64(n).inv());  // L0 = (1/n) sum X^k  (utils.rs:150-159)
  std::vector<Fr> l04 = to4(l0c);
  std::vector<Fr> num4(n4);
  Fr w4 = root_of_unity(c.log_n + 2), alpha2 = alpha.sqr();
  std::vector<Fr> xs(n4);
  xs[0] = Fr::one();
  for (size_t i = 1; i < n4; i++) xs[i] = xs[i - 1] * w4;
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n4; i++) {
    Fr a = a4[i], b = b4[i], cc = c4[i], zz = z4[i], zw = z4[(i + 4) % n4], x = xs[i];
    Fr gate = s4[0][i] * a + s4[1][i] * b - s4[2][i] * cc + s4[3][i] * a * b + s4[4][i] + pi4[i];
    Fr l2 = (a + beta * c.k[0] * x + gamma) * (b + beta * c.k[1] * x + gamma) * (cc + beta * c.k[2] * x + gamma) * zz;
    Fr l3 = (a + beta * g4[0][i] + gamma) * (b + beta * g4[1][i] + gamma) * (cc + beta * g4[2][i] + gamma) * zw;
    Fr l4 = (zz - Fr::one()) * l04[i];
    num4[i] = gate + alpha * (l2 - l3) + alpha2 * l4;
  }
  ntt(num4.data(), c.log_n + 2, true);
  std::vector<Fr> t(3 * n);
  for (size_t k = 0; k < n; k++) { Fr t2 = num4[k + 3 * n], t1 = num4[k + 2 * n] + t2, t0 = num4[k + n] + t1; t[k + 2 * n] = t2; t[k + n] = t1; t[k] = t0; }

  lap("quotient");
  // openings
  Fr omega = root_of_unity(c.log_n);
  Fr y[5], pts5[5] = {zeta, zeta, zeta, zeta, zeta * omega};
  const std::vector<Fr>* polys[5] = {&co[0], &co[1], &co[2], &zc, &zc};
  Aff wit[6];
  for (int i = 0; i < 5; i++) { std::vector<Fr> q = horner_div(*polys[i], pts5[i], &y[i]); wit[i] = commit(c, q); }
  // linearisation (proof.rs:376-439)
  Fr sb0 = eval(c.sig_coef[0], zeta), sb1 = eval(c.sig_coef[1], zeta), pib = eval(pic, zeta);
  Fr l2 = Fr::one();
  for (int i = 0; i < 3; i++) l2 = l2 * (y[i] + c.k[i] * beta * zeta + gamma);
  Fr perm_ab = (y[0] + beta * sb0 + gamma) * (y[1] + beta * sb1 + gamma);
  Fr zn = zeta.pow64(n), zh = zn - Fr::one();
  Fr l0 = zeta == Fr::one() ? Fr::one() : zh * (Fr::from_u64(n) * (zeta - Fr::one())).inv();
  Fr abz = alpha * perm_ab * y[4];
  Fr s_z = alpha * l2 + alpha2 * l0, s_s3 = (abz * beta).neg(), s_t0 = zh.neg(), s_t1 = (zh * zn).neg(), s_t2 = (zh * zn * zn).neg();
  std::vector<Fr> r(n);
#pragma omp parallel for schedule(static)
  for (size_t k = 0; k < n; k++)
    r[k] = y[0] * c.sel_coef[0][k] + y[1] * c.sel_coef[1][k] - y[2] * c.sel_coef[2][k] + y[0] * y[1] * c.sel_coef[3][k] + c.sel_coef[4][k] +
           s_z * zc[k] + s_s3 * c.sig_coef[2][k] + s_t0 * t[k] + s_t1 * t[k + n] + s_t2 * t[k + 2 * n];
  r[0] = r[0] + pib - abz * (gamma + y[2]) - alpha2 * l0;
  Fr ry;
  { std::vector<Fr> q = horner_div(r, zeta, &ry); wit[5] = commit(c, q); }
  Aff tcom[3];
  for (int i = 0; i < 3; i++) { std::vector<Fr> s(t.begin() + i * n, t.begin() + (i + 1) * n); tcom[i] = commit(c, s); }

  lap("open+commit");
  uint8_t* w = out;
  auto pg = [&](const Aff& p) { ser_g1(p, w); w += 96; };
  auto pf = [&](const Fr& x) { Fr cn = x.canonical(); memcpy(w, cn.l, 32); w += 32; };
  for (int i = 0; i < 3; i++) { pg(com[i]); pg(wit[i]); pf(y[i]); }
  pg(com[3]); pg(wit[3]); pf(y[3]); pg(wit[4]); pf(y[4]);
  pf(zeta);
  for (int i = 0; i < 3; i++) pg(tcom[i]);
  pg(wit[5]); pf(ry);
  return 0;
}

// ---- "reference as written" pieces, for the B0 timing in BASELINE.md -------------------------
// commit exactly as kzg/src/lib.rs:46-53: per-point double-and-add, affine conversion, sum.
void oracle_commit_as_written(void* cv, const uint64_t* coeffs, size_t len, uint8_t out[97]) {
  Circuit& c = *(Circuit*)cv;
  Jac acc = Jac::id();
  for (size_t i = 0; i < len && i < c.srs.size(); i++) {
    Fr k; memcpy(k.l, coeffs + 4 * i, 32);
    Aff a = to_aff(scalar_mul(c.srs[i], k));
    acc = jmadd(acc, a);
  }
  Aff r = to_aff(acc);
  memcpy(out, r.x.l, 48); memcpy(out + 48, r.y.l, 48); out[96] = r.inf;
}
// DensePolynomial::naive_mul (plonk/src/proof.rs:317-359); out has la + lb - 1 entries
void oracle_naive_mul(const uint64_t* a, size_t la, const uint64_t* b, size_t lb, uint64_t* out) {
  std::vector<Fr> r(la + lb - 1, Fr::zero());
  const Fr* A = (const Fr*)a; const Fr* B = (const Fr*)b;
  for (size_t i = 0; i < la; i++) for (size_t j = 0; j < lb; j++) r[i + j] = r[i + j] + A[i] * B[j];
  memcpy(out, r.data(), r.size() * 32);
}
double oracle_now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // extern "C"
