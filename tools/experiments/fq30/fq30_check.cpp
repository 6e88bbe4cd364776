// CPU unit test of the reduced-radix field / curve code (typlonk_b200/csrc/fq30.cuh, compiled for the
// host) against the 64-bit-limb host reference (host_field.h).  Exit code 0 = all checks passed.
#include <stdio.h>
#include <stdlib.h>

#include "../../typlonk_b200/csrc/fq30_host.h"

using namespace tp;
using namespace tph;

static uint64_t seed = 0x9e3779b97f4a7c15ull;
static uint64_t rnd() { seed ^= seed << 13; seed ^= seed >> 7; seed ^= seed << 17; return seed; }
static int fails = 0;
#define CHECK(c, msg) do { if (!(c)) { fails++; if (fails < 20) printf("FAIL %s (line %d)\n", msg, __LINE__); } } while (0)

static HFq rand_hfq() {
  uint64_t c[6];
  for (int i = 0; i < 6; i++) c[i] = rnd();
  c[5] &= 0x0fffffffffffffffull;
  return HFq::to_mont(c);  // arbitrary element
}
static Fq30 to30(const HFq& x) { Fq30 r; fq30_from_hfq(x, r.l); return r; }
static HFq from30(const Fq30& a) { return hfq_from_fq30(a.l); }
static bool limbs_ok(const Fq30& a) {
  for (int i = 0; i < 12; i++) if (a.l[i] >> 30) return false;
  return (a.l[12] >> 26) == 0;  // value < 2^386
}
// a lazily-reduced representative: add k*q
static Fq30 lazy(const Fq30& a, int k) {
  Fq30Tables T = fq30_tables();
  Fq30 kq; for (int i = 0; i < 13; i++) kq.l[i] = T.kq[k][i];
  return fq30_add(a, kq);
}
static const uint64_t GXc[6] = {0xfb3af00adb22c6bbull, 0x6c55e83ff97a1aefull, 0xa14e3a3f171bac58ull, 0xc3688c4f9774b905ull, 0x2695638c4fa9ac0full, 0x17f1d3a73197d794ull};
static const uint64_t GYc[6] = {0x0caa232946c5e7e1ull, 0xd03cc744a2888ae4ull, 0x00db18cb2c04b3edull, 0xfcf5e095d5d00af6ull, 0xa09e30ed741d8ae4ull, 0x08b3f481e3aaa0f1ull};

static G1Aff30 aff30(const HG1& p) { HFq x, y; g1_to_affine(p, &x, &y); G1Aff30 r; r.x = to30(x); r.y = to30(y); return r; }
static bool same_point(const G1Xyzz30& d, const HG1& expect) {
  HG1 got = xyzz30_is_identity(d) ? HG1::identity() : g1_from_xyzz(from30(d.x), from30(d.y), from30(d.zz), from30(d.zzz));
  HFq gx, gy, ex, ey;
  bool gi = !g1_to_affine(got, &gx, &gy), ei = !g1_to_affine(expect, &ex, &ey);
  if (gi || ei) return gi == ei;
  return gx == ex && gy == ey;
}
static bool bounds_ok(const G1Xyzz30& p) { return limbs_ok(p.x) && limbs_ok(p.y) && limbs_ok(p.zz) && limbs_ok(p.zzz); }

int main() {
  // --- field ---
  for (int it = 0; it < 20000; it++) {
    HFq a = rand_hfq(), b = rand_hfq();
    if (it == 0) a = HFq::zero();
    if (it == 1) { a = HFq::one().neg(); b = a; }
    Fq30 A = to30(a), B = to30(b);
    CHECK(from30(A) == a, "roundtrip");
    int ka = rnd() % 15, kb = rnd() % 15;
    Fq30 Al = lazy(A, ka), Bl = lazy(B, kb);
    Fq30 m = fq30_mul(Al, Bl);
    CHECK(limbs_ok(m), "mul limbs");
    CHECK(from30(m) == a * b, "mul");
    CHECK(from30(fq30_add(Al, Bl)) == a + b, "add");
    CHECK(from30(fq30_sub<16>(Al, Bl)) == a - b, "sub16");
    if (kb < 8) CHECK(from30(fq30_sub<8>(Al, Bl)) == a - b, "sub8");
    CHECK(limbs_ok(fq30_sub<16>(Al, Bl)), "sub limbs");
    CHECK(from30(fq30_neg<1>(B)) == b.neg(), "neg");
    CHECK(fq30_is_zero_mod_q(fq30_sub<8>(lazy(A, ka % 8), lazy(A, kb % 8))), "zero test true");
    CHECK(fq30_is_zero_mod_q(fq30_sub<16>(Al, Bl)) == (a == b), "zero test false");
  }
  // --- curve ---
  HG1 g = g1_from_affine(HFq::to_mont(GXc), HFq::to_mont(GYc));
  const int NP = 40;
  HG1 hp[NP];
  G1Aff30 ap[NP];
  for (int i = 0; i < NP; i++) { uint64_t k[1] = {rnd() | 1}; hp[i] = g1_mul_u64limbs(g, k, 1); ap[i] = aff30(hp[i]); }
  // long mixed-add chain with signs, bounds tracked
  G1Xyzz30 acc = xyzz30_identity();
  HG1 ref = HG1::identity();
  for (int it = 0; it < 400; it++) {
    int i = rnd() % NP; bool neg = rnd() & 1;
    HG1 p = hp[i]; if (neg) p.y = p.y.neg();
    xyzz30_madd(acc, ap[i], neg);
    ref = g1_add(ref, p);
    CHECK(bounds_ok(acc), "madd bounds");
    if (it % 50 == 0) CHECK(same_point(acc, ref), "madd chain");
  }
  CHECK(same_point(acc, ref), "madd chain end");
  // special cases: P + P, P - P, identity handling
  { G1Xyzz30 a = xyzz30_identity(); xyzz30_madd(a, ap[0], false); xyzz30_madd(a, ap[0], false); CHECK(same_point(a, g1_dbl(hp[0])), "madd doubling");
    xyzz30_madd(a, ap[0], false); CHECK(same_point(a, g1_add(g1_dbl(hp[0]), hp[0])), "3P");
    G1Xyzz30 b = xyzz30_identity(); xyzz30_madd(b, ap[1], false); xyzz30_madd(b, ap[1], true); CHECK(xyzz30_is_identity(b), "P - P");
    xyzz30_madd(b, ap[2], true); HG1 n2 = hp[2]; n2.y = n2.y.neg(); CHECK(same_point(b, n2), "identity + (-P)"); }
  // general add / dbl / mul_small chains
  { G1Xyzz30 a = xyzz30_identity(), b = xyzz30_identity(); HG1 ra = HG1::identity(), rb = HG1::identity();
    for (int i = 0; i < 5; i++) { xyzz30_madd(a, ap[i], false); ra = g1_add(ra, hp[i]); xyzz30_madd(b, ap[10 + i], i & 1); HG1 p = hp[10 + i]; if (i & 1) p.y = p.y.neg(); rb = g1_add(rb, p); }
    for (int it = 0; it < 60; it++) {
      xyzz30_add(a, b); ra = g1_add(ra, rb); CHECK(bounds_ok(a), "add bounds");
      if (it % 7 == 0) { xyzz30_dbl(b); rb = g1_dbl(rb); CHECK(bounds_ok(b), "dbl bounds"); }
      if (it % 11 == 0) { xyzz30_add(a, a); ra = g1_dbl(ra); }     // aliasing + equal points -> doubling branch
    }
    CHECK(same_point(a, ra), "add chain"); CHECK(same_point(b, rb), "dbl chain");
    G1Xyzz30 c = a; xyzz30_mul_small(c, 1000003u); uint64_t k[1] = {1000003}; 
    HFq ax, ay; g1_to_affine(ra, &ax, &ay); CHECK(same_point(c, g1_mul_u64limbs(g1_from_affine(ax, ay), k, 1)), "mul_small");
    G1Xyzz30 d = a, e = a; e.y = fq30_neg<4>(e.y); xyzz30_add(d, e); CHECK(xyzz30_is_identity(d), "P + (-P) general");
    G1Xyzz30 z = xyzz30_identity(); xyzz30_add(z, a); CHECK(same_point(z, ra), "identity + P"); xyzz30_add(z, xyzz30_identity()); CHECK(same_point(z, ra), "P + identity"); }
  printf("fq30 host check: %d failures\n", fails);
  return fails ? 1 : 0;
}
