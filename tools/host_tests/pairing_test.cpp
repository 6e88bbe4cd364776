// Host unit test of typlonk_b200/csrc/pairing.h: tower identities, q^2-Frobenius against a plain power, G2 group law,
// bilinearity and non-degeneracy of the pairing product check.  Prints "ok" lines; exit status = number of failures.
//   g++ -O2 -std=c++17 -o pairing_test pairing_test.cpp
#include <cstdio>
#include "../../typlonk_b200/csrc/pairing.h"
using namespace tph;
static int failures = 0;
#define CHECK(name, ...) do { bool ok_ = (__VA_ARGS__); printf("%s %s\n", ok_ ? "ok  " : "FAIL", name); if (!ok_) failures++; } while (0)
static uint64_t lcg = 0x9e3779b97f4a7c15ull;
static HFq rnd_fq() { HFq r; for (int i = 0; i < 6; i++) { lcg = lcg * 6364136223846793005ull + 1442695040888963407ull; r.v[i] = lcg; } r.v[5] &= 0x0fffffffffffffffull; return r; }
static Fq2 rnd2() { return {rnd_fq(), rnd_fq()}; }
static Fq12 rnd12() { return {{rnd2(), rnd2(), rnd2()}, {rnd2(), rnd2(), rnd2()}}; }
int main() {
  Fq12 x = rnd12(), y = rnd12(), z = rnd12();
  CHECK("fq12 associativity", (x * y) * z == x * (y * z));
  CHECK("fq12 sqr == mul", x.sqr() == x * x);
  CHECK("fq12 inv", (x * x.inv()).is_one());
  Fq6 s = x.c0;
  CHECK("fq6 inv", s * s.inv() == Fq6::one());
  CHECK("fq6 mul_v", s.mul_v() == s * Fq6{Fq2::zero(), Fq2::one(), Fq2::zero()});
  // w^2 == v
  Fq12 w = {Fq6::zero(), Fq6::one()};
  CHECK("w^2 == v", w * w == Fq12{{Fq2::zero(), Fq2::one(), Fq2::zero()}, Fq6::zero()});
  // q^2 as limbs
  uint64_t q2[12] = {0};
  for (int i = 0; i < 6; i++) { uint64_t c = 0; for (int j = 0; j < 6; j++) { u128 t = (u128)FQ_PARAMS.mod[i] * FQ_PARAMS.mod[j] + q2[i + j] + c; q2[i + j] = (uint64_t)t; c = (uint64_t)(t >> 64); } q2[i + 6] += c; }
  CHECK("frob2 == x^(q^2)", x.frob2() == x.pow(q2, 12));
  uint64_t q6[36] = {0};  // q^6 = (q^2)^3
  { uint64_t q4[24] = {0}; for (int i = 0; i < 12; i++) { uint64_t c = 0; for (int j = 0; j < 12; j++) { u128 t = (u128)q2[i] * q2[j] + q4[i + j] + c; q4[i + j] = (uint64_t)t; c = (uint64_t)(t >> 64); } q4[i + 12] += c; }
    for (int i = 0; i < 24; i++) { uint64_t c = 0; for (int j = 0; j < 12; j++) { u128 t = (u128)q4[i] * q2[j] + q6[i + j] + c; q6[i + j] = (uint64_t)t; c = (uint64_t)(t >> 64); } q6[i + 12] += c; } }
  CHECK("conj == x^(q^6)", x.conj() == x.pow(q6, 36));
  CHECK("frob1 == x^q", x.frob1() == x.pow(FQ_PARAMS.mod, 6));
  CHECK("frob1 o frob1 == frob2", x.frob1().frob1() == x.frob2());
  {
    Fq12 c = x.conj() * x.inv();   // easy part: lands in the cyclotomic subgroup
    c = c.frob2() * c;
    CHECK("cyclotomic: inverse == conjugate", (c * c.conj()).is_one());
    CHECK("cyclotomic_sqr == sqr", c.cyclotomic_sqr() == c.sqr());
    Fq12 c2 = c.sqr() * y.conj() * y.inv();
    c2 = c2.frob2() * c2;
    CHECK("cyclotomic_sqr == sqr (2)", c2.cyclotomic_sqr() == c2.sqr());
    uint64_t xabs[1] = {PAIRING_X_ABS};
    CHECK("cyclotomic_exp_x == conj(pow |x|)", c.cyclotomic_exp_x() == c.pow(xabs, 1).conj());
    uint64_t three[1] = {3};
    CHECK("fast final exponentiation == reference^3", final_exponentiation(x) == final_exponentiation_reference(x).pow(three, 1));
    CHECK("fast final exponentiation == reference^3 (2)", final_exponentiation(z) == final_exponentiation_reference(z).pow(three, 1));
  }
  G2Aff g2 = g2_generator();
  CHECK("g2 generator on twist", g2_on_curve(g2));
  G2Aff rg2 = g2_mul(g2, FR_PARAMS.mod, 4);
  CHECK("r * G2 == infinity", rg2.inf);
  uint64_t k5[1] = {5}, k7[1] = {7}, k35[1] = {35};
  G2Aff g5 = g2_mul(g2, k5, 1), g35 = g2_mul(g5, k7, 1), g35b = g2_mul(g2, k35, 1);
  CHECK("g2 scalar mul composes", !g35.inf && g35.x == g35b.x && g35.y == g35b.y && g2_on_curve(g35));
  G1Aff g1 = {HFq::to_mont(G1_GEN_X), HFq::to_mont(G1_GEN_Y), false};
  CHECK("g1 generator on curve", g1aff_on_curve(g1));
  G2Prepared pg2 = g2_prepare(g2), pg5 = g2_prepare(g5), pg35 = g2_prepare(g35);
  CHECK("68 line steps", pg2.lines.size() == 68);
  const G2Prepared* q1[1] = {&pg2};
  Fq12 e = final_exponentiation(miller_loop(&g1, q1, 1));
  CHECK("e(G1, G2) != 1", !e.is_one());
  CHECK("e(G1, G2)^r == 1", e.pow(FR_PARAMS.mod, 4).is_one());
  // e(5 G1, G2) == e(G1, 5 G2) == e^5
  HG1 j1 = g1_from_aff(g1);
  G1Aff a5 = g1_to_aff(g1_mul_u64limbs(j1, k5, 1)), a7 = g1_to_aff(g1_mul_u64limbs(j1, k7, 1));
  Fq12 e5a = final_exponentiation(miller_loop(&a5, q1, 1));
  const G2Prepared* q5[1] = {&pg5};
  Fq12 e5b = final_exponentiation(miller_loop(&g1, q5, 1));
  CHECK("e(5P, Q) == e(P, Q)^5", e5a == e.pow(k5, 1));
  CHECK("e(P, 5Q) == e(P, Q)^5", e5b == e.pow(k5, 1));
  // product check: e(7P, 5Q) * e(-P, 35Q) == 1
  G1Aff neg1 = g1; neg1.y = neg1.y.neg();
  G1Aff ps[2] = {a7, neg1};
  const G2Prepared* qs[2] = {&pg5, &pg35};
  CHECK("e(7P, 5Q) e(-P, 35Q) == 1", pairing_product_is_one(ps, qs, 2));
  G1Aff ps2[2] = {a7, g1};
  CHECK("e(7P, 5Q) e(P, 35Q) != 1", !pairing_product_is_one(ps2, qs, 2));
  G1Aff inf = {HFq::zero(), HFq::one(), true};
  G1Aff ps3[2] = {inf, inf};
  CHECK("infinity pairs -> 1", pairing_product_is_one(ps3, qs, 2));
  printf("%d failures\n", failures);
  return failures;
}
