// Host-side BLS12-381 pairing for the verifier (kzg/src/lib.rs:66-81 `KzgScheme::verify`, which the reference
// delegates to ark-ec 0.3 `Bls12_381::pairing`) and G2 for `Srs::g2` (kzg/src/srs.rs:25-28).  A proof is checked
// with 12 pairings over the two fixed G2 points of the SRS, so the design is: line coefficients of a G2 point are
// computed ONCE (affine steps, `G2Prepared`), several pairs share one Miller loop (one squaring per bit for the
// whole product) and one final exponentiation decides `prod e(P_i, Q_i) == 1`.
//
// Tower: Fq2 = Fq[u]/(u^2 + 1), Fq6 = Fq2[v]/(v^3 - xi), Fq12 = Fq6[w]/(w^2 - v), xi = 1 + u.
// Twist E': y^2 = x^3 + 4 xi, untwisted by (x, y) -> (x / w^2, y / w^3).
// Product code -- independent of oracle/.
#pragma once
#include <vector>

#include "host_field.h"
#include "pairing_consts.h"

namespace tph {

struct Fq2 {
  HFq a, b;  // a + b u
  static Fq2 zero() { return {HFq::zero(), HFq::zero()}; }
  static Fq2 one() { return {HFq::one(), HFq::zero()}; }
  bool is_zero() const { return a.is_zero() && b.is_zero(); }
  bool operator==(const Fq2& o) const { return a == o.a && b == o.b; }
  bool operator!=(const Fq2& o) const { return !(*this == o); }
  Fq2 operator+(const Fq2& o) const { return {a + o.a, b + o.b}; }
  Fq2 operator-(const Fq2& o) const { return {a - o.a, b - o.b}; }
  Fq2 neg() const { return {a.neg(), b.neg()}; }
  Fq2 dbl() const { return {a.dbl(), b.dbl()}; }
  Fq2 conj() const { return {a, b.neg()}; }
  Fq2 operator*(const Fq2& o) const {  // three base-field products
    HFq t0 = a * o.a, t1 = b * o.b;
    return {t0 - t1, (a + b) * (o.a + o.b) - t0 - t1};
  }
  Fq2 sqr() const {
    HFq ab = a * b;
    return {(a + b) * (a - b), ab.dbl()};
  }
  Fq2 scale(const HFq& k) const { return {a * k, b * k}; }
  Fq2 mul_xi() const { return {a - b, a + b}; }  // (a + b u)(1 + u)
  Fq2 inv() const {
    HFq n = (a.sqr() + b.sqr()).inv();
    return {a * n, (b * n).neg()};
  }
};

struct Fq6 {
  Fq2 c0, c1, c2;  // c0 + c1 v + c2 v^2
  static Fq6 zero() { return {Fq2::zero(), Fq2::zero(), Fq2::zero()}; }
  static Fq6 one() { return {Fq2::one(), Fq2::zero(), Fq2::zero()}; }
  bool operator==(const Fq6& o) const { return c0 == o.c0 && c1 == o.c1 && c2 == o.c2; }
  Fq6 operator+(const Fq6& o) const { return {c0 + o.c0, c1 + o.c1, c2 + o.c2}; }
  Fq6 operator-(const Fq6& o) const { return {c0 - o.c0, c1 - o.c1, c2 - o.c2}; }
  Fq6 neg() const { return {c0.neg(), c1.neg(), c2.neg()}; }
  Fq6 mul_v() const { return {c2.mul_xi(), c0, c1}; }
  Fq6 operator*(const Fq6& o) const {  // six Fq2 products
    Fq2 t0 = c0 * o.c0, t1 = c1 * o.c1, t2 = c2 * o.c2;
    Fq2 r0 = ((c1 + c2) * (o.c1 + o.c2) - t1 - t2).mul_xi() + t0;
    Fq2 r1 = (c0 + c1) * (o.c0 + o.c1) - t0 - t1 + t2.mul_xi();
    Fq2 r2 = (c0 + c2) * (o.c0 + o.c2) - t0 - t2 + t1;
    return {r0, r1, r2};
  }
  Fq6 inv() const {
    Fq2 d0 = c0.sqr() - (c1 * c2).mul_xi();
    Fq2 d1 = c2.sqr().mul_xi() - c0 * c1;
    Fq2 d2 = c1.sqr() - c0 * c2;
    Fq2 t = (c0 * d0 + (c2 * d1 + c1 * d2).mul_xi()).inv();
    return {d0 * t, d1 * t, d2 * t};
  }
};

struct Fq12 {
  Fq6 c0, c1;  // c0 + c1 w
  static Fq12 one() { return {Fq6::one(), Fq6::zero()}; }
  bool operator==(const Fq12& o) const { return c0 == o.c0 && c1 == o.c1; }
  bool is_one() const { return *this == one(); }
  Fq12 operator*(const Fq12& o) const {
    Fq6 t0 = c0 * o.c0, t1 = c1 * o.c1;
    return {t0 + t1.mul_v(), (c0 + c1) * (o.c0 + o.c1) - t0 - t1};
  }
  Fq12 sqr() const {
    Fq6 t = c0 * c1;
    return {(c0 + c1) * (c0 + c1.mul_v()) - t - t.mul_v(), t + t};
  }
  Fq12 conj() const { return {c0, c1.neg()}; }  // x -> x^(q^6)
  Fq12 inv() const {
    Fq6 t = (c0 * c0 - (c1 * c1).mul_v()).inv();
    return {c0 * t, (c1 * t).neg()};
  }
  // x -> x^(q^2): the coefficient of w^k (k = 0..5, Fq2-valued, fixed by this map) picks up gamma^k, gamma = w^(q^2 - 1) in Fq
  Fq12 frob2() const {
    HFq g1 = HFq::to_mont(PAIRING_FROB2_GAMMA);
    HFq g2 = g1 * g1, g3 = g2 * g1, g4 = g2 * g2, g5 = g4 * g1;
    return {{c0.c0, c0.c1.scale(g2), c0.c2.scale(g4)}, {c1.c0.scale(g1), c1.c1.scale(g3), c1.c2.scale(g5)}};
  }
  // x -> x^q: coefficient k of w^k is conjugated (the Fq2 Frobenius) and picks up w^(k (q - 1)) = gamma1^k,
  // gamma1 = xi^((q - 1) / 6) in Fq2, computed once.
  static const Fq2* frob1_gammas() {
    static const struct Table {
      Fq2 g[6];
      Table() {
        uint64_t e[6];  // (q - 1) / 6 by long division
        memcpy(e, FQ_PARAMS.mod, 48);
        e[0] -= 1;
        u128 rem = 0;
        for (int i = 5; i >= 0; i--) {
          u128 cur = (rem << 64) | e[i];
          e[i] = (uint64_t)(cur / 6);
          rem = cur % 6;
        }
        Fq2 xi = {HFq::one(), HFq::one()}, r = Fq2::one();
        for (int i = 5; i >= 0; i--)
          for (int b = 63; b >= 0; b--) {
            r = r.sqr();
            if ((e[i] >> b) & 1) r = r * xi;
          }
        g[0] = Fq2::one();
        for (int k = 1; k < 6; k++) g[k] = g[k - 1] * r;
      }
    } t;
    return t.g;
  }
  Fq12 frob1() const {
    const Fq2* g = frob1_gammas();
    return {{c0.c0.conj(), c0.c1.conj() * g[2], c0.c2.conj() * g[4]},
            {c1.c0.conj() * g[1], c1.c1.conj() * g[3], c1.c2.conj() * g[5]}};
  }
  // Squaring of an element of the cyclotomic subgroup (x^(q^6 + 1) has been divided out by the easy part of the final
  // exponentiation; Granger-Scott).  Seen as Fq4 = Fq2[s]/(s^2 - xi) pairs (c0.c0, c1.c1), (c1.c0, c0.c2),
  // (c0.c1, c1.c2), the square needs only the three Fq4 squarings: 9 Fq2 products instead of 18.
  static void fq4_sqr(const Fq2& a, const Fq2& b, Fq2* re, Fq2* im) {
    Fq2 ab = a * b;
    *re = (a + b) * (a + b.mul_xi()) - ab - ab.mul_xi();  // a^2 + xi b^2
    *im = ab.dbl();
  }
  Fq12 cyclotomic_sqr() const {
    Fq2 a0, a1, b0, b1, d0, d1;
    fq4_sqr(c0.c0, c1.c1, &a0, &a1);
    fq4_sqr(c1.c0, c0.c2, &b0, &b1);
    fq4_sqr(c0.c1, c1.c2, &d0, &d1);
    auto three_minus_two = [](const Fq2& t, const Fq2& z) { return (t - z).dbl() + t; };  // 3 t - 2 z
    auto three_plus_two = [](const Fq2& t, const Fq2& z) { return (t + z).dbl() + t; };   // 3 t + 2 z
    Fq12 r;
    r.c0.c0 = three_minus_two(a0, c0.c0);
    r.c1.c1 = three_plus_two(a1, c1.c1);
    r.c1.c0 = three_plus_two(d1.mul_xi(), c1.c0);
    r.c0.c2 = three_minus_two(d0, c0.c2);
    r.c0.c1 = three_minus_two(b0, c0.c1);
    r.c1.c2 = three_plus_two(b1, c1.c2);
    return r;
  }
  // x^(-|X|) = x^(curve parameter) for x in the cyclotomic subgroup (where the inverse is the conjugate)
  Fq12 cyclotomic_exp_x() const {
    Fq12 r = *this;
    for (int b = 62; b >= 0; b--) {  // PAIRING_X_ABS has bit 63 set
      r = r.cyclotomic_sqr();
      if ((PAIRING_X_ABS >> b) & 1) r = r * *this;
    }
    return r.conj();
  }
  Fq12 pow(const uint64_t* e, int n) const {
    Fq12 r = one();
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
};

// ---- G2: E'(Fq2), y^2 = x^3 + 4(1 + u) ---------------------------------------------------------------------------
struct G2Aff {
  Fq2 x, y;
  bool inf;
};
static inline G2Aff g2_generator() {
  return {{HFq::to_mont(G2_GEN_X0), HFq::to_mont(G2_GEN_X1)}, {HFq::to_mont(G2_GEN_Y0), HFq::to_mont(G2_GEN_Y1)}, false};
}
static inline bool g2_on_curve(const G2Aff& p) {
  if (p.inf) return true;
  Fq2 b = {HFq::from_u64(4), HFq::from_u64(4)};
  return p.y.sqr() == p.x.sqr() * p.x + b;
}
struct G2Jac {
  Fq2 x, y, z;  // z == 0 -> identity
};
static inline G2Jac g2_dbl(const G2Jac& p) {
  if (p.z.is_zero() || p.y.is_zero()) return {Fq2::one(), Fq2::one(), Fq2::zero()};
  Fq2 a = p.x.sqr(), b = p.y.sqr(), c = b.sqr();
  Fq2 d = ((p.x + b).sqr() - a - c).dbl();
  Fq2 e = a.dbl() + a, f = e.sqr();
  G2Jac r;
  r.x = f - d.dbl();
  r.y = e * (d - r.x) - c.dbl().dbl().dbl();
  r.z = (p.y * p.z).dbl();
  return r;
}
static inline G2Jac g2_add_affine(const G2Jac& p, const G2Aff& q) {
  if (q.inf) return p;
  if (p.z.is_zero()) return {q.x, q.y, Fq2::one()};
  Fq2 z1z1 = p.z.sqr();
  Fq2 u2 = q.x * z1z1, s2 = q.y * p.z * z1z1;
  if (p.x == u2) {
    if (p.y == s2) return g2_dbl(p);
    return {Fq2::one(), Fq2::one(), Fq2::zero()};
  }
  Fq2 h = u2 - p.x, hh = h.sqr(), i = hh.dbl().dbl(), j = h * i;
  Fq2 rr = (s2 - p.y).dbl(), v = p.x * i;
  G2Jac r;
  r.x = rr.sqr() - j - v.dbl();
  r.y = rr * (v - r.x) - (p.y * j).dbl();
  r.z = (p.z + h).sqr() - z1z1 - hh;
  return r;
}
static inline G2Aff g2_to_affine(const G2Jac& p) {
  if (p.z.is_zero()) return {Fq2::zero(), Fq2::one(), true};
  Fq2 zi = p.z.inv(), zi2 = zi.sqr();
  return {p.x * zi2, p.y * zi2 * zi, false};
}
// k: canonical (non-Montgomery) little-endian limbs
static inline G2Aff g2_mul(const G2Aff& p, const uint64_t* k, int n) {
  G2Jac r = {Fq2::one(), Fq2::one(), Fq2::zero()};
  for (int i = n - 1; i >= 0; i--)
    for (int b = 63; b >= 0; b--) {
      r = g2_dbl(r);
      if ((k[i] >> b) & 1) r = g2_add_affine(r, p);
    }
  return g2_to_affine(r);
}
// 193-byte ABI record: x.c0 | x.c1 | y.c0 | y.c1 (Montgomery limbs, 48 B each) | infinity flag
static inline void g2_encode(const G2Aff& p, uint8_t out[193]) {
  G2Aff q = p.inf ? G2Aff{Fq2::zero(), Fq2::one(), true} : p;
  memcpy(out, q.x.a.v, 48);
  memcpy(out + 48, q.x.b.v, 48);
  memcpy(out + 96, q.y.a.v, 48);
  memcpy(out + 144, q.y.b.v, 48);
  out[192] = q.inf ? 1 : 0;
}
static inline G2Aff g2_decode(const uint8_t in[193]) {
  G2Aff p;
  memcpy(p.x.a.v, in, 48);
  memcpy(p.x.b.v, in + 48, 48);
  memcpy(p.y.a.v, in + 96, 48);
  memcpy(p.y.b.v, in + 144, 48);
  p.inf = in[192] != 0;
  return p;
}

// ---- Miller loop over prepared G2 points ---------------------------------------------------------------------------
// A step through T (and Q) with slope s on the twist contributes, up to factors the final exponentiation removes,
//   l(P) = (s x_T - y_T) - s x_P v + y_P v w :  only (s, s x_T - y_T) depend on the G2 point.
struct G2Prepared {
  struct Line {
    Fq2 slope, c;
  };
  std::vector<Line> lines;
  bool inf = true;
};
static inline G2Prepared g2_prepare(const G2Aff& q) {
  G2Prepared out;
  out.inf = q.inf;
  if (q.inf) return out;
  Fq2 tx = q.x, ty = q.y;
  auto step = [&](const Fq2& s) { out.lines.push_back({s, s * tx - ty}); };  // recorded BEFORE T moves
  for (int b = 62; b >= 0; b--) {  // |x| has 64 bits, the top one starts T = Q
    Fq2 s = (tx.sqr().dbl() + tx.sqr()) * ty.dbl().inv();
    step(s);
    Fq2 nx = s.sqr() - tx.dbl();
    ty = s * (tx - nx) - ty;
    tx = nx;
    if ((PAIRING_X_ABS >> b) & 1) {
      Fq2 s2 = (ty - q.y) * (tx - q.x).inv();
      step(s2);
      Fq2 ax = s2.sqr() - tx - q.x;
      ty = s2 * (tx - ax) - ty;
      tx = ax;
    }
  }
  return out;
}
struct G1Aff {
  HFq x, y;
  bool inf;
};
static inline Fq12 line_at(const G2Prepared::Line& l, const G1Aff& p) {
  Fq12 r;
  r.c0 = {l.c, l.slope.scale(p.x).neg(), Fq2::zero()};
  r.c1 = {Fq2::zero(), {p.y, HFq::zero()}, Fq2::zero()};
  return r;
}
// prod_i f_{|x|, Q_i}(P_i), conjugated because the curve parameter is negative
static inline Fq12 miller_loop(const G1Aff* ps, const G2Prepared* const* qs, int count) {
  Fq12 f = Fq12::one();
  size_t idx = 0;
  for (int b = 62; b >= 0; b--) {
    f = f.sqr();
    for (int i = 0; i < count; i++)
      if (!ps[i].inf && !qs[i]->inf) f = f * line_at(qs[i]->lines[idx], ps[i]);
    idx++;
    if ((PAIRING_X_ABS >> b) & 1) {
      for (int i = 0; i < count; i++)
        if (!ps[i].inf && !qs[i]->inf) f = f * line_at(qs[i]->lines[idx], ps[i]);
      idx++;
    }
  }
  return f.conj();
}
// Plain square-and-multiply by (q^4 - q^2 + 1) / r: the definition the fast chain below is tested against.
static inline Fq12 final_exponentiation_reference(const Fq12& f) {
  Fq12 e = f.conj() * f.inv();  // ^(q^6 - 1)
  e = e.frob2() * e;            // ^(q^2 + 1)
  return e.pow(PAIRING_HARD_EXP, PAIRING_HARD_LIMBS);
}
// f^(3 (q^12 - 1) / r).  With x the (negative) curve parameter, 3 (q^4 - q^2 + 1) / r = (x - 1)^2 (x + q)(x^2 + q^2 - 1) + 3
// (checked as an integer identity in tools/gen_pairing_consts.py), so the hard part is five exponentiations by x
// (63 cyclotomic squarings + 5 products each), two Frobenius maps and a handful of products instead of a 1268-bit
// square-and-multiply.  The extra factor 3 is coprime to r: `== 1` and bilinearity are unaffected; the VALUE is the
// cube of final_exponentiation_reference's.
static inline Fq12 final_exponentiation(const Fq12& f) {
  Fq12 e = f.conj() * f.inv();  // ^(q^6 - 1)
  e = e.frob2() * e;            // ^(q^2 + 1): now in the cyclotomic subgroup, inverse = conjugate
  Fq12 t = e.cyclotomic_exp_x() * e.conj();              // e^(x - 1)
  t = t.cyclotomic_exp_x() * t.conj();                   // e^((x - 1)^2)
  t = t.cyclotomic_exp_x() * t.frob1();                  // ^(x + q)
  t = t.cyclotomic_exp_x().cyclotomic_exp_x() * t.frob2() * t.conj();  // ^(x^2 + q^2 - 1)
  return t * e.cyclotomic_sqr() * e;                     // * e^3
}
static inline bool pairing_product_is_one(const G1Aff* ps, const G2Prepared* const* qs, int count) {
  return final_exponentiation(miller_loop(ps, qs, count)).is_one();
}

// G1 affine helpers on top of the Jacobian group of host_field.h
static inline G1Aff g1aff_decode(const uint8_t in[97]) {
  G1Aff p;
  memcpy(p.x.v, in, 48);
  memcpy(p.y.v, in + 48, 48);
  p.inf = in[96] != 0;
  return p;
}
static inline HG1 g1_from_aff(const G1Aff& p) { return p.inf ? HG1::identity() : g1_from_affine(p.x, p.y); }
static inline G1Aff g1_to_aff(const HG1& p) {
  G1Aff r;
  r.inf = !g1_to_affine(p, &r.x, &r.y);
  if (r.inf) {
    r.x = HFq::zero();
    r.y = HFq::one();
  }
  return r;
}
static inline HG1 g1_neg(const HG1& p) { return {p.x, p.y.neg(), p.z}; }
static inline bool g1aff_on_curve(const G1Aff& p) { return p.inf || p.y.sqr() == p.x.sqr() * p.x + HFq::from_u64(4); }
// k: Montgomery Fr
static inline HG1 g1_mul_fr(const HG1& p, const HFr& k) {
  HFr c = k.from_mont();
  return g1_mul_u64limbs(p, c.v, 4);
}

}  // namespace tph
