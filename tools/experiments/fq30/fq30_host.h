// Host-side conversions between the reduced-radix internal form of fq30.cuh (x * 2^390 mod q as
// 13 x 30-bit limbs, lazily reduced) and the standard Montgomery form (x * 2^384 mod q, 6 x u64)
// that crosses the C ABI.  Product code.
#pragma once
#include "fq30.cuh"
#include "host_field.h"

namespace tph {

// integer value of the limbs (< 2^390, possibly a few multiples of q too large) -> HFq holding
// x = value * 2^-390 mod q in standard Montgomery form
static inline HFq hfq_from_fq30(const uint32_t l[13]) {
  uint64_t w[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 13; i++) {
    int bit = 30 * i;
    uint64_t v = l[i];
    w[bit >> 6] |= v << (bit & 63);
    if ((bit & 63) > 34 && (bit >> 6) + 1 < 7) w[(bit >> 6) + 1] |= v >> (64 - (bit & 63));
  }
  // limbs may overlap only through carries already resolved on the device (limbs < 2^30), so OR is exact
  HFq lo = HFq::to_mont(w);                 // (value mod 2^384) * 2^384
  static const HFq c384 = HFq::to_mont(FQ_PARAMS.one);  // element "2^384 mod q"
  HFq e = lo + HFq::from_u64(w[6]) * c384;  // element "value"
  static const uint64_t inv390_canon[6] = FQ30_INV390_CANON_U64;
  static const HFq inv390 = HFq::to_mont(inv390_canon);  // element "2^-390"
  return e * inv390;
}

// standard Montgomery HFq (x * 2^384) -> fully reduced internal limbs (x * 2^390 mod q)
static inline void fq30_from_hfq(const HFq& x_std, uint32_t out[13]) {
  HFq v = x_std;
  for (int i = 0; i < 6; i++) v = v.dbl();  // * 2^6, canonical limbs = x * 2^390 mod q
  for (int i = 0; i < 13; i++) {
    int bit = 30 * i;
    uint64_t lo = v.v[bit >> 6] >> (bit & 63);
    if ((bit & 63) > 34 && (bit >> 6) + 1 < 6) lo |= v.v[(bit >> 6) + 1] << (64 - (bit & 63));
    out[i] = (uint32_t)(lo & 0x3fffffffu);
  }
}

}  // namespace tph
