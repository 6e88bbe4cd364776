// Device memory layouts and load/store helpers for the reduced-radix (fq30.cuh) MSM domain.
// An Fq30 occupies 16 words (13 limbs + 3 zero pad) so every access is four aligned 128-bit
// transactions; affine SRS points are 128 B, XYZZ points 256 B.
#pragma once
#include "ec.cuh"
#include "fq30.cuh"

namespace tp {

struct alignas(16) Fq30Mem {
  uint32_t w[16];
};
struct alignas(16) G1Aff30Mem {
  Fq30Mem x, y;
};
struct alignas(16) G1Xyzz30Mem {
  Fq30Mem x, y, zz, zzz;
};

__device__ __forceinline__ Fq30 fq30_load(const Fq30Mem* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1], c = q[2], d = q[3];
  Fq30 r;
  r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
  r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
  r.l[8] = c.x; r.l[9] = c.y; r.l[10] = c.z; r.l[11] = c.w;
  r.l[12] = d.x;
  return r;
}
__device__ __forceinline__ void fq30_store(Fq30Mem* p, const Fq30& a) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(a.l[0], a.l[1], a.l[2], a.l[3]);
  q[1] = make_uint4(a.l[4], a.l[5], a.l[6], a.l[7]);
  q[2] = make_uint4(a.l[8], a.l[9], a.l[10], a.l[11]);
  q[3] = make_uint4(a.l[12], 0, 0, 0);
}
__device__ __forceinline__ G1Aff30 aff30_load(const G1Aff30Mem* p) {
  G1Aff30 r;
  r.x = fq30_load(&p->x);
  r.y = fq30_load(&p->y);
  return r;
}
__device__ __forceinline__ G1Xyzz30 xyzz30_load(const G1Xyzz30Mem* p) {
  G1Xyzz30 r;
  r.x = fq30_load(&p->x); r.y = fq30_load(&p->y); r.zz = fq30_load(&p->zz); r.zzz = fq30_load(&p->zzz);
  return r;
}
__device__ __forceinline__ void xyzz30_store(G1Xyzz30Mem* p, const G1Xyzz30& a) {
  fq30_store(&p->x, a.x); fq30_store(&p->y, a.y); fq30_store(&p->zz, a.zz); fq30_store(&p->zzz, a.zzz);
}

// standard Montgomery Fq (x * 2^384 mod q, 12 x 32 bits, fully reduced) -> internal x * 2^390 mod q
__device__ __forceinline__ Fq30 fq30_from_std(const Fq& s) {
  Fq v = s;
#pragma unroll
  for (int i = 0; i < 6; i++) v = fq_dbl(v);  // * 2^6 mod q
  Fq30 r;
#pragma unroll
  for (int i = 0; i < FQ30_L; i++) {
    const int bit = 30 * i, w = bit >> 5, off = bit & 31;
    unsigned long long two = v.v[w];
    if (w + 1 < 12) two |= (unsigned long long)v.v[w + 1] << 32;
    r.l[i] = (uint32_t)(two >> off) & FQ30_MASK;
  }
  return r;
}

}  // namespace tp
