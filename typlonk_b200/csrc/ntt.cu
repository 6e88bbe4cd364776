// Radix-2 NTT / iNTT / coset NTT over BLS12-381 Fr for sm_100a.
//
// Replaces ark-poly 0.3.0 `Radix2EvaluationDomain::{fft,ifft}_in_place` behind
// `Evaluations::interpolate` / `evaluate_over_domain` (reference call sites
// plonk/src/proof.rs:50,106,115,125,128,337,415; plonk/src/builder.rs:85;
// permutation/src/lib.rs:171,188).  Natural order in, natural order out.
//
// Structure: decimation-in-time with the bit-reversal folded into the first pass's
// gather, ceil(log N / 8) passes over HBM.  Each CTA stages a tile of 2^(k+cw) <= 1024
// elements (32 KB) in shared memory as two uint4 planes, runs k butterfly stages on
// 2^cw independent columns, and writes back with 128-byte-contiguous accesses.  Twiddles
// come from one table omega_N^i, i in [0, N/2]; the inverse transform reads the same table
// mirrored (omega^-i = -omega^(N/2-i)) so no second table is kept.
#include "common.cuh"

namespace tp {

static const uint64_t ROOT_OF_UNITY_CANON[4] = {0x3829971f439f0d2bull, 0xb63683508c2280b9ull, 0xd09b681922c813b4ull,
                                                0x16a2a19edfe81f20ull};  // 7^((r-1)/2^32)

tph::HFr omega_for_log(unsigned log_n) {
  tph::HFr w = tph::HFr::to_mont(ROOT_OF_UNITY_CANON);
  for (unsigned i = log_n; i < 32; i++) w = w.sqr();
  return w;
}

// out[i] = base^(i * stride_exp) ... generic power table: out[i] = base^i for i < count.
__global__ void k_pow_table(Fr* out, Fr base, size_t count) {
  const int PER = 16;
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t start = t * PER;
  if (start >= count) return;
  Fr cur = fr_pow_u64(base, (unsigned long long)start);
  for (int i = 0; i < PER && start + i < count; i++) {
    fr_store(out + start + i, cur);
    cur = fr_mul(cur, base);
  }
}

__device__ __forceinline__ unsigned bitrev(unsigned x, unsigned bits) { return bits == 0 ? 0u : (__brev(x) >> (32 - bits)); }

struct NttPassArgs {
  const Fr* in;
  Fr* out;
  const Fr* tw;
  unsigned L, s0, k, cw;
  int first, inverse, last;
  Fr ninv;
  const Fr* coset_lo;  // nullptr when no coset
  const Fr* coset_hi;
};

#define NTT_COSET_LO_BITS 12

__device__ __forceinline__ Fr coset_power(const NttPassArgs& a, unsigned idx) {
  Fr lo = fr_load(a.coset_lo + (idx & ((1u << NTT_COSET_LO_BITS) - 1)));
  Fr hi = fr_load(a.coset_hi + (idx >> NTT_COSET_LO_BITS));
  return fr_mul(lo, hi);
}

__global__ void __launch_bounds__(512) k_ntt_pass(NttPassArgs a) {
  __shared__ uint4 s_lo[1024];
  __shared__ uint4 s_hi[1024];
  const unsigned T = 1u << (a.k + a.cw);
  const unsigned cmask = (1u << a.cw) - 1;
  const unsigned K = 1u << a.k;
  unsigned base = 0, lo_grp = 0;
  if (!a.first) {
    unsigned hi_idx = blockIdx.x >> (a.s0 - a.cw);
    lo_grp = blockIdx.x & ((1u << (a.s0 - a.cw)) - 1);
    base = (hi_idx << (a.s0 + a.k)) + (lo_grp << a.cw);
  }
  // ---- load -----------------------------------------------------------------------
  for (unsigned e = threadIdx.x; e < T; e += blockDim.x) {
    unsigned l = e & cmask, m = e >> a.cw;
    unsigned src, spos;
    if (a.first) {
      src = (m << (a.L - a.k)) + (blockIdx.x << a.cw) + l;
      spos = (l << a.k) + bitrev(m, a.k);
    } else {
      src = base + (m << a.s0) + l;
      spos = (l << a.k) + m;
    }
    Fr x = fr_load(a.in + src);
    if (a.first && a.coset_lo && !a.inverse) x = fr_mul(x, coset_power(a, src));
    s_lo[spos] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
    s_hi[spos] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
  }
  __syncthreads();
  // ---- butterflies ------------------------------------------------------------------
  const unsigned half_n = 1u << (a.L - 1);
  for (unsigned t = 1; t <= a.k; t++) {
    for (unsigned b = threadIdx.x; b < (T >> 1); b += blockDim.x) {
      unsigned col = b >> (a.k - 1);
      unsigned q = b & ((K >> 1) - 1);
      unsigned j = q & ((1u << (t - 1)) - 1);
      unsigned blk = q >> (t - 1);
      unsigned p0 = (col << a.k) + (blk << t) + j;
      unsigned p1 = p0 + (1u << (t - 1));
      unsigned s = a.s0 + t;
      unsigned lo_val = a.first ? 0u : ((lo_grp << a.cw) + col);
      unsigned idx = ((j << a.s0) + lo_val) << (a.L - s);
      Fr w = fr_load(a.tw + (a.inverse ? (half_n - idx) : idx));
      uint4 ul = s_lo[p0], uh = s_hi[p0], vl = s_lo[p1], vh = s_hi[p1];
      Fr u, v;
      u.v[0] = ul.x; u.v[1] = ul.y; u.v[2] = ul.z; u.v[3] = ul.w; u.v[4] = uh.x; u.v[5] = uh.y; u.v[6] = uh.z; u.v[7] = uh.w;
      v.v[0] = vl.x; v.v[1] = vl.y; v.v[2] = vl.z; v.v[3] = vl.w; v.v[4] = vh.x; v.v[5] = vh.y; v.v[6] = vh.z; v.v[7] = vh.w;
      Fr wv = fr_mul(w, v);
      Fr r0, r1;
      if (a.inverse) {
        r0 = fr_sub(u, wv);
        r1 = fr_add(u, wv);
      } else {
        r0 = fr_add(u, wv);
        r1 = fr_sub(u, wv);
      }
      s_lo[p0] = make_uint4(r0.v[0], r0.v[1], r0.v[2], r0.v[3]);
      s_hi[p0] = make_uint4(r0.v[4], r0.v[5], r0.v[6], r0.v[7]);
      s_lo[p1] = make_uint4(r1.v[0], r1.v[1], r1.v[2], r1.v[3]);
      s_hi[p1] = make_uint4(r1.v[4], r1.v[5], r1.v[6], r1.v[7]);
    }
    __syncthreads();
  }
  // ---- store ------------------------------------------------------------------------
  for (unsigned e = threadIdx.x; e < T; e += blockDim.x) {
    unsigned dst, spos;
    if (a.first) {
      unsigned l = e >> a.k, pos = e & (K - 1);
      unsigned bidx = bitrev((blockIdx.x << a.cw) + l, a.L - a.k);
      dst = (bidx << a.k) + pos;
      spos = e;
    } else {
      unsigned l = e & cmask, m = e >> a.cw;
      dst = base + (m << a.s0) + l;
      spos = (l << a.k) + m;
    }
    uint4 xl = s_lo[spos], xh = s_hi[spos];
    if (a.last && (a.inverse)) {
      Fr x;
      x.v[0] = xl.x; x.v[1] = xl.y; x.v[2] = xl.z; x.v[3] = xl.w; x.v[4] = xh.x; x.v[5] = xh.y; x.v[6] = xh.z; x.v[7] = xh.w;
      x = fr_mul(x, a.ninv);
      if (a.coset_lo) x = fr_mul(x, coset_power(a, dst));
      fr_store(a.out + dst, x);
    } else {
      uint4* o = reinterpret_cast<uint4*>(a.out + dst);
      o[0] = xl;
      o[1] = xh;
    }
  }
}

int ntt_get_twiddles(tp_ctx* ctx, unsigned log_n, const Fr** tw) {
  auto it = ctx->ntt_tables.find(log_n);
  if (it == ctx->ntt_tables.end()) {
    NttTables t;
    size_t count = ((size_t)1 << log_n) / 2 + 1;
    TP_CUDA_OK(ctx, cudaMalloc(&t.tw, count * sizeof(Fr)));
    Fr w = to_dev(omega_for_log(log_n));
    size_t threads = (count + 15) / 16;
    k_pow_table<<<(unsigned)((threads + 127) / 128), 128, 0, ctx->stream>>>(t.tw, w, count);
    TP_LAUNCH(ctx, "k_pow_table");
    it = ctx->ntt_tables.emplace(log_n, t).first;
  }
  *tw = it->second.tw;
  return TP_OK;
}

static int get_coset_table(tp_ctx* ctx, unsigned log_n, const tph::HFr& g, const Fr** lo, const Fr** hi) {
  for (auto& c : ctx->coset_tables) {
    if (c.log_n == log_n && memcmp(c.g, g.v, 32) == 0) {
      *lo = c.lo;
      *hi = c.hi;
      return TP_OK;
    }
  }
  CosetTable c;
  c.log_n = log_n;
  memcpy(c.g, g.v, 32);
  size_t nlo = (size_t)1 << NTT_COSET_LO_BITS;
  size_t nhi = (((size_t)1 << log_n) >> NTT_COSET_LO_BITS) + 1;
  TP_CUDA_OK(ctx, cudaMalloc(&c.lo, nlo * sizeof(Fr)));
  TP_CUDA_OK(ctx, cudaMalloc(&c.hi, nhi * sizeof(Fr)));
  k_pow_table<<<(unsigned)((nlo / 16 + 127) / 128), 128, 0, ctx->stream>>>(c.lo, to_dev(g), nlo);
  TP_LAUNCH(ctx, "k_pow_table");
  tph::HFr gh = g.pow_u64((uint64_t)1 << NTT_COSET_LO_BITS);
  k_pow_table<<<(unsigned)(((nhi + 15) / 16 + 127) / 128), 128, 0, ctx->stream>>>(c.hi, to_dev(gh), nhi);
  TP_LAUNCH(ctx, "k_pow_table");
  ctx->coset_tables.push_back(c);
  *lo = c.lo;
  *hi = c.hi;
  return TP_OK;
}

int ntt_dev(tp_ctx* ctx, const Fr* in, Fr* out, unsigned log_n, bool inverse, const uint64_t* coset) {
  if (log_n == 0) {
    if (in != out) TP_CUDA_OK(ctx, cudaMemcpyAsync(out, in, sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
    return TP_OK;
  }
  if (log_n > 28) return fail(ctx, TP_ERR_INVALID_ARG, "ntt: log_n > 28");
  ProfScope prof(ctx, TP_PHASE_NTT);
  const Fr* tw;
  TP_TRY(ntt_get_twiddles(ctx, log_n, &tw));
  const Fr *clo = nullptr, *chi = nullptr;
  if (coset) {
    tph::HFr g;
    memcpy(g.v, coset, 32);
    if (inverse) g = g.inv();
    TP_TRY(get_coset_table(ctx, log_n, g, &clo, &chi));
  }
  size_t n = (size_t)1 << log_n;
  // pass plan
  unsigned ks[8];
  int npass;
  if (log_n <= 10) {
    npass = 1;
    ks[0] = log_n;
  } else {
    npass = (log_n + 7) / 8;
    unsigned q = log_n / npass, r = log_n % npass;
    for (int i = 0; i < npass; i++) ks[i] = q + (i < (int)r ? 1 : 0);
  }
  Fr* scratch = nullptr;
  if (npass > 1 && in == out) {
    TP_TRY(ensure(ctx, ctx->ntt_scratch, n * sizeof(Fr)));
    scratch = (Fr*)ctx->ntt_scratch.p;
  }
  tph::HFr ninv = tph::HFr::from_u64((uint64_t)n).inv();
  const Fr* src = in;
  unsigned s0 = 0;
  for (int p = 0; p < npass; p++) {
    NttPassArgs a;
    a.in = src;
    a.out = (p == 0 && scratch) ? scratch : out;
    a.tw = tw;
    a.L = log_n;
    a.s0 = s0;
    a.k = ks[p];
    a.first = (p == 0);
    a.inverse = inverse ? 1 : 0;
    a.last = (p == npass - 1);
    a.ninv = to_dev(ninv);
    a.coset_lo = clo;
    a.coset_hi = chi;
    unsigned cw = (npass == 1) ? 0 : 2;
    if (a.first && log_n - a.k < cw) cw = log_n - a.k;
    a.cw = cw;
    unsigned T = 1u << (a.k + a.cw);
    unsigned threads = T / 2 < 32 ? 32 : T / 2;
    unsigned grid = (unsigned)(n >> (a.k + a.cw));
    k_ntt_pass<<<grid, threads, 0, ctx->stream>>>(a);
    TP_LAUNCH(ctx, "k_ntt_pass");
    src = a.out;
    s0 += a.k;
  }
  return TP_OK;
}

}  // namespace tp
