// KZG commit as a Pippenger G1 multi-scalar multiplication for sm_100a.
//
// Replaces `KzgScheme::evaluate_in_s` (kzg/src/lib.rs:41-54), which the reference computes as
// n independent double-and-add scalar multiplications folded with `Sum`.  The affine result of
// a group sum is unique, so the bucket method is bit-exact with it.
//
// Pipeline (all on the ctx stream):
//   1. k_msm_digits   : scalar Montgomery -> canonical, signed-digit windows of c bits; each
//                       non-zero digit takes a rank in its bucket with one atomicAdd
//                       (histogram and within-bucket rank in a single pass)
//   2. k_scan_*       : exclusive scan of the bucket histogram -> bucket offsets
//   3. k_msm_scatter  : counting-sort scatter of (point index | sign) and bucket key
//   4. k_msm_accumulate: load-balanced segmented accumulation -- every thread owns a fixed-size
//                       chunk of the sorted entries (not a bucket), adds SRS points in XYZZ mixed
//                       coordinates, writes complete runs straight to their bucket and emits
//                       head/tail partials for runs that cross chunk boundaries
//   5. k_msm_merge    : stitches the boundary partials
//   6. k_msm_bucket_reduce / k_msm_window_reduce: running-sum reduction of each window in
//                       parallel segments, then a shared-memory tree
//   7. host           : Horner over the <= 64 window sums + affine conversion (serial tail; one
//                       CPU thread is ~10x faster than one GPU thread at 381-bit arithmetic), and
//                       the all-gather of partial points when the MSM is sharded over GPUs.
#include <stdlib.h>

#include <utility>

#include "common.cuh"

namespace tp {

#define MSM_NONE 0x7fffffffu
#define MSM_IDENT 0x80000000u   // partial slot holds the identity (point not stored)

// Tunables (defaults chosen on B200; TP_MSM_CHUNK / TP_MSM_SEG / TP_MSM_C override for sweeps).
static unsigned env_uint(const char* name, unsigned dflt) {
  const char* v = getenv(name);
  if (!v || !*v) return dflt;
  long x = strtol(v, nullptr, 10);
  return x > 0 ? (unsigned)x : dflt;
}
static unsigned msm_chunk() { static unsigned v = env_uint("TP_MSM_CHUNK", 64); return v; }   // sorted entries per accumulate thread
static unsigned msm_seg() { static unsigned v = env_uint("TP_MSM_SEG", 16); return v; }        // buckets per bucket-reduce thread

struct MsmPlan {
  unsigned c;        // window bits
  unsigned nwin;     // number of windows
  unsigned nbuck;    // buckets per window = 2^(c-1)
};

static MsmPlan msm_plan(size_t n) {
  MsmPlan best = {0, 0, 0};
  double best_cost = 1e300;
  static unsigned force_c = env_uint("TP_MSM_C", 0);
  for (unsigned c = 2; c <= 21; c++) {
    if (force_c && c != force_c) continue;
    unsigned nwin = (255 + c - 1) / c;
    unsigned top_bits = 255 - (nwin - 1) * c;
    if (top_bits == c) nwin++;  // signed carry out of a full top window
    double nb = (double)(1u << (c - 1));
    double cost = (double)n * nwin * 10.0 + nwin * nb * 40.0 + nwin * 3000.0;
    if (top_bits + 6 < c && n > 4096) cost += 25e6;  // top window funnels ~n points into < nb/32 buckets
    if (cost < best_cost) {
      best_cost = cost;
      best = {c, nwin, 1u << (c - 1)};
    }
  }
  return best;
}

// ---- 1. digits + histogram/rank ----------------------------------------------------------
#define MSM_MAX_BATCH 16
struct MsmScalarSets {
  const Fr* p[MSM_MAX_BATCH];
};
// blockIdx.y = batch element; its windows are numbered b * nwin + w in every later stage.
__global__ void k_msm_digits(MsmScalarSets sets, size_t n, unsigned c, unsigned nwin, unsigned nbuck,
                             unsigned* __restrict__ hist, unsigned* __restrict__ keys, unsigned* __restrict__ ranks) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned wbase = blockIdx.y * nwin;
  Fr s = fr_from_mont(fr_load(sets.p[blockIdx.y] + i));
  unsigned carry = 0;
  for (unsigned w = 0; w < nwin; w++) {
    unsigned bit = w * c;
    unsigned raw = 0;
    if (bit < 256) {
      unsigned limb = bit >> 5, off = bit & 31;
      unsigned long long two = s.v[limb];
      if (limb + 1 < 8) two |= (unsigned long long)s.v[limb + 1] << 32;
      raw = (unsigned)(two >> off) & ((1u << c) - 1);
    }
    raw += carry;
    unsigned key = MSM_NONE, rank = 0;
    carry = 0;
    if (raw != 0) {
      unsigned mag = raw, neg = 0;
      if (raw > (1u << (c - 1))) {
        mag = (1u << c) - raw;
        neg = 1;
        carry = 1;
      }
      if (mag != 0) {
        unsigned k = (wbase + w) * nbuck + (mag - 1);
        rank = atomicAdd(&hist[k], 1u);
        key = k | (neg << 31);
      }
    }
    keys[(size_t)(wbase + w) * n + i] = key;
    ranks[(size_t)(wbase + w) * n + i] = rank;
  }
}

// ---- 2. exclusive scan of u32 (3 kernels) --------------------------------------------------
#define SCAN_BLOCK 1024
__global__ void k_scan_local(const unsigned* __restrict__ in, unsigned* __restrict__ out, unsigned* __restrict__ sums,
                             size_t n) {
  __shared__ unsigned s[SCAN_BLOCK];
  size_t g = (size_t)blockIdx.x * SCAN_BLOCK + threadIdx.x;
  unsigned v = g < n ? in[g] : 0;
  s[threadIdx.x] = v;
  __syncthreads();
  for (unsigned d = 1; d < SCAN_BLOCK; d <<= 1) {
    unsigned t = threadIdx.x >= d ? s[threadIdx.x - d] : 0;
    __syncthreads();
    s[threadIdx.x] += t;
    __syncthreads();
  }
  if (g < n) out[g] = s[threadIdx.x] - v;
  if (threadIdx.x == SCAN_BLOCK - 1) sums[blockIdx.x] = s[threadIdx.x];
}
__global__ void k_scan_sums(unsigned* sums, size_t nblocks) {
  // single block, serial over tiles of SCAN_BLOCK
  __shared__ unsigned s[SCAN_BLOCK];
  __shared__ unsigned carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (size_t base = 0; base < nblocks; base += SCAN_BLOCK) {
    size_t g = base + threadIdx.x;
    unsigned v = g < nblocks ? sums[g] : 0;
    s[threadIdx.x] = v;
    __syncthreads();
    for (unsigned d = 1; d < SCAN_BLOCK; d <<= 1) {
      unsigned t = threadIdx.x >= d ? s[threadIdx.x - d] : 0;
      __syncthreads();
      s[threadIdx.x] += t;
      __syncthreads();
    }
    if (g < nblocks) sums[g] = s[threadIdx.x] - v + carry;
    __syncthreads();
    if (threadIdx.x == SCAN_BLOCK - 1) carry += s[threadIdx.x];
    __syncthreads();
  }
}
__global__ void k_scan_add(unsigned* out, const unsigned* sums, size_t n) {
  size_t g = (size_t)blockIdx.x * SCAN_BLOCK + threadIdx.x;
  if (g < n) out[g] += sums[blockIdx.x];
}

// ---- 3. scatter ----------------------------------------------------------------------------
__global__ void k_msm_scatter(const unsigned* __restrict__ keys, const unsigned* __restrict__ ranks,
                              const unsigned* __restrict__ offsets, size_t n, size_t total,
                              unsigned* __restrict__ sorted_idx, unsigned* __restrict__ sorted_key) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  unsigned key = keys[e];
  if ((key & MSM_NONE) == MSM_NONE) return;
  unsigned k = key & MSM_NONE;
  unsigned pos = offsets[k] + ranks[e];
  unsigned i = (unsigned)(e % n);
  sorted_idx[pos] = i | (key & 0x80000000u);
  sorted_key[pos] = k;
}

// ---- 4. segmented accumulation ---------------------------------------------------------------
// part_keys[2t], part_keys[2t+1]: keys of the head / tail partial of chunk t (MSM_NONE if absent)
#ifndef TP_ACC_MIN_BLOCKS
#define TP_ACC_MIN_BLOCKS 1
#endif
__global__ void __launch_bounds__(128, TP_ACC_MIN_BLOCKS) k_msm_accumulate(const G1Affine* __restrict__ bases,
                                                        const unsigned* __restrict__ sorted_idx,
                                                        const unsigned* __restrict__ sorted_key, unsigned m_total,
                                                        G1Xyzz* __restrict__ buckets, unsigned* __restrict__ part_keys,
                                                        G1Xyzz* __restrict__ part_pts, unsigned chunk) {
  unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned start = t * chunk;
  if (start >= m_total) return;
  unsigned end = start + chunk < m_total ? start + chunk : m_total;
  G1Xyzz acc = xyzz_identity();
  unsigned cur = sorted_key[start];
  bool is_first_run = true;
  for (unsigned e = start; e < end; e++) {
    unsigned key = sorted_key[e];
    if (key != cur) {
      if (is_first_run) {
        part_keys[2 * t] = cur;
        xyzz_store(part_pts + 2 * t, acc);
        is_first_run = false;
      } else {
        xyzz_store(buckets + cur, acc);  // complete interior run: exclusive owner of the bucket
      }
      acc = xyzz_identity();
      cur = key;
    }
    unsigned idx = sorted_idx[e];
    G1Affine p = affine_load(bases + (idx & 0x7fffffffu));
    if (!affine_is_identity(p)) xyzz_madd(acc, p, (idx >> 31) != 0);
  }
  if (is_first_run) {
    part_keys[2 * t] = cur;
    xyzz_store(part_pts + 2 * t, acc);
    part_keys[2 * t + 1] = cur | MSM_IDENT;  // single-run chunk: empty tail partial
  } else {
    part_keys[2 * t + 1] = cur;
    xyzz_store(part_pts + 2 * t + 1, acc);
  }
}

// ---- 5. boundary merge -----------------------------------------------------------------------
// The partial list (2 slots per chunk, keys non-decreasing) is reduced level by level: every
// thread takes MERGE_B consecutive slots, sums runs of equal key with general XYZZ additions,
// writes runs that lie strictly inside its block to their bucket (it is their only owner) and
// emits a head / tail partial for the next level.  Run length -- i.e. bucket size -- only costs
// extra levels of ~MERGE_B additions, so a bucket holding millions of points (small top window,
// repeated scalars) is as parallel as a uniform one.  The last level is one thread.
#define MERGE_B 16
__global__ void __launch_bounds__(128) k_msm_merge_level(const unsigned* __restrict__ keys_in,
                                                         const G1Xyzz* __restrict__ pts_in, unsigned nslots,
                                                         G1Xyzz* __restrict__ buckets, unsigned* __restrict__ keys_out,
                                                         G1Xyzz* __restrict__ pts_out, int final_level) {
  unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned start = final_level ? 0 : t * MERGE_B;
  if (start >= nslots || (final_level && t != 0)) return;
  unsigned end = final_level ? nslots : (start + MERGE_B < nslots ? start + MERGE_B : nslots);
  unsigned k0 = keys_in[start];
  unsigned cur = k0 & MSM_NONE;
  G1Xyzz acc = (k0 & MSM_IDENT) ? xyzz_identity() : xyzz_load(pts_in + start);
  bool is_first_run = true;
  for (unsigned e = start + 1; e < end; e++) {
    unsigned kraw = keys_in[e];
    unsigned key = kraw & MSM_NONE;
    if (key != cur) {
      if (is_first_run && !final_level) {
        keys_out[2 * t] = cur;
        xyzz_store(pts_out + 2 * t, acc);
        is_first_run = false;
      } else {
        xyzz_store(buckets + cur, acc);
      }
      cur = key;
      acc = (kraw & MSM_IDENT) ? xyzz_identity() : xyzz_load(pts_in + e);
    } else if (!(kraw & MSM_IDENT)) {
      G1Xyzz o = xyzz_load(pts_in + e);
      xyzz_add(acc, o);
    }
  }
  if (final_level) {
    xyzz_store(buckets + cur, acc);
  } else if (is_first_run) {
    keys_out[2 * t] = cur;
    xyzz_store(pts_out + 2 * t, acc);
    keys_out[2 * t + 1] = cur | MSM_IDENT;
  } else {
    keys_out[2 * t + 1] = cur;
    xyzz_store(pts_out + 2 * t + 1, acc);
  }
}

// ---- 6. bucket reduction ---------------------------------------------------------------------
// Each thread reduces MSM_SEG consecutive buckets of one window: sum_v v * B_v over its segment.
__global__ void __launch_bounds__(128) k_msm_bucket_reduce(const G1Xyzz* __restrict__ buckets,
                                                           const unsigned* __restrict__ hist, unsigned nbuck,
                                                           unsigned seg_len, unsigned segs_per_win, unsigned nwin,
                                                           G1Xyzz* __restrict__ seg_out) {
  unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= segs_per_win * nwin) return;
  unsigned w = t / segs_per_win, sidx = t % segs_per_win;
  unsigned lo = sidx * seg_len;  // bucket index b holds value v = b + 1
  G1Xyzz running = xyzz_identity(), total = xyzz_identity();
  for (int b = (int)seg_len - 1; b >= 0; b--) {
    unsigned k = w * nbuck + lo + b;
    if (hist[k] != 0) {
      G1Xyzz p = xyzz_load(buckets + k);
      xyzz_add(running, p);
    }
    xyzz_add(total, running);
  }
  // total = sum (b+1) B ; add lo * running
  if (lo != 0) {
    xyzz_mul_small(running, lo);
    xyzz_add(total, running);
  }
  xyzz_store(seg_out + t, total);
}
// One CTA per window: strided serial sum then shared-memory tree.
__global__ void __launch_bounds__(128) k_msm_window_reduce(const G1Xyzz* __restrict__ seg, unsigned segs_per_win,
                                                           G1Xyzz* __restrict__ win_out) {
  __shared__ G1Xyzz sh[128];
  unsigned w = blockIdx.x;
  G1Xyzz acc = xyzz_identity();
  for (unsigned s = threadIdx.x; s < segs_per_win; s += blockDim.x) {
    G1Xyzz p = xyzz_load(seg + (size_t)w * segs_per_win + s);
    xyzz_add(acc, p);
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (unsigned d = blockDim.x / 2; d > 0; d >>= 1) {
    if (threadIdx.x < d) {
      G1Xyzz a = sh[threadIdx.x];
      G1Xyzz b = sh[threadIdx.x + d];
      xyzz_add(a, b);
      sh[threadIdx.x] = a;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) xyzz_store(win_out + w, sh[0]);
}

// ---- host side ---------------------------------------------------------------------------------
void encode_g1(const tph::HG1& p, uint8_t out[TP_G1_BYTES]) {
  tph::HFq x, y;
  if (tph::g1_to_affine(p, &x, &y)) {
    memcpy(out, x.v, 48);
    memcpy(out + 48, y.v, 48);
    out[96] = 0;
  } else {
    tph::HFq zero = tph::HFq::zero(), one = tph::HFq::one();
    memcpy(out, zero.v, 48);
    memcpy(out + 48, one.v, 48);
    out[96] = 1;
  }
}

static int exclusive_scan_u32(tp_ctx* ctx, const unsigned* in, unsigned* out, size_t n) {
  size_t nblocks = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
  TP_TRY(ensure(ctx, ctx->msm_blocksums, nblocks * sizeof(unsigned)));
  unsigned* sums = (unsigned*)ctx->msm_blocksums.p;
  k_scan_local<<<(unsigned)nblocks, SCAN_BLOCK, 0, ctx->stream>>>(in, out, sums, n);
  TP_LAUNCH(ctx, "k_scan_local");
  k_scan_sums<<<1, SCAN_BLOCK, 0, ctx->stream>>>(sums, nblocks);
  TP_LAUNCH(ctx, "k_scan_sums");
  k_scan_add<<<(unsigned)nblocks, SCAN_BLOCK, 0, ctx->stream>>>(out, sums, n);
  TP_LAUNCH(ctx, "k_scan_add");
  return TP_OK;
}

// This rank's partial sums over `len` bases for `batch` scalar vectors at once: all vectors share
// one sort / accumulate / merge / reduce pipeline (window index = b * nwin + w), which amortises the
// latency-bound reduction tail and the launch overhead over the batch.
static int msm_local(tp_ctx* ctx, const G1Affine* bases, const Fr* const* scalars, int batch, size_t len,
                     tph::HG1* results) {
  for (int b = 0; b < batch; b++) results[b] = tph::HG1::identity();
  if (len == 0 || batch == 0) return TP_OK;
  if (batch > MSM_MAX_BATCH) return fail(ctx, TP_ERR_INVALID_ARG, "msm: batch too large");
  if (len >= ((size_t)1 << 27)) return fail(ctx, TP_ERR_INVALID_ARG, "msm: more than 2^27 points per call");
  MsmPlan pl = msm_plan(len);
  if ((size_t)pl.nwin * len * batch >= ((size_t)1 << 31) || (size_t)pl.nwin * pl.nbuck * batch >= ((size_t)1 << 30)) {
    if (batch == 1) return fail(ctx, TP_ERR_INVALID_ARG, "msm: too many window entries");
    int half = batch / 2;  // split the batch until the entry count fits 31 bits
    TP_TRY(msm_local(ctx, bases, scalars, half, len, results));
    return msm_local(ctx, bases, scalars + half, batch - half, len, results + half);
  }
  const unsigned nwin_total = pl.nwin * batch;
  size_t nkeys = (size_t)nwin_total * pl.nbuck;
  size_t total = (size_t)nwin_total * len;
  TP_TRY(ensure(ctx, ctx->msm_hist, (nkeys + 1) * sizeof(unsigned)));
  TP_TRY(ensure(ctx, ctx->msm_offsets, (nkeys + 1) * sizeof(unsigned)));
  TP_TRY(ensure(ctx, ctx->msm_keys, total * sizeof(unsigned)));
  TP_TRY(ensure(ctx, ctx->msm_ranks, total * sizeof(unsigned)));
  TP_TRY(ensure(ctx, ctx->msm_sorted, total * sizeof(unsigned)));
  TP_TRY(ensure(ctx, ctx->msm_sorted_keys, total * sizeof(unsigned)));
  TP_TRY(ensure(ctx, ctx->msm_buckets, nkeys * sizeof(G1Xyzz)));
  unsigned* hist = (unsigned*)ctx->msm_hist.p;
  unsigned* offsets = (unsigned*)ctx->msm_offsets.p;
  unsigned* keys = (unsigned*)ctx->msm_keys.p;
  unsigned* ranks = (unsigned*)ctx->msm_ranks.p;
  unsigned* sorted = (unsigned*)ctx->msm_sorted.p;
  unsigned* sorted_keys = (unsigned*)ctx->msm_sorted_keys.p;
  G1Xyzz* buckets = (G1Xyzz*)ctx->msm_buckets.p;
  unsigned m_total = 0;
  {
    ProfScope prof(ctx, TP_PHASE_MSM_SORT);
    TP_CUDA_OK(ctx, cudaMemsetAsync(hist, 0, (nkeys + 1) * sizeof(unsigned), ctx->stream));
    MsmScalarSets sets;
    for (int b = 0; b < MSM_MAX_BATCH; b++) sets.p[b] = scalars[b < batch ? b : 0];
    dim3 grid((unsigned)((len + 255) / 256), (unsigned)batch);
    k_msm_digits<<<grid, 256, 0, ctx->stream>>>(sets, len, pl.c, pl.nwin, pl.nbuck, hist, keys, ranks);
    TP_LAUNCH(ctx, "k_msm_digits");
    TP_TRY(exclusive_scan_u32(ctx, hist, offsets, nkeys + 1));
    k_msm_scatter<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(keys, ranks, offsets, len, total, sorted,
                                                                           sorted_keys);
    TP_LAUNCH(ctx, "k_msm_scatter");
    TP_CUDA_OK(ctx, cudaMemcpyAsync(ctx->pinned, offsets + nkeys, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    TP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    m_total = *(unsigned*)ctx->pinned;
  }
  if (m_total == 0) return TP_OK;  // all scalars zero
  const unsigned chunk = msm_chunk();
  unsigned nchunks = (m_total + chunk - 1) / chunk;
  // first half: accumulate's partials; second half: ping-pong space for the merge levels
  TP_TRY(ensure(ctx, ctx->msm_part_keys, ((size_t)2 * nchunks + (size_t)nchunks / 4 + 64) * sizeof(unsigned)));
  TP_TRY(ensure(ctx, ctx->msm_part_pts, ((size_t)2 * nchunks + (size_t)nchunks / 4 + 64) * sizeof(G1Xyzz)));
  {
    ProfScope prof(ctx, TP_PHASE_MSM_ACCUM);
    k_msm_accumulate<<<(nchunks + 127) / 128, 128, 0, ctx->stream>>>(bases, sorted, sorted_keys, m_total, buckets,
                                                                     (unsigned*)ctx->msm_part_keys.p,
                                                                     (G1Xyzz*)ctx->msm_part_pts.p, chunk);
    TP_LAUNCH(ctx, "k_msm_accumulate");
  }
  unsigned seg_len = pl.nbuck < msm_seg() ? pl.nbuck : msm_seg();
  unsigned segs_per_win = pl.nbuck / seg_len;
  TP_TRY(ensure(ctx, ctx->msm_seg, (size_t)segs_per_win * nwin_total * sizeof(G1Xyzz)));
  TP_TRY(ensure(ctx, ctx->msm_winsums, (size_t)nwin_total * sizeof(G1Xyzz)));
  {
    ProfScope prof(ctx, TP_PHASE_MSM_REDUCE);
    {
      unsigned* ka = (unsigned*)ctx->msm_part_keys.p;
      G1Xyzz* pa = (G1Xyzz*)ctx->msm_part_pts.p;
      unsigned* kb = ka + 2 * (size_t)nchunks;
      G1Xyzz* pb = pa + 2 * (size_t)nchunks;
      unsigned nslots = 2 * nchunks;
      while (nslots > 2 * MERGE_B) {
        unsigned nth = (nslots + MERGE_B - 1) / MERGE_B;
        k_msm_merge_level<<<(nth + 127) / 128, 128, 0, ctx->stream>>>(ka, pa, nslots, buckets, kb, pb, 0);
        TP_LAUNCH(ctx, "k_msm_merge_level");
        std::swap(ka, kb);
        std::swap(pa, pb);
        nslots = 2 * nth;
      }
      k_msm_merge_level<<<1, 32, 0, ctx->stream>>>(ka, pa, nslots, buckets, kb, pb, 1);
      TP_LAUNCH(ctx, "k_msm_merge_level");
    }
    unsigned nthreads = segs_per_win * nwin_total;
    k_msm_bucket_reduce<<<(nthreads + 127) / 128, 128, 0, ctx->stream>>>(buckets, hist, pl.nbuck, seg_len, segs_per_win,
                                                                         nwin_total, (G1Xyzz*)ctx->msm_seg.p);
    TP_LAUNCH(ctx, "k_msm_bucket_reduce");
    k_msm_window_reduce<<<nwin_total, 128, 0, ctx->stream>>>((G1Xyzz*)ctx->msm_seg.p, segs_per_win,
                                                             (G1Xyzz*)ctx->msm_winsums.p);
    TP_LAUNCH(ctx, "k_msm_window_reduce");
  }
  if ((size_t)nwin_total * sizeof(G1Xyzz) > ctx->pinned_cap) return fail(ctx, TP_ERR_INVALID_ARG, "msm: staging buffer too small");
  TP_CUDA_OK(ctx, cudaMemcpyAsync(ctx->pinned, ctx->msm_winsums.p, (size_t)nwin_total * sizeof(G1Xyzz),
                                  cudaMemcpyDeviceToHost, ctx->stream));
  TP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  // serial tail on the host: result_b = sum_w 2^(c w) * W_{b,w}
  const uint8_t* ws = (const uint8_t*)ctx->pinned;
  for (int b = 0; b < batch; b++) {
    tph::HG1 acc = tph::HG1::identity();
    for (int w = (int)pl.nwin - 1; w >= 0; w--) {
      for (unsigned d = 0; d < pl.c; d++) acc = tph::g1_dbl(acc);
      const uint8_t* q = ws + ((size_t)b * pl.nwin + w) * 192;
      tph::HFq x, y, zz, zzz;
      memcpy(x.v, q, 48);
      memcpy(y.v, q + 48, 48);
      memcpy(zz.v, q + 96, 48);
      memcpy(zzz.v, q + 144, 48);
      acc = tph::g1_add(acc, tph::g1_from_xyzz(x, y, zz, zzz));
    }
    results[b] = acc;
  }
  return TP_OK;
}

int msm_batch_dev(tp_ctx* ctx, const tp_srs* srs, const Fr* const* scalars_dev, int batch, size_t len,
                  uint8_t (*out)[TP_G1_BYTES]) {
  if (len > srs->len) return fail(ctx, TP_ERR_SRS_TOO_SHORT, "commit: polynomial longer than the SRS");
  if (batch <= 0 || batch > MSM_MAX_BATCH) return fail(ctx, TP_ERR_INVALID_ARG, "msm: bad batch size");
  ProfScope prof(ctx, TP_PHASE_MSM_TOTAL);
  tph::HG1 res[MSM_MAX_BATCH];
  if (ctx->world <= 1) {
    TP_TRY(msm_local(ctx, srs->g1, scalars_dev, batch, len, res));
  } else {
    // contiguous point-range shard; every rank holds the full SRS and scalar vectors
    size_t per = (len + ctx->world - 1) / ctx->world;
    size_t first = per * ctx->rank;
    size_t cnt = first >= len ? 0 : (len - first < per ? len - first : per);
    const Fr* shifted[MSM_MAX_BATCH];
    for (int b = 0; b < batch; b++) shifted[b] = scalars_dev[b] + first;
    tph::HG1 part[MSM_MAX_BATCH];
    TP_TRY(msm_local(ctx, srs->g1 + first, shifted, batch, cnt, part));
    const size_t per_rank = (size_t)144 * batch;
    std::vector<uint8_t> send(per_rank), recv(per_rank * ctx->world);
    for (int b = 0; b < batch; b++) {
      memcpy(&send[(size_t)b * 144], part[b].x.v, 48);
      memcpy(&send[(size_t)b * 144 + 48], part[b].y.v, 48);
      memcpy(&send[(size_t)b * 144 + 96], part[b].z.v, 48);
    }
    if (!ctx->allgather || ctx->allgather(ctx->allgather_user, send.data(), recv.data(), per_rank) != 0)
      return fail(ctx, TP_ERR_COLLECTIVE, "msm: all-gather of partial points failed");
    for (int b = 0; b < batch; b++) {
      res[b] = tph::HG1::identity();
      for (int r = 0; r < ctx->world; r++) {
        const uint8_t* q = recv.data() + (size_t)r * per_rank + (size_t)b * 144;
        tph::HG1 p;
        memcpy(p.x.v, q, 48);
        memcpy(p.y.v, q + 48, 48);
        memcpy(p.z.v, q + 96, 48);
        res[b] = tph::g1_add(res[b], p);
      }
    }
  }
  for (int b = 0; b < batch; b++) encode_g1(res[b], out[b]);
  return TP_OK;
}

int msm_dev(tp_ctx* ctx, const tp_srs* srs, const Fr* scalars_dev, size_t len, uint8_t out[TP_G1_BYTES]) {
  const Fr* one[1] = {scalars_dev};
  return msm_batch_dev(ctx, srs, one, 1, len, (uint8_t(*)[TP_G1_BYTES])out);
}

}  // namespace tp
