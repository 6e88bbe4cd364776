// Elementwise and scan kernels of the prover (all HBM-streaming integer work over Fr):
//   * permutation grand product   -- permutation/src/proving.rs:7-31 (one inversion per cell in
//     the reference) as a prefix-product scan of the numerators and a suffix-product scan of the
//     denominators sharing their launches, with ONE inversion per proof
//   * opening polynomial          -- kzg/src/lib.rs:55-64: Horner evaluation and the division
//     by (X - z) are one linear recurrence q[k-1] = p[k] + z q[k], solved as a hierarchical scan
//   * gate check                  -- the `vanishes(line1)` assert, plonk/src/proof.rs:317-321
//   * quotient numerator on the 4n domain and division by X^n - 1 -- plonk/src/proof.rs:292-375
//     (the reference builds these with O(n^2) `naive_mul`)
//   * linear combination for r(X) -- plonk/src/proof.rs:376-439
//   * sigma / id tables           -- permutation/src/lib.rs:101-128
#include "common.cuh"

namespace tp {

// Element-wise kernels with many field products (the quotient numerator has ~30) are straight-line code far larger than
// the instruction cache; with TP_POLY_MUL_CALL they go through one shared copy of the multiplier (A/B in profiles/).
#ifdef TP_POLY_MUL_CALL
static __device__ __noinline__ Fr pmul(Fr a, Fr b) { return fr_mul(a, b); }
#else
__device__ __forceinline__ Fr pmul(const Fr& a, const Fr& b) { return fr_mul(a, b); }
#endif


#define EW_THREADS 256
static inline unsigned ew_grid(size_t n) { return (unsigned)((n + EW_THREADS - 1) / EW_THREADS); }

// =====================================================================================
// exclusive prefix-product scan (in place).  chunk per thread, recursive on chunk totals.
// =====================================================================================
#define SCAN_CH 32
#define SCAN_BASE 64

// two independent scans of the same length side by side (blockIdx.y picks the array): the upper levels of a scan are
// chains of small launches, so two cost what one costs
struct ScanPair {
  Fr* a[2];
  Fr* tot[2];
};
__global__ void k_mulscan2_serial(ScanPair p, size_t n) {   // exclusive
  if (threadIdx.x != 0) return;
  Fr* a = p.a[blockIdx.x];
  Fr acc = fr_one();
  for (size_t i = 0; i < n; i++) {
    Fr x = fr_load(a + i);
    fr_store(a + i, acc);
    acc = fr_mul(acc, x);
  }
}
__global__ void k_mulscan2_chunk(ScanPair p, size_t n) {    // exclusive
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t lo = t * SCAN_CH;
  if (lo >= n) return;
  Fr* a = p.a[blockIdx.y];
  size_t hi = lo + SCAN_CH < n ? lo + SCAN_CH : n;
  Fr acc = fr_one();
  for (size_t i = lo; i < hi; i++) {
    Fr x = fr_load(a + i);
    fr_store(a + i, acc);
    acc = fr_mul(acc, x);
  }
  fr_store(p.tot[blockIdx.y] + t, acc);
}
__global__ void k_mulscan2_apply(ScanPair p, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  size_t ch = i / SCAN_CH;
  if (ch == 0) return;
  Fr* a = p.a[blockIdx.y];
  fr_store(a + i, fr_mul(fr_load(a + i), fr_load(p.tot[blockIdx.y] + ch)));
}
// exclusive prefix products of a0[0..n) and a1[0..n), in place
static int mulscan2_exclusive(tp_ctx* ctx, Fr* a0, Fr* a1, size_t n, int level) {
  ScanPair p;
  p.a[0] = a0;
  p.a[1] = a1;
  p.tot[0] = p.tot[1] = nullptr;
  if (n <= SCAN_BASE) {
    k_mulscan2_serial<<<2, 32, 0, ctx->stream>>>(p, n);
    TP_LAUNCH(ctx, "k_mulscan2_serial");
    return TP_OK;
  }
  size_t nch = (n + SCAN_CH - 1) / SCAN_CH;
  TP_TRY(ensure(ctx, ctx->scan_tmp[level], 2 * nch * sizeof(Fr)));
  p.tot[0] = (Fr*)ctx->scan_tmp[level].p;
  p.tot[1] = p.tot[0] + nch;
  k_mulscan2_chunk<<<dim3(ew_grid(nch), 2), EW_THREADS, 0, ctx->stream>>>(p, n);
  TP_LAUNCH(ctx, "k_mulscan2_chunk");
  TP_TRY(mulscan2_exclusive(ctx, p.tot[0], p.tot[1], nch, level + 1));
  k_mulscan2_apply<<<dim3(ew_grid(n), 2), EW_THREADS, 0, ctx->stream>>>(p, n);
  TP_LAUNCH(ctx, "k_mulscan2_apply");
  return TP_OK;
}

// =====================================================================================
// grand product  (permutation/src/proving.rs:7-31)
// =====================================================================================
// z[0] = 1, z[j+1] = prod_{k<=j} num_k / den_k with num_k = prod_i (v_ik + beta id_ik + gamma), den_k likewise over
// sigma.  The reference inverts every denominator; here NOTHING is inverted per element:
//     1 / prod_{k<=j} den_k = Dtot^-1 * prod_{k>j} den_k,
// so z[j+1] = (prefix product of num up to j) * (suffix product of den after j) * Dtot^-1 -- two product scans that
// share their launches and ONE inversion per proof, done by the host between the up- and the down-sweep (it also
// replaces the zero-denominator flag: a zero denominator makes Dtot zero).  About 15 field products per row in all.
#define PERM_CH 16   // rows per thread
struct PermArgs {
  const Fr* v[3];
  const Fr* id[3];
  const Fr* sg[3];
  Fr beta, gamma;
  size_t n;
  Fr* pre;     // n: prefix product of num inside the row's chunk (inclusive)
  Fr* den;     // n: den of the row
  Fr* tot_num; // nch + 1: chunk totals of num, then a trailing one
  Fr* tot_den; // nch + 1: chunk totals of den in REVERSE chunk order, then a trailing one
  size_t nch;
};
__global__ void __launch_bounds__(128) k_perm_chunks(PermArgs a) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t > a.nch) return;
  if (t == a.nch) {   // the trailing ones: the exclusive scans then leave the grand totals there
    fr_store(a.tot_num + a.nch, fr_one());
    fr_store(a.tot_den + a.nch, fr_one());
    return;
  }
  const size_t lo = t * PERM_CH;
  const size_t hi = lo + PERM_CH < a.n ? lo + PERM_CH : a.n;
  Fr pn = fr_one(), pd = fr_one();
  for (size_t j = lo; j < hi; j++) {
    Fr num, den;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      Fr v = fr_add(fr_load(a.v[i] + j), a.gamma);
      Fr nu = fr_add(v, pmul(a.beta, fr_load(a.id[i] + j)));
      Fr de = fr_add(v, pmul(a.beta, fr_load(a.sg[i] + j)));
      num = i == 0 ? nu : pmul(num, nu);
      den = i == 0 ? de : pmul(den, de);
    }
    pn = j == lo ? num : pmul(pn, num);
    pd = j == lo ? den : pmul(pd, den);
    fr_store(a.pre + j, pn);
    fr_store(a.den + j, den);
  }
  fr_store(a.tot_num + t, pn);
  fr_store(a.tot_den + (a.nch - 1 - t), pd);
}
// carry_num[t] = product of num over the chunks before t; carry_den[nch-1-t] = product of den over the chunks after t
__global__ void __launch_bounds__(128) k_perm_finish(PermArgs a, Fr dtot_inv, Fr* __restrict__ z /* n + 1 */) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.nch) return;
  const size_t lo = t * PERM_CH;
  const size_t hi = lo + PERM_CH < a.n ? lo + PERM_CH : a.n;
  if (t == 0) fr_store(z, fr_one());
  Fr k = pmul(pmul(fr_load(a.tot_num + t), fr_load(a.tot_den + (a.nch - 1 - t))), dtot_inv);
  // walk the chunk backwards: k = (carries) * Dtot^-1 * product of the chunk's den after row j
  for (size_t j = hi; j-- > lo;) {
    fr_store(z + j + 1, pmul(fr_load(a.pre + j), k));
    if (j > lo) k = pmul(k, fr_load(a.den + j));
  }
}

int perm_grand_product_dev(tp_ctx* ctx, const Fr* const values[3], const Fr* const id[3], const Fr* const sigma[3],
                           size_t n, const Fr& beta, const Fr& gamma, Fr* out, bool* closes) {
  ProfScope prof(ctx, TP_PHASE_PERM);
  const size_t nch = (n + PERM_CH - 1) / PERM_CH;
  TP_TRY(ensure(ctx, ctx->misc[0], n * sizeof(Fr)));
  TP_TRY(ensure(ctx, ctx->misc[1], n * sizeof(Fr)));
  TP_TRY(ensure(ctx, ctx->misc[2], 2 * (nch + 1) * sizeof(Fr)));
  PermArgs a;
  for (int i = 0; i < 3; i++) {
    a.v[i] = values[i];
    a.id[i] = id[i];
    a.sg[i] = sigma[i];
  }
  a.beta = beta;
  a.gamma = gamma;
  a.n = n;
  a.nch = nch;
  a.pre = (Fr*)ctx->misc[0].p;
  a.den = (Fr*)ctx->misc[1].p;
  a.tot_num = (Fr*)ctx->misc[2].p;
  a.tot_den = a.tot_num + (nch + 1);
  k_perm_chunks<<<(unsigned)((nch + 1 + 127) / 128), 128, 0, ctx->stream>>>(a);
  TP_LAUNCH(ctx, "k_perm_chunks");
  TP_TRY(mulscan2_exclusive(ctx, a.tot_num, a.tot_den, nch + 1, 0));
  // the one inversion: Dtot = product of every denominator (the last slot of the reversed den scan)
  TP_CUDA_OK(ctx, cudaMemcpyAsync(ctx->pinned, a.tot_den + nch, sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
  TP_CUDA_OK(ctx, cudaMemcpyAsync((uint8_t*)ctx->pinned + 32, a.tot_num + nch, sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
  TP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  tph::HFr dtot, ntot;
  memcpy(dtot.v, ctx->pinned, 32);
  memcpy(ntot.v, (uint8_t*)ctx->pinned + 32, 32);
  if (dtot.is_zero()) return fail(ctx, TP_ERR_ZERO_DENOMINATOR, "permutation prove: zero denominator");
  const tph::HFr dinv = dtot.inv();
  if (closes) *closes = ntot * dinv == tph::HFr::one();   // out[n], the value the caller pops (proof.rs:120)
  k_perm_finish<<<(unsigned)((nch + 127) / 128), 128, 0, ctx->stream>>>(a, to_dev(dinv), out);
  TP_LAUNCH(ctx, "k_perm_finish");
  return TP_OK;
}

// =====================================================================================
// linear recurrence  q[k-1] = p[k] + z q[k]   (Horner + division by X - z)
// =====================================================================================
#define LIN_CH 32
#define LIN_BASE 64
#define LIN_MAX_BATCH 8
// Up to LIN_MAX_BATCH independent recurrences of the same length run side by side (blockIdx.y):
// the scans are latency-bound chains of small launches, so the prover's eight evaluations /
// openings at one Fiat-Shamir point cost what one costs.
struct LinBatch {
  const Fr* p[LIN_MAX_BATCH];
  Fr* q[LIN_MAX_BATCH];       // may be null: evaluation only
  Fr z[LIN_MAX_BATCH];
};

// h[t] = sum_{k in chunk t} p[k] z^(k - lo)
__global__ void k_linrec_reduce(LinBatch a, size_t n, Fr* h, size_t h_stride) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t lo = t * LIN_CH;
  if (lo >= n) return;
  const Fr* p = a.p[blockIdx.y];
  const Fr z = a.z[blockIdx.y];
  size_t hi = lo + LIN_CH < n ? lo + LIN_CH : n;
  Fr acc = fr_zero();
  for (size_t k = hi; k-- > lo;) acc = fr_add(fr_load(p + k), fr_mul(z, acc));
  fr_store(h + blockIdx.y * h_stride + t, acc);
}
// q[k-1] = p[k] + z q[k] inside chunk t, starting from carry = qprime[t] (q at index hi-1)
__global__ void k_linrec_apply(LinBatch a, size_t n, const Fr* qprime, size_t qp_stride) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t lo = t * LIN_CH;
  if (lo >= n) return;
  Fr* q = a.q[blockIdx.y];
  if (!q) return;
  const Fr* p = a.p[blockIdx.y];
  const Fr z = a.z[blockIdx.y];
  size_t hi = lo + LIN_CH < n ? lo + LIN_CH : n;
  Fr acc = fr_load(qprime + blockIdx.y * qp_stride + t);
  if (hi == n) fr_store(q + n - 1, fr_zero());
  for (size_t k = hi; k-- > lo;) {
    acc = fr_add(fr_load(p + k), fr_mul(z, acc));
    if (k >= 1) fr_store(q + k - 1, acc);
  }
}
// serial base case (one thread per recurrence): q and y = p(z) -> yout[b]
__global__ void k_linrec_serial(LinBatch a, size_t n, Fr* yout) {
  const unsigned b = threadIdx.x;
  if (b >= LIN_MAX_BATCH) return;
  const Fr* p = a.p[b];
  if (!p) return;
  Fr* q = a.q[b];
  const Fr z = a.z[b];
  Fr acc = fr_zero();
  if (q) fr_store(q + n - 1, fr_zero());
  for (size_t k = n; k-- > 0;) {
    acc = fr_add(fr_load(p + k), fr_mul(z, acc));
    if (q && k >= 1) fr_store(q + k - 1, acc);
  }
  fr_store(yout + b, acc);
}

// y[b] is written to the device words yout[b].
static int linrec(tp_ctx* ctx, const LinBatch& in, int batch, size_t n, Fr* yout, int level) {
  if (n <= LIN_BASE) {
    LinBatch a = in;
    for (int b = batch; b < LIN_MAX_BATCH; b++) a.p[b] = nullptr;
    k_linrec_serial<<<1, LIN_MAX_BATCH, 0, ctx->stream>>>(a, n, yout);
    TP_LAUNCH(ctx, "k_linrec_serial");
    return TP_OK;
  }
  size_t nch = (n + LIN_CH - 1) / LIN_CH;
  // level buffers: h (nch per recurrence) and qprime (nch per recurrence)
  TP_TRY(ensure(ctx, ctx->scan_tmp[level], 2 * nch * LIN_MAX_BATCH * sizeof(Fr)));
  Fr* h = (Fr*)ctx->scan_tmp[level].p;
  Fr* qp = h + nch * LIN_MAX_BATCH;
  k_linrec_reduce<<<dim3(ew_grid(nch), (unsigned)batch), EW_THREADS, 0, ctx->stream>>>(in, n, h, nch);
  TP_LAUNCH(ctx, "k_linrec_reduce");
  LinBatch next;
  bool any_q = false;
  for (int b = 0; b < LIN_MAX_BATCH; b++) {
    const int s = b < batch ? b : 0;
    next.p[b] = h + (size_t)s * nch;
    next.q[b] = in.q[s] ? qp + (size_t)s * nch : nullptr;
    next.z[b] = to_dev(to_host(in.z[s]).pow_u64(LIN_CH));
    any_q = any_q || (b < batch && in.q[b]);
  }
  TP_TRY(linrec(ctx, next, batch, nch, yout, level + 1));
  if (any_q) {
    k_linrec_apply<<<dim3(ew_grid(nch), (unsigned)batch), EW_THREADS, 0, ctx->stream>>>(in, n, qp, nch);
    TP_LAUNCH(ctx, "k_linrec_apply");
  }
  return TP_OK;
}

// Evaluate (and, where q_out[b] is non-null, divide by X - z[b]) `batch` polynomials of `len` coefficients.
int poly_open_batch_dev(tp_ctx* ctx, const Fr* const* p, size_t len, const Fr* z, Fr* const* q_out, int batch,
                        tph::HFr* y) {
  if (len == 0) return fail(ctx, TP_ERR_EMPTY_POLY, "open: empty polynomial");
  if (batch < 1 || batch > LIN_MAX_BATCH) return fail(ctx, TP_ERR_INVALID_ARG, "open: bad batch size");
  ProfScope prof(ctx, TP_PHASE_SCAN);
  TP_TRY(ensure(ctx, ctx->misc[2], LIN_MAX_BATCH * sizeof(Fr)));
  Fr* yd = (Fr*)ctx->misc[2].p;
  LinBatch in;
  for (int b = 0; b < LIN_MAX_BATCH; b++) {
    const int s = b < batch ? b : 0;
    in.p[b] = p[s];
    in.q[b] = q_out ? q_out[s] : nullptr;
    in.z[b] = z[s];
  }
  TP_TRY(linrec(ctx, in, batch, len, yd, 0));
  TP_CUDA_OK(ctx, cudaMemcpyAsync(ctx->pinned, yd, batch * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
  TP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  for (int b = 0; b < batch; b++) memcpy(y[b].v, (const uint8_t*)ctx->pinned + 32 * b, 32);
  return TP_OK;
}
int poly_open_dev(tp_ctx* ctx, const Fr* p, size_t len, const Fr& z, Fr* q_out, tph::HFr* y) {
  const Fr* ps[1] = {p};
  Fr* qs[1] = {q_out};
  return poly_open_batch_dev(ctx, ps, len, &z, qs, 1, y);
}

// =====================================================================================
// gate check on the n rows
// =====================================================================================
struct GateArgs {
  const Fr* sel[5];
  const Fr* adv[3];
  const Fr* pi;
  size_t n;
  unsigned* flag;
};
__global__ void k_gate_check(GateArgs g) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= g.n) return;
  Fr a = fr_load(g.adv[0] + j), b = fr_load(g.adv[1] + j), c = fr_load(g.adv[2] + j);
  Fr acc = fr_mul(fr_load(g.sel[0] + j), a);
  acc = fr_add(acc, fr_mul(fr_load(g.sel[1] + j), b));
  acc = fr_sub(acc, fr_mul(fr_load(g.sel[2] + j), c));
  acc = fr_add(acc, fr_mul(fr_mul(fr_load(g.sel[3] + j), a), b));
  acc = fr_add(acc, fr_load(g.sel[4] + j));
  const Fr pi = fr_load(g.pi + j);
  acc = fr_add(acc, pi);
  // bit 0: the gate equation fails on some row; bit 1: some public input is non-zero
  const unsigned f = (fr_is_zero(acc) ? 0u : 1u) | (fr_is_zero(pi) ? 0u : 2u);
  if (f) atomicOr(g.flag, f);
}
int gate_check_dev(tp_ctx* ctx, const Fr* const sel_evals[5], const Fr* const adv[3], const Fr* pi, size_t n,
                   bool* ok, bool* pi_is_zero) {
  TP_TRY(ensure(ctx, ctx->flag, sizeof(unsigned)));
  GateArgs g;
  for (int i = 0; i < 5; i++) g.sel[i] = sel_evals[i];
  for (int i = 0; i < 3; i++) g.adv[i] = adv[i];
  g.pi = pi;
  g.n = n;
  g.flag = (unsigned*)ctx->flag.p;
  TP_CUDA_OK(ctx, cudaMemsetAsync(g.flag, 0, sizeof(unsigned), ctx->stream));
  k_gate_check<<<ew_grid(n), EW_THREADS, 0, ctx->stream>>>(g);
  TP_LAUNCH(ctx, "k_gate_check");
  TP_CUDA_OK(ctx, cudaMemcpyAsync(ctx->pinned, g.flag, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
  TP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  const unsigned f = *(unsigned*)ctx->pinned;
  *ok = (f & 1u) == 0;
  if (pi_is_zero) *pi_is_zero = (f & 2u) == 0;
  return TP_OK;
}

// =====================================================================================
// quotient numerator on the 4n domain, one coset of H at a time
// =====================================================================================
// The 4n-point domain <omega_4n> is the union of the four cosets omega_4n^k H, k = 0..3.  Every
// 4n-sized table here is COSET-MAJOR: entry k * n + i belongs to the point omega_4n^(4i + k).  A
// coset is a self-contained unit of work (five size-n coset NTTs, this kernel, one inverse coset
// NTT), which is what lets the quotient be sharded over GPUs by coset (api.cu).
struct QuotKernelArgs {
  const Fr* sel4[5];
  const Fr* sig4[3];
  const Fr* adv4[3];
  const Fr* z4;
  const Fr* pi4;
  const Fr* l0_4;
  const Fr* tw4;
  Fr alpha, alpha2, beta, gamma;
  Fr bk[3];  // beta * k_i
  Fr* out;
  size_t n;
  unsigned cosets[4];  // blockIdx.y selects the coset
};
__global__ void __launch_bounds__(EW_THREADS) k_quotient_numerator(QuotKernelArgs q) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= q.n) return;
  const unsigned k = q.cosets[blockIdx.y];
  const size_t half = q.n * 2;
  const size_t i4 = 4 * i + k;               // exponent of omega_4n at this point
  const size_t o = (size_t)k * q.n + i;      // coset-major slot
  Fr x = i4 < half ? fr_load(q.tw4 + i4) : fr_neg(fr_load(q.tw4 + (i4 - half)));
  Fr a = fr_load(q.adv4[0] + o), b = fr_load(q.adv4[1] + o), c = fr_load(q.adv4[2] + o);
  Fr z = fr_load(q.z4 + o);
  Fr zw = fr_load(q.z4 + (size_t)k * q.n + (i + 1 < q.n ? i + 1 : 0));   // z(omega x): next point of the same coset
  // gate line (proof.rs:317-320)
  Fr acc = pmul(fr_load(q.sel4[0] + o), a);
  acc = fr_add(acc, pmul(fr_load(q.sel4[1] + o), b));
  acc = fr_sub(acc, pmul(fr_load(q.sel4[2] + o), c));
  acc = fr_add(acc, pmul(pmul(fr_load(q.sel4[3] + o), a), b));
  acc = fr_add(acc, fr_load(q.sel4[4] + o));
  acc = fr_add(acc, fr_load(q.pi4 + o));
  // permutation lines (proof.rs:323-354)
  Fr ag = fr_add(a, q.gamma), bg = fr_add(b, q.gamma), cg = fr_add(c, q.gamma);
  Fr l2 = pmul(fr_add(ag, pmul(q.bk[0], x)), fr_add(bg, pmul(q.bk[1], x)));
  l2 = pmul(l2, fr_add(cg, pmul(q.bk[2], x)));
  l2 = pmul(l2, z);
  Fr l3 = pmul(fr_add(ag, pmul(q.beta, fr_load(q.sig4[0] + o))), fr_add(bg, pmul(q.beta, fr_load(q.sig4[1] + o))));
  l3 = pmul(l3, fr_add(cg, pmul(q.beta, fr_load(q.sig4[2] + o))));
  l3 = pmul(l3, zw);
  acc = fr_add(acc, pmul(q.alpha, fr_sub(l2, l3)));
  // L0 line (proof.rs:355-360)
  Fr l4 = pmul(fr_sub(z, fr_one()), fr_load(q.l0_4 + o));
  acc = fr_add(acc, pmul(q.alpha2, l4));
  fr_store(q.out + o, acc);
}
int quotient_numerator_dev(tp_ctx* ctx, const QuotientArgs& a, const unsigned* cosets, int ncosets) {
  if (ncosets <= 0) return TP_OK;
  ProfScope prof(ctx, TP_PHASE_QUOTIENT);
  QuotKernelArgs q;
  for (int i = 0; i < 5; i++) q.sel4[i] = a.sel4[i];
  for (int i = 0; i < 3; i++) {
    q.sig4[i] = a.sig4[i];
    q.adv4[i] = a.adv4[i];
  }
  q.z4 = a.z4;
  q.pi4 = a.pi4;
  q.l0_4 = a.l0_4;
  q.tw4 = a.tw4;
  tph::HFr al = to_host(a.alpha), be = to_host(a.beta);
  q.alpha = a.alpha;
  q.alpha2 = to_dev(al.sqr());
  q.beta = a.beta;
  q.gamma = a.gamma;
  for (int i = 0; i < 3; i++) q.bk[i] = to_dev(be * to_host(a.k[i]));
  q.out = a.out;
  q.n = a.n;
  for (int i = 0; i < 4; i++) q.cosets[i] = cosets[i < ncosets ? i : 0];
  k_quotient_numerator<<<dim3(ew_grid(a.n), (unsigned)ncosets), EW_THREADS, 0, ctx->stream>>>(q);
  TP_LAUNCH(ctx, "k_quotient_numerator");
  return TP_OK;
}

// Numerator coefficients from its four per-coset interpolants, and the floor division by X^n - 1
// (proof.rs:373, 504-508; the remainder is discarded as the reference does).  Write
// N(X) = sum_j X^(jn) N_j(X), deg N_j < n.  On coset k, x^n = iota^k (iota = omega_4n^n, iota^2 = -1),
// so the interpolant of coset k is C_k = sum_j iota^(kj) N_j and N_j = 1/4 sum_k iota^(-kj) C_k.
// Then t_2 = N_3, t_1 = N_2 + t_2, t_0 = N_1 + t_1.
// c0_is_zero: the numerator vanishes on H itself (always, for a witness that satisfies the copy constraints), so
// coset 0 was never evaluated and its slot holds nothing.
__global__ void k_quotient_combine(const Fr* __restrict__ c4, size_t n, Fr iota_inv, Fr quarter, int c0_is_zero,
                                   Fr* __restrict__ t) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Fr c0 = c0_is_zero ? fr_zero() : fr_load(c4 + i);
  const Fr c1 = fr_load(c4 + n + i), c2 = fr_load(c4 + 2 * n + i), c3 = fr_load(c4 + 3 * n + i);
  const Fr e = fr_add(c0, c2), f = fr_sub(c0, c2), g = fr_add(c1, c3);
  const Fr h = fr_mul(iota_inv, fr_sub(c1, c3));
  const Fr n1 = fr_mul(quarter, fr_add(f, h));
  const Fr n2 = fr_mul(quarter, fr_sub(e, g));
  const Fr n3 = fr_mul(quarter, fr_sub(f, h));
  const Fr t1 = fr_add(n2, n3);
  fr_store(t + 2 * n + i, n3);
  fr_store(t + n + i, t1);
  fr_store(t + i, fr_add(n1, t1));
}
int quotient_combine_dev(tp_ctx* ctx, const Fr* c4, size_t n, bool c0_is_zero, Fr* t) {
  ProfScope prof(ctx, TP_PHASE_QUOTIENT);
  // iota = omega_4n^n is a primitive fourth root of unity: iota^-1 = -iota; quarter = 4^-1
  unsigned log_n = 0;
  while (((size_t)1 << log_n) < n) log_n++;
  tph::HFr iota_inv = omega_for_log(log_n + 2).pow_u64((uint64_t)n).neg();
  tph::HFr quarter = tph::HFr::from_u64(4).inv();
  k_quotient_combine<<<ew_grid(n), EW_THREADS, 0, ctx->stream>>>(c4, n, to_dev(iota_inv), to_dev(quarter), c0_is_zero ? 1 : 0, t);
  TP_LAUNCH(ctx, "k_quotient_combine");
  return TP_OK;
}

// L0(x) = (x^n - 1) / (n (x - 1)) on the 4n domain (coset-major output); chunked batch inversion (8 per thread).
#define L0_CH 8
__global__ void k_l0_evals(const Fr* tw4, size_t n, Fr ninv, Fr* out) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t n4 = 4 * n, half = 2 * n;
  size_t lo = t * L0_CH;
  if (lo >= n4) return;
  Fr xs[L0_CH], pre[L0_CH];
  Fr acc = fr_one();
  Fr one = fr_one();
  for (int i = 0; i < L0_CH; i++) {
    size_t idx = lo + i;
    Fr x = idx < half ? fr_load(tw4 + idx) : fr_neg(fr_load(tw4 + (idx - half)));
    Fr d = fr_sub(x, one);
    if ((idx & 3) == 0) d = one;  // x in H: handled separately (avoid the zero at x = 1)
    xs[i] = d;
    pre[i] = acc;
    acc = fr_mul(acc, d);
  }
  Fr inv = fr_inv(acc);
  // x^n = iota^(idx mod 4), iota = omega_4n^n
  Fr iota = fr_load(tw4 + n);
  for (int i = L0_CH - 1; i >= 0; i--) {
    size_t idx = lo + i;
    Fr di = fr_mul(inv, pre[i]);
    inv = fr_mul(inv, xs[i]);
    Fr r;
    unsigned m = (unsigned)(idx & 3);
    if (m == 0) {
      r = idx == 0 ? one : fr_zero();
    } else {
      Fr xn = m == 1 ? iota : (m == 2 ? fr_neg(one) : fr_neg(iota));
      r = fr_mul(fr_mul(fr_sub(xn, one), ninv), di);
    }
    fr_store(out + (size_t)m * n + (idx >> 2), r);
  }
}
int l0_evals_4n_dev(tp_ctx* ctx, const Fr* tw4, size_t n, Fr* out) {
  tph::HFr ninv = tph::HFr::from_u64((uint64_t)n).inv();
  size_t nth = (4 * n + L0_CH - 1) / L0_CH;
  if ((4 * n) % L0_CH != 0) return fail(ctx, TP_ERR_INVALID_ARG, "l0: 4n must be a multiple of 8");
  k_l0_evals<<<(unsigned)((nth + 127) / 128), 128, 0, ctx->stream>>>(tw4, n, to_dev(ninv), out);
  TP_LAUNCH(ctx, "k_l0_evals");
  return TP_OK;
}

// =====================================================================================
// linear combination out = constant (at coeff 0) + sum_t s_t * p_t
// =====================================================================================
#define LIN_MAX_TERMS 12
struct LinArgs {
  const Fr* p[LIN_MAX_TERMS];
  Fr s[LIN_MAX_TERMS];
  int nterms;
  Fr constant;
  size_t n;
  Fr* out;
};
__global__ void k_lincomb(LinArgs a) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  Fr acc = i == 0 ? a.constant : fr_zero();
  for (int t = 0; t < a.nterms; t++) acc = fr_add(acc, fr_mul(a.s[t], fr_load(a.p[t] + i)));
  fr_store(a.out + i, acc);
}
int lincomb_dev(tp_ctx* ctx, const LinTerm* terms, int nterms, const Fr& constant, size_t n, Fr* out) {
  if (nterms > LIN_MAX_TERMS) return fail(ctx, TP_ERR_INVALID_ARG, "lincomb: too many terms");
  LinArgs a;
  for (int t = 0; t < nterms; t++) {
    a.p[t] = terms[t].p;
    a.s[t] = terms[t].s;
  }
  a.nterms = nterms;
  a.constant = constant;
  a.n = n;
  a.out = out;
  k_lincomb<<<ew_grid(n), EW_THREADS, 0, ctx->stream>>>(a);
  TP_LAUNCH(ctx, "k_lincomb");
  return TP_OK;
}

// =====================================================================================
// sigma / id tables
// =====================================================================================
struct SigmaArgs {
  const uint64_t* perm;
  size_t n;
  const Fr* tw;  // omega_n^i, i in [0, n/2]
  Fr k[3];
  Fr* id[3];
  Fr* sg[3];
};
__device__ __forceinline__ Fr root_pow(const Fr* tw, size_t n, size_t j) {
  size_t half = n >> 1;
  if (n == 1) return fr_one();
  return j < half ? fr_load(tw + j) : fr_neg(fr_load(tw + (j - half)));
}
__global__ void k_sigma_tables(SigmaArgs a) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 3 * a.n) return;
  size_t i = idx / a.n, j = idx % a.n;
  uint64_t p = a.perm[idx];
  size_t ti = (size_t)(p / a.n), tj = (size_t)(p % a.n);
  Fr ki = i == 0 ? a.k[0] : (i == 1 ? a.k[1] : a.k[2]);
  Fr kt = ti == 0 ? a.k[0] : (ti == 1 ? a.k[1] : a.k[2]);
  Fr idv = fr_mul(ki, root_pow(a.tw, a.n, j));
  Fr sgv = fr_mul(kt, root_pow(a.tw, a.n, tj));
  Fr* idp = i == 0 ? a.id[0] : (i == 1 ? a.id[1] : a.id[2]);
  Fr* sgp = i == 0 ? a.sg[0] : (i == 1 ? a.sg[1] : a.sg[2]);
  fr_store(idp + j, idv);
  fr_store(sgp + j, sgv);
}
int sigma_tables_dev(tp_ctx* ctx, const uint64_t* perm_dev, size_t n, const Fr* tw, const Fr k[3], Fr* id[3],
                     Fr* sigma[3]) {
  SigmaArgs a;
  a.perm = perm_dev;
  a.n = n;
  a.tw = tw;
  for (int i = 0; i < 3; i++) {
    a.k[i] = k[i];
    a.id[i] = id[i];
    a.sg[i] = sigma[i];
  }
  k_sigma_tables<<<ew_grid(3 * n), EW_THREADS, 0, ctx->stream>>>(a);
  TP_LAUNCH(ctx, "k_sigma_tables");
  return TP_OK;
}

__global__ void k_pad_copy(const Fr* in, size_t len, Fr* out, size_t out_len) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= out_len) return;
  uint4* o = reinterpret_cast<uint4*>(out + i);
  if (i < len) {
    const uint4* s = reinterpret_cast<const uint4*>(in + i);
    o[0] = s[0];
    o[1] = s[1];
  } else {
    o[0] = make_uint4(0, 0, 0, 0);
    o[1] = make_uint4(0, 0, 0, 0);
  }
}
int pad_copy_dev(tp_ctx* ctx, const Fr* in, size_t len, Fr* out, size_t out_len) {
  k_pad_copy<<<ew_grid(out_len), EW_THREADS, 0, ctx->stream>>>(in, len, out, out_len);
  TP_LAUNCH(ctx, "k_pad_copy");
  return TP_OK;
}
__global__ void k_rotate_copy(const Fr* in, size_t n, size_t shift, Fr* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  size_t s = i + shift;
  if (s >= n) s -= n;
  const uint4* src = reinterpret_cast<const uint4*>(in + s);
  uint4* o = reinterpret_cast<uint4*>(out + i);
  o[0] = src[0];
  o[1] = src[1];
}
int rotate_copy_dev(tp_ctx* ctx, const Fr* in, size_t n, size_t shift, Fr* out) {
  k_rotate_copy<<<ew_grid(n), EW_THREADS, 0, ctx->stream>>>(in, n, shift % n, out);
  TP_LAUNCH(ctx, "k_rotate_copy");
  return TP_OK;
}

}  // namespace tp
