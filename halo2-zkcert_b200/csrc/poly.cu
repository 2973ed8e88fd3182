// Device polynomial / column primitives of the create_proof pipeline.
//
// What each replaces in halo2_proofs 0.2.0 @4b42325 (un-vendored; /root/reference/Cargo.lock:1320-1336):
//   fr_scan(MUL)        the serial running products of permutation::prover / lookup::prover (a8, a9)
//   fr_kate_division    arithmetic::kate_division (serial synthetic division) as a blocked linear-recurrence scan (a11)
//   fr_eval_batch       arithmetic::eval_polynomial (serial Horner), one CTA tree per (poly, point) (a11)
//   fr_lincomb          the `poly * scalar + poly` folds of ProverSHPLONK / ProverGWC (a12)
//   eval_program        plonk::evaluation::GraphEvaluator over Lagrange rows / the extended coset (a7, a9)
//   sort_u256           the `sort` inside lookup::prover::permute_expression_pair (a9) as a bitonic network
#include "poly.cuh"

namespace zkc {

// ---- scans ---------------------------------------------------------------------------------------------
#define SCAN_CH 16

template <int OP> __device__ __forceinline__ Fr scan_op(const Fr& a, const Fr& b) { return OP == SCAN_MUL ? fe_mul(a, b) : fe_add(a, b); }
template <int OP> __device__ __forceinline__ Fr scan_id() { return OP == SCAN_MUL ? fe_one<FrP>() : fe_zero<FrP>(); }

template <int OP>
__global__ void k_scan_chunk_totals(const Fr* in, Fr* tot, uint64_t n, int reverse) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t lo = t * SCAN_CH;
  if (lo >= n) return;
  const uint64_t hi = lo + SCAN_CH < n ? lo + SCAN_CH : n;
  Fr acc = scan_id<OP>();
  for (uint64_t i = lo; i < hi; ++i) acc = scan_op<OP>(acc, fe_load(in + (reverse ? n - 1 - i : i)));
  fe_store(tot + t, acc);
}

// single CTA: tot[c] <- start (op) fold_{c' < c} tot[c']
template <int OP>
__global__ void __launch_bounds__(1024) k_scan_totals(Fr* tot, uint64_t nchunks, Fr start) {
  extern __shared__ uint4 smraw[];
  Fr* sm = reinterpret_cast<Fr*>(smraw);
  const uint32_t t = threadIdx.x;
  const uint64_t per = (nchunks + 1023) / 1024;
  const uint64_t lo = (uint64_t)t * per, hi = lo + per < nchunks ? lo + per : nchunks;
  Fr acc = scan_id<OP>();
  for (uint64_t i = lo; i < hi; ++i) acc = scan_op<OP>(acc, fe_load(tot + i));
  fe_store(sm + t, acc);
  __syncthreads();
  for (uint32_t d = 1; d < 1024; d <<= 1) {
    Fr v = scan_id<OP>();
    const bool act = t >= d;
    if (act) v = fe_load(sm + t - d);
    __syncthreads();
    if (act) fe_store(sm + t, scan_op<OP>(v, fe_load(sm + t)));
    __syncthreads();
  }
  Fr run = t ? scan_op<OP>(start, fe_load(sm + t - 1)) : start;
  for (uint64_t i = lo; i < hi; ++i) { Fr v = fe_load(tot + i); fe_store(tot + i, run); run = scan_op<OP>(run, v); }
}

template <int OP>
__global__ void k_scan_apply(const Fr* in, Fr* out, const Fr* tot, uint64_t n, int reverse) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t lo = t * SCAN_CH;
  if (lo >= n) return;
  const uint64_t hi = lo + SCAN_CH < n ? lo + SCAN_CH : n;
  Fr run = fe_load(tot + t);
  for (uint64_t i = lo; i < hi; ++i) {
    const uint64_t idx = reverse ? n - 1 - i : i;
    Fr v = fe_load(in + idx);
    fe_store(out + idx, run);
    run = scan_op<OP>(run, v);
  }
}

// chunk totals beyond this count are scanned by a recursive pass instead of the single-CTA kernel
#define SCAN_SINGLE_MAX 2048   // (one 1024-thread CTA is one SM: beyond a few totals per thread a parallel level is faster)
static int fr_scan_impl(zkc_ctx* ctx, const Fr* in, Fr* out, uint64_t n, int op, int reverse, const Fr& start, Fr* scratch) {
  const uint64_t nchunks = (n + SCAN_CH - 1) / SCAN_CH;
  Fr* tot = scratch;
  const unsigned grid = (unsigned)((nchunks + 127) / 128);
  cudaStream_t st = ctx->stream;
  if (op == SCAN_MUL) { k_scan_chunk_totals<SCAN_MUL><<<grid, 128, 0, st>>>(in, tot, n, reverse); }
  else { k_scan_chunk_totals<SCAN_ADD><<<grid, 128, 0, st>>>(in, tot, n, reverse); }
  ZKC_LAUNCH_CHECK(ctx);
  if (nchunks > SCAN_SINGLE_MAX) {
    // totals are already in scan order: a forward exclusive scan (with the start value folded in) one level up
    ZKC_TRY(fr_scan_impl(ctx, tot, tot, nchunks, op, 0, start, scratch + nchunks));
  } else {
    if (op == SCAN_MUL) { k_scan_totals<SCAN_MUL><<<1, 1024, 1024 * sizeof(Fr), st>>>(tot, nchunks, start); }
    else { k_scan_totals<SCAN_ADD><<<1, 1024, 1024 * sizeof(Fr), st>>>(tot, nchunks, start); }
    ZKC_LAUNCH_CHECK(ctx);
  }
  if (op == SCAN_MUL) { k_scan_apply<SCAN_MUL><<<grid, 128, 0, st>>>(in, out, tot, n, reverse); }
  else { k_scan_apply<SCAN_ADD><<<grid, 128, 0, st>>>(in, out, tot, n, reverse); }
  ZKC_LAUNCH_CHECK(ctx);
  return ZKC_OK;
}
int fr_scan(zkc_ctx* ctx, const Fr* in, Fr* out, uint64_t n, int op, int reverse, const Fr& start) {
  if (n == 0) return ZKC_OK;
  ProfScope _p(ctx, op == SCAN_MUL ? "scan.mul" : "scan.add");
  uint64_t need = 0;
  for (uint64_t m = n; ; ) { m = (m + SCAN_CH - 1) / SCAN_CH; need += m; if (m <= SCAN_SINGLE_MAX) break; }
  Fr* scratch;
  ZKC_TRY(scratch_reserve(ctx, SCR_MISC, need * sizeof(Fr), (void**)&scratch));
  return fr_scan_impl(ctx, in, out, n, op, reverse, start, scratch);
}

// ---- batch inversion by two scans: inv_i = (prod_{j<i} a_j) (prod_{j>i} a_j) / prod_j a_j, zeros passed through ----
__global__ void k_bi_prep(const Fr* a, Fr* t, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr v = fe_load(a + i);
  fe_store(t + i, fe_is_zero(v) ? fe_one<FrP>() : v);
}
__global__ void k_bi_finish(const Fr* a, const Fr* prefix, const Fr* suffix, Fr tinv, Fr* out, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Fr v = fe_load(a + i);
  if (fe_is_zero(v)) { fe_store(out + i, v); return; }
  fe_store(out + i, fe_mul(fe_mul(fe_load(prefix + i), fe_load(suffix + i)), tinv));
}
int fr_batch_invert_scan(zkc_ctx* ctx, const Fr* a, Fr* out, uint64_t n) {
  ProfScope _p(ctx, "batch_invert");
  Fr* buf = nullptr;
  cudaStream_t st = ctx->stream;
  ZKC_CUDA_TRY(ctx, cudaMallocAsync((void**)&buf, 3 * n * sizeof(Fr), st));
  Fr *t = buf, *pf = buf + n, *sf = buf + 2 * n;
  const unsigned grid = (unsigned)((n + 255) / 256);
  int status = ZKC_OK;
  k_bi_prep<<<grid, 256, 0, st>>>(a, t, n); ctx->launches++;
  status = fr_scan(ctx, t, pf, n, SCAN_MUL, 0, fe_one<FrP>());
  if (status == ZKC_OK) status = fr_scan(ctx, t, sf, n, SCAN_MUL, 1, fe_one<FrP>());
  if (status == ZKC_OK) {
    // the one true inversion runs on the host: a lone GPU thread needs ~380 dependent products (160 us) for it, the
    // host 10 us plus a 64-byte round trip
    Fr last[2];
    cudaError_t e = cudaMemcpyAsync(&last[0], pf + n - 1, sizeof(Fr), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&last[1], t + n - 1, sizeof(Fr), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) {
      k_bi_finish<<<grid, 256, 0, st>>>(a, pf, sf, fe_inv(fe_mul(last[0], last[1])), out, n); ctx->launches++;
      e = cudaGetLastError();
    }
    if (e != cudaSuccess) status = set_err(ctx, ZKC_ERR_CUDA, cudaGetErrorString(e));
  }
  cudaFreeAsync(buf, st);
  return status;
}

// u32 exclusive scan: three phases with 1024-element chunks per CTA
__global__ void __launch_bounds__(256) k_u32_block_sums(const uint32_t* in, uint32_t* sums, uint64_t n) {
  __shared__ uint32_t sm[256];
  const uint64_t base = (uint64_t)blockIdx.x * 1024;
  uint32_t s = 0;
  for (uint32_t e = threadIdx.x; e < 1024; e += 256) { const uint64_t i = base + e; if (i < n) s += in[i]; }
  sm[threadIdx.x] = s;
  __syncthreads();
  for (uint32_t d = 128; d > 0; d >>= 1) { if (threadIdx.x < d) sm[threadIdx.x] += sm[threadIdx.x + d]; __syncthreads(); }
  if (threadIdx.x == 0) sums[blockIdx.x] = sm[0];
}
__global__ void __launch_bounds__(1024) k_u32_scan_sums(uint32_t* sums, uint64_t nb, uint32_t* total) {
  __shared__ uint32_t part[1024];
  const uint32_t t = threadIdx.x;
  const uint64_t per = (nb + 1023) / 1024;
  const uint64_t lo = (uint64_t)t * per, hi = lo + per < nb ? lo + per : nb;
  uint32_t s = 0;
  for (uint64_t i = lo; i < hi; ++i) s += sums[i];
  part[t] = s;
  __syncthreads();
  for (uint32_t d = 1; d < 1024; d <<= 1) {
    uint32_t v = t >= d ? part[t - d] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  uint32_t run = t ? part[t - 1] : 0;
  for (uint64_t i = lo; i < hi; ++i) { uint32_t v = sums[i]; sums[i] = run; run += v; }
  if (t == 1023 && total) *total = part[1023];
}
__global__ void __launch_bounds__(256) k_u32_apply(const uint32_t* in, uint32_t* out, const uint32_t* sums, uint64_t n) {
  __shared__ uint32_t sm[256];
  const uint64_t base = (uint64_t)blockIdx.x * 1024 + (uint64_t)threadIdx.x * 4;
  uint32_t v[4], s = 0;
#pragma unroll
  for (int e = 0; e < 4; ++e) { v[e] = base + e < n ? in[base + e] : 0; s += v[e]; }
  sm[threadIdx.x] = s;
  __syncthreads();
  for (uint32_t d = 1; d < 256; d <<= 1) {
    uint32_t x = threadIdx.x >= d ? sm[threadIdx.x - d] : 0;
    __syncthreads();
    sm[threadIdx.x] += x;
    __syncthreads();
  }
  uint32_t run = sums[blockIdx.x] + (threadIdx.x ? sm[threadIdx.x - 1] : 0);
#pragma unroll
  for (int e = 0; e < 4; ++e) { if (base + e < n) out[base + e] = run; run += v[e]; }
}

int u32_scan(zkc_ctx* ctx, const uint32_t* in, uint32_t* out, uint64_t n, uint32_t* total_dev) {
  if (n == 0) { if (total_dev) ZKC_CUDA_TRY(ctx, cudaMemsetAsync(total_dev, 0, 4, ctx->stream)); return ZKC_OK; }
  const uint64_t nb = (n + 1023) / 1024;
  uint32_t* sums;
  ZKC_TRY(scratch_reserve(ctx, SCR_MISC2, nb * 4, (void**)&sums));
  cudaStream_t st = ctx->stream;
  k_u32_block_sums<<<(unsigned)nb, 256, 0, st>>>(in, sums, n); ZKC_LAUNCH_CHECK(ctx);
  k_u32_scan_sums<<<1, 1024, 0, st>>>(sums, nb, total_dev); ZKC_LAUNCH_CHECK(ctx);
  k_u32_apply<<<(unsigned)nb, 256, 0, st>>>(in, out, sums, n); ZKC_LAUNCH_CHECK(ctx);
  return ZKC_OK;
}

// ---- element-wise helpers -------------------------------------------------------------------------------
__global__ void k_powers64(Fr* out, Fr base, Fr first, uint64_t n) {
  const uint64_t start = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 64;
  if (start >= n) return;
  Fr w = fe_mul(first, fe_pow_u64(base, start));
  const uint64_t end = start + 64 < n ? start + 64 : n;
  for (uint64_t i = start; i < end; ++i) { fe_store(out + i, w); w = fe_mul(w, base); }
}
int fr_powers(zkc_ctx* ctx, Fr* out, uint64_t n, const Fr& base, const Fr& first) {
  if (n == 0) return ZKC_OK;
  const uint64_t threads = (n + 63) / 64;
  k_powers64<<<(unsigned)((threads + 127) / 128), 128, 0, ctx->stream>>>(out, base, first, n);
  ZKC_LAUNCH_CHECK(ctx);
  return ZKC_OK;
}

#define LINCOMB_MAX 32
struct LincombArgs { const Fr* polys[LINCOMB_MAX]; Fr coefs[LINCOMB_MAX]; uint32_t count; int accumulate; };
__global__ void k_lincomb(Fr* out, uint64_t n, LincombArgs a) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr acc = a.accumulate ? fe_load(out + i) : fe_zero<FrP>();
  for (uint32_t j = 0; j < a.count; ++j) acc = fe_add(acc, fe_mul(fe_load(a.polys[j] + i), a.coefs[j]));
  fe_store(out + i, acc);
}
int fr_lincomb(zkc_ctx* ctx, Fr* out, uint64_t n, const std::vector<const Fr*>& polys, const std::vector<Fr>& coefs) {
  ProfScope _p(ctx, "lincomb");
  if (polys.empty()) { ZKC_CUDA_TRY(ctx, cudaMemsetAsync(out, 0, n * sizeof(Fr), ctx->stream)); return ZKC_OK; }
  for (size_t off = 0; off < polys.size(); off += LINCOMB_MAX) {
    LincombArgs a;
    a.count = (uint32_t)std::min<size_t>(LINCOMB_MAX, polys.size() - off);
    a.accumulate = off != 0;
    for (uint32_t j = 0; j < a.count; ++j) { a.polys[j] = polys[off + j]; a.coefs[j] = coefs[off + j]; }
    k_lincomb<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(out, n, a);
    ZKC_LAUNCH_CHECK(ctx);
  }
  return ZKC_OK;
}

#define SUBLOW_MAX 16
struct SubLowArgs { Fr v[SUBLOW_MAX]; uint32_t m; };
__global__ void k_sub_low(Fr* a, SubLowArgs s) {
  const uint32_t i = threadIdx.x;
  if (i < s.m) fe_store(a + i, fe_sub(fe_load(a + i), s.v[i]));
}
int fr_sub_low(zkc_ctx* ctx, Fr* a, const std::vector<Fr>& low) {
  if (low.empty()) return ZKC_OK;
  if (low.size() > SUBLOW_MAX) return set_err(ctx, ZKC_ERR_BAD_ARG, "fr_sub_low: too many coefficients");
  SubLowArgs s; s.m = (uint32_t)low.size();
  for (size_t i = 0; i < low.size(); ++i) s.v[i] = low[i];
  k_sub_low<<<1, 32, 0, ctx->stream>>>(a, s);
  ZKC_LAUNCH_CHECK(ctx);
  return ZKC_OK;
}

__global__ void k_scale(Fr* a, uint64_t n, Fr s) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) fe_store(a + i, fe_mul(fe_load(a + i), s));
}
int fr_scale(zkc_ctx* ctx, Fr* a, uint64_t n, const Fr& s) {
  if (!n) return ZKC_OK;
  k_scale<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(a, n, s);
  ZKC_LAUNCH_CHECK(ctx);
  return ZKC_OK;
}
__global__ void k_mul_add(Fr* a, const Fr* b, uint64_t n, Fr s) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) fe_store(a + i, fe_add(fe_mul(fe_load(a + i), s), fe_load(b + i)));
}
int fr_mul_add(zkc_ctx* ctx, Fr* a, const Fr* b, uint64_t n, const Fr& s) {
  if (!n) return ZKC_OK;
  k_mul_add<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(a, b, n, s);
  ZKC_LAUNCH_CHECK(ctx);
  return ZKC_OK;
}

// ---- kate division: q_j = sum_{i>j} a_i z^(i-j-1), blocked synthetic division -----------------------------
// phase 1: per 16-coefficient chunk P_c = sum_e a[c0+e] z^e; phase 2 (one CTA per job, or a recursion): carry_c = value of all
// higher chunks at z (a linear recurrence with constant multiplier Z = z^16, solved by a Hillis-Steele
// scan over Z^(per*2^s)); phase 3: q_{i-1} = a_i + z q_i inside each chunk starting from its carry.
#define KD_CH 16
int fr_kate_division(zkc_ctx* ctx, const Fr* a, Fr* q, uint64_t n, const Fr& z, Fr* tmp1, Fr* tmp2) {
  (void)tmp2;
  if (n == 0) return ZKC_OK;
  if (a != q) ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(q, a, n * sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
  return fr_kate_division_batch(ctx, std::vector<Fr*>{q}, std::vector<Fr>{z}, n, tmp1);
}

// Several independent in-place divisions (different polynomials, different roots) in one set of launches:
// blockIdx.y selects the job.  Used per "round" of SHPLONK (one root of every rotation set) and for all GWC points.
struct KdBatch { Fr* a[KD_MAX_JOBS]; Fr z[KD_MAX_JOBS]; Fr Z[KD_MAX_JOBS]; uint32_t njobs; };
__global__ void k_kd_chunk_b(KdBatch b, Fr* P, uint64_t n, uint64_t nchunks) {
  const uint32_t j = blockIdx.y;
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t lo = t * KD_CH;
  if (lo >= n) return;
  const uint64_t hi = lo + KD_CH < n ? lo + KD_CH : n;
  const Fr z = b.z[j];
  const Fr* a = b.a[j];
  Fr acc = fe_zero<FrP>();
  for (uint64_t i = hi; i-- > lo;) acc = fe_add(fe_mul(acc, z), fe_load(a + i));
  fe_store(P + (uint64_t)j * nchunks + t, acc);
}
__global__ void __launch_bounds__(1024) k_kd_carry_b(KdBatch b, Fr* Pall, uint64_t nchunks) {
  extern __shared__ uint4 smraw[];
  Fr* sm = reinterpret_cast<Fr*>(smraw);
  Fr* P = Pall + (uint64_t)blockIdx.x * nchunks;
  const Fr Z = b.Z[blockIdx.x];
  const uint32_t t = threadIdx.x;
  const uint64_t per = (nchunks + 1023) / 1024;
  const uint64_t lo = (uint64_t)t * per, hi = lo + per < nchunks ? lo + per : nchunks;
  Fr L = fe_zero<FrP>();
  for (uint64_t c = hi; c-- > lo && hi > lo;) L = fe_add(fe_mul(L, Z), fe_load(P + c));
  fe_store(sm + t, L);
  __syncthreads();
  Fr M = fe_pow_u64(Z, per);
  for (uint32_t d = 1; d < 1024; d <<= 1) {
    Fr v = fe_zero<FrP>();
    const bool act = t + d < 1024;
    if (act) v = fe_load(sm + t + d);
    __syncthreads();
    if (act) fe_store(sm + t, fe_add(fe_load(sm + t), fe_mul(M, v)));
    M = fe_sqr(M);
    __syncthreads();
  }
  Fr carry = t + 1 < 1024 ? fe_load(sm + t + 1) : fe_zero<FrP>();
  for (uint64_t c = hi; c-- > lo && hi > lo;) {
    const Fr pc = fe_load(P + c);
    fe_store(P + c, carry);
    carry = fe_add(pc, fe_mul(Z, carry));
  }
}
__global__ void k_kd_apply_b(KdBatch b, const Fr* carry, uint64_t n, uint64_t nchunks) {
  const uint32_t j = blockIdx.y;
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t lo = t * KD_CH;
  if (lo >= n) return;
  const uint64_t hi = lo + KD_CH < n ? lo + KD_CH : n;
  const Fr z = b.z[j];
  Fr* a = b.a[j];
  Fr r = fe_load(carry + (uint64_t)j * nchunks + t);
  for (uint64_t i = hi; i-- > lo;) {
    const Fr ai = fe_load(a + i);
    fe_store(a + i, r);
    r = fe_add(ai, fe_mul(z, r));
  }
}
// The carries of the 16-coefficient chunks are themselves a synthetic division — of the chunk values P_c by (X - z^16) —
// so large inputs recurse (every level is a grid-wide launch) and only the last <= KD_SINGLE_MAX partials go through the
// single-CTA scan.  `tmp` holds the partials of all levels: njobs * n / 15 elements at most (callers pass n + n / 8).
#define KD_SINGLE_MAX 2048
static int kd_batch_level(zkc_ctx* ctx, KdBatch b, uint64_t n, Fr* tmp) {
  const uint64_t nchunks = (n + KD_CH - 1) / KD_CH;
  const unsigned gx = (unsigned)((nchunks + 127) / 128);
  dim3 grid(gx, b.njobs);
  k_kd_chunk_b<<<grid, 128, 0, ctx->stream>>>(b, tmp, n, nchunks); ZKC_LAUNCH_CHECK(ctx);
  if (nchunks > KD_SINGLE_MAX) {
    KdBatch up;
    up.njobs = b.njobs;
    for (uint32_t j = 0; j < b.njobs; ++j) { up.a[j] = tmp + (uint64_t)j * nchunks; up.z[j] = b.Z[j]; up.Z[j] = fe_pow_u64(b.Z[j], KD_CH); }
    ZKC_TRY(kd_batch_level(ctx, up, nchunks, tmp + (uint64_t)b.njobs * nchunks));
  } else {
    k_kd_carry_b<<<b.njobs, 1024, 1024 * sizeof(Fr), ctx->stream>>>(b, tmp, nchunks); ZKC_LAUNCH_CHECK(ctx);
  }
  k_kd_apply_b<<<grid, 128, 0, ctx->stream>>>(b, tmp, n, nchunks); ZKC_LAUNCH_CHECK(ctx);
  return ZKC_OK;
}
int fr_kate_division_batch(zkc_ctx* ctx, const std::vector<Fr*>& polys, const std::vector<Fr>& roots, uint64_t n, Fr* tmp /* n + n / 8 elements */) {
  if (n == 0 || polys.empty()) return ZKC_OK;
  ProfScope _p(ctx, "kate_division");
  for (size_t off = 0; off < polys.size(); off += KD_MAX_JOBS) {
    KdBatch b;
    b.njobs = (uint32_t)std::min<size_t>(KD_MAX_JOBS, polys.size() - off);   // KD_MAX_JOBS * n / 15 <= n + n / 8
    for (uint32_t j = 0; j < b.njobs; ++j) { b.a[j] = polys[off + j]; b.z[j] = roots[off + j]; b.Z[j] = fe_pow_u64(roots[off + j], KD_CH); }
    ZKC_TRY(kd_batch_level(ctx, b, n, tmp));
  }
  return ZKC_OK;
}

// q[j] += C * z^(len-1-j): the contribution of everything above a coefficient slice to its synthetic-division
// quotient (team proving: each rank divides its own slice, the tails' values arrive as one field element)
__global__ void k_add_geometric(Fr* q, uint64_t len, Fr C, Fr z) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t hi = len - (t * 64 < len ? t * 64 : len);    // this thread covers j in (hi - 64, hi], walking down from exponent t*64
  if (hi == 0) return;
  Fr w = fe_mul(C, fe_pow_u64(z, t * 64));
  const uint64_t lo = hi > 64 ? hi - 64 : 0;
  for (uint64_t j = hi; j-- > lo;) { fe_store(q + j, fe_add(fe_load(q + j), w)); w = fe_mul(w, z); }
}
int fr_add_geometric(zkc_ctx* ctx, Fr* q, uint64_t len, const Fr& C, const Fr& z) {
  if (len == 0 || fe_is_zero(C)) return ZKC_OK;
  const uint64_t threads = (len + 63) / 64;
  k_add_geometric<<<(unsigned)((threads + 127) / 128), 128, 0, ctx->stream>>>(q, len, C, z);
  ZKC_LAUNCH_CHECK(ctx);
  return ZKC_OK;
}

// ---- batched polynomial evaluation ---------------------------------------------------------------------------
// grid = (blocks_per_poly, n_evals); each CTA of 256 threads evaluates a 4096-coefficient slice at
// the point (16-coefficient Horner per thread, then a shared-memory tree with x^(16*2^l) factors).
#define EV_PER_THREAD 16
#define EV_THREADS 256
struct EvalJob { const Fr* poly; Fr x; Fr xp[8]; Fr X; Fr X32; };   // xp[l] = x^(16 * 2^l), X = x^4096, X32 = X^32
__global__ void __launch_bounds__(EV_THREADS) k_eval_slices(const EvalJob* jobs, uint64_t n, Fr* partial, uint32_t blocks_per_poly) {
  __shared__ uint4 smraw[EV_THREADS * 2];
  Fr* sm = reinterpret_cast<Fr*>(smraw);
  const EvalJob* job = jobs + blockIdx.y;
  const Fr x = fe_load_nc(&job->x);
  const Fr* poly = job->poly;
  const uint64_t base = ((uint64_t)blockIdx.x * EV_THREADS + threadIdx.x) * EV_PER_THREAD;
  Fr acc = fe_zero<FrP>();
  for (int e = EV_PER_THREAD - 1; e >= 0; --e) {
    const uint64_t i = base + e;
    Fr c = i < n ? fe_load(poly + i) : fe_zero<FrP>();
    acc = fe_add(fe_mul(acc, x), c);
  }
  fe_store(sm + threadIdx.x, acc);
  __syncthreads();
  int l = 0;
  for (uint32_t d = 1; d < EV_THREADS; d <<= 1, ++l) {
    if ((threadIdx.x & (2 * d - 1)) == 0) {
      Fr lo = fe_load(sm + threadIdx.x), hi = fe_load(sm + threadIdx.x + d);
      fe_store(sm + threadIdx.x, fe_add(lo, fe_mul(hi, fe_load_nc(&job->xp[l]))));
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) fe_store(partial + (uint64_t)blockIdx.y * blocks_per_poly + blockIdx.x, fe_load(sm));
}
// one warp per evaluation: sum_b X^b * partial[b], X = x^4096: lane l folds blocks l, l+32, ... by Horner in X^32
__global__ void __launch_bounds__(32) k_eval_combine(const EvalJob* jobs, const Fr* partial, uint32_t blocks_per_poly, Fr* out) {
  __shared__ uint4 smraw[32 * 2];
  Fr* sm = reinterpret_cast<Fr*>(smraw);
  const Fr X = fe_load_nc(&jobs[blockIdx.x].X), X32 = fe_load_nc(&jobs[blockIdx.x].X32);
  const Fr* p = partial + (uint64_t)blockIdx.x * blocks_per_poly;
  Fr acc = fe_zero<FrP>();
  int cnt = ((int)blocks_per_poly - (int)threadIdx.x + 31) / 32;
  for (int m = cnt - 1; m >= 0; --m) acc = fe_add(fe_mul(acc, X32), fe_load(p + threadIdx.x + 32 * m));
  fe_store(sm + threadIdx.x, fe_mul(acc, fe_pow_u64(X, threadIdx.x)));
  __syncwarp();
  if (threadIdx.x == 0) {
    Fr s = fe_zero<FrP>();
    for (int l = 0; l < 32; ++l) s = fe_add(s, fe_load(sm + l));
    fe_store(out + blockIdx.x, s);
  }
}
int fr_eval_batch(zkc_ctx* ctx, const std::vector<const Fr*>& polys, uint64_t n, const std::vector<Fr>& points, std::vector<Fr>& out) {
  const size_t m = polys.size();
  out.resize(m);
  if (m == 0) return ZKC_OK;
  ProfScope _p(ctx, "eval_batch");
  const uint32_t bpp = (uint32_t)((n + EV_THREADS * EV_PER_THREAD - 1) / (EV_THREADS * EV_PER_THREAD));
  std::vector<EvalJob> jobs(m);
  for (size_t i = 0; i < m; ++i) {
    jobs[i].poly = polys[i]; jobs[i].x = points[i];
    bool reuse = false;   // many queries share a point: copy the power table of the previous identical point
    for (size_t j = i; j-- > 0 && !reuse;) if (fe_eq(points[j], points[i])) { memcpy(jobs[i].xp, jobs[j].xp, sizeof(jobs[i].xp)); jobs[i].X = jobs[j].X; jobs[i].X32 = jobs[j].X32; reuse = true; }
    if (reuse) continue;
    Fr pw = fe_pow_u64(points[i], EV_PER_THREAD);
    for (int l = 0; l < 8; ++l) { jobs[i].xp[l] = pw; pw = fe_sqr(pw); }
    jobs[i].X = pw;                       // x^(16 * 2^8) = x^4096
    jobs[i].X32 = fe_pow_u64(pw, 32);
  }
  char* base;
  const size_t o_jobs = 0, o_part = (m * sizeof(EvalJob) + 255) & ~(size_t)255, o_out = o_part + ((m * bpp * sizeof(Fr) + 255) & ~(size_t)255);
  ZKC_TRY(scratch_reserve(ctx, SCR_MISC3, o_out + m * sizeof(Fr), (void**)&base));
  ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(base + o_jobs, jobs.data(), m * sizeof(EvalJob), cudaMemcpyHostToDevice, ctx->stream));
  // the pageable source above is consumed synchronously by cudaMemcpyAsync (staged), safe to reuse
  for (size_t off = 0; off < m; off += 65535) {
    const uint32_t cnt = (uint32_t)std::min<size_t>(65535, m - off);
    dim3 grid(bpp, cnt);
    k_eval_slices<<<grid, EV_THREADS, 0, ctx->stream>>>((const EvalJob*)(base + o_jobs) + off, n, (Fr*)(base + o_part) + off * bpp, bpp);
    ZKC_LAUNCH_CHECK(ctx);
    k_eval_combine<<<cnt, 32, 0, ctx->stream>>>((const EvalJob*)(base + o_jobs) + off, (const Fr*)(base + o_part) + off * bpp, bpp,
                                               (Fr*)(base + o_out) + off);
    ZKC_LAUNCH_CHECK(ctx);
  }
  ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(out.data(), base + o_out, m * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
  ZKC_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return ZKC_OK;
}

// ---- constraint-system program interpreter ---------------------------------------------------------------------
#define PROG_STACK 12
struct ProgPows { Fr p[8]; };   // mult^L for the factored groups of the program (host/cs.h optimize_program)
template <bool GROUPS>   // GROUPS: the stream holds factored runs (ops 9 / 10); programs without any run the leaner instantiation
__global__ void __launch_bounds__(128) k_eval_program(const uint32_t* words, uint32_t npairs, const Fr* consts, DevQueries q, Fr* out,
                                                      uint64_t rows, uint32_t rot_scale, Fr mult, ProgPows pows, int accumulate, uint64_t row0,
                                                      uint64_t cnt) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= cnt) return;
  const uint64_t i = row0 + tid;
  // The top of the evaluation stack lives in registers (t0); only the values underneath it go to the local-memory array.  With
  // the operand-fused opcodes of host/cs.h optimize_program most operations never touch that array.
  Fr stack[PROG_STACK];
  Fr t0 = fe_zero<FrP>();
  int sp = 0;    // live values: t0 (when sp > 0) and stack[0 .. sp - 1)
  Fr acc = accumulate ? fe_load(out + i) : fe_zero<FrP>();
  Fr saved = fe_zero<FrP>();
  auto column = [&](uint32_t kind, uint32_t arg) -> Fr {   // kind: 0 advice, 1 fixed, 2 instance
    const uint32_t col = kind == 0 ? q.aq_col[arg] : (kind == 1 ? q.fq_col[arg] : q.iq_col[arg]);
    const int32_t rot = kind == 0 ? q.aq_rot[arg] : (kind == 1 ? q.fq_rot[arg] : q.iq_rot[arg]);
    const Fr* colp = kind == 0 ? q.advice[col] : (kind == 1 ? q.fixed[col] : q.instance[col]);
    // rows (a power of two) is the cyclic length: the column itself (Lagrange values), or one residue class of the
    // class-major extended coset, in which case i >= rows selects the class and rotations stay inside it
    const uint64_t idx = (i & ~(rows - 1)) | ((i + rows + (int64_t)rot * (int64_t)rot_scale) & (rows - 1));
    return fe_load(colp + idx);
  };
  for (uint32_t pc = 0; pc < npairs; ++pc) {
    const uint32_t op = words[2 * pc], arg = words[2 * pc + 1];
    switch (op) {
      case 0: if (sp) stack[sp - 1] = t0; t0 = fe_load_nc(consts + arg); ++sp; break;
      case 1: case 2: case 3: if (sp) stack[sp - 1] = t0; t0 = column(op - 1, arg); ++sp; break;
      case 4: t0 = fe_neg(t0); break;
      case 5: t0 = fe_add(stack[sp - 2], t0); --sp; break;
      case 6: t0 = fe_mul(stack[sp - 2], t0); --sp; break;
      case 7: t0 = fe_mul(t0, fe_load_nc(consts + arg)); break;
      case 9: if (GROUPS) { saved = acc; acc = fe_zero<FrP>(); } break;                             // GROUP_BEGIN
      case 10: if (GROUPS) {                                                                        // GROUP_END
        Fr pw;   // static indices: a dynamically indexed kernel parameter would be copied to local memory by every thread
        switch (arg & 7) {
          case 0: pw = pows.p[0]; break; case 1: pw = pows.p[1]; break; case 2: pw = pows.p[2]; break; case 3: pw = pows.p[3]; break;
          case 4: pw = pows.p[4]; break; case 5: pw = pows.p[5]; break; case 6: pw = pows.p[6]; break; default: pw = pows.p[7]; break;
        }
        acc = fe_add(fe_mul(saved, pw), fe_mul(t0, acc));
        sp = 0;
        break;
      }
      case 11: case 12: case 13: t0 = fe_mul(t0, column(op - 11, arg)); break;                      // operand-fused forms
      case 14: case 15: case 16: t0 = fe_add(t0, column(op - 14, arg)); break;
      case 17: case 18: case 19: t0 = fe_sub(t0, column(op - 17, arg)); break;
      case 20: t0 = fe_sub(stack[sp - 2], t0); --sp; break;
      case 21: t0 = fe_add(t0, fe_load_nc(consts + arg)); break;
      default: acc = fe_add(fe_mul(acc, mult), t0); sp = 0; break;   // OP_END
    }
  }
  fe_store(out + i, acc);
}
int eval_program(zkc_ctx* ctx, const DevProgram& prog, const DevQueries& q, Fr* out, uint64_t rows, uint32_t rot_scale, const Fr& mult,
                 int accumulate, uint64_t row0, uint64_t cnt) {
  ProfScope _p(ctx, "eval_program");
  if (cnt == UINT64_MAX) { row0 = 0; cnt = rows; }
  if (cnt == 0) return ZKC_OK;
  if (prog.npairs == 0) {
    if (!accumulate) ZKC_CUDA_TRY(ctx, cudaMemsetAsync(out + row0, 0, cnt * sizeof(Fr), ctx->stream));
    return ZKC_OK;
  }
  ProgPows pows;
  for (uint32_t s = 0; s < 8; ++s) pows.p[s] = s < prog.npows ? fe_pow_u64(mult, prog.pow_len[s]) : fe_zero<FrP>();
  if (prog.npows) k_eval_program<true><<<(unsigned)((cnt + 127) / 128), 128, 0, ctx->stream>>>(prog.words, prog.npairs, prog.consts, q, out, rows, rot_scale,
                                                                                                mult, pows, accumulate, row0, cnt);
  else k_eval_program<false><<<(unsigned)((cnt + 127) / 128), 128, 0, ctx->stream>>>(prog.words, prog.npairs, prog.consts, q, out, rows, rot_scale,
                                                                                      mult, pows, accumulate, row0, cnt);
  ZKC_LAUNCH_CHECK(ctx);
  return ZKC_OK;
}

// ---- bitonic sort of 256-bit canonical keys -------------------------------------------------------------------------
__device__ __forceinline__ bool u256_gt(const Fr& a, const Fr& b) {
#pragma unroll
  for (int i = 7; i >= 0; --i) {
    if (a.v[i] > b.v[i]) return true;
    if (a.v[i] < b.v[i]) return false;
  }
  return false;
}
#define SORT_TILE 1024
// all (k, j) steps with j < SORT_TILE for k in [k_lo, k_hi] done inside shared memory
__global__ void __launch_bounds__(512) k_bitonic_local(Fr* keys, uint64_t k_lo, uint64_t k_hi, int only_tail) {
  __shared__ uint4 smraw[SORT_TILE * 2];
  Fr* sm = reinterpret_cast<Fr*>(smraw);
  const uint64_t base = (uint64_t)blockIdx.x * SORT_TILE;
  for (uint32_t e = threadIdx.x; e < SORT_TILE; e += 512) fe_store(sm + e, fe_load(keys + base + e));
  __syncthreads();
  for (uint64_t k = k_lo; k <= k_hi; k <<= 1) {
    uint64_t j0 = only_tail ? (SORT_TILE >> 1) : (k >> 1);
    if (j0 > (SORT_TILE >> 1)) j0 = SORT_TILE >> 1;
    for (uint64_t j = j0; j > 0; j >>= 1) {
      const uint32_t t = threadIdx.x;                 // 512 threads <-> 512 pairs
      const uint32_t lo = (uint32_t)(((t & ~(uint32_t)(j - 1)) << 1) | (t & (uint32_t)(j - 1)));
      const uint32_t hi = lo + (uint32_t)j;
      const bool asc = (((base + lo) & k) == 0);
      Fr a = fe_load(sm + lo), b = fe_load(sm + hi);
      if (u256_gt(a, b) == asc) { fe_store(sm + lo, b); fe_store(sm + hi, a); }
      __syncthreads();
    }
  }
  for (uint32_t e = threadIdx.x; e < SORT_TILE; e += 512) fe_store(keys + base + e, fe_load(sm + e));
}
__global__ void k_bitonic_global(Fr* keys, uint64_t n, uint64_t j, uint64_t k) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (n >> 1)) return;
  const uint64_t lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
  const uint64_t hi = lo + j;
  const bool asc = ((lo & k) == 0);
  Fr a = fe_load(keys + lo), b = fe_load(keys + hi);
  if (u256_gt(a, b) == asc) { fe_store(keys + lo, b); fe_store(keys + hi, a); }
}
__global__ void k_sort_small(Fr* keys, uint32_t n) {   // n < SORT_TILE: single thread block, global memory
  for (uint32_t k = 2; k <= n; k <<= 1)
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
        const uint32_t lo = ((t & ~(j - 1)) << 1) | (t & (j - 1)), hi = lo + j;
        const bool asc = ((lo & k) == 0);
        Fr a = fe_load(keys + lo), b = fe_load(keys + hi);
        if (u256_gt(a, b) == asc) { fe_store(keys + lo, b); fe_store(keys + hi, a); }
      }
      __syncthreads();
    }
}
int sort_u256(zkc_ctx* ctx, Fr* keys, uint64_t n) {
  if (n < 2) return ZKC_OK;
  if (n & (n - 1)) return set_err(ctx, ZKC_ERR_BAD_ARG, "sort_u256: n must be a power of two");
  ProfScope _p(ctx, "sort_u256");
  cudaStream_t st = ctx->stream;
  if (n < SORT_TILE) {
    k_sort_small<<<1, 256, 0, st>>>(keys, (uint32_t)n);
    ZKC_LAUNCH_CHECK(ctx);
    return ZKC_OK;
  }
  const unsigned tiles = (unsigned)(n / SORT_TILE);
  k_bitonic_local<<<tiles, 512, 0, st>>>(keys, 2, SORT_TILE, 0);
  ZKC_LAUNCH_CHECK(ctx);
  for (uint64_t k = SORT_TILE * 2; k <= n; k <<= 1) {
    for (uint64_t j = k >> 1; j >= SORT_TILE; j >>= 1) {
      k_bitonic_global<<<(unsigned)(((n >> 1) + 255) / 256), 256, 0, st>>>(keys, n, j, k);
      ZKC_LAUNCH_CHECK(ctx);
    }
    k_bitonic_local<<<tiles, 512, 0, st>>>(keys, k, k, 1);
    ZKC_LAUNCH_CHECK(ctx);
  }
  return ZKC_OK;
}


// ---- counting sort for small keys --------------------------------------------------------------------------------
// Range-check lookups (halo2-base RangeChip: values < 2^lookup_bits, /root/reference/src/bin/cli.rs:421) sort columns whose
// canonical values fit a few dozen bits.  keys[0..U) are canonical values, keys[U..n) the all-ones sentinel written by
// k_lookup_prepare (already in final position).  If every value is below 2^24 the sort is histogram -> scan -> expand;
// otherwise the generic bitonic network runs.  Same output either way.
#define COUNT_SORT_MAX_BITS 24
__global__ void k_sort_probe(const Fr* keys, uint64_t U, uint32_t* probe /* [0] max low limb, [1] OR of the high limbs */) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t lo = 0, hi = 0;
  if (i < U) {
    const Fr v = fe_load(keys + i);
    lo = v.v[0];
#pragma unroll
    for (int j = 1; j < 8; ++j) hi |= v.v[j];
  }
  lo = __reduce_max_sync(0xffffffffu, lo);
  hi = __reduce_or_sync(0xffffffffu, hi);
  if ((threadIdx.x & 31) == 0) { if (lo) atomicMax(probe, lo); if (hi) atomicOr(probe + 1, hi); }
}
__global__ void k_count_hist(const Fr* keys, uint64_t U, uint32_t* hist) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= U) return;
  const uint32_t key = keys[i].v[0];
  const uint32_t peers = __match_any_sync(__activemask(), key);
  if ((threadIdx.x & 31) == (uint32_t)(__ffs(peers) - 1)) atomicAdd(hist + key, (uint32_t)__popc(peers));
}
__global__ void k_count_expand(const uint32_t* offsets, uint32_t bins, Fr* keys, uint64_t U) {
  const uint64_t pos = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= U) return;
  uint32_t lo = 0, hi = bins;          // first bin whose exclusive offset exceeds pos
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (offsets[mid] > pos) hi = mid; else lo = mid + 1;
  }
  Fr v = fe_zero<FrP>();
  v.v[0] = lo - 1;
  fe_store(keys + pos, v);
}
int sort_u256_padded(zkc_ctx* ctx, Fr* keys, uint64_t n, uint64_t U) {
  if (U > n) return set_err(ctx, ZKC_ERR_BAD_ARG, "sort_u256_padded: U > n");
  if (U < 2048 || U >= (1ull << 32)) return sort_u256(ctx, keys, n);
  cudaStream_t st = ctx->stream;
  uint32_t* w;
  ZKC_TRY(scratch_reserve(ctx, SCR_MISC3, 256, (void**)&w));
  uint32_t probe[2];
  {
    ProfScope _p(ctx, "sort_u256");
    ZKC_CUDA_TRY(ctx, cudaMemsetAsync(w, 0, 8, st));
    k_sort_probe<<<(unsigned)((U + 255) / 256), 256, 0, st>>>(keys, U, w); ZKC_LAUNCH_CHECK(ctx);
    ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(probe, w, 8, cudaMemcpyDeviceToHost, st));
    ZKC_CUDA_TRY(ctx, cudaStreamSynchronize(st));
  }
  if (probe[1] || probe[0] >= (1u << COUNT_SORT_MAX_BITS)) return sort_u256(ctx, keys, n);
  ProfScope _p(ctx, "sort_u256");
  const uint32_t bins = probe[0] + 1;
  ZKC_TRY(scratch_reserve(ctx, SCR_MISC3, 256 + 2 * (size_t)bins * 4, (void**)&w));
  uint32_t* hist = w + 64; uint32_t* offsets = hist + bins;
  ZKC_CUDA_TRY(ctx, cudaMemsetAsync(hist, 0, (size_t)bins * 4, st));
  k_count_hist<<<(unsigned)((U + 255) / 256), 256, 0, st>>>(keys, U, hist); ZKC_LAUNCH_CHECK(ctx);
  ZKC_TRY(u32_scan(ctx, hist, offsets, bins, nullptr));
  k_count_expand<<<(unsigned)((U + 255) / 256), 256, 0, st>>>(offsets, bins, keys, U); ZKC_LAUNCH_CHECK(ctx);
  return ZKC_OK;
}

}  // namespace zkc
