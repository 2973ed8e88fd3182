// Radix-2^s Fr NTT for sm_100a: replaces halo2_proofs::arithmetic::best_fft and the
// EvaluationDomain conversions built on it (halo2_proofs 0.2.0 @4b42325 src/arithmetic.rs,
// src/poly/domain.rs — un-vendored, pinned at /root/reference/Cargo.lock:1320-1336; SURVEY.md §8a
// rows a4/a5, Appendix A.4).  Same I/O convention as upstream: natural order in, natural order out.
//
// Decomposition (four-step, recursive): N = L1 * L2 [* L3].  Every pass stages a tile of
// 2^s x C elements in shared memory as two 16-byte planes (limbs 0-3 / limbs 4-7) so that a warp
// touching consecutive elements is bank-conflict free, runs s in-place DIF stages there, and writes
// the tile back with the in-tile bit reversal undone for free:
//   * strided pass  : tile = all L rows (stride B) x C contiguous columns; in place; the result is
//                     multiplied by the inter-pass twiddle w_{L*B}^(b*k) on the way out.
//   * last pass     : tile = C rows of L contiguous elements; output is written transposed
//                     (digit-reversed), C contiguous elements per output row, which is what makes
//                     the whole transform natural-order without a separate permutation pass.
// Coset scaling (zeta^(i mod 3), EvaluationDomain::distribute_powers_zeta), zero padding to the
// extended domain and the 1/n scaling of inverse transforms are fused into the first load / last
// store.  One twiddle table w_N^i (i < N/2) per size serves forward and inverse transforms.
#include "common.cuh"
#include <algorithm>
#include <cstdlib>

namespace zkc {

struct NttPass {
  const Fr* src;
  Fr* dst;
  uint64_t src_stride, dst_stride;  // per-column strides in elements
  uint64_t n_in;                    // valid input elements per column (rest read as zero)
  const Fr* tw;                     // w_N^i, i < N/2
  uint32_t log_n, s, logC;
  uint32_t logB;          // strided pass: inner stride
  uint32_t logN1, logN2;  // last pass: rows A = N1*N2, output index = k1 + N1*k2 + A*kp
  int inverse, pre, post;
  Fr pre1, pre2;          // pre == 1: in[i] *= pre^(i mod 3)
  // residue classes of the extended coset (dom_coeff_to_classes): blockIdx.y = column * cls_div + class slot.  The first pass
  // reads column y / cls_div and multiplies element g by pre_tab[class * N + g] (pre == 2); the last pass writes to
  // dst + column * dst_stride + class * N.  cls_div == 0: plain batches (blockIdx.y = column on both sides).
  uint32_t src_cls_div, dst_cls_div, cls0;
  const Fr* pre_tab;
  const Fr* post_tab;     // post == 2: out[o] *= post_tab[(cls0 + blockIdx.y) * N + o]
  Fr post0, post1, post2; // out[i] *= post[i mod 3]
};

__device__ __forceinline__ Fr tw_get(const Fr* tw, uint64_t e, uint32_t log_n, int inverse) {
  const uint64_t N = 1ull << log_n;
  if (inverse) e = (N - e) & (N - 1);
  const uint64_t half = N >> 1;
  const bool neg = e >= half;
  Fr w = fe_load_nc(tw + (neg ? e - half : e));
  return neg ? fe_neg(w) : w;
}

__device__ __forceinline__ void butterfly_dif(uint4* lo, uint4* hi, uint32_t i0, uint32_t i1, const Fr* tw, uint64_t e,
                                              uint32_t log_n, int inverse) {
  Fr a = fe_from_halves<FrP>(lo[i0], hi[i0]);
  Fr b = fe_from_halves<FrP>(lo[i1], hi[i1]);
  Fr sum = fe_add(a, b);
  Fr diff = fe_sub(a, b);
  if (e != 0) diff = fe_mul(diff, tw_get(tw, e, log_n, inverse));
  lo[i0] = fe_lo(sum); hi[i0] = fe_hi(sum);
  lo[i1] = fe_lo(diff); hi[i1] = fe_hi(diff);
}

__device__ __forceinline__ const Fr* pass_src(const NttPass& p) {
  return p.src + (uint64_t)(p.src_cls_div ? blockIdx.y / p.src_cls_div : blockIdx.y) * p.src_stride;
}
__device__ __forceinline__ Fr* pass_dst(const NttPass& p) {
  if (!p.dst_cls_div) return p.dst + (uint64_t)blockIdx.y * p.dst_stride;
  return p.dst + (uint64_t)(blockIdx.y / p.dst_cls_div) * p.dst_stride + ((uint64_t)(p.cls0 + blockIdx.y % p.dst_cls_div) << p.log_n);
}
__device__ __forceinline__ Fr pass_pre(const NttPass& p, Fr x, uint64_t g) {
  if (p.pre == 1) {
    const uint32_t m = (uint32_t)(g % 3);
    if (m == 1) x = fe_mul(x, p.pre1); else if (m == 2) x = fe_mul(x, p.pre2);
  } else if (p.pre == 2) {
    x = fe_mul(x, fe_load_nc(p.pre_tab + ((uint64_t)(p.cls0 + blockIdx.y % p.src_cls_div) << p.log_n) + g));
  }
  return x;
}

#define NTT_THREADS 256
#define NTT_PLANE_PAD 4  // uint4 units: offsets the high plane by 64 B so paired accesses hit distinct banks

__global__ void __launch_bounds__(NTT_THREADS) k_ntt_strided(NttPass p) {
  extern __shared__ uint4 sm[];
  const uint32_t L = 1u << p.s, C = 1u << p.logC, T = L << p.logC;
  uint4* lo = sm;
  uint4* hi = sm + T + NTT_PLANE_PAD;
  const Fr* src = pass_src(p);
  Fr* dst = pass_dst(p);
  const uint32_t tiles_per_a = 1u << (p.logB - p.logC);
  const uint64_t a = blockIdx.x >> (p.logB - p.logC);
  const uint64_t b0 = (uint64_t)(blockIdx.x & (tiles_per_a - 1)) << p.logC;
  const uint64_t base = ((a << p.s) << p.logB) + b0;

  for (uint32_t e = threadIdx.x; e < T; e += NTT_THREADS) {
    const uint32_t l = e >> p.logC, c = e & (C - 1);
    const uint64_t g = base + ((uint64_t)l << p.logB) + c;
    Fr x = fe_zero<FrP>();
    if (g < p.n_in) {
      x = fe_load(src + g);
      if (p.pre) x = pass_pre(p, x, g);
    }
    lo[e] = fe_lo(x); hi[e] = fe_hi(x);
  }
  __syncthreads();

  for (uint32_t logh = p.s; logh-- > 0;) {
    const uint32_t h = 1u << logh;
    for (uint32_t bt = threadIdx.x; bt < (T >> 1); bt += NTT_THREADS) {
      const uint32_t c = bt & (C - 1), pr = bt >> p.logC;
      const uint32_t j = pr & (h - 1);
      const uint32_t i = ((pr - j) << 1) | j;
      const uint32_t i0 = (i << p.logC) + c;
      butterfly_dif(lo, hi, i0, i0 + (h << p.logC), p.tw, (uint64_t)j << (p.log_n - 1 - logh), p.log_n, p.inverse);
    }
    __syncthreads();
  }

  const uint32_t tw_shift = p.log_n - p.s - p.logB;
  for (uint32_t e = threadIdx.x; e < T; e += NTT_THREADS) {
    const uint32_t k = e >> p.logC, c = e & (C - 1);
    const uint32_t row = __brev(k) >> (32 - p.s);
    Fr x = fe_from_halves<FrP>(lo[(row << p.logC) + c], hi[(row << p.logC) + c]);
    const uint64_t ex = ((b0 + c) * (uint64_t)k) << tw_shift;
    if (ex != 0) x = fe_mul(x, tw_get(p.tw, ex, p.log_n, p.inverse));
    fe_store(dst + base + ((uint64_t)k << p.logB) + c, x);
  }
}

__global__ void __launch_bounds__(NTT_THREADS) k_ntt_last(NttPass p) {
  extern __shared__ uint4 sm[];
  const uint32_t L = 1u << p.s, C = 1u << p.logC, T = L << p.logC;
  const uint32_t pitch = L + (C > 1 ? 1u : 0u);
  uint4* lo = sm;
  uint4* hi = sm + pitch * C + NTT_PLANE_PAD;
  const Fr* src = pass_src(p);
  Fr* dst = pass_dst(p);
  const uint64_t N2 = 1ull << p.logN2;
  const uint64_t kb = blockIdx.x >> p.logN2, k2 = blockIdx.x & (N2 - 1);

  for (uint32_t e = threadIdx.x; e < T; e += NTT_THREADS) {
    const uint32_t c = e >> p.s, i = e & (L - 1);
    const uint64_t row = (((kb << p.logC) + c) << p.logN2) + k2;
    const uint64_t g = (row << p.s) + i;
    Fr x = fe_zero<FrP>();
    if (g < p.n_in) {
      x = fe_load(src + g);
      if (p.pre) x = pass_pre(p, x, g);
    }
    lo[c * pitch + i] = fe_lo(x); hi[c * pitch + i] = fe_hi(x);
  }
  __syncthreads();

  for (uint32_t logh = p.s; logh-- > 0;) {
    const uint32_t h = 1u << logh;
    for (uint32_t bt = threadIdx.x; bt < (T >> 1); bt += NTT_THREADS) {
      // columns fastest: the 8 lanes of a quarter-warp touch 8 rows of odd pitch — distinct 16-byte banks at every stage (with
      // the butterfly index fastest the last three stages, h < 8, were two-way conflicted) — and share one twiddle
      const uint32_t c = bt & (C - 1), pr = bt >> p.logC;
      const uint32_t j = pr & (h - 1);
      const uint32_t i = ((pr - j) << 1) | j;
      const uint32_t i0 = c * pitch + i;
      butterfly_dif(lo, hi, i0, i0 + h, p.tw, (uint64_t)j << (p.log_n - 1 - logh), p.log_n, p.inverse);
    }
    __syncthreads();
  }

  const uint32_t logA = p.logN1 + p.logN2;
  for (uint32_t e = threadIdx.x; e < T; e += NTT_THREADS) {
    const uint32_t c = e & (C - 1), kp = e >> p.logC;
    const uint32_t pos = __brev(kp) >> (32 - p.s);
    Fr x = fe_from_halves<FrP>(lo[c * pitch + pos], hi[c * pitch + pos]);
    const uint64_t o = ((kb << p.logC) + c) + (k2 << p.logN1) + ((uint64_t)kp << logA);
    if (p.post == 1) {
      const uint32_t m = (uint32_t)(o % 3);
      x = fe_mul(x, m == 0 ? p.post0 : (m == 1 ? p.post1 : p.post2));
    } else if (p.post == 2) {   // per-transform table: transform blockIdx.y is residue class cls0 + blockIdx.y
      x = fe_mul(x, fe_load_nc(p.post_tab + ((uint64_t)(p.cls0 + blockIdx.y) << p.log_n) + o));
    }
    fe_store(dst + o, x);
  }
}

// tw[i] = omega^i, i < half_n; each thread fills a run of 64 entries.
__global__ void k_gen_twiddles(Fr* tw, Fr omega, uint64_t half_n) {
  const uint64_t start = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 64;
  if (start >= half_n) return;
  Fr w = fe_pow_u64(omega, start);
  const uint64_t end = start + 64 < half_n ? start + 64 : half_n;
  for (uint64_t i = start; i < end; ++i) { fe_store(tw + i, w); w = fe_mul(w, omega); }
}

__global__ void k_scale_periodic(Fr* a, const Fr* t, uint32_t mask, uint64_t row0, uint64_t cnt) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= cnt) return;
  const uint64_t i = row0 + tid;
  fe_store(a + i, fe_mul(fe_load(a + i), fe_load_nc(t + (i & mask))));
}

// pre[c * n + i] = scale * zeta^(i mod 3) * w^(c * i)  (w = extended_omega, scale = 1; or their inverses and scale = 1/n): the coefficient scaling that turns a size-n transform into
// the evaluations on residue class c of the extended coset.  Each thread fills a run of 64 entries of one class.
__global__ void k_class_pre(Fr* pre, Fr w_ext, Fr zeta, Fr scale, uint32_t log_n, uint32_t ncls) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t runs = 1ull << (log_n > 6 ? log_n - 6 : 0), n = 1ull << log_n;
  if (t >= runs * ncls) return;
  const uint64_t c = t / runs, start = (t % runs) * 64;
  const Fr wc = fe_pow_u64(w_ext, c);
  Fr cur = fe_mul(fe_pow_u64(wc, start), scale);
  const Fr z1 = zeta, z2 = fe_sqr(zeta);
  const uint64_t end = start + 64 < n ? start + 64 : n;
  for (uint64_t i = start; i < end; ++i) {
    const uint32_t m = (uint32_t)(i % 3);
    fe_store(pre + c * n + i, m == 0 ? cur : fe_mul(cur, m == 1 ? z1 : z2));
    cur = fe_mul(cur, wc);
  }
}
// class-major <-> natural order of extended-coset rows: natural row c + 2^e * m  <->  class-major row c * n + m
__global__ void k_classes_to_natural(const Fr* cm, Fr* nat, uint32_t log_n, uint32_t e, uint64_t rows) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;   // natural index: coalesced stores
  if (i >= rows) return;
  const uint64_t c = i & ((1ull << e) - 1), m = i >> e;
  fe_store(nat + i, fe_load(cm + (c << log_n) + m));
}
__global__ void k_natural_to_classes(const Fr* nat, Fr* cm, uint32_t log_n, uint32_t e, uint64_t rows) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;   // class-major index
  if (i >= rows) return;
  const uint64_t c = i >> log_n, m = i & ((1ull << log_n) - 1);
  fe_store(cm + i, fe_load(nat + c + (m << e)));
}
// class-major rows [row0, row0 + cnt): a[i] *= t[class of i]
__global__ void k_scale_by_class(Fr* a, const Fr* t, uint32_t log_n, uint64_t row0, uint64_t cnt) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= cnt) return;
  const uint64_t i = row0 + tid;
  fe_store(a + i, fe_mul(fe_load(a + i), fe_load_nc(t + (i >> log_n))));
}

static int get_twiddles(zkc_ctx* ctx, uint32_t log_n, const Fr** out) {
  auto it = ctx->twiddles.find(log_n);
  if (it != ctx->twiddles.end()) { *out = it->second; return ZKC_OK; }
  const uint64_t half = log_n == 0 ? 1 : (1ull << (log_n - 1));
  Fr* tw = nullptr;
  ZKC_CUDA_TRY(ctx, cudaMalloc(&tw, half * sizeof(Fr)));
  const Fr omega = fr_root_of_unity(log_n);
  const uint64_t threads = (half + 63) / 64;
  k_gen_twiddles<<<(unsigned)((threads + 127) / 128), 128, 0, ctx->stream>>>(tw, omega, half);
  ZKC_LAUNCH_CHECK(ctx);
  // tables are shared by both streams of the ctx: make the one-time generation visible to either
  ZKC_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->twiddles[log_n] = tw;
  *out = tw;
  return ZKC_OK;
}

int ntt_twiddles(zkc_ctx* ctx, uint32_t log_n, const Fr** out) { return get_twiddles(ctx, log_n, out); }

struct NttOpts {
  int inverse = 0;
  uint64_t n_in = 0;      // 0 = full
  int pre = 0; Fr pre1, pre2;
  int post = 0; Fr post0, post1, post2;
  // residue-class batches: every source column is transformed once per class in [cls0, cls0 + ncls), pre-scaled by
  // pre_tab[class * N + i]; class c of column j lands at dst + j * dst_stride + c * N
  uint32_t ncls = 0, cls0 = 0;
  const Fr* pre_tab = nullptr;
  const Fr* post_tab = nullptr;   // post == 2 (plain batches: transform y is class cls0 + y)
};

static const uint32_t LOG_TILE = 11;  // 2^11 elements = 64 KiB of shared memory per CTA
#define NTT_TWO_PASS_MAX 20   // B200 sweep (tools/nttsweep.py): 2^19-2^20 gain 4-5 % from the saved pass, 2^21-2^22 break even

static int launch_pass(zkc_ctx* ctx, bool last, const NttPass& p, uint32_t grid_x, uint32_t ncols) {

  const uint32_t L = 1u << p.s, C = 1u << p.logC;
  size_t smem = last ? (size_t)(2 * (L + (C > 1 ? 1 : 0)) * C + NTT_PLANE_PAD) * sizeof(uint4)
                     : (size_t)(2 * L * C + NTT_PLANE_PAD) * sizeof(uint4);
  {
    ZKC_CUDA_TRY(ctx, cudaFuncSetAttribute(last ? (const void*)k_ntt_last : (const void*)k_ntt_strided,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));

  }
  dim3 grid(grid_x, ncols);
  {
    // field products of this pass (the roofline numerator of bench.py): N/2 per butterfly stage, N inter-pass twiddles after a
    // strided pass, N for a fused post-scaling, ~2/3 n_in for the coset pre-scaling
    const uint64_t N = 1ull << p.log_n;
    ctx->stats["ntt.muls"] += (uint64_t)ncols * ((N >> 1) * p.s + (last ? 0 : N) + (p.post ? N : 0) + (p.pre == 1 ? (p.n_in * 2) / 3 : (p.pre == 2 ? p.n_in : 0)));
    ctx->stats["ntt.bytes"] += (uint64_t)ncols * (std::min<uint64_t>(p.n_in, N) + N) * sizeof(Fr);   // read the valid inputs, write N
  }
  ProfScope _p(ctx, last ? "ntt.last" : "ntt.strided");
  if (last) k_ntt_last<<<grid, NTT_THREADS, smem, ctx->stream>>>(p);
  else k_ntt_strided<<<grid, NTT_THREADS, smem, ctx->stream>>>(p);
  ZKC_LAUNCH_CHECK(ctx);
  return ZKC_OK;
}

// Transform `ncols` columns: src (column stride src_stride, n_in valid elements) -> dst (stride
// dst_stride, N elements).  src == dst (with equal strides) is allowed.
int ntt_run(zkc_ctx* ctx, const Fr* src, uint64_t src_stride, Fr* dst, uint64_t dst_stride, uint32_t log_n, uint32_t ncols,
            const NttOpts& o) {
  if (log_n == 0 || ncols == 0) {
    if (log_n == 0 && src != dst && ncols) {
      ZKC_CUDA_TRY(ctx, cudaMemcpy2DAsync(dst, dst_stride * sizeof(Fr), src, src_stride * sizeof(Fr), sizeof(Fr), ncols,
                                          cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return ZKC_OK;
  }
  if (log_n > 27) return set_err(ctx, ZKC_ERR_BAD_ARG, "ntt: log_n > 27 unsupported");
  const uint64_t N = 1ull << log_n;
  const uint32_t mult = o.ncls ? o.ncls : 1;   // transforms per source column
  const Fr* tw;
  ZKC_TRY(get_twiddles(ctx, log_n, &tw));
  NttPass base{};
  base.tw = tw; base.log_n = log_n; base.inverse = o.inverse;
  base.n_in = N;
  auto first = [&](NttPass& p) { p.src = src; p.src_stride = src_stride; p.n_in = o.n_in ? o.n_in : N; p.pre = o.pre; p.pre1 = o.pre1; p.pre2 = o.pre2;
                            p.src_cls_div = o.ncls; p.cls0 = o.cls0; p.pre_tab = o.pre_tab; };
  auto final_ = [&](NttPass& p) { p.dst = dst; p.dst_stride = dst_stride; p.post = o.post; p.post0 = o.post0; p.post1 = o.post1; p.post2 = o.post2;
                             p.dst_cls_div = o.ncls; p.cls0 = o.cls0; p.post_tab = o.post_tab; };

  if (log_n <= LOG_TILE) {
    NttPass p = base; first(p); final_(p);
    p.s = log_n; p.logC = 0; p.logN1 = 0; p.logN2 = 0;
    return launch_pass(ctx, true, p, 1, ncols * mult);
  }
  // Two passes up to 2^two_pass_max, three above.  A pass boundary costs one twiddle product per element and one trip
  // through HBM; a tile is 2^11 elements, so two passes of a 2^22 transform read single 32-byte elements at large
  // strides (one DRAM sector each) — affordable because the transform is bound by the integer pipe, not by HBM.
  uint32_t two_pass_max = NTT_TWO_PASS_MAX;
  if (const int v = ctx->tune.ntt_two_pass_max) { if (v <= 2 * (int)LOG_TILE) two_pass_max = (uint32_t)v; }
  // bound the scratch: process columns in chunks of <= 1 GiB
  uint32_t chunk = (uint32_t)std::max<uint64_t>(1, (1ull << 30) / (N * sizeof(Fr) * mult));
  if (chunk > ncols) chunk = ncols;
  Fr* tmp;
  ZKC_TRY(scratch_reserve(ctx, SCR_NTT, (size_t)chunk * mult * N * sizeof(Fr), (void**)&tmp));
  for (uint32_t c0 = 0; c0 < ncols; c0 += chunk) {
    const uint32_t nc = std::min(chunk, ncols - c0) * mult;   // grid.y of this chunk
    const Fr* csrc = src + (uint64_t)c0 * src_stride;
    Fr* cdst = dst + (uint64_t)c0 * dst_stride;
    if (log_n <= two_pass_max) {
      const uint32_t s1 = (log_n + 1) / 2, s2 = log_n - s1;
      NttPass p1 = base; first(p1); p1.src = csrc;
      p1.dst = tmp; p1.dst_stride = N; p1.s = s1; p1.logB = s2; p1.logC = std::min(3u, std::min(LOG_TILE - s1, s2));
      ZKC_TRY(launch_pass(ctx, false, p1, 1u << (s2 - p1.logC), nc));
      NttPass p2 = base; final_(p2); p2.dst = cdst;
      p2.src = tmp; p2.src_stride = N; p2.s = s2; p2.logN1 = s1; p2.logN2 = 0; p2.logC = std::min(3u, std::min(LOG_TILE - s2, s1));
      ZKC_TRY(launch_pass(ctx, true, p2, 1u << (s1 - p2.logC), nc));
    } else {
      const uint32_t s1 = (log_n + 2) / 3, s2 = (log_n - s1 + 1) / 2, s3 = log_n - s1 - s2;
      NttPass p1 = base; first(p1); p1.src = csrc;
      p1.dst = tmp; p1.dst_stride = N; p1.s = s1; p1.logB = s2 + s3; p1.logC = std::min(3u, LOG_TILE - s1);
      ZKC_TRY(launch_pass(ctx, false, p1, 1u << (s2 + s3 - p1.logC), nc));
      NttPass p2 = base;
      p2.src = tmp; p2.src_stride = N; p2.dst = tmp; p2.dst_stride = N; p2.s = s2; p2.logB = s3; p2.logC = std::min(3u, std::min(LOG_TILE - s2, s3));
      ZKC_TRY(launch_pass(ctx, false, p2, 1u << (s1 + s3 - p2.logC), nc));
      NttPass p3 = base; final_(p3); p3.dst = cdst;
      p3.src = tmp; p3.src_stride = N; p3.s = s3; p3.logN1 = s1; p3.logN2 = s2; p3.logC = std::min(3u, std::min(LOG_TILE - s3, s1));
      ZKC_TRY(launch_pass(ctx, true, p3, 1u << (s1 + s2 - p3.logC), nc));
    }
  }
  return ZKC_OK;
}

}  // namespace zkc

using namespace zkc;

// ---- EvaluationDomain --------------------------------------------------------------------------
struct zkc_domain {
  zkc_ctx* ctx;
  uint32_t k, extended_k, j;
  int zeta_choice;
  Fr omega, omega_inv, extended_omega, extended_omega_inv, g_coset, g_coset_inv, ifft_divisor, extended_ifft_divisor;
  Fr* t_inv_dev = nullptr;  // 2^(extended_k - k) inverted vanishing evaluations
  Fr* class_pre = nullptr;  // [class][i < n]: zeta^(i mod 3) * extended_omega^(class * i), built on first use (dom_coeff_to_classes)
  Fr* class_post = nullptr; // [class < j - 1][i < n]: (zeta * extended_omega^class)^-i / n (dom_classes_to_pieces)
  Fr* mix_dev = nullptr;    // (j - 1)^2 inverse Vandermonde entries of dom_classes_to_pieces
};

static inline void fr_to_abi(const Fr& f, zkc_fr* o) { memcpy(o, f.v, 32); }
static inline Fr fr_from_abi(const zkc_fr* i) { Fr f; memcpy(f.v, i, 32); return f; }

extern "C" int zkc_domain_create(zkc_ctx* ctx, uint32_t j, uint32_t k, int zeta_choice, zkc_domain** out) {
  if (!ctx || !out || j < 2 || k > 26) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_domain_create: bad arguments");
  CtxLock lock(ctx);
  zkc_domain* d = new zkc_domain();
  d->ctx = ctx; d->k = k; d->j = j; d->zeta_choice = zeta_choice;
  const uint64_t n = 1ull << k;
  d->extended_k = k;
  while ((1ull << d->extended_k) < n * (j - 1)) d->extended_k++;
  if (d->extended_k > 27) { delete d; return set_err(ctx, ZKC_ERR_BAD_ARG, "extended domain too large"); }
  d->extended_omega = fr_root_of_unity(d->extended_k);
  d->extended_omega_inv = fe_inv(d->extended_omega);
  d->omega = fr_root_of_unity(k);
  d->omega_inv = fe_inv(d->omega);
  d->g_coset = fr_from_raw_words(zeta_choice == 0 ? FR_ZETA_RAW : FR_ZETA_ALT_RAW);
  d->g_coset_inv = fe_sqr(d->g_coset);
  d->ifft_divisor = fe_inv(fr_from_u64(n));
  d->extended_ifft_divisor = fe_inv(fr_from_u64(1ull << d->extended_k));
  const uint32_t tl = 1u << (d->extended_k - k);
  std::vector<Fr> t(tl);
  Fr cur = d->g_coset;
  for (uint32_t i = 0; i < tl; ++i) {
    t[i] = fe_inv(fe_sub(fe_pow_u64(cur, n), fe_one<FrP>()));
    cur = fe_mul(cur, d->extended_omega);
  }
  cudaError_t e = cudaMalloc(&d->t_inv_dev, tl * sizeof(Fr));
  if (e == cudaSuccess) e = cudaMemcpy(d->t_inv_dev, t.data(), tl * sizeof(Fr), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { delete d; return set_err(ctx, ZKC_ERR_CUDA, std::string("zkc_domain_create: ") + cudaGetErrorString(e)); }
  *out = d;
  return ZKC_OK;
}

extern "C" void zkc_domain_free(zkc_domain* d) {
  if (!d) return;
  if (d->t_inv_dev) cudaFree(d->t_inv_dev);
  if (d->class_pre) cudaFree(d->class_pre);
  if (d->class_post) cudaFree(d->class_post);
  if (d->mix_dev) cudaFree(d->mix_dev);
  delete d;
}

extern "C" int zkc_domain_get_info(const zkc_domain* d, zkc_domain_info* o) {
  if (!d || !o) return ZKC_ERR_BAD_ARG;
  o->k = d->k; o->extended_k = d->extended_k; o->j = d->j;
  fr_to_abi(d->omega, &o->omega); fr_to_abi(d->omega_inv, &o->omega_inv);
  fr_to_abi(d->extended_omega, &o->extended_omega); fr_to_abi(d->extended_omega_inv, &o->extended_omega_inv);
  fr_to_abi(d->g_coset, &o->g_coset); fr_to_abi(d->g_coset_inv, &o->g_coset_inv);
  fr_to_abi(d->ifft_divisor, &o->ifft_divisor); fr_to_abi(d->extended_ifft_divisor, &o->extended_ifft_divisor);
  return ZKC_OK;
}

namespace zkc {
// internal entry points used by the prover pipeline as well
int dom_lagrange_to_coeff(zkc_ctx* ctx, const zkc_domain* d, Fr* a, uint32_t ncols) {
  NttOpts o; o.inverse = 1; o.post = 1; o.post0 = o.post1 = o.post2 = d->ifft_divisor;
  return ntt_run(ctx, a, 1ull << d->k, a, 1ull << d->k, d->k, ncols, o);
}
int dom_coeff_to_lagrange(zkc_ctx* ctx, const zkc_domain* d, Fr* a, uint32_t ncols) {
  NttOpts o;
  return ntt_run(ctx, a, 1ull << d->k, a, 1ull << d->k, d->k, ncols, o);
}
int dom_coeff_to_extended(zkc_ctx* ctx, const zkc_domain* d, const Fr* in, uint64_t in_stride, Fr* out, uint32_t ncols) {
  NttOpts o; o.n_in = 1ull << d->k; o.pre = 1; o.pre1 = d->g_coset; o.pre2 = fe_sqr(d->g_coset);
  return ntt_run(ctx, in, in_stride, out, 1ull << d->extended_k, d->extended_k, ncols, o);
}
int dom_extended_to_coeff(zkc_ctx* ctx, const zkc_domain* d, Fr* a, uint32_t ncols) {
  NttOpts o; o.inverse = 1; o.post = 1;
  o.post0 = d->extended_ifft_divisor;
  o.post1 = fe_mul(d->extended_ifft_divisor, d->g_coset_inv);
  o.post2 = fe_mul(d->extended_ifft_divisor, fe_sqr(d->g_coset_inv));
  const uint64_t en = 1ull << d->extended_k;
  ZKC_TRY(ntt_run(ctx, a, en, a, en, d->extended_k, ncols, o));
  // upstream truncates to n*(j-1) coefficients; the buffer keeps its size, so zero the tail
  const uint64_t keep = (1ull << d->k) * (d->j - 1);
  if (keep < en) {
    // per column: a 2-D memset's pitch is limited to cudaDeviceProp::memPitch (2^31 - 1), which en * 32 bytes exceeds from extended_k = 26
    for (uint32_t c = 0; c < ncols; ++c)
      ZKC_CUDA_TRY(ctx, cudaMemsetAsync(a + (uint64_t)c * en + keep, 0, (en - keep) * sizeof(Fr), ctx->stream));
  }
  return ZKC_OK;
}
// ---- residue classes of the extended coset --------------------------------------------------------------------------------------
// The extended coset { zeta * w_ext^r } splits into 2^e classes r = c + 2^e * m (e = extended_k - k), and class c is the coset
// (zeta * w_ext^c) * <omega> of the size-n subgroup: p on class c is ONE size-n transform of the coefficients scaled by
// zeta^(i mod 3) * w_ext^(c i).  This is the four-step decomposition of the size-2^extended_k transform with N1 = n, N2 = 2^e,
// specialised to an input whose upper (2^e - 1) n coefficients are zero: the N2-point column transforms have one non-zero input
// each (nothing to compute), the twiddle step is the pre-scaling, and the N1-point row transforms are the classes.  Rotations by
// omega stay inside a class (row m -> m + 1), so h(X) is evaluated class by class and a team of GPUs needs no exchange of coset
// rows at all: rank j transforms the classes its row block touches from the (replicated) coefficient forms.
// CLASS-MAJOR layout: class c of a column occupies [c * n, (c + 1) * n).
int dom_class_pre(zkc_ctx* ctx, const zkc_domain* d, const Fr** out) {
  zkc_domain* dm = const_cast<zkc_domain*>(d);
  if (!dm->class_pre) {
    const uint32_t ncls = 1u << (d->extended_k - d->k);
    Fr* p;
    ZKC_CUDA_TRY(ctx, cudaMalloc(&p, (sizeof(Fr) << d->k) * ncls));
    const uint64_t threads = (1ull << (d->k > 6 ? d->k - 6 : 0)) * ncls;
    k_class_pre<<<(unsigned)((threads + 127) / 128), 128, 0, ctx->stream>>>(p, d->extended_omega, d->g_coset, fe_one<FrP>(), d->k, ncls);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);   // shared by both streams of the ctx
    if (e != cudaSuccess) { cudaFree(p); return set_err(ctx, ZKC_ERR_CUDA, std::string("dom_class_pre: ") + cudaGetErrorString(e)); }
    dm->class_pre = p;
  }
  *out = dm->class_pre;
  return ZKC_OK;
}
// classes [c0, c1) of `ncols` coefficient-form columns (stride in_stride) -> out (column stride 2^extended_k, class-major)
int dom_coeff_to_classes(zkc_ctx* ctx, const zkc_domain* d, const Fr* in, uint64_t in_stride, Fr* out, uint32_t ncols, uint32_t c0, uint32_t c1) {
  if (c1 <= c0 || !ncols) return ZKC_OK;
  NttOpts o; o.pre = 2; o.ncls = c1 - c0; o.cls0 = c0;
  ZKC_TRY(dom_class_pre(ctx, d, &o.pre_tab));
  return ntt_run(ctx, in, in_stride, out, 1ull << d->extended_k, d->k, ncols, o);
}
int dom_classes_to_natural(zkc_ctx* ctx, const zkc_domain* d, const Fr* cm, Fr* nat) {
  const uint64_t en = 1ull << d->extended_k;
  k_classes_to_natural<<<(unsigned)((en + 255) / 256), 256, 0, ctx->stream>>>(cm, nat, d->k, d->extended_k - d->k, en);
  ZKC_LAUNCH_CHECK(ctx);
  return ZKC_OK;
}
int dom_natural_to_classes(zkc_ctx* ctx, const zkc_domain* d, const Fr* nat, Fr* cm) {
  const uint64_t en = 1ull << d->extended_k;
  k_natural_to_classes<<<(unsigned)((en + 255) / 256), 256, 0, ctx->stream>>>(nat, cm, d->k, d->extended_k - d->k, en);
  ZKC_LAUNCH_CHECK(ctx);
  return ZKC_OK;
}
// ---- the way back: h(X) from its values on q = j - 1 residue classes ----------------------------------------------------------
// h has fewer than q n coefficients: h = sum_{i1 < q} X^(n i1) H_i1 with deg H_i1 < n (the pieces that get committed).  On class c
// X^n is the constant tau_c = (zeta w_ext^c)^n, so h restricted to class c is the degree < n polynomial g_c = sum_i1 tau_c^i1 H_i1:
// q classes determine h (upstream evaluates on all 2^e >= q of them only because its FFT wants a power of two).  So:
//   g_c   = coset-iNTT of class c     (size-n inverse transform, then (zeta w_ext^c)^-i / n: the fused post table)
//   H_i1  = sum_c (V^-1)[i1][c] g_c   (V[c][i1] = tau_c^i1: a q x q Vandermonde matrix, inverted once on the host)
// — exact, so the pieces are upstream's extended_to_coeff output coefficient for coefficient.
#define MIX_MAX_Q 8
__global__ void k_mix_classes(const Fr* g, Fr* out, const Fr* minv, uint32_t q, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr v[MIX_MAX_Q];
  for (uint32_t c = 0; c < q; ++c) v[c] = fe_load(g + (uint64_t)c * n + i);
  for (uint32_t p = 0; p < q; ++p) {
    Fr acc = fe_mul(v[0], fe_load_nc(minv + p * q));
    for (uint32_t c = 1; c < q; ++c) acc = fe_add(acc, fe_mul(v[c], fe_load_nc(minv + p * q + c)));
    fe_store(out + (uint64_t)p * n + i, acc);
  }
}
static int dom_pieces_setup(zkc_ctx* ctx, const zkc_domain* d) {
  zkc_domain* dm = const_cast<zkc_domain*>(d);
  if (dm->class_post) return ZKC_OK;
  const uint32_t q = d->j - 1;
  if (q > MIX_MAX_Q || q > (1u << (d->extended_k - d->k))) return set_err(ctx, ZKC_ERR_BAD_ARG, "dom_classes_to_pieces: degree too large");
  // inverse of V[c][p] = tau_c^p by Gauss-Jordan over Fr (q <= 8)
  const uint64_t n = 1ull << d->k;
  std::vector<Fr> a((size_t)q * 2 * q);
  Fr zc = d->g_coset;
  for (uint32_t c = 0; c < q; ++c) {
    const Fr tau = fe_pow_u64(zc, n);
    Fr pw = fe_one<FrP>();
    for (uint32_t p = 0; p < q; ++p) { a[(size_t)c * 2 * q + p] = pw; a[(size_t)c * 2 * q + q + p] = p == c ? fe_one<FrP>() : fe_zero<FrP>(); pw = fe_mul(pw, tau); }
    zc = fe_mul(zc, d->extended_omega);
  }
  for (uint32_t col = 0; col < q; ++col) {
    uint32_t piv = col;
    while (piv < q && fe_is_zero(a[(size_t)piv * 2 * q + col])) ++piv;
    if (piv == q) return set_err(ctx, ZKC_ERR_BAD_ARG, "dom_classes_to_pieces: singular class matrix");
    if (piv != col) for (uint32_t t = 0; t < 2 * q; ++t) std::swap(a[(size_t)piv * 2 * q + t], a[(size_t)col * 2 * q + t]);
    const Fr inv = fe_inv(a[(size_t)col * 2 * q + col]);
    for (uint32_t t = 0; t < 2 * q; ++t) a[(size_t)col * 2 * q + t] = fe_mul(a[(size_t)col * 2 * q + t], inv);
    for (uint32_t r = 0; r < q; ++r) {
      if (r == col) continue;
      const Fr f = a[(size_t)r * 2 * q + col];
      if (fe_is_zero(f)) continue;
      for (uint32_t t = 0; t < 2 * q; ++t) a[(size_t)r * 2 * q + t] = fe_sub(a[(size_t)r * 2 * q + t], fe_mul(f, a[(size_t)col * 2 * q + t]));
    }
  }
  std::vector<Fr> minv((size_t)q * q);
  for (uint32_t p = 0; p < q; ++p) for (uint32_t c = 0; c < q; ++c) minv[(size_t)p * q + c] = a[(size_t)p * 2 * q + q + c];
  Fr *mix = nullptr, *post = nullptr;
  ZKC_CUDA_TRY(ctx, cudaMalloc(&mix, minv.size() * sizeof(Fr)));
  cudaError_t e = cudaMemcpy(mix, minv.data(), minv.size() * sizeof(Fr), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&post, (sizeof(Fr) << d->k) * q);
  if (e == cudaSuccess) {
    const uint64_t threads = (1ull << (d->k > 6 ? d->k - 6 : 0)) * q;
    k_class_pre<<<(unsigned)((threads + 127) / 128), 128, 0, ctx->stream>>>(post, d->extended_omega_inv, d->g_coset_inv, d->ifft_divisor, d->k, q);
    ctx->launches++;
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  }
  if (e != cudaSuccess) { cudaFree(mix); if (post) cudaFree(post); return set_err(ctx, ZKC_ERR_CUDA, std::string("dom_classes_to_pieces: ") + cudaGetErrorString(e)); }
  dm->mix_dev = mix; dm->class_post = post;
  return ZKC_OK;
}
// values of h on classes 0 .. q-1 (class-major, `vals`: q * n, destroyed) -> the q pieces of h in coefficient form (`out`: q * n),
// in two steps so that a team can deal the classes: g_c for classes [c0, c1) in place, then the mix over all q classes
int dom_classes_inverse(zkc_ctx* ctx, const zkc_domain* d, Fr* vals, uint32_t c0, uint32_t c1) {
  ZKC_TRY(dom_pieces_setup(ctx, d));
  if (c1 <= c0) return ZKC_OK;
  const uint64_t n = 1ull << d->k;
  NttOpts o; o.inverse = 1; o.post = 2; o.post_tab = d->class_post; o.cls0 = c0;
  return ntt_run(ctx, vals + (uint64_t)c0 * n, n, vals + (uint64_t)c0 * n, n, d->k, c1 - c0, o);
}
int dom_classes_mix(zkc_ctx* ctx, const zkc_domain* d, const Fr* g, Fr* out) {
  ZKC_TRY(dom_pieces_setup(ctx, d));
  const uint32_t q = d->j - 1;
  const uint64_t n = 1ull << d->k;
  k_mix_classes<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(g, out, d->mix_dev, q, n);
  ZKC_LAUNCH_CHECK(ctx);
  ctx->stats["ntt.muls"] += (uint64_t)q * q * n;
  return ZKC_OK;
}
int dom_classes_to_pieces(zkc_ctx* ctx, const zkc_domain* d, Fr* vals, Fr* out) {
  ZKC_TRY(dom_classes_inverse(ctx, d, vals, 0, d->j - 1));
  return dom_classes_mix(ctx, d, vals, out);
}
// division by X^n - 1 on class-major rows [row0, row0 + cnt): the vanishing polynomial is constant on a class
int dom_divide_by_vanishing_classes(zkc_ctx* ctx, const zkc_domain* d, Fr* a, uint64_t row0, uint64_t cnt) {
  if (cnt == 0) return ZKC_OK;
  k_scale_by_class<<<(unsigned)((cnt + 255) / 256), 256, 0, ctx->stream>>>(a, d->t_inv_dev, d->k, row0, cnt);
  ZKC_LAUNCH_CHECK(ctx);
  return ZKC_OK;
}
int dom_divide_by_vanishing(zkc_ctx* ctx, const zkc_domain* d, Fr* a, uint64_t row0, uint64_t cnt) {   // rows [row0, row0 + cnt)
  if (cnt == 0) return ZKC_OK;
  k_scale_periodic<<<(unsigned)((cnt + 255) / 256), 256, 0, ctx->stream>>>(a, d->t_inv_dev, (1u << (d->extended_k - d->k)) - 1, row0, cnt);
  ZKC_LAUNCH_CHECK(ctx);
  return ZKC_OK;
}

// best_fft with an arbitrary root: only omega = canonical root or its inverse occur upstream
int fft_any(zkc_ctx* ctx, Fr* a, const Fr& omega, uint32_t log_n, uint32_t ncols) {
  if (log_n > 27) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_fft_fr: log_n > 27");
  const Fr root = fr_root_of_unity(log_n);
  NttOpts o;
  if (fe_eq(omega, root)) o.inverse = 0;
  else if (fe_eq(fe_mul(omega, root), fe_one<FrP>())) o.inverse = 1;
  else return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_fft_fr: omega is not the domain generator or its inverse");
  return ntt_run(ctx, a, 1ull << log_n, a, 1ull << log_n, log_n, ncols, o);
}
}  // namespace zkc

// host-buffer wrapper: upload, run, download
template <class F>
static int with_host_buffer(zkc_ctx* ctx, const void* in, size_t in_bytes, void* out, size_t out_bytes, size_t dev_bytes, F body) {
  void* d;
  ZKC_TRY(scratch_reserve(ctx, SCR_HOSTIO, dev_bytes, &d));
  if (in_bytes) ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(d, in, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
  ZKC_TRY(body((Fr*)d));
  if (out_bytes) ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(out, d, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  ZKC_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return ZKC_OK;
}

extern "C" int zkc_fft_fr_dev(zkc_ctx* ctx, zkc_fr* a, const zkc_fr* omega, uint32_t log_n, uint32_t ncols) {
  if (!ctx || !a || !omega) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_fft_fr_dev: null argument");
  CtxLock lock(ctx);
  return fft_any(ctx, (Fr*)a, fr_from_abi(omega), log_n, ncols);
}
extern "C" int zkc_fft_fr(zkc_ctx* ctx, zkc_fr* a, const zkc_fr* omega, uint32_t log_n) {
  if (!ctx || !a || !omega) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_fft_fr: null argument");
  CtxLock lock(ctx);
  const size_t bytes = sizeof(Fr) << log_n;
  const Fr w = fr_from_abi(omega);
  return with_host_buffer(ctx, a, bytes, a, bytes, bytes, [&](Fr* d) { return fft_any(ctx, d, w, log_n, 1); });
}

extern "C" int zkc_lagrange_to_coeff_dev(zkc_ctx* ctx, const zkc_domain* d, zkc_fr* a, uint32_t ncols) {
  if (!ctx || !d || !a) return set_err(ctx, ZKC_ERR_BAD_ARG, "null argument");
  CtxLock lock(ctx); return dom_lagrange_to_coeff(ctx, d, (Fr*)a, ncols);
}
extern "C" int zkc_coeff_to_lagrange_dev(zkc_ctx* ctx, const zkc_domain* d, zkc_fr* a, uint32_t ncols) {
  if (!ctx || !d || !a) return set_err(ctx, ZKC_ERR_BAD_ARG, "null argument");
  CtxLock lock(ctx); return dom_coeff_to_lagrange(ctx, d, (Fr*)a, ncols);
}
extern "C" int zkc_coeff_to_extended_dev(zkc_ctx* ctx, const zkc_domain* d, const zkc_fr* in, zkc_fr* out, uint32_t ncols) {
  if (!ctx || !d || !in || !out) return set_err(ctx, ZKC_ERR_BAD_ARG, "null argument");
  CtxLock lock(ctx); return dom_coeff_to_extended(ctx, d, (const Fr*)in, 1ull << d->k, (Fr*)out, ncols);
}
extern "C" int zkc_coeff_to_extended_classes_dev(zkc_ctx* ctx, const zkc_domain* d, const zkc_fr* in, zkc_fr* out, uint32_t ncols, uint32_t c0,
                                                 uint32_t c1) {
  if (!ctx || !d || !in || !out) return set_err(ctx, ZKC_ERR_BAD_ARG, "null argument");
  if (c0 > c1 || c1 > (1u << (d->extended_k - d->k))) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_coeff_to_extended_classes_dev: class range outside [0, 2^(extended_k - k)]");
  CtxLock lock(ctx); return dom_coeff_to_classes(ctx, d, (const Fr*)in, 1ull << d->k, (Fr*)out, ncols, c0, c1);
}
extern "C" int zkc_extended_classes_to_natural_dev(zkc_ctx* ctx, const zkc_domain* d, const zkc_fr* cm, zkc_fr* nat) {
  if (!ctx || !d || !cm || !nat || cm == nat) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_extended_classes_to_natural_dev: null or aliased argument");
  CtxLock lock(ctx); return dom_classes_to_natural(ctx, d, (const Fr*)cm, (Fr*)nat);
}
extern "C" int zkc_extended_to_coeff_dev(zkc_ctx* ctx, const zkc_domain* d, zkc_fr* a, uint32_t ncols) {
  if (!ctx || !d || !a) return set_err(ctx, ZKC_ERR_BAD_ARG, "null argument");
  CtxLock lock(ctx); return dom_extended_to_coeff(ctx, d, (Fr*)a, ncols);
}
extern "C" int zkc_divide_by_vanishing_dev(zkc_ctx* ctx, const zkc_domain* d, zkc_fr* a) {
  if (!ctx || !d || !a) return set_err(ctx, ZKC_ERR_BAD_ARG, "null argument");
  CtxLock lock(ctx); return dom_divide_by_vanishing(ctx, d, (Fr*)a, 0, 1ull << d->extended_k);
}

extern "C" int zkc_lagrange_to_coeff(zkc_ctx* ctx, const zkc_domain* d, zkc_fr* a) {
  if (!ctx || !d || !a) return set_err(ctx, ZKC_ERR_BAD_ARG, "null argument");
  CtxLock lock(ctx);
  const size_t bytes = sizeof(Fr) << d->k;
  return with_host_buffer(ctx, a, bytes, a, bytes, bytes, [&](Fr* dv) { return dom_lagrange_to_coeff(ctx, d, dv, 1); });
}
extern "C" int zkc_coeff_to_lagrange(zkc_ctx* ctx, const zkc_domain* d, zkc_fr* a) {
  if (!ctx || !d || !a) return set_err(ctx, ZKC_ERR_BAD_ARG, "null argument");
  CtxLock lock(ctx);
  const size_t bytes = sizeof(Fr) << d->k;
  return with_host_buffer(ctx, a, bytes, a, bytes, bytes, [&](Fr* dv) { return dom_coeff_to_lagrange(ctx, d, dv, 1); });
}
extern "C" int zkc_coeff_to_extended(zkc_ctx* ctx, const zkc_domain* d, const zkc_fr* coeffs, zkc_fr* out) {
  if (!ctx || !d || !coeffs || !out) return set_err(ctx, ZKC_ERR_BAD_ARG, "null argument");
  CtxLock lock(ctx);
  const size_t nb = sizeof(Fr) << d->k, eb = sizeof(Fr) << d->extended_k;
  void* din;
  ZKC_TRY(scratch_reserve(ctx, SCR_HOSTIO2, nb, &din));
  ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(din, coeffs, nb, cudaMemcpyHostToDevice, ctx->stream));
  return with_host_buffer(ctx, nullptr, 0, out, eb, eb, [&](Fr* dv) { return dom_coeff_to_extended(ctx, d, (const Fr*)din, 1ull << d->k, dv, 1); });
}
extern "C" int zkc_extended_to_coeff(zkc_ctx* ctx, const zkc_domain* d, zkc_fr* a) {
  if (!ctx || !d || !a) return set_err(ctx, ZKC_ERR_BAD_ARG, "null argument");
  CtxLock lock(ctx);
  const size_t eb = sizeof(Fr) << d->extended_k;
  return with_host_buffer(ctx, a, eb, a, eb, eb, [&](Fr* dv) { return dom_extended_to_coeff(ctx, d, dv, 1); });
}
