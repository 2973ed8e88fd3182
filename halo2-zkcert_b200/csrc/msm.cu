// Signed-window Pippenger G1 MSM for sm_100a: replaces halo2_proofs::arithmetic::best_multiexp and
// ParamsKZG::{commit, commit_lagrange} (halo2_proofs 0.2.0 @4b42325 src/arithmetic.rs,
// src/poly/kzg/commitment.rs — un-vendored, pinned at /root/reference/Cargo.lock:1320-1336;
// SURVEY.md §8a rows a3/a6, Appendix A.3).  The sum of [s_i]P_i is a unique group element, so the
// bytes after normalisation equal the CPU prover's whatever the algorithm.
//
// Pipeline (all device-resident, one stream, no host round trip until the final 2 KB):
//   1 k_msm_digits   scalar -> canonical -> signed c-bit digits; histogram of bucket sizes (atomics)
//   2 u32_scan       multi-CTA exclusive scan of the histogram (bucket offsets)
//   3 k_msm_scatter  counting sort: entries (point index | sign, bucket key) grouped by bucket
//   4 k_msm_accum    SEGMENTED FLAT WALK: thread t owns entries [tT, (t+1)T) whatever buckets they
//                    fall in, accumulates with XYZZ mixed adds and flushes one partial per bucket it
//                    touches into slot (key + t).  Work per thread is constant, so bit-, byte- and
//                    limb-valued witness columns (a few giant buckets) run as fast as uniform scalars.
//   5 k_msm_gather   one thread per bucket sums its (usually 1-3) partials; buckets with many
//                    partials are queued and reduced by a whole CTA (k_msm_gather_heavy).
//   6 k_msm_reduce1/2 sum_b b*S_b per window as c independent two-level tree reductions U_t = sum of
//                    buckets whose index has bit t set (no serial running sum), then 2^t weights by Horner.
// With a resident SRS the bases are expanded once into W tables 2^(c*w) * P_i, so all windows of
// all points share ONE bucket set per column and step 6 shrinks by a factor W.
#include "msm.cuh"
#include "dist.cuh"
#include <algorithm>
#include <cstdlib>

namespace zkc {

// signed digits of the canonical scalar; writes dig[(col*W + w)*n + i] = mag | sign << 31
__global__ void k_msm_digits(const Fr* scalars, uint32_t* dig, uint32_t* counts, MsmGeom g) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= g.n * g.ncols) return;
  const uint32_t col = (uint32_t)(idx / g.n);
  const uint64_t i = idx - (uint64_t)col * g.n;
  const Fr s = fe_to_canonical(fe_load(scalars + (uint64_t)col * g.sstride + i));
  uint32_t carry = 0;
  const uint32_t full = 1u << g.c;
  for (uint32_t w = 0; w < g.W; ++w) {
    const uint32_t bit = w * g.c;
    const uint32_t word = bit >> 5, sh = bit & 31;
    uint64_t v = word < 8 ? s.v[word] : 0;
    if (word + 1 < 8) v |= (uint64_t)s.v[word + 1] << 32;
    uint32_t raw = ((uint32_t)(v >> sh) & (full - 1)) + carry;
    uint32_t sign = 0;
    carry = 0;
    if (raw > g.NB) { raw = full - raw; sign = 1; carry = 1; }
    dig[((uint64_t)col * g.W + w) * g.n + i] = raw | (sign << 31);
    if (raw) {
      // warp-aggregated histogram update: lanes that hit the same bucket (bit / byte valued witness columns put
      // most of a warp on one counter) elect a leader that adds their count once
      const uint32_t key = (uint32_t)(((uint64_t)col * g.sets + (g.sets == 1 ? 0 : w)) * g.NB + (raw - 1));
      const uint32_t peers = __match_any_sync(__activemask(), key);
      if ((threadIdx.x & 31) == (uint32_t)(__ffs(peers) - 1)) atomicAdd(counts + key, (uint32_t)__popc(peers));
    }
  }
}

__global__ void k_msm_scatter(const uint32_t* dig, uint32_t* cursor, uint32_t* ent_pt, uint32_t* ent_key, MsmGeom g) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= g.emax()) return;
  const uint32_t d = dig[idx];
  const uint32_t mag = d & 0x7fffffffu;
  if (!mag) return;
  const uint64_t cw = idx / g.n;            // col*W + w
  const uint64_t i = idx - cw * g.n;
  const uint32_t col = (uint32_t)(cw / g.W), w = (uint32_t)(cw - (uint64_t)col * g.W);
  const uint32_t key = (uint32_t)(((uint64_t)col * g.sets + (g.sets == 1 ? 0 : w)) * g.NB + (mag - 1));
  // warp-aggregated slot allocation (same reason as in k_msm_digits)
  const uint32_t peers = __match_any_sync(__activemask(), key);
  const uint32_t lane = threadIdx.x & 31, leader = (uint32_t)(__ffs(peers) - 1);
  uint32_t basepos = 0;
  if (lane == leader) basepos = atomicAdd(cursor + key, (uint32_t)__popc(peers));
  basepos = __shfl_sync(peers, basepos, leader);
  const uint32_t pos = basepos + (uint32_t)__popc(peers & ((1u << lane) - 1));
  const uint64_t pt = g.sets == 1 ? (uint64_t)w * g.bstride + i : i;   // precomputed tables are laid out [w][i]
  ent_pt[pos] = (uint32_t)pt | (d & 0x80000000u);
  ent_key[pos] = (uint32_t)key;
}

// Entries per accumulate thread for THIS batch: g.T when the batch fills the machine, halved (down to 4) while the actual
// number of non-zero digits E leaves fewer than MSM_MIN_THREADS chains — witness columns of bits / bytes / small limbs
// have a fraction of the worst-case entries, and a short batch is bound by the length of the dependent chain, not by work.
#define MSM_MIN_THREADS (148u * 512u)
__device__ __forceinline__ uint32_t msm_T(const MsmGeom& g, uint64_t E) {
  uint32_t T = g.T;
  while (T > 4 && E / T < MSM_MIN_THREADS) T >>= 1;
  return T;
}

__global__ void __launch_bounds__(128) k_msm_accum(const G1Affine* bases, const uint32_t* ent_pt, const uint32_t* ent_key,
                                                   const uint32_t* offsets, G1Xyzz* partial, MsmGeom g) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t E = offsets[g.nbtot()];
  const uint32_t T = msm_T(g, E);
  const uint64_t p0 = t * T;
  if (p0 >= E) return;
  const uint64_t p1 = p0 + T < E ? p0 + T : E;
  G1Xyzz acc = xyzz_identity();
  uint32_t cur = ent_key[p0];
  // software pipeline: the next entry's base point is requested before the current mixed add is issued
  uint32_t e = ent_pt[p0];
  G1Affine q = affine_load_nc(bases + (e & 0x7fffffffu));
  for (uint64_t p = p0; p < p1; ++p) {
    const uint32_t k = ent_key[p];
    const uint32_t e_cur = e;
    const G1Affine q_cur = q;
    if (p + 1 < p1) { e = ent_pt[p + 1]; q = affine_load_nc(bases + (e & 0x7fffffffu)); }
    if (k != cur) { xyzz_store(partial + cur + t, acc); acc = xyzz_identity(); cur = k; }
    if (!affine_is_identity(q_cur)) xyzz_madd(acc, q_cur, (e_cur >> 31) != 0);
  }
  xyzz_store(partial + cur + t, acc);
}

#define MSM_HEAVY_PER_LANE 12
#define MSM_GIANT 4096        // partials: above this a bucket is sliced over MSM_GIANT_SLICES CTAs
#define MSM_GIANT_SLICES 64
#define MSM_GIANT_CAP 512
__device__ __forceinline__ G1Xyzz xyzz_shfl_xor(const G1Xyzz& p, int mask) {
  G1Xyzz r;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    r.x.v[i] = __shfl_xor_sync(0xffffffffu, p.x.v[i], mask);
    r.y.v[i] = __shfl_xor_sync(0xffffffffu, p.y.v[i], mask);
    r.zz.v[i] = __shfl_xor_sync(0xffffffffu, p.zz.v[i], mask);
    r.zzz.v[i] = __shfl_xor_sync(0xffffffffu, p.zzz.v[i], mask);
  }
  return r;
}
__global__ void __launch_bounds__(128) k_msm_gather(const uint32_t* offsets, const G1Xyzz* partial, G1Xyzz* buckets,
                                                    uint32_t* heavy_list, uint32_t* heavy_count, MsmGeom g, uint32_t logG) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t key = tid >> logG;
  const uint32_t G = 1u << logG, lane = (uint32_t)tid & (G - 1);
  const bool valid = key < g.nbtot();     // whole groups are valid or not (nbtot * G is a multiple of the warp's group span)
  uint32_t off = 0, cnt = 0;
  if (valid) { off = offsets[key]; cnt = offsets[key + 1] - off; }
  uint64_t first = 0, np = 0;
  const uint32_t T = msm_T(g, offsets[g.nbtot()]);
  if (cnt) { first = key + off / T; np = key + (off + cnt - 1) / T - first + 1; }
  const bool heavy = np > (uint64_t)MSM_HEAVY_PER_LANE * G;
  G1Xyzz acc = xyzz_identity();
  if (cnt && !heavy)
    for (uint64_t s = lane; s < np; s += G) xyzz_add(acc, xyzz_load(partial + first + s));
  for (uint32_t d = G >> 1; d > 0; d >>= 1) { G1Xyzz o = xyzz_shfl_xor(acc, (int)d); xyzz_add(acc, o); }
  if (valid && lane == 0) {
    if (heavy) {
      // heavy_count[0] / heavy_list[0..nbt): one CTA per bucket; heavy_count[1] / giant list (after nbt): sliced over many CTAs
      if (np > MSM_GIANT && heavy_count[1] < MSM_GIANT_CAP) {
        const uint32_t gi = atomicAdd(heavy_count + 1, 1u);
        if (gi < MSM_GIANT_CAP) heavy_list[g.nbtot() + gi] = (uint32_t)key;
        else heavy_list[atomicAdd(heavy_count, 1u)] = (uint32_t)key;
      } else {
        heavy_list[atomicAdd(heavy_count, 1u)] = (uint32_t)key;
      }
    } else {
      xyzz_store(buckets + key, acc);
    }
  }
}

// block-wide XYZZ tree reduction through shared memory; result valid in thread 0
__device__ __forceinline__ G1Xyzz block_reduce_xyzz(G1Xyzz acc, G1Xyzz* sm) {
  const uint32_t t = threadIdx.x;
  xyzz_store(sm + t, acc);
  __syncthreads();
  for (uint32_t d = blockDim.x >> 1; d > 0; d >>= 1) {
    if (t < d) { G1Xyzz a = xyzz_load(sm + t); xyzz_add(a, xyzz_load(sm + t + d)); xyzz_store(sm + t, a); }
    __syncthreads();
  }
  G1Xyzz r = xyzz_load(sm);
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(128) k_msm_gather_heavy(const uint32_t* offsets, const G1Xyzz* partial, G1Xyzz* buckets,
                                                          const uint32_t* heavy_list, const uint32_t* heavy_count, MsmGeom g) {
  extern __shared__ uint4 smraw[];
  G1Xyzz* sm = reinterpret_cast<G1Xyzz*>(smraw);
  const uint32_t nh = *heavy_count;
  const uint32_t T = msm_T(g, offsets[g.nbtot()]);
  for (uint32_t h = blockIdx.x; h < nh; h += gridDim.x) {
    const uint64_t key = heavy_list[h];
    const uint32_t off = offsets[key], cnt = offsets[key + 1] - off;
    const uint64_t first = key + off / T, last = key + (off + cnt - 1) / T;
    G1Xyzz acc = xyzz_identity();
    for (uint64_t s = first + threadIdx.x; s <= last; s += blockDim.x) xyzz_add(acc, xyzz_load(partial + s));
    G1Xyzz r = block_reduce_xyzz(acc, sm);
    if (threadIdx.x == 0) xyzz_store(buckets + key, r);
  }
}

// giant buckets, pass 1: slice j of giant bucket gi -> hpart[gi * SLICES + j];  grid = (SLICES, lanes)
__global__ void __launch_bounds__(128) k_msm_gather_giant1(const uint32_t* offsets, const G1Xyzz* partial, G1Xyzz* hpart,
                                                           const uint32_t* heavy_list, const uint32_t* heavy_count, MsmGeom g) {
  extern __shared__ uint4 smraw[];
  G1Xyzz* sm = reinterpret_cast<G1Xyzz*>(smraw);
  const uint32_t ng = min(heavy_count[1], (uint32_t)MSM_GIANT_CAP);
  const uint32_t T = msm_T(g, offsets[g.nbtot()]);
  for (uint32_t gi = blockIdx.y; gi < ng; gi += gridDim.y) {
    const uint64_t key = heavy_list[g.nbtot() + gi];
    const uint32_t off = offsets[key], cnt = offsets[key + 1] - off;
    const uint64_t first = key + off / T, np = key + (off + cnt - 1) / T - first + 1;
    const uint64_t per = (np + MSM_GIANT_SLICES - 1) / MSM_GIANT_SLICES;
    const uint64_t lo = (uint64_t)blockIdx.x * per, hi = lo + per < np ? lo + per : np;
    G1Xyzz acc = xyzz_identity();
    for (uint64_t s = lo + threadIdx.x; s < hi; s += blockDim.x) xyzz_add(acc, xyzz_load(partial + first + s));
    G1Xyzz r = block_reduce_xyzz(acc, sm);
    if (threadIdx.x == 0) xyzz_store(hpart + (uint64_t)gi * MSM_GIANT_SLICES + blockIdx.x, r);
  }
}
// pass 2: fold the slices
__global__ void __launch_bounds__(MSM_GIANT_SLICES) k_msm_gather_giant2(const G1Xyzz* hpart, G1Xyzz* buckets, const uint32_t* heavy_list,
                                                                        const uint32_t* heavy_count, MsmGeom g) {
  extern __shared__ uint4 smraw[];
  G1Xyzz* sm = reinterpret_cast<G1Xyzz*>(smraw);
  const uint32_t ng = min(heavy_count[1], (uint32_t)MSM_GIANT_CAP);
  for (uint32_t gi = blockIdx.x; gi < ng; gi += gridDim.x) {
    G1Xyzz r = block_reduce_xyzz(xyzz_load(hpart + (uint64_t)gi * MSM_GIANT_SLICES + threadIdx.x), sm);
    if (threadIdx.x == 0) xyzz_store(buckets + heavy_list[g.nbtot() + gi], r);
  }
}

// Level 1: P[set][t][chunk] = sum of buckets b of the chunk whose index has bit t set.  grid = (c, nchunks, nsets)
// A chunk is the aligned range b in [ch * 1024, ch * 1024 + 1024); the buckets with bit t set are ENUMERATED (the j-th one
// directly), not filtered: every lane adds in every iteration, where a filter would leave half of each warp idle and
// double the dependent chain.  b = NB = 2^(c-1), the one bucket outside the aligned ranges, has only the top bit.
#define RED_LOG 10
#define RED_CHUNK (1u << RED_LOG)
__global__ void __launch_bounds__(64) k_msm_reduce1(const G1Xyzz* buckets, G1Xyzz* P, MsmGeom g, uint32_t nchunks) {
  extern __shared__ uint4 smraw[];
  G1Xyzz* sm = reinterpret_cast<G1Xyzz*>(smraw);
  const uint32_t t = blockIdx.x, ch = blockIdx.y;
  const uint64_t set = blockIdx.z;
  const G1Xyzz* bk = buckets + set * g.NB;     // bucket b lives at bk[b - 1]
  G1Xyzz acc = xyzz_identity();
  const uint32_t lo = ch * RED_CHUNK;
  if (t < RED_LOG) {
    for (uint32_t j = threadIdx.x; j < RED_CHUNK / 2; j += blockDim.x) {
      const uint32_t b = lo + (((j >> t) << (t + 1)) | (1u << t) | (j & ((1u << t) - 1)));
      if (b < g.NB) xyzz_add(acc, xyzz_load(bk + (b - 1)));
    }
  } else {
    // whole chunks qualify or not: the two CTAs of a chunk pair (ch with / without bit t) take half of the qualifying chunk each
    const uint32_t bit = 1u << (t - RED_LOG);
    const uint32_t base = (ch | bit) * RED_CHUNK + ((ch & bit) ? RED_CHUNK / 2 : 0);
    for (uint32_t j = threadIdx.x; j < RED_CHUNK / 2; j += blockDim.x)
      if (base + j < g.NB) xyzz_add(acc, xyzz_load(bk + (base + j - 1)));
  }
  if (ch == 0 && threadIdx.x == 0 && t == g.c - 1) xyzz_add(acc, xyzz_load(bk + (g.NB - 1)));
  G1Xyzz r = block_reduce_xyzz(acc, sm);
  if (threadIdx.x == 0) xyzz_store(P + (set * g.c + t) * nchunks + ch, r);
}
// Level 2: U[set][t] = sum over chunks.  grid = (c, nsets), one warp
__global__ void __launch_bounds__(32) k_msm_reduce2(const G1Xyzz* P, G1Xyzz* U, MsmGeom g, uint32_t nchunks) {
  extern __shared__ uint4 smraw[];
  G1Xyzz* sm = reinterpret_cast<G1Xyzz*>(smraw);
  const uint64_t idx = (uint64_t)blockIdx.y * g.c + blockIdx.x;
  const G1Xyzz* p = P + idx * nchunks;
  G1Xyzz acc = xyzz_identity();
  for (uint32_t c = threadIdx.x; c < nchunks; c += 32) xyzz_add(acc, xyzz_load(p + c));
  G1Xyzz r = block_reduce_xyzz(acc, sm);
  if (threadIdx.x == 0) xyzz_store(U + idx, r);
}

// ---- host-side epilogue: Horner over bit sums and windows (a few hundred point ops) -------------
static G1Xyzz host_combine(const G1Xyzz* U, const MsmGeom& g) {
  G1Xyzz total = xyzz_identity();
  for (int set = (int)g.sets - 1; set >= 0; --set) {
    G1Xyzz r = xyzz_identity();
    for (int t = (int)g.c - 1; t >= 0; --t) { r = xyzz_dbl(r); xyzz_add(r, U[(size_t)set * g.c + t]); }
    for (uint32_t d = 0; d < g.c; ++d) total = xyzz_dbl(total);
    xyzz_add(total, r);
  }
  return total;
}

static void xyzz_to_abi(const G1Xyzz& p, zkc_g1* out) {
  G1Affine a = xyzz_to_affine(p);
  Fq one = fe_one<FqP>(), zero = fe_zero<FqP>();
  if (affine_is_identity(a)) {  // halo2curves G1::identity() = (0, 1, 0)
    memcpy(&out->x, zero.v, 32); memcpy(&out->y, one.v, 32); memcpy(&out->z, zero.v, 32);
  } else {
    memcpy(&out->x, a.x.v, 32); memcpy(&out->y, a.y.v, 32); memcpy(&out->z, one.v, 32);
  }
}

int u32_scan(zkc_ctx* ctx, const uint32_t* in, uint32_t* out, uint64_t n, uint32_t* total_dev);   // poly.cu

uint32_t msm_pick_c(const zkc_ctx* ctx, uint64_t n, bool precomputed) {
  if (const int v = precomputed ? ctx->tune.msm_c_pre : ctx->tune.msm_c) { const int W = (255 + v - 1) / v; return (uint32_t)((255 + W - 1) / W); }
  uint32_t lg = 0;
  while ((1ull << (lg + 1)) <= n) ++lg;
  // measured on B200: shared-bucket (precomputed) layout wants ~4-8 partials per bucket for the
  // gather phase: c = lg-2 up to 2^18, lg-3 at 2^19, lg-4 from 2^20; the per-window layout uses lg-4 throughout
  int c = (int)lg - 4;
  if (precomputed && lg <= 18) c = (int)lg - 2;
  else if (precomputed && lg == 19) c = (int)lg - 3;
  c = std::max(3, std::min(20, c));
  const int W = (255 + c - 1) / c;
  return (uint32_t)((255 + W - 1) / W);   // same number of windows, evenly filled (see msm_geom)
}

MsmGeom msm_geom(const zkc_ctx* ctx, uint64_t n, uint32_t ncols, uint32_t c, bool precomputed) {
  MsmGeom g;
  g.W = (255 + c - 1) / c;
  c = (255 + g.W - 1) / g.W;   // same window count, evenly filled: a nearly empty top window would pile n/2^few entries on a handful of buckets
  g.c = c; g.NB = 1u << (c - 1); g.sets = precomputed ? 1 : g.W; g.n = n; g.ncols = ncols; g.sstride = n; g.bstride = n;
  const uint64_t e = g.emax();
  // entries per accumulate thread (B200 sweep): short chunks keep more warps in flight (T=8 reaches
  // 0.99 of the IMAD.WIDE peak) but multiply the partials the gather phase must fold; 32 / 64 minimise the sum
  uint32_t T = e >= (1ull << 26) ? 64 : 32;
  if (ctx->tune.msm_T) T = (uint32_t)ctx->tune.msm_T;
  while (T > 4 && e / T < 148ull * 512) T >>= 1;
  g.T = T;
  return g;
}

// All kernels of one batch on ctx->stream: bit-plane sums to U_out[nc * sets * c], number of non-zero digits to *total_out.
static int msm_kernels(zkc_ctx* ctx, const Fr* scalars, const G1Affine* bases, const MsmGeom& g, G1Xyzz* U, uint32_t* total_out) {
  const uint64_t n = g.n, nbt = g.nbtot(), em = g.emax();
  const uint32_t nc = g.ncols;
  if (nbt + em / g.T + 2ull * MSM_MIN_THREADS + 1 >= (1ull << 32) || em >= (1ull << 32) || (uint64_t)g.W * g.bstride >= (1ull << 31))
    return set_err(ctx, ZKC_ERR_BAD_ARG, "msm: batch too large");
  if ((uint64_t)nc * g.sets > 65535) return set_err(ctx, ZKC_ERR_BAD_ARG, "msm: too many bucket sets in one batch");
  // accumulate threads / partial slots: worst case em / T chains; a sparse batch runs shorter chains (msm_T), never more
  // than 2 * MSM_MIN_THREADS of them
  const uint64_t nthreads = std::max<uint64_t>((em + g.T - 1) / g.T, 2ull * MSM_MIN_THREADS);
  const uint64_t nslots = nbt + nthreads + 1;
  size_t o = 0;
  auto carve = [&](size_t bytes) { size_t r = o; o += (bytes + 255) & ~(size_t)255; return r; };
  const size_t o_counts = carve(nbt * 4), o_offsets = carve((nbt + 1) * 4), o_cursor = carve(nbt * 4), o_heavyc = carve(8),
               o_heavy = carve((nbt + MSM_GIANT_CAP) * 4), o_hpart = carve((size_t)MSM_GIANT_CAP * MSM_GIANT_SLICES * sizeof(G1Xyzz)), o_dig = carve(em * 4), o_pt = carve(em * 4), o_key = carve(em * 4),
               o_part = carve(nslots * sizeof(G1Xyzz)), o_bk = carve(nbt * sizeof(G1Xyzz)),
               o_redp = carve((size_t)nc * g.sets * g.c * ((g.NB + 1023) / 1024) * sizeof(G1Xyzz));
  char* base;
  ZKC_TRY(scratch_reserve(ctx, SCR_MSM, o, (void**)&base));
  uint32_t* counts = (uint32_t*)(base + o_counts); uint32_t* offsets = (uint32_t*)(base + o_offsets);
  uint32_t* cursor = (uint32_t*)(base + o_cursor); uint32_t* heavyc = (uint32_t*)(base + o_heavyc);
  uint32_t* heavy = (uint32_t*)(base + o_heavy); uint32_t* dig = (uint32_t*)(base + o_dig);
  uint32_t* ent_pt = (uint32_t*)(base + o_pt); uint32_t* ent_key = (uint32_t*)(base + o_key);
  G1Xyzz* partial = (G1Xyzz*)(base + o_part); G1Xyzz* buckets = (G1Xyzz*)(base + o_bk); G1Xyzz* redp = (G1Xyzz*)(base + o_redp); G1Xyzz* hpart = (G1Xyzz*)(base + o_hpart);
  cudaStream_t st = ctx->stream;
  ZKC_CUDA_TRY(ctx, cudaMemsetAsync(counts, 0, nbt * 4, st));
  ZKC_CUDA_TRY(ctx, cudaMemsetAsync(heavyc, 0, 8, st));
  const uint64_t npts = n * nc;
  { ProfScope _p(ctx, "msm.digits");
    k_msm_digits<<<(unsigned)((npts + 127) / 128), 128, 0, st>>>(scalars, dig, counts, g);
    ZKC_LAUNCH_CHECK(ctx); }
  { ProfScope _p(ctx, "msm.scan");
    ZKC_TRY(u32_scan(ctx, counts, offsets, nbt, offsets + nbt));
    ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(cursor, offsets, nbt * 4, cudaMemcpyDeviceToDevice, st)); }
  { ProfScope _p(ctx, "msm.scatter");
    k_msm_scatter<<<(unsigned)((em + 255) / 256), 256, 0, st>>>(dig, cursor, ent_pt, ent_key, g);
    ZKC_LAUNCH_CHECK(ctx); }
  { ProfScope _p(ctx, "msm.accum");
    k_msm_accum<<<(unsigned)((nthreads + 127) / 128), 128, 0, st>>>(bases, ent_pt, ent_key, offsets, partial, g);
    ZKC_LAUNCH_CHECK(ctx); }
  {
    // group width from the expected number of partials per bucket (entries per bucket / T)
    const double avg_partials = (double)g.W * (double)n / (double)g.NB / (double)g.sets / (double)g.T + 1.0;
    uint32_t logG = 0;
    while (logG < 5 && (double)(1u << logG) * 3.0 < avg_partials) ++logG;
    { ProfScope _p(ctx, "msm.gather");
      const uint64_t nthr = nbt << logG;
      k_msm_gather<<<(unsigned)((nthr + 127) / 128), 128, 0, st>>>(offsets, partial, buckets, heavy, heavyc, g, logG);
      ZKC_LAUNCH_CHECK(ctx); }
    { ProfScope _p(ctx, "msm.gather_heavy");
      k_msm_gather_heavy<<<ctx->sm_count * 4, 128, 128 * sizeof(G1Xyzz), st>>>(offsets, partial, buckets, heavy, heavyc, g);
      ZKC_LAUNCH_CHECK(ctx);
      k_msm_gather_giant1<<<dim3(MSM_GIANT_SLICES, 16), 128, 128 * sizeof(G1Xyzz), st>>>(offsets, partial, hpart, heavy, heavyc, g);
      ZKC_LAUNCH_CHECK(ctx);
      k_msm_gather_giant2<<<64, MSM_GIANT_SLICES, MSM_GIANT_SLICES * sizeof(G1Xyzz), st>>>(hpart, buckets, heavy, heavyc, g);
      ZKC_LAUNCH_CHECK(ctx); }
  }
  {
    ProfScope _p(ctx, "msm.reduce");
    const uint32_t nchunks = (g.NB + RED_CHUNK - 1) / RED_CHUNK;
    dim3 g1(g.c, nchunks, nc * g.sets), g2(g.c, nc * g.sets);
    k_msm_reduce1<<<g1, 64, 64 * sizeof(G1Xyzz), st>>>(buckets, redp, g, nchunks);
    ZKC_LAUNCH_CHECK(ctx);
    k_msm_reduce2<<<g2, 32, 32 * sizeof(G1Xyzz), st>>>(redp, U, g, nchunks);
    ZKC_LAUNCH_CHECK(ctx);
  }
  ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(total_out, offsets + nbt, 4, cudaMemcpyDeviceToDevice, st));
  return ZKC_OK;
}

// Enqueue one batch (all kernels + the async D2H of the c bit-plane sums per column) on ctx->stream.
int msm_enqueue(zkc_ctx* ctx, const Fr* scalars, const G1Affine* bases, uint64_t n, uint32_t nc, uint32_t c, bool precomputed, MsmPending* pend,
                int result_slot, bool team) {
  team = team && team_active(ctx) && n >= (uint64_t)ctx->team_world;
  const int shards = team ? ctx->team_world : 1;
  const MsmGeom g0 = msm_geom(ctx, n, nc, c, precomputed);
  const size_t ubytes = (size_t)nc * g0.sets * g0.c * sizeof(G1Xyzz), slot_bytes = (ubytes + 4 + 255) & ~(size_t)255;   // U, then the digit count
  char* dU;
  ZKC_TRY(scratch_reserve(ctx, SCR_MSM2, slot_bytes * shards, (void**)&dU));
  if (!team) {
    ZKC_TRY(msm_kernels(ctx, scalars, bases, g0, (G1Xyzz*)dU, (uint32_t*)(dU + ubytes)));
  } else {
    // point-range shards: rank r walks points [lo, hi) of every column against the same slice of every window table
    for (int r : team_ranks(ctx)) {
      uint64_t lo, hi;
      shard_range(n, shards, r, &lo, &hi);
      MsmGeom g = msm_geom(ctx, hi - lo, nc, c, precomputed);
      g.sstride = n; g.bstride = n;
      char* slot = dU + (size_t)r * slot_bytes;
      ZKC_TRY(msm_kernels(ctx, scalars + lo, bases + lo, g, (G1Xyzz*)slot, (uint32_t*)(slot + ubytes)));
    }
    ProfScope _p(ctx, "team.msm_allgather");
    ZKC_TRY(team_allgather(ctx, dU, slot_bytes));
  }
  void* hU;
  ZKC_TRY(pinned_reserve(ctx, slot_bytes * shards, &hU, result_slot));   // slot 1: a batch whose result is consumed later
  ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(hU, dU, slot_bytes * shards, cudaMemcpyDeviceToHost, ctx->stream));
  cudaEvent_t ev = result_slot ? ctx->ev_msm_side : ctx->ev_msm_main;
  ZKC_CUDA_TRY(ctx, cudaEventRecord(ev, ctx->stream));
  pend->g = g0; pend->nc = nc; pend->n = n; pend->hU = hU; pend->ubytes = ubytes; pend->shards = shards; pend->done = ev; pend->active = true;
  return ZKC_OK;
}

// Wait for the batch, then the host-side epilogue (sum of the shards' bit planes, Horner over bit planes / windows, normalisation).
int msm_finish(zkc_ctx* ctx, MsmPending* pend, zkc_g1* out) {
  if (!pend->active) return set_err(ctx, ZKC_ERR_BAD_ARG, "msm_finish: nothing pending");
  ZKC_CUDA_TRY(ctx, cudaEventSynchronize(pend->done));
  pend->active = false;
  const size_t slot_bytes = (pend->ubytes + 4 + 255) & ~(size_t)255;
  const size_t per_col = (size_t)pend->g.sets * pend->g.c;
  std::vector<G1Xyzz> sum;
  const G1Xyzz* U = (const G1Xyzz*)pend->hU;
  uint64_t madds = 0;
  for (int r = 0; r < pend->shards; ++r) {
    uint32_t e; memcpy(&e, (char*)pend->hU + (size_t)r * slot_bytes + pend->ubytes, 4);
    if (ctx->team_emulate || pend->shards == 1 || r == ctx->team_rank) madds += e;    // this GPU's own work
  }
  ctx->stats["msm.madds"] += madds; ctx->stats["msm.points"] += pend->n * pend->nc;
  if (pend->shards > 1) {
    sum.assign(U, U + per_col * pend->nc);
    for (int r = 1; r < pend->shards; ++r) {
      const G1Xyzz* Ur = (const G1Xyzz*)((const char*)pend->hU + (size_t)r * slot_bytes);
      for (size_t i = 0; i < sum.size(); ++i) xyzz_add(sum[i], Ur[i]);
    }
    U = sum.data();
  }
  for (uint32_t col = 0; col < pend->nc; ++col) {
    G1Xyzz r = host_combine(U + (size_t)col * per_col, pend->g);
    xyzz_to_abi(r, out + col);
  }
  return ZKC_OK;
}

// Core: `ncols` scalar columns (n each, contiguous) against `bases` (n points, or W tables of n when
// precomputed).  Writes ncols results to `out` (host).
int msm_run(zkc_ctx* ctx, const Fr* scalars, const G1Affine* bases, uint64_t n, uint32_t ncols, uint32_t c, bool precomputed, zkc_g1* out,
            bool team = false) {
  if (ncols == 0) return ZKC_OK;
  if (n == 0) {
    for (uint32_t i = 0; i < ncols; ++i) xyzz_to_abi(xyzz_identity(), out + i);
    return ZKC_OK;
  }
  if (n >= (1ull << 31) / 32) return set_err(ctx, ZKC_ERR_BAD_ARG, "msm: n too large");
  // bound memory: process columns in chunks
  MsmGeom g1 = msm_geom(ctx, n, 1, c, precomputed);
  const uint64_t per_col = g1.emax() * 12 + g1.nbtot() * (128 + 12) + (g1.nbtot() + g1.emax() / g1.T + 1) * 128;
  uint32_t chunk = (uint32_t)std::max<uint64_t>(1, (3ull << 30) / per_col);
  chunk = std::min(chunk, ncols);
  for (uint32_t c0 = 0; c0 < ncols; c0 += chunk) {
    const uint32_t nc = std::min(chunk, ncols - c0);
    MsmPending pend;
    ZKC_TRY(msm_enqueue(ctx, scalars + (uint64_t)c0 * n, bases, n, nc, c, precomputed, &pend, 0, team));
    ZKC_TRY(msm_finish(ctx, &pend, out + c0));
  }
  return ZKC_OK;
}

// ---- SRS: resident bases + window tables ----------------------------------------------------------
// table[w][i] = 2^(c*w) * P_i in affine form; thread per point walks the windows.
__global__ void k_srs_expand(G1Affine* table, uint64_t n, uint32_t c, uint32_t W) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G1Affine p = affine_load_nc(table + i);
  for (uint32_t w = 1; w < W; ++w) {
    G1Xyzz q = xyzz_from_affine(p);
    for (uint32_t d = 0; d < c; ++d) q = xyzz_dbl(q);
    p = xyzz_to_affine(q);
    fe_store(&table[(uint64_t)w * n + i].x, p.x);
    fe_store(&table[(uint64_t)w * n + i].y, p.y);
  }
}

// out[i] = [scalar_i] G via 8-bit fixed windows over a table tab[w][d] = [d * 256^w] G
__global__ void k_fixed_base_table(G1Affine* tab) {
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;   // w*256 + d
  if (idx >= 32 * 256) return;
  const uint32_t w = idx >> 8, d = idx & 255;
  G1Affine gen; gen.x = fe_zero<FqP>(); gen.x.v[0] = 1; gen.x = fe_from_canonical(gen.x);
  gen.y = fe_zero<FqP>(); gen.y.v[0] = 2; gen.y = fe_from_canonical(gen.y);
  G1Xyzz acc = xyzz_identity();
  // [d]G by double-and-add, then 8*w doublings
  for (int bit = 7; bit >= 0; --bit) { acc = xyzz_dbl(acc); if ((d >> bit) & 1) xyzz_madd(acc, gen, false); }
  for (uint32_t k = 0; k < 8 * w; ++k) acc = xyzz_dbl(acc);
  G1Affine a = xyzz_to_affine(acc);
  fe_store(&tab[idx].x, a.x); fe_store(&tab[idx].y, a.y);
}
__global__ void k_fixed_base_mul(const Fr* scalars, const G1Affine* tab, G1Affine* out, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Fr s = fe_to_canonical(fe_load(scalars + i));
  G1Xyzz acc = xyzz_identity();
  for (uint32_t w = 0; w < 32; ++w) {
    const uint32_t d = (s.v[w >> 2] >> ((w & 3) * 8)) & 255;
    if (d) xyzz_madd(acc, affine_load_nc(tab + w * 256 + d), false);
  }
  G1Affine a = xyzz_to_affine(acc);
  fe_store(&out[i].x, a.x); fe_store(&out[i].y, a.y);
}
// pw[i] = s^i
__global__ void k_powers(Fr* pw, Fr s, uint64_t n) {
  const uint64_t start = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 64;
  if (start >= n) return;
  Fr w = fe_pow_u64(s, start);
  const uint64_t end = start + 64 < n ? start + 64 : n;
  for (uint64_t i = start; i < end; ++i) { fe_store(pw + i, w); w = fe_mul(w, s); }
}
// den[i] = s - omega^i (omega powers read from pw)
__global__ void k_lagrange_den(const Fr* wpow, Fr s, Fr* den, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  fe_store(den + i, fe_sub(s, fe_load(wpow + i)));
}
// sc[i] = zn * omega^i * deninv[i]
__global__ void k_lagrange_scalars(const Fr* wpow, const Fr* deninv, Fr zn, Fr* sc, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  fe_store(sc + i, fe_mul(fe_mul(zn, fe_load(wpow + i)), fe_load(deninv + i)));
}

int fr_batch_invert(zkc_ctx* ctx, const Fr* a, Fr* out, size_t n);

}  // namespace zkc

using namespace zkc;

struct zkc_srs {
  zkc_ctx* ctx;
  uint32_t k, c, W;
  uint64_t n;
  G1Affine* tab[2] = {nullptr, nullptr};  // [basis] -> W tables of n affine points ([0] = the SRS itself)
};

static int srs_expand(zkc_ctx* ctx, zkc_srs* s) {
  for (int b = 0; b < 2; ++b) {
    if (s->W > 1) {
      k_srs_expand<<<(unsigned)((s->n + 127) / 128), 128, 0, ctx->stream>>>(s->tab[b], s->n, s->c, s->W);
      ZKC_LAUNCH_CHECK(ctx);
    }
  }
  ZKC_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return ZKC_OK;
}

static int srs_alloc(zkc_ctx* ctx, uint32_t k, zkc_srs** out) {
  zkc_srs* s = new zkc_srs();
  s->ctx = ctx; s->k = k; s->n = 1ull << k;
  // window width from the points ONE GPU walks: a team rank sees n / world points per commitment (point-range shards), and
  // the bucket phases (gather, reduce) do not shrink with the shard, so a team wants narrower windows than a single GPU
  const uint64_t n_eff = team_active(ctx) ? std::max<uint64_t>(s->n / (uint64_t)ctx->team_world, 1024) : s->n;
  s->c = msm_pick_c(ctx, n_eff, true);
  s->W = (255 + s->c - 1) / s->c;
  for (int b = 0; b < 2; ++b) {
    cudaError_t e = cudaMalloc(&s->tab[b], (size_t)s->W * s->n * sizeof(G1Affine));
    if (e != cudaSuccess) {
      if (s->tab[0]) cudaFree(s->tab[0]);
      delete s;
      return set_err(ctx, ZKC_ERR_OOM, std::string("zkc_srs: ") + cudaGetErrorString(e));
    }
  }
  *out = s;
  return ZKC_OK;
}

extern "C" uint32_t zkc_srs_k(const zkc_srs* s) { return s ? s->k : 0; }

extern "C" void zkc_srs_free(zkc_srs* s) {
  if (!s) return;
  for (int b = 0; b < 2; ++b) if (s->tab[b]) cudaFree(s->tab[b]);
  delete s;
}

extern "C" int zkc_srs_load(zkc_ctx* ctx, uint32_t k, const zkc_g1_affine* g, const zkc_g1_affine* gl, zkc_srs** out) {
  if (!ctx || !g || !gl || !out || k > 26) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_srs_load: bad arguments");
  CtxLock lock(ctx);
  zkc_srs* s;
  ZKC_TRY(srs_alloc(ctx, k, &s));
  const size_t bytes = s->n * sizeof(G1Affine);
  cudaError_t e = cudaMemcpyAsync(s->tab[0], g, bytes, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(s->tab[1], gl, bytes, cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) { zkc_srs_free(s); return set_err(ctx, ZKC_ERR_CUDA, cudaGetErrorString(e)); }
  int st = srs_expand(ctx, s);
  if (st != ZKC_OK) { zkc_srs_free(s); return st; }
  *out = s;
  return ZKC_OK;
}

extern "C" int zkc_srs_setup(zkc_ctx* ctx, uint32_t k, const zkc_fr* s_abi, zkc_srs** out) {
  if (!ctx || !s_abi || !out || k > 26) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_srs_setup: bad arguments");
  CtxLock lock(ctx);
  zkc_srs* s;
  ZKC_TRY(srs_alloc(ctx, k, &s));
  const uint64_t n = s->n;
  Fr sec; memcpy(sec.v, s_abi, 32);
  // private temporaries: the helpers called below (batch inversion -> scans) own the ctx scratch arenas
  Fr *pw = nullptr, *den = nullptr;
  G1Affine* tab = nullptr;
  cudaStream_t stream = ctx->stream;
  int st = ZKC_OK;
  if (cudaMallocAsync((void**)&pw, n * sizeof(Fr), stream) != cudaSuccess || cudaMallocAsync((void**)&den, n * sizeof(Fr), stream) != cudaSuccess ||
      cudaMallocAsync((void**)&tab, 32 * 256 * sizeof(G1Affine), stream) != cudaSuccess)
    st = set_err(ctx, ZKC_ERR_OOM, "zkc_srs_setup: out of memory");
  auto release = [&]() { if (pw) cudaFreeAsync(pw, stream); if (den) cudaFreeAsync(den, stream); if (tab) cudaFreeAsync(tab, stream); };
  if (st != ZKC_OK) { release(); zkc_srs_free(s); return st; }
  const unsigned gb = (unsigned)((n + 127) / 128), gp = (unsigned)(((n + 63) / 64 + 127) / 128);
  k_fixed_base_table<<<64, 128, 0, stream>>>(tab); ctx->launches++;
  // g[i] = [s^i] G
  k_powers<<<gp, 128, 0, stream>>>(pw, sec, n); ctx->launches++;
  k_fixed_base_mul<<<gb, 128, 0, stream>>>(pw, tab, s->tab[0], n); ctx->launches++;
  // g_lagrange[i] = [ (s^n - 1)/n * w^i / (s - w^i) ] G
  const Fr omega = fr_root_of_unity(k);
  const Fr zn = fe_mul(fe_sub(fe_pow_u64(sec, n), fe_one<FrP>()), fe_inv(fr_from_u64(n)));
  k_powers<<<gp, 128, 0, stream>>>(pw, omega, n); ctx->launches++;
  k_lagrange_den<<<gb, 128, 0, stream>>>(pw, sec, den, n); ctx->launches++;
  st = fr_batch_invert(ctx, den, den, n);
  if (st == ZKC_OK) {
    k_lagrange_scalars<<<gb, 128, 0, stream>>>(pw, den, zn, den, n); ctx->launches++;
    k_fixed_base_mul<<<gb, 128, 0, stream>>>(den, tab, s->tab[1], n); ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) st = set_err(ctx, ZKC_ERR_CUDA, cudaGetErrorString(e));
  }
  if (st == ZKC_OK) st = srs_expand(ctx, s);
  release();
  if (st != ZKC_OK) { zkc_srs_free(s); return st; }
  *out = s;
  return ZKC_OK;
}

extern "C" int zkc_srs_get(zkc_ctx* ctx, const zkc_srs* s, int basis, zkc_g1_affine* out) {
  if (!ctx || !s || !out || basis < 0 || basis > 1) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_srs_get: bad arguments");
  CtxLock lock(ctx);
  ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(out, s->tab[basis], s->n * sizeof(G1Affine), cudaMemcpyDeviceToHost, ctx->stream));
  ZKC_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return ZKC_OK;
}

namespace zkc {
// asynchronous single-column commitment of a full-length polynomial (finish with msm_finish)
int srs_commit_enqueue(zkc_ctx* ctx, const zkc_srs* s, int basis, const Fr* poly, uint64_t len, MsmPending* pend) {
  if (len != s->n) return set_err(ctx, ZKC_ERR_BAD_ARG, "srs_commit_enqueue: full-length polynomials only");
  return msm_enqueue(ctx, poly, s->tab[basis], len, 1, s->c, true, pend, 1, true);
}
// commit `ncols` device-resident polynomials of `len` <= n coefficients (column stride = len)
int srs_commit_dev(zkc_ctx* ctx, const zkc_srs* s, int basis, const Fr* polys, uint64_t len, uint32_t ncols, zkc_g1* out) {
  if (len > s->n) return set_err(ctx, ZKC_ERR_BAD_ARG, "commit: polynomial longer than the SRS");
  if (len == s->n) return msm_run(ctx, polys, s->tab[basis], len, ncols, s->c, true, out, true);   // team: point-range shards
  // shorter polynomials: the window tables are laid out with stride n, so fall back to the generic
  // (non-precomputed) walk over the first `len` bases.
  return msm_run(ctx, polys, s->tab[basis], len, ncols, msm_pick_c(ctx, len, false), false, out);
}
}  // namespace zkc

extern "C" int zkc_commit_dev(zkc_ctx* ctx, const zkc_srs* s, int basis, const zkc_fr* polys, size_t len, uint32_t ncols, zkc_g1* out) {
  if (!ctx || !s || !polys || !out || basis < 0 || basis > 1) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_commit_dev: bad arguments");
  CtxLock lock(ctx);
  return srs_commit_dev(ctx, s, basis, (const Fr*)polys, len, ncols, out);
}
extern "C" int zkc_commit(zkc_ctx* ctx, const zkc_srs* s, int basis, const zkc_fr* poly, size_t len, zkc_g1* out) {
  if (!ctx || !s || !poly || !out || basis < 0 || basis > 1) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_commit: bad arguments");
  CtxLock lock(ctx);
  void* d;
  ZKC_TRY(scratch_reserve(ctx, SCR_HOSTIO, std::max<size_t>(len, 1) * sizeof(Fr), &d));
  if (len) ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(d, poly, len * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
  return srs_commit_dev(ctx, s, basis, (const Fr*)d, len, 1, out);
}

extern "C" int zkc_msm_g1_dev(zkc_ctx* ctx, const zkc_fr* scalars, const zkc_g1_affine* bases, size_t n, uint32_t ncols, zkc_g1* out) {
  if (!ctx || !out || (n && (!scalars || !bases))) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_msm_g1_dev: bad arguments");
  CtxLock lock(ctx);
  return msm_run(ctx, (const Fr*)scalars, (const G1Affine*)bases, n, ncols, msm_pick_c(ctx, n ? n : 1, false), false, out);
}
extern "C" int zkc_msm_g1(zkc_ctx* ctx, const zkc_fr* scalars, const zkc_g1_affine* bases, size_t n, zkc_g1* out) {
  if (!ctx || !out || (n && (!scalars || !bases))) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_msm_g1: bad arguments");
  CtxLock lock(ctx);
  void *ds, *db;
  ZKC_TRY(scratch_reserve(ctx, SCR_HOSTIO, std::max<size_t>(n, 1) * sizeof(Fr), &ds));
  ZKC_TRY(scratch_reserve(ctx, SCR_HOSTIO2, std::max<size_t>(n, 1) * sizeof(G1Affine), &db));
  if (n) {
    ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(ds, scalars, n * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(db, bases, n * sizeof(G1Affine), cudaMemcpyHostToDevice, ctx->stream));
  }
  return msm_run(ctx, (const Fr*)ds, (const G1Affine*)db, n, 1, msm_pick_c(ctx, n ? n : 1, false), false, out);
}

// Host-side epilogue of a point-sharded MSM: sum of `n` normalised Jacobian partial results,
// one per rank.  No device work; exact group arithmetic on the host.
extern "C" int zkc_g1_sum(const zkc_g1* pts, size_t n, zkc_g1* out) {
  if (!out || (n && !pts)) return ZKC_ERR_BAD_ARG;
  G1Xyzz acc = xyzz_identity();
  for (size_t i = 0; i < n; ++i) {
    Fq z; memcpy(z.v, &pts[i].z, 32);
    if (fe_is_zero(z)) continue;
    if (!fe_eq(z, fe_one<FqP>())) return ZKC_ERR_BAD_ARG;   // only normalised points are accepted
    G1Affine a; memcpy(a.x.v, &pts[i].x, 32); memcpy(a.y.v, &pts[i].y, 32);
    xyzz_madd(acc, a, false);
  }
  xyzz_to_abi(acc, out);
  return ZKC_OK;
}
