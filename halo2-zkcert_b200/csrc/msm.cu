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
//   6 k_msm_tree     sum_b b*S_b per bucket set as c bit-plane sums U_t = sum of buckets whose index has bit t set,
//                    computed by ONE binary tree over the bucket indices whose nodes carry (U_0..U_{t-1}, S): about two
//                    additions per bucket, depth c - 1, no serial running sum; the 2^t weights are applied by Horner on the host.
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
#define MSM_MIN_THREADS (148u * 256u)
__device__ __forceinline__ uint32_t msm_T(const MsmGeom& g, uint64_t E) {
  uint32_t T = g.T;
  while (T > 4 && E / T < MSM_MIN_THREADS) T >>= 1;
  return T;
}

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

template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_msm_accum(const G1Affine* bases, const uint32_t* ent_pt, const uint32_t* ent_key,
                                                   const uint32_t* offsets, G1Xyzz* partial, MsmGeom g) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t E = offsets[g.nbtot()];
  const uint32_t T = msm_T(g, E);
  const uint64_t p0 = t * T;
  const bool active = p0 < E;
  const uint64_t p1 = !active ? p0 : (p0 + T < E ? p0 + T : E);
  G1Xyzz acc = xyzz_identity();
  uint32_t cur = 0xffffffffu;
  if (active) {
    cur = ent_key[p0];
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
  }
  // A warp-wide convergence point before the final store: with it ptxas fits the loop into 126 registers without spilling,
  // without it the same code takes 128 + 80 bytes of spills (four CTAs per SM either way; checked with -Xptxas -v).
  const uint32_t k0 = __shfl_sync(0xffffffffu, cur, 0);
  if (k0 == 0xfffffffeu) acc = xyzz_identity();   // never true: bucket keys are < 2^31
  if (active) xyzz_store(partial + cur + t, acc);
}

// A warp of accumulate threads that sits entirely inside ONE bucket (bit / byte valued witness columns and the top window of
// range-checked cells pile thousands of entries on a few buckets) folds its 32 partials with a shuffle tree and leaves the
// identity in the other 31 slots, so such a bucket hands 32x fewer non-trivial partials to the bucket-sum kernels: their dependent
// chain is what costs there, and adding the identity is free.  A kernel of its own: inside k_msm_accum the 32 extra live
// registers of the shuffled operand pushed the accumulate loop over the 128 registers that four CTAs per SM allow.
__global__ void __launch_bounds__(128) k_msm_fold(const uint32_t* ent_key, const uint32_t* offsets, G1Xyzz* partial, MsmGeom g) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t E = offsets[g.nbtot()];
  const uint32_t T = msm_T(g, E);
  const uint64_t p0 = t * T;
  const bool active = p0 < E;
  uint32_t cur = 0xffffffffu;
  bool single = false;
  if (active) {
    const uint64_t p1 = p0 + T < E ? p0 + T : E;
    cur = ent_key[p1 - 1];
    single = ent_key[p0] == cur;       // the thread's whole range lies in one bucket: one partial, at slot cur + t
  }
  const uint32_t k0 = __shfl_sync(0xffffffffu, cur, 0);
  if (!__all_sync(0xffffffffu, active && single && cur == k0)) return;
  G1Xyzz acc = xyzz_load(partial + cur + t);
#pragma unroll 1
  for (int d = 16; d > 0; d >>= 1) { const G1Xyzz o = xyzz_shfl_xor(acc, d); xyzz_add(acc, o); }
  if (threadIdx.x & 31) acc = xyzz_identity();
  xyzz_store(partial + cur + t, acc);
}

// ---- bucket sums: the partials of a bucket are folded by QUADS of lanes (xyzz_add_quad: this phase is a chain of dependent
// additions on a mostly idle machine, so four lanes share each addition).  A bucket gets G quads (G from the expected number of
// partials); buckets with more than MSM_HEAVY_PER_QUAD * G partials are queued for a whole CTA, above MSM_GIANT partials they
// are sliced over MSM_GIANT_SLICES CTAs.
#define MSM_HEAVY_PER_QUAD 6
#define MSM_GIANT 1024        // partials: above this a bucket is sliced over MSM_GIANT_SLICES CTAs
#define MSM_GIANT_SLICES 32
#define MSM_GIANT_CAP 1024
__global__ void __launch_bounds__(128) k_msm_gather(const uint32_t* offsets, const G1Xyzz* partial, G1Xyzz* buckets,
                                                    uint32_t* heavy_list, uint32_t* heavy_count, MsmGeom g, uint32_t logG) {
  const uint64_t qd = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;      // quad index
  const uint64_t key = qd >> logG;
  const uint32_t G = 1u << logG, slot = (uint32_t)qd & (G - 1);
  const bool valid = key < g.nbtot();     // whole groups are valid or not (nbtot * G is a multiple of the warp's group span)
  uint32_t off = 0, cnt = 0;
  if (valid) { off = offsets[key]; cnt = offsets[key + 1] - off; }
  uint64_t first = 0;
  uint32_t np = 0;
  const uint32_t T = msm_T(g, offsets[g.nbtot()]);
  if (cnt) { first = key + off / T; np = (uint32_t)(key + (off + cnt - 1) / T - first + 1); }
  const bool heavy = np > (uint32_t)MSM_HEAVY_PER_QUAD * G;
  // the additions are warp-wide collectives: every quad of the warp makes the same number of steps (the longest chain among
  // its buckets, at most MSM_HEAVY_PER_QUAD), quads that have run out add the identity
  const uint32_t mine = heavy ? 0 : np;
  const uint32_t steps = (__reduce_max_sync(0xffffffffu, mine) + G - 1) >> logG;
  G1Xyzz acc = xyzz_identity();
  for (uint32_t it = 0; it < steps; ++it) {
    const uint32_t sidx = it * G + slot;
    G1Xyzz o = xyzz_identity();
    if (sidx < mine) o = xyzz_load(partial + first + sidx);
    xyzz_add_quad(acc, o);
  }
  for (uint32_t d = G >> 1; d > 0; d >>= 1) { G1Xyzz o = xyzz_shfl_xor(acc, (int)(d << 2)); xyzz_add_quad(acc, o); }
  if (valid && slot == 0 && (threadIdx.x & 3) == 0) {
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

// The same with ONE lane per slot (plain additions): when a batch has enough buckets to fill the machine the bucket sums are bound
// by multiplier throughput, where the quad variant's four-fold lane count loses; the launcher picks by grid size.
__global__ void __launch_bounds__(128) k_msm_gather_plain(const uint32_t* offsets, const G1Xyzz* partial, G1Xyzz* buckets,
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
  const bool heavy = np > (uint64_t)MSM_HEAVY_PER_QUAD * G;
  G1Xyzz acc = xyzz_identity();
  if (cnt && !heavy)
    for (uint64_t s = lane; s < np; s += G) xyzz_add(acc, xyzz_load(partial + first + s));
  for (uint32_t d = G >> 1; d > 0; d >>= 1) { G1Xyzz o = xyzz_shfl_xor(acc, (int)d); xyzz_add(acc, o); }
  if (valid && lane == 0) {
    if (heavy) {
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

// CTA-wide XYZZ sum of one value per QUAD (the four lanes of a quad hold the same value): shuffle tree inside the warps,
// shared memory across them; every step is a quad addition.  Result valid in thread 0.  blockDim.x a multiple of 32, <= 256.
__device__ __forceinline__ G1Xyzz block_reduce_quads(G1Xyzz acc, G1Xyzz* sm) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll 1
  for (int d = 4; d < 32; d <<= 1) { const G1Xyzz o = xyzz_shfl_xor(acc, d); xyzz_add_quad(acc, o); }
  if (nwarps == 1) return acc;
  if (lane == 0) xyzz_store(sm + warp, acc);
  __syncthreads();
  if (warp == 0) {
    const uint32_t q = lane >> 2;
    acc = q < nwarps ? xyzz_load(sm + q) : xyzz_identity();
#pragma unroll 1
    for (uint32_t d = 4; (d >> 2) < nwarps; d <<= 1) { const G1Xyzz o = xyzz_shfl_xor(acc, (int)d); xyzz_add_quad(acc, o); }
  }
  __syncthreads();
  return acc;
}
// sum of partial[first + lo .. first + hi) by the CTA's quads: each quad walks a strided share (uniform trip count), then the tree
__device__ __forceinline__ G1Xyzz block_sum_partials(const G1Xyzz* partial, uint64_t first, uint64_t lo, uint64_t hi, G1Xyzz* sm) {
  const uint32_t nq = blockDim.x >> 2, qd = threadIdx.x >> 2;
  G1Xyzz acc = xyzz_identity();
  for (uint64_t base = lo; base < hi; base += nq) {
    G1Xyzz o = xyzz_identity();
    if (base + qd < hi) o = xyzz_load(partial + first + base + qd);
    xyzz_add_quad(acc, o);
  }
  return block_reduce_quads(acc, sm);
}

// heavy buckets (a few dozen to MSM_GIANT partials): one CTA per bucket, its quads walk the partials, then the tree
__global__ void __launch_bounds__(128) k_msm_gather_heavy(const uint32_t* offsets, const G1Xyzz* partial, G1Xyzz* buckets,
                                                          const uint32_t* heavy_list, const uint32_t* heavy_count, MsmGeom g) {
  extern __shared__ uint4 smraw[];
  G1Xyzz* sm = reinterpret_cast<G1Xyzz*>(smraw);
  const uint32_t nh = *heavy_count;
  const uint32_t T = msm_T(g, offsets[g.nbtot()]);
  for (uint32_t h = blockIdx.x; h < nh; h += gridDim.x) {
    const uint64_t key = heavy_list[h];
    const uint32_t off = offsets[key], cnt = offsets[key + 1] - off;
    const uint64_t first = key + off / T, np = key + (off + cnt - 1) / T - first + 1;
    G1Xyzz r = block_sum_partials(partial, first, 0, np, sm);
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
    G1Xyzz r = block_sum_partials(partial, first, lo < np ? lo : np, hi, sm);
    if (threadIdx.x == 0) xyzz_store(hpart + (uint64_t)gi * MSM_GIANT_SLICES + blockIdx.x, r);
  }
}
// pass 2: fold the slices (one quad per slice)
__global__ void __launch_bounds__(MSM_GIANT_SLICES * 4) k_msm_gather_giant2(const G1Xyzz* hpart, G1Xyzz* buckets, const uint32_t* heavy_list,
                                                                            const uint32_t* heavy_count, MsmGeom g) {
  extern __shared__ uint4 smraw[];
  G1Xyzz* sm = reinterpret_cast<G1Xyzz*>(smraw);
  const uint32_t ng = min(heavy_count[1], (uint32_t)MSM_GIANT_CAP);
  for (uint32_t gi = blockIdx.x; gi < ng; gi += gridDim.x) {
    G1Xyzz r = block_reduce_quads(xyzz_load(hpart + (uint64_t)gi * MSM_GIANT_SLICES + (threadIdx.x >> 2)), sm);
    if (threadIdx.x == 0) xyzz_store(buckets + heavy_list[g.nbtot() + gi], r);
  }
}

// ---- sum_b b * S_b as c bit-plane sums U_t = sum of the buckets whose index has bit t set, WORK-EFFICIENTLY -------------------
// A binary tree over the bucket indices b in [0, NB) (b = 0 is an empty slot; NB = 2^(c-1) itself is handled at the end).  A
// node of level t covers 2^t consecutive indices and carries (U_0 .. U_{t-1}, S): the bit-plane sums restricted to its range
// and the plain sum.  Merging the siblings L (bit t clear) and R (bit t set):
//        U_i = L.U_i + R.U_i  (i < t),     U_t = R.S,     S = L.S + R.S
// i.e. t + 1 additions per merge, about 2 per bucket over the whole tree (the per-plane tree reductions it replaces took
// c / 2 = 7.5), all additions of one level independent, depth c - 1.  One CTA merges `m` levels of 2^m sibling nodes through
// shared memory (ping-pong), one output point per QUAD of lanes and step (xyzz_add_quad: the chain of dependent additions is
// what this phase costs, so four lanes share each addition); the host chains launches until one node per bucket set is
// left, and the last launch writes U_0 .. U_{c-2} and U_{c-1} = S_NB where the host epilogue expects them.
// Node layout in memory: [U_0 .. U_{t-1}, S], t + 1 points.
template <bool QUAD>
__global__ void __launch_bounds__(256, 2) k_msm_tree(const G1Xyzz* in, G1Xyzz* out, const G1Xyzz* buckets, G1Xyzz* U, uint32_t t0, uint32_t m,
                                                           uint32_t nodes_in, uint32_t NB, uint32_t c) {
  extern __shared__ uint4 smraw[];
  G1Xyzz* buf0 = reinterpret_cast<G1Xyzz*>(smraw);
  const uint32_t grp = blockIdx.x, groups = nodes_in >> m;
  const uint64_t set = blockIdx.y;
  const uint32_t nloc = 1u << m, P0 = t0 + 1;
  const bool leaf = t0 == 0;
  const G1Xyzz* bk = buckets + set * NB;           // bucket b lives at bk[b - 1]
  const G1Xyzz* src_g = in + ((uint64_t)set * nodes_in + (uint64_t)grp * nloc) * P0;
  const uint32_t leaf0 = grp * nloc;               // first bucket index of this CTA (leaf launch)
  // buffers: a level with `nodes` input nodes of P points produces nodes / 2 nodes of P + 1 points
  const uint32_t cap0 = (nloc >> 1) * (P0 + 1);    // largest output level is the first one
  G1Xyzz* buf1 = buf0 + cap0;
  G1Xyzz* cur = nullptr;                            // level input (nullptr: read the launch input from global memory)
  G1Xyzz* nxt = buf0;
  uint32_t nodes = nloc, P = P0;
  for (uint32_t lvl = 0; lvl < m; ++lvl) {
    const uint32_t T = P - 1, Pout = P + 1, nout = nodes >> 1;
    const bool last = lvl + 1 == m;
    // (a) the T + 1 sums of every merge: one per quad of lanes (QUAD: xyzz_add_quad, whole warps stay together) or per lane
    const uint32_t total = nout * P, nslots = QUAD ? blockDim.x >> 2 : blockDim.x;
    for (uint32_t base = 0; base < total; base += nslots) {
      const uint32_t o = base + (QUAD ? threadIdx.x >> 2 : threadIdx.x);
      const bool valid = o < total;
      const uint32_t j = valid ? o / P : 0, ii = valid ? o - j * P : 0;
      const uint32_t i = ii < T ? ii : T + 1;            // output slot: U_ii, or S behind U_T
      G1Xyzz l = xyzz_identity(), r = xyzz_identity();
      const uint32_t li = (2 * j) * P + ii, ri = (2 * j + 1) * P + ii;   // input slot ii of both children (ii == T: their S)
      if (!valid) {
      } else if (cur) {
        l = xyzz_load(cur + li); r = xyzz_load(cur + ri);
      } else if (leaf) {      // P = 1: the node is the bucket itself
        const uint32_t b0 = leaf0 + 2 * j, b1 = b0 + 1;
        if (b0) l = xyzz_load(bk + (b0 - 1));
        r = xyzz_load(bk + (b1 - 1));
      } else {
        l = xyzz_load(src_g + li); r = xyzz_load(src_g + ri);
      }
      if (QUAD) xyzz_add_quad(l, r);
      else xyzz_add(l, r);
      if (!valid || (QUAD && (threadIdx.x & 3) != 0)) continue;    // the four lanes of a quad hold the same sum: one of them stores it
      if (!last) xyzz_store(nxt + (uint64_t)j * Pout + i, l);
      else if (groups > 1) xyzz_store(out + ((uint64_t)set * groups + grp) * Pout + i, l);
      else if (i < T) xyzz_store(U + set * c + i, l);     // the root: U_0 .. U_{c-3} here, U_{c-2} below; its S is not needed
    }
    // (b) U_T = R.S: a copy
    for (uint32_t j = threadIdx.x; j < nout; j += blockDim.x) {
      G1Xyzz r;
      const uint32_t ri = (2 * j + 1) * P + T;
      if (cur) r = xyzz_load(cur + ri);
      else if (leaf) r = xyzz_load(bk + (leaf0 + 2 * j));   // bucket index leaf0 + 2 j + 1
      else r = xyzz_load(src_g + ri);
      if (!last) xyzz_store(nxt + (uint64_t)j * Pout + T, r);
      else if (groups > 1) xyzz_store(out + ((uint64_t)set * groups + grp) * Pout + T, r);
      else xyzz_store(U + set * c + T, r);
    }
    __syncthreads();
    cur = nxt; nxt = (nxt == buf0) ? buf1 : buf0;
    nodes = nout; P = Pout;
  }
  if (groups == 1 && grp == 0 && threadIdx.x == 0) xyzz_store(U + set * c + (c - 1), xyzz_load(bk + (NB - 1)));   // b = NB: top bit only
}
// levels one launch can merge: input (2^m nodes of t0 + 1 points) is read from global memory; the two shared buffers hold the
// outputs of the first and second level
static uint32_t tree_levels(uint32_t t0, uint32_t remaining, size_t smem_cap) {
  uint32_t m = 1;
  while (m < remaining && m < 7) {
    const uint32_t mm = m + 1;
    const size_t pts = (size_t)(1u << (mm - 1)) * (t0 + 2) + (mm > 1 ? (size_t)(1u << (mm - 2)) * (t0 + 3) : 0);
    if (pts * sizeof(G1Xyzz) > smem_cap) break;
    m = mm;
  }
  return m;
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
  // measured on B200 (tools/sweep_c.py, round 2: the bucket back end costs about two additions per bucket now): shared-bucket
  // (precomputed) layout c = lg-1 up to 2^17, lg-2 at 2^18, lg-3 from 2^19; the per-window layout uses lg-4 throughout
  int c = (int)lg - 4;
  if (precomputed && lg <= 17) c = (int)lg - 1;
  else if (precomputed && lg == 18) c = (int)lg - 2;
  else if (precomputed) c = (int)lg - 3;
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
  while (T > 4 && e / T < MSM_MIN_THREADS) T >>= 1;
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
               o_redp = carve((size_t)2 * nc * g.sets * (g.NB / 16 + g.c + 8) * sizeof(G1Xyzz));   // two node arrays of the reduction tree
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
    if (ctx->tune.msm_accum_occ == 3) k_msm_accum<3><<<(unsigned)((nthreads + 127) / 128), 128, 0, st>>>(bases, ent_pt, ent_key, offsets, partial, g);
    else k_msm_accum<4><<<(unsigned)((nthreads + 127) / 128), 128, 0, st>>>(bases, ent_pt, ent_key, offsets, partial, g);
    ZKC_LAUNCH_CHECK(ctx);
    k_msm_fold<<<(unsigned)((nthreads + 127) / 128), 128, 0, st>>>(ent_key, offsets, partial, g);
    ZKC_LAUNCH_CHECK(ctx); }
  {
    // quads per bucket from the expected number of partials per bucket (entries per bucket / T)
    const double avg_partials = (double)g.W * (double)n / (double)g.NB / (double)g.sets / (double)g.T + 1.0;
    uint32_t logG = 0;
    while (logG < 3 && (double)(1u << logG) * 3.0 < avg_partials) ++logG;
    { ProfScope _p(ctx, "msm.gather");
      const uint64_t nslots = nbt << logG;
      if (nslots * 4 <= (uint64_t)ctx->sm_count * 256) {     // room for four lanes per addition (latency-bound batch)
        k_msm_gather<<<(unsigned)((nslots * 4 + 127) / 128), 128, 0, st>>>(offsets, partial, buckets, heavy, heavyc, g, logG);
      } else {
        k_msm_gather_plain<<<(unsigned)((nslots + 127) / 128), 128, 0, st>>>(offsets, partial, buckets, heavy, heavyc, g, logG);
      }
      ZKC_LAUNCH_CHECK(ctx); }
    { ProfScope _p(ctx, "msm.gather_heavy");
      k_msm_gather_heavy<<<ctx->sm_count * 4, 128, 8 * sizeof(G1Xyzz), st>>>(offsets, partial, buckets, heavy, heavyc, g);
      ZKC_LAUNCH_CHECK(ctx);
      k_msm_gather_giant1<<<dim3(MSM_GIANT_SLICES, 32), 128, 8 * sizeof(G1Xyzz), st>>>(offsets, partial, hpart, heavy, heavyc, g);
      ZKC_LAUNCH_CHECK(ctx);
      k_msm_gather_giant2<<<64, MSM_GIANT_SLICES * 4, 8 * sizeof(G1Xyzz), st>>>(hpart, buckets, heavy, heavyc, g);
      ZKC_LAUNCH_CHECK(ctx); }
  }
  {
    ProfScope _p(ctx, "msm.reduce");
    const uint32_t nsets = nc * g.sets;
    G1Xyzz* node[2] = {redp, redp + (size_t)nsets * (g.NB / 16 + g.c + 8)};
    uint32_t t0 = 0, nodes = g.NB, which = 0;      // NB = 2^(c-1) leaves: c - 1 levels
    const G1Xyzz* in = nullptr;
    while (nodes > 1) {
      const uint32_t remaining = g.c - 1 - t0;
      const uint32_t m = tree_levels(t0, remaining, 48 * 1024);
      const size_t pts = (size_t)(1u << (m - 1)) * (t0 + 2) + (m > 1 ? (size_t)(1u << (m - 2)) * (t0 + 3) : 0);
      // Four lanes per addition (xyzz_add_quad) while the launch leaves the machine mostly idle — the chain of dependent additions
      // is what it costs then —, one lane per addition when there are enough merges to fill it (many columns: throughput-bound).
      const uint32_t first = (1u << (m - 1)) * (t0 + 1);           // additions of the first, widest level of one CTA
      const uint64_t ctas = (uint64_t)(nodes >> m) * nsets;
      const bool quad = ctas * std::min<uint32_t>(first, 64) * 4 <= (uint64_t)ctx->sm_count * 512;
      if (quad) {
        const uint32_t threads = first <= 16 ? 64 : first <= 32 ? 128 : 256;
        k_msm_tree<true><<<dim3(nodes >> m, nsets), threads, pts * sizeof(G1Xyzz), st>>>(in, node[which], buckets, U, t0, m, nodes, g.NB, g.c);
      } else {
        const uint32_t threads = first <= 32 ? 32 : first <= 64 ? 64 : 128;
        k_msm_tree<false><<<dim3(nodes >> m, nsets), threads, pts * sizeof(G1Xyzz), st>>>(in, node[which], buckets, U, t0, m, nodes, g.NB, g.c);
      }
      ZKC_LAUNCH_CHECK(ctx);
      in = node[which]; which ^= 1;
      t0 += m; nodes >>= m;
    }
  }
  ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(total_out, offsets + nbt, 4, cudaMemcpyDeviceToDevice, st));
  return ZKC_OK;
}

// Enqueue one batch (all kernels + the async D2H of the c bit-plane sums per column) on ctx->stream.
int msm_enqueue(zkc_ctx* ctx, const Fr* scalars, const G1Affine* bases, uint64_t n, uint32_t nc, uint32_t c, bool precomputed, MsmPending* pend,
                int result_slot, bool team) {
  team = team && team_active(ctx) && n >= (uint64_t)ctx->team_world;
  const int shards = team ? ctx->team_world : 1;
  // Two ways to share a batch of commitments: by POINT RANGE (default: every rank walks 1/world of every column, the
  // split best_multiexp makes across threads) or, for batches of at least `world` columns and on request (tunable
  // team_commit_by_column), by COLUMN: rank r commits columns shard_range(nc, world, r) whole.  The column split runs the
  // bucket phases of nc / world bucket sets per rank instead of nc, but its accumulate work is only balanced when world divides nc.
  const bool by_column = team && ctx->tune.team_commit_by_column && nc >= (uint32_t)shards;
  const MsmGeom g0 = msm_geom(ctx, n, nc, c, precomputed);
  const uint32_t slot_cols = by_column ? (nc + (uint32_t)shards - 1) / (uint32_t)shards : nc;
  const size_t ubytes = (size_t)slot_cols * g0.sets * g0.c * sizeof(G1Xyzz), slot_bytes = (ubytes + 4 + 255) & ~(size_t)255;   // U, then the digit count
  char* dU;
  ZKC_TRY(scratch_reserve(ctx, SCR_MSM2, slot_bytes * shards, (void**)&dU));
  if (!team) {
    ZKC_TRY(msm_kernels(ctx, scalars, bases, g0, (G1Xyzz*)dU, (uint32_t*)(dU + ubytes)));
  } else if (by_column) {
    for (int r : team_ranks(ctx)) {
      uint64_t a, b;
      shard_range(nc, shards, r, &a, &b);
      char* slot = dU + (size_t)r * slot_bytes;
      ZKC_TRY(msm_kernels(ctx, scalars + a * n, bases, msm_geom(ctx, n, (uint32_t)(b - a), c, precomputed), (G1Xyzz*)slot, (uint32_t*)(slot + ubytes)));
    }
    ProfScope _p(ctx, "team.msm_allgather");
    ZKC_TRY(team_allgather(ctx, dU, slot_bytes));
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
  pend->g = g0; pend->nc = nc; pend->n = n; pend->hU = hU; pend->ubytes = ubytes; pend->shards = shards; pend->by_column = by_column; pend->done = ev; pend->active = true;
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
  if (pend->by_column) {
    for (int r = 0; r < pend->shards; ++r) {
      uint64_t a, b;
      shard_range(pend->nc, pend->shards, r, &a, &b);
      const G1Xyzz* Ur = (const G1Xyzz*)((const char*)pend->hU + (size_t)r * slot_bytes);
      for (uint64_t col = a; col < b; ++col) xyzz_to_abi(host_combine(Ur + (size_t)(col - a) * per_col, pend->g), out + col);
    }
    return ZKC_OK;
  }
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
  // bound memory: process columns in chunks sized by what ONE GPU holds — a team rank walks n / world points of every column
  // (point-range split) or all n points of its own columns (column split)
  team = team && team_active(ctx) && n >= (uint64_t)ctx->team_world;
  const uint32_t world = team ? (uint32_t)ctx->team_world : 1;
  const bool by_column = team && ctx->tune.team_commit_by_column && ncols >= world;
  const uint64_t n_rank = team && !by_column ? (n + world - 1) / world : n;
  MsmGeom g1 = msm_geom(ctx, n_rank, 1, c, precomputed);
  const uint64_t per_col = g1.emax() * 12 + g1.nbtot() * (128 + 12) + (g1.nbtot() + g1.emax() / g1.T + 1) * 128;
  uint32_t chunk = (uint32_t)std::max<uint64_t>(1, (3ull << 30) / per_col);
  if (by_column) chunk *= world;   // columns per batch: every rank takes at most chunk / world of them
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
