// Kernels specific to the create_proof pipeline: permutation / lookup grand-product inputs, the
// lookup permutation (permute_expression_pair), the permutation and lookup terms of h(X) on the
// extended coset, and bulk Fr::random generation.  SURVEY.md §8a rows a7-a10, Appendix A.6-A.8.
#pragma once
#include "poly.cuh"

namespace zkc {

#define PERM_MAX_CHUNK 8

// ---- Fr::random in bulk -------------------------------------------------------------------------------------
// rand_chacha hands its keystream out word by word (rand_core BlockRng): Fr::random (halo2curves from_u512 of 8 x
// next_u64) takes 16 consecutive words, fill_bytes(32) eight.  Draw #i of a run that starts at keystream word
// `first_word` covers words [first_word + 16 i, first_word + 16 i + 16): one block when the run is block aligned, the
// upper half of one block and the lower half of the next when it starts 8 words in (after an odd number of 32-byte seeds).
__device__ __forceinline__ uint32_t rotl32_d(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
struct ChaChaKey { uint32_t k[8]; };
__device__ __forceinline__ void chacha_block_dev(const uint32_t* key, uint64_t ctr, int double_rounds, uint32_t* out) {
  uint32_t s[16] = {0x61707865, 0x3320646e, 0x79622d32, 0x6b206574, key[0], key[1], key[2], key[3], key[4], key[5], key[6],
                    key[7], (uint32_t)ctr, (uint32_t)(ctr >> 32), 0, 0};
  uint32_t x[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) x[j] = s[j];
#define ZKC_QR(a, b, c, d) \
  x[a] += x[b]; x[d] = rotl32_d(x[d] ^ x[a], 16); x[c] += x[d]; x[b] = rotl32_d(x[b] ^ x[c], 12); \
  x[a] += x[b]; x[d] = rotl32_d(x[d] ^ x[a], 8);  x[c] += x[d]; x[b] = rotl32_d(x[b] ^ x[c], 7);
#pragma unroll 1
  for (int r = 0; r < double_rounds; ++r) {
    ZKC_QR(0, 4, 8, 12) ZKC_QR(1, 5, 9, 13) ZKC_QR(2, 6, 10, 14) ZKC_QR(3, 7, 11, 15)
    ZKC_QR(0, 5, 10, 15) ZKC_QR(1, 6, 11, 12) ZKC_QR(2, 7, 8, 13) ZKC_QR(3, 4, 9, 14)
  }
#undef ZKC_QR
#pragma unroll
  for (int j = 0; j < 16; ++j) out[j] = x[j] + s[j];
}
// from_u512: lo * R^2 / R + hi * R^3 / R.  The row multiplier (second operand) of fe_mul may be any
// 256-bit value: each CIOS row adds a * b_i with a < r, so the running value stays < 2r.
__device__ __forceinline__ Fr fr_from_u512_dev(const Fr& lo, const Fr& hi) {
  const Fr r2 = fe_r2<FrP>();
  const Fr r3 = fe_mul(r2, r2);
  return fe_add(fe_mul(r2, lo), fe_mul(r3, hi));
}
// first_word % 16 must be 0 or 8 (the host falls back to its own generator otherwise)
__global__ void k_chacha_fr(Fr* out, ChaChaKey key, uint64_t first_word, uint64_t count, int double_rounds) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const uint64_t w = first_word + 16 * i;
  uint32_t b0[16];
  chacha_block_dev(key.k, w >> 4, double_rounds, b0);
  Fr lo, hi;
  if ((w & 15) == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { lo.v[j] = b0[j]; hi.v[j] = b0[8 + j]; }
  } else {
    uint32_t b1[16];
    chacha_block_dev(key.k, (w >> 4) + 1, double_rounds, b1);
#pragma unroll
    for (int j = 0; j < 8; ++j) { lo.v[j] = b0[8 + j]; hi.v[j] = b1[j]; }
  }
  fe_store(out + i, fr_from_u512_dev(lo, hi));
}
// chunk j of `chunk_len` elements is the Fr::random stream of ChaCha20Rng::from_seed(keys[j]) from its start
// (the per-thread generators of the vanishing argument's random polynomial; SURVEY OPEN-3)
__global__ void k_chacha_fr_chunked(Fr* out, const ChaChaKey* keys, uint64_t chunk_len, uint64_t count) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const uint64_t c = i / chunk_len;
  uint32_t key[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) key[j] = keys[c].k[j];
  uint32_t b0[16];
  chacha_block_dev(key, i - c * chunk_len, 10, b0);
  Fr lo, hi;
#pragma unroll
  for (int j = 0; j < 8; ++j) { lo.v[j] = b0[j]; hi.v[j] = b0[8 + j]; }
  fe_store(out + i, fr_from_u512_dev(lo, hi));
}

// ---- permutation argument: numerators / denominators of the grand-product ratio ------------------------
struct PermSetArgs {
  const Fr* cols[PERM_MAX_CHUNK];    // column values (Lagrange rows or extended coset)
  const Fr* sigmas[PERM_MAX_CHUNK];  // sigma values in the same domain
  Fr delta_beta[PERM_MAX_CHUNK];     // beta * DELTA^(global column index)  [* zeta on the coset]
  uint32_t m;
};
// rows i < U:  num[i] = prod_t (v_t + delta_beta_t * omega^i + gamma),  den[i] = prod_t (v_t + beta * sigma_t + gamma)
// (rows [row0, row0 + cnt) of the set; team proving splits the rows across ranks)
__global__ void k_perm_num_den(PermSetArgs a, const Fr* omega_pows, Fr beta, Fr gamma, Fr* num, Fr* den, uint64_t row0, uint64_t cnt) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= cnt) return;
  const uint64_t i = row0 + tid;
  const Fr w = fe_load_nc(omega_pows + i);
  Fr nacc = fe_one<FrP>(), dacc = fe_one<FrP>();
  for (uint32_t t = 0; t < a.m; ++t) {
    const Fr v = fe_load(a.cols[t] + i);
    nacc = fe_mul(nacc, fe_add(fe_add(v, fe_mul(a.delta_beta[t], w)), gamma));
    dacc = fe_mul(dacc, fe_add(fe_add(v, fe_mul(beta, fe_load(a.sigmas[t] + i))), gamma));
  }
  fe_store(num + i, nacc);
  fe_store(den + i, dacc);
}
// lookup: num = (a + beta)(s + gamma), den = (a' + beta)(s' + gamma)
__global__ void k_lookup_num_den(const Fr* a, const Fr* s, const Fr* ap, const Fr* sp, Fr beta, Fr gamma, Fr* num, Fr* den, uint64_t row0,
                                 uint64_t cnt) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= cnt) return;
  const uint64_t i = row0 + tid;
  fe_store(num + i, fe_mul(fe_add(fe_load(a + i), beta), fe_add(fe_load(s + i), gamma)));
  fe_store(den + i, fe_mul(fe_add(fe_load(ap + i), beta), fe_add(fe_load(sp + i), gamma)));
}
// z_set[i] = scan[set*U + i] for i <= U; rows U+1.. from tails[set*bf + (i-U-1)]
__global__ void k_assemble_z(const Fr* scan, const Fr* tails, Fr* z, uint64_t n, uint64_t U, uint32_t bf, uint32_t nsets) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * nsets) return;
  const uint64_t set = idx / n, i = idx - set * n;
  Fr v = i <= U ? fe_load(scan + set * U + i) : fe_load(tails + set * bf + (i - U - 1));
  fe_store(z + idx, v);
}

// ---- lookup permutation ------------------------------------------------------------------------------------
// canonical copies padded with the all-ones sentinel (greater than any field element) up to n
__global__ void k_lookup_prepare(const Fr* in, Fr* out, uint64_t U, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr v;
  if (i < U) v = fe_to_canonical(fe_load(in + i));
  else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v.v[j] = 0xffffffffu;
  }
  fe_store(out + i, v);
}
__device__ __forceinline__ int u256_cmp(const Fr& a, const Fr& b) {
#pragma unroll
  for (int i = 7; i >= 0; --i) {
    if (a.v[i] > b.v[i]) return 1;
    if (a.v[i] < b.v[i]) return -1;
  }
  return 0;
}
// A, T sorted ascending (canonical).  rep[i] = 1 if row i repeats the previous input value;
// left[j] = 1 if table entry j is NOT consumed by a first occurrence; counts[0] += #first rows, counts[1] += #consumed
__global__ void k_lookup_flags(const Fr* A, const Fr* T, uint32_t* rep, uint32_t* left, uint32_t* counts, uint64_t U) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= U) return;
  const Fr a = fe_load(A + i);
  const bool firstA = (i == 0) || u256_cmp(a, fe_load(A + i - 1)) != 0;
  rep[i] = firstA ? 0u : 1u;
  const Fr t = fe_load(T + i);
  const bool firstT = (i == 0) || u256_cmp(t, fe_load(T + i - 1)) != 0;
  bool used = false;
  if (firstT) {   // binary search t in A[0..U)
    uint64_t lo = 0, hi = U;
    while (lo < hi) {
      const uint64_t mid = (lo + hi) >> 1;
      const int c = u256_cmp(fe_load(A + mid), t);
      if (c == 0) { used = true; break; }
      if (c < 0) lo = mid + 1; else hi = mid;
    }
  }
  left[i] = used ? 0u : 1u;
  if (firstA) atomicAdd(counts, 1u);
  if (used) atomicAdd(counts + 1, 1u);
}
__global__ void k_lookup_replist(const uint32_t* rep, const uint32_t* rep_rank, uint32_t* replist, uint64_t U) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= U) return;
  if (rep[i]) replist[rep_rank[i]] = (uint32_t)i;
}
// S'[i] = A[i] on first occurrences; the l-th leftover table value (ascending) goes to the (R-1-l)-th repeated row
// (PSE: `repeated_input_rows.pop()`), or to the l-th repeated row when `ascending` (axiom fork's rayon variant; SURVEY OPEN-9)
__global__ void k_lookup_assign(const Fr* A, const Fr* T, const uint32_t* rep, const uint32_t* left, const uint32_t* left_rank,
                                const uint32_t* replist, const uint32_t* totals /* [R] */, Fr* Sp, uint64_t U, int ascending) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= U) return;
  if (!rep[i]) fe_store(Sp + i, fe_load(A + i));
  if (left[i]) {
    const uint32_t R = totals[0];
    const uint32_t l = left_rank[i];
    if (l < R) fe_store(Sp + replist[ascending ? l : R - 1 - l], fe_load(T + i));
  }
}
// canonical -> Montgomery for rows < U, tails (already Montgomery) for rows >= U
__global__ void k_lookup_finish(const Fr* canon, const Fr* tail, Fr* out, uint64_t U, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  fe_store(out + i, i < U ? fe_from_canonical(fe_load(canon + i)) : fe_load(tail + (i - U)));
}

// ---- h(X): permutation and lookup terms on the extended coset ----------------------------------------------
#define PERM_MAX_SETS 32
struct PermFixedArgs { const Fr* z[PERM_MAX_SETS]; uint32_t nsets; };
// value = value*y + l0*(1 - z_0);  value*y + l_last*(z_l^2 - z_l);  for s >= 1: value*y + l0*(z_s - z_{s-1}[i + last_rot])
// (all h(X) kernels work on the CLASS-MAJOR extended coset (ntt.cu, dom_coeff_to_classes): `rows` = n is the length of one residue
//  class, row i belongs to class i / rows, a rotation moves inside the class; [row0, row0 + cnt) is the block this launch evaluates)
__device__ __forceinline__ uint64_t class_rot(uint64_t i, uint64_t rows, int64_t off) { return (i & ~(rows - 1)) | ((i + rows + off) & (rows - 1)); }
__global__ void k_quot_perm_fixed(Fr* value, PermFixedArgs a, const Fr* l0, const Fr* l_last, Fr y, uint64_t rows, int64_t last_off, uint64_t row0,
                                  uint64_t cnt) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= cnt) return;
  const uint64_t i = row0 + tid;
  Fr v = fe_load(value + i);
  const Fr L0 = fe_load(l0 + i), LL = fe_load(l_last + i);
  const Fr z0 = fe_load(a.z[0] + i);
  v = fe_add(fe_mul(v, y), fe_mul(fe_sub(fe_one<FrP>(), z0), L0));
  const Fr zl = fe_load(a.z[a.nsets - 1] + i);
  v = fe_add(fe_mul(v, y), fe_mul(fe_sub(fe_sqr(zl), zl), LL));
  const uint64_t r = class_rot(i, rows, last_off);
  for (uint32_t s = 1; s < a.nsets; ++s) {
    const Fr d = fe_sub(fe_load(a.z[s] + i), fe_load(a.z[s - 1] + r));
    v = fe_add(fe_mul(v, y), fe_mul(d, L0));
  }
  fe_store(value + i, v);
}
// value = value*y + l_active * ( z(wX) prod (v + beta sigma + gamma) - z(X) prod (v + delta_beta_t X + gamma) )
// X = zeta * w_ext^r is folded into delta_beta (zeta) and the twiddle table (w_ext^r), r = class + 2^e * (row inside the class).
__global__ void k_quot_perm_set(Fr* value, PermSetArgs a, const Fr* z, const Fr* l_active, const Fr* tw_ext, uint32_t ext_k, uint32_t log_rows,
                                Fr beta, Fr gamma, Fr y, uint64_t row0, uint64_t cnt) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= cnt) return;
  const uint64_t i = row0 + tid;
  const uint64_t rows = 1ull << log_rows, half = 1ull << (ext_k - 1);
  const uint64_t r = (i >> log_rows) + ((i & (rows - 1)) << (ext_k - log_rows));
  Fr w = fe_load_nc(tw_ext + (r >= half ? r - half : r));
  if (r >= half) w = fe_neg(w);
  Fr left = fe_load(z + class_rot(i, rows, 1));
  Fr right = fe_load(z + i);
  for (uint32_t t = 0; t < a.m; ++t) {
    const Fr v = fe_load(a.cols[t] + i);
    left = fe_mul(left, fe_add(fe_add(v, fe_mul(beta, fe_load(a.sigmas[t] + i))), gamma));
    right = fe_mul(right, fe_add(fe_add(v, fe_mul(a.delta_beta[t], w)), gamma));
  }
  const Fr v0 = fe_load(value + i);
  fe_store(value + i, fe_add(fe_mul(v0, y), fe_mul(fe_sub(left, right), fe_load(l_active + i))));
}
// the five lookup terms (A.7 / evaluation.rs order)
__global__ void k_quot_lookup(Fr* value, const Fr* zc, const Fr* ac, const Fr* sc, const Fr* comp_in, const Fr* comp_tab, const Fr* l0,
                              const Fr* l_last, const Fr* l_active, Fr beta, Fr gamma, Fr y, uint64_t rows, uint64_t row0, uint64_t cnt) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= cnt) return;
  const uint64_t i = row0 + tid;
  Fr v = fe_load(value + i);
  const Fr L0 = fe_load(l0 + i), LL = fe_load(l_last + i), LA = fe_load(l_active + i);
  const Fr z = fe_load(zc + i), zn = fe_load(zc + class_rot(i, rows, 1));
  const Fr ap = fe_load(ac + i), sp = fe_load(sc + i), apm = fe_load(ac + class_rot(i, rows, -1));
  const Fr table_value = fe_mul(fe_add(fe_load(comp_in + i), beta), fe_add(fe_load(comp_tab + i), gamma));
  const Fr a_minus_s = fe_sub(ap, sp);
  v = fe_add(fe_mul(v, y), fe_mul(fe_sub(fe_one<FrP>(), z), L0));
  v = fe_add(fe_mul(v, y), fe_mul(fe_sub(fe_sqr(z), z), LL));
  const Fr lhs = fe_mul(zn, fe_mul(fe_add(ap, beta), fe_add(sp, gamma)));
  v = fe_add(fe_mul(v, y), fe_mul(fe_sub(lhs, fe_mul(z, table_value)), LA));
  v = fe_add(fe_mul(v, y), fe_mul(a_minus_s, L0));
  v = fe_add(fe_mul(v, y), fe_mul(fe_mul(a_minus_s, fe_sub(ap, apm)), LA));
  fe_store(value + i, v);
}

__global__ void k_pk_mul_vec(const Fr* a, const Fr* b, Fr* out, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) fe_store(out + i, fe_mul(fe_load(a + i), fe_load(b + i)));
}

// compact witness columns (bit / byte / u16 / u64 cells) -> Montgomery Fr; kind: 1 bits (LSB first), 2 u8, 3 u16, 4 u64
__global__ void k_expand_compact(const uint8_t* src, Fr* out, uint64_t n, int kind) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t v;
  if (kind == 1) v = (src[i >> 3] >> (i & 7)) & 1;
  else if (kind == 2) v = src[i];
  else if (kind == 3) v = reinterpret_cast<const uint16_t*>(src)[i];
  else v = reinterpret_cast<const uint64_t*>(src)[i];
  Fr r;
  if (v == 0) r = fe_zero<FrP>();
  else if (v == 1) r = fe_one<FrP>();
  else { Fr c = fe_zero<FrP>(); c.v[0] = (uint32_t)v; c.v[1] = (uint32_t)(v >> 32); r = fe_from_canonical(c); }
  fe_store(out + i, r);
}

// rows [row0, row0 + rows) of each of `ncols` columns (stride n) := tails[c * rows + r] (or `value` when tails is null):
// the advice blinding policies of create_proof (SURVEY OPEN-1)
__global__ void k_fill_rows(Fr* cols, uint64_t n, uint64_t row0, uint32_t rows, uint32_t ncols, const Fr* tails, Fr value) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (uint64_t)rows * ncols) return;
  const uint64_t c = idx / rows, r = idx - c * rows;
  fe_store(cols + c * n + row0 + r, tails ? fe_load(tails + idx) : value);
}

// l_active = 1 - l_last - l_blind (extended coset)
__global__ void k_l_active(const Fr* l_last, const Fr* l_blind, Fr* out, uint64_t rows) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  fe_store(out + i, fe_sub(fe_sub(fe_one<FrP>(), fe_load(l_last + i)), fe_load(l_blind + i)));
}

}  // namespace zkc
