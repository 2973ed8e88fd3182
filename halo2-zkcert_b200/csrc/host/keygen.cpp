// Host-side pieces of keygen behind the C ABI, so that snark-verifier-sdk's `gen_pk` / `read_pk`
// (/root/reference/src/helpers.rs:213,265; src/bin/cli.rs:247,268,294,312,335,362,402,455) are a drop-in too:
//   * permutation::keygen::Assembly — the cycle structure the copy constraints build (sigma columns follow from it)
//   * ConstraintSystem::compress_selectors — which simple selectors share a fixed column, and the column values
//   * ProvingKey::{write, read} — the `.pk` files the CLI keeps next to the SRS, SerdeFormat::RawBytes[Unchecked]
// Upstream: halo2_proofs 0.2.0 "halo2-axiom" @4b42325 src/plonk/{permutation/keygen.rs, circuit/compress_selectors.rs, keygen.rs}
// and src/plonk.rs (un-vendored; /root/reference/Cargo.lock:1320-1336).  No device work in this file.
#include <algorithm>
#include <cstring>
#include <vector>
#include "../../../include/zkcert_cuda.h"
#include "../ff.cuh"

using namespace zkc;

// ---- permutation::keygen::Assembly::copy, replayed over the whole list of copy constraints -------------------------------------
// Cells are (index into cs.permutation.columns, row).  `mapping` starts as the identity, `aux` names each cell's cycle, `sizes`
// the cycle lengths; copy(left, right) merges the smaller cycle into the larger one (ties: right into left) by relabelling it
// and swapping the two cells' successors.  The result depends on the ORDER of the copies, as upstream's does.
extern "C" int zkc_keygen_permutation_mapping(uint32_t k, uint32_t num_columns, const uint32_t* copies, size_t num_copies, uint64_t* mapping_out) {
  if (k > 28 || !mapping_out || (num_copies && !copies)) return ZKC_ERR_BAD_ARG;
  const uint64_t n = 1ull << k, total = n * num_columns;
  std::vector<uint64_t> aux(total), sizes(total, 1);
  uint64_t* mp = mapping_out;
  for (uint64_t i = 0; i < total; ++i) { mp[i] = i; aux[i] = i; }
  for (size_t c = 0; c < num_copies; ++c) {
    const uint32_t* e = copies + 4 * c;
    if (e[0] >= num_columns || e[2] >= num_columns || e[1] >= n || e[3] >= n) return ZKC_ERR_BOUNDS_FAILURE;   // plonk::Error::BoundsFailure
    uint64_t left = (uint64_t)e[0] * n + e[1], right = (uint64_t)e[2] * n + e[3];
    if (aux[left] == aux[right]) continue;
    if (sizes[aux[left]] < sizes[aux[right]]) std::swap(left, right);
    const uint64_t lcyc = aux[left], rcyc = aux[right];
    sizes[lcyc] += sizes[rcyc];
    uint64_t i = rcyc;
    do { aux[i] = lcyc; i = mp[i]; } while (i != rcyc);
    std::swap(mp[left], mp[right]);
  }
  return ZKC_OK;
}

// ---- compress_selectors::process -----------------------------------------------------------------------------------------------
// activations: num_selectors columns of n bytes (0 / 1); max_degrees[s] = the largest degree of a gate the selector multiplies
// (0: the selector appears in no gate and gets a column of its own).  Greedy, in selector order: a selector opens a combination
// and takes in every later selector that is never active on the same row as a member and keeps
// (largest member degree - 1) + (number of members) <= max_degree.  Member j (1-based) of a combination of L selectors is
// "on" where the combination's fixed column holds j; the substitution expression upstream builds is
//     q * prod_{i = 1..L, i != j} (i - q)
// Outputs: combination_of[s], root_of[s] (the j above), columns_out[c * n + row] = 0 or the root active there, *num_combinations.
extern "C" int zkc_keygen_compress_selectors(uint32_t k, uint32_t num_selectors, const uint8_t* activations, const uint32_t* max_degrees,
                                             uint32_t max_degree, uint32_t* combination_of, uint32_t* root_of, uint32_t* combination_len,
                                             uint32_t* columns_out, uint32_t* num_combinations) {
  if (k > 28 || !num_combinations || (num_selectors && (!activations || !max_degrees || !combination_of || !root_of || !combination_len || !columns_out)))
    return ZKC_ERR_BAD_ARG;
  const uint64_t n = 1ull << k;
  const uint32_t S = num_selectors;
  uint32_t ncomb = 0;
  std::vector<char> added(S, 0);
  auto column = [&](uint32_t s) { return activations + (uint64_t)s * n; };
  auto emit = [&](const std::vector<uint32_t>& members) {
    uint32_t* col = columns_out + (uint64_t)ncomb * n;
    memset(col, 0, n * sizeof(uint32_t));
    for (uint32_t j = 0; j < members.size(); ++j) {
      const uint32_t s = members[j];
      combination_of[s] = ncomb; root_of[s] = j + 1; combination_len[s] = (uint32_t)members.size();
      const uint8_t* a = column(s);
      for (uint64_t r = 0; r < n; ++r) if (a[r]) col[r] = j + 1;
    }
    ++ncomb;
  };
  // selectors of degree 0 first: a fixed column each
  for (uint32_t s = 0; s < S; ++s) {
    if (max_degrees[s] > max_degree) return ZKC_ERR_BAD_ARG;
    if (max_degrees[s] == 0) { added[s] = 1; emit({s}); }
  }
  // exclusion matrix (lower triangle): two selectors active on one row cannot share a column
  std::vector<std::vector<char>> excl(S);
  for (uint32_t i = 0; i < S; ++i) {
    excl[i].assign(i, 0);
    if (added[i]) continue;
    for (uint32_t j = 0; j < i; ++j) {
      if (added[j]) continue;
      const uint8_t *a = column(i), *b = column(j);
      for (uint64_t r = 0; r < n; ++r) if (a[r] && b[r]) { excl[i][j] = 1; break; }
    }
  }
  for (uint32_t i = 0; i < S; ++i) {
    if (added[i]) continue;
    added[i] = 1;
    uint32_t d = max_degrees[i] - 1;          // the virtual selector's own contribution is counted through the member count
    std::vector<uint32_t> members{i};
    for (uint32_t j = i + 1; j < S; ++j) {
      if (d + members.size() == max_degree) break;       // nothing fits any more
      if (added[j]) continue;
      bool clash = false;
      for (uint32_t m : members) if (excl[j][m]) { clash = true; break; }
      if (clash) continue;
      const uint32_t nd = std::max(d, max_degrees[j] - 1);
      if (nd + members.size() + 1 > max_degree) continue;
      d = nd;
      members.push_back(j);
      added[j] = 1;
    }
    emit(members);
  }
  *num_combinations = ncomb;
  return ZKC_OK;
}

// ---- ProvingKey files ------------------------------------------------------------------------------------------------------------
// Layout of ProvingKey::write (SerdeFormat::RawBytes / RawBytesUnchecked: field elements and point coordinates as their
// in-memory Montgomery limbs, 32 B each) as recalled (SURVEY OPEN-8; `be` = 1: the u32 counts are big-endian, as upstream
// wrote them at this revision; 0: little-endian, the later convention):
//   VerifyingKey:  k u32 | num_fixed_commitments u32 | fixed commitments (64 B each) | permutation commitments (64 B each, count
//                  from the constraint system) | selectors: one bit per row, ceil(n / 8) bytes per selector (count from the cs)
//   l0 | l_last | l_active_row                          each:  len u32 | len * 32 B       (extended domain)
//   fixed_values | fixed_polys | fixed_cosets           each:  count u32 | count * (len u32 | len * 32 B)
//   permutation: permutations | polys | cosets          same shape, count = permutation columns
// zkc_pk_file_layout computes the byte offsets of every section; the reader (zkc_pk_read, prover.cu) streams fixed_values and
// permutations into the device and rebuilds everything else there.
namespace {
void put_u32(uint8_t* p, uint32_t v, int be) { for (int i = 0; i < 4; ++i) p[i] = (uint8_t)(v >> (be ? 24 - 8 * i : 8 * i)); }
uint32_t get_u32(const uint8_t* p, int be) { uint32_t v = 0; for (int i = 0; i < 4; ++i) v |= (uint32_t)p[i] << (be ? 24 - 8 * i : 8 * i); return v; }
}  // namespace

extern "C" int zkc_pk_file_layout(uint32_t k, uint32_t extended_k, uint32_t num_fixed, uint32_t num_perm, uint32_t num_selectors, zkc_pk_file_layout_t* out) {
  if (!out || k > 28 || extended_k < k || extended_k > 30) return ZKC_ERR_BAD_ARG;
  const uint64_t n = 1ull << k, en = 1ull << extended_k;
  uint64_t o = 0;
  out->k_off = o; o += 4;
  out->num_fixed_off = o; o += 4;
  out->fixed_commitments_off = o; o += 64ull * num_fixed;
  out->perm_commitments_off = o; o += 64ull * num_perm;
  out->selectors_off = o; o += (uint64_t)num_selectors * ((n + 7) / 8);
  auto poly = [&](uint64_t len) { const uint64_t at = o; o += 4 + 32 * len; return at; };
  auto slice = [&](uint32_t count, uint64_t len) { const uint64_t at = o; o += 4 + (uint64_t)count * (4 + 32 * len); return at; };
  out->l0_off = poly(en); out->l_last_off = poly(en); out->l_active_row_off = poly(en);
  out->fixed_values_off = slice(num_fixed, n); out->fixed_polys_off = slice(num_fixed, n); out->fixed_cosets_off = slice(num_fixed, en);
  out->perm_values_off = slice(num_perm, n); out->perm_polys_off = slice(num_perm, n); out->perm_cosets_off = slice(num_perm, en);
  out->total = o;
  return ZKC_OK;
}

// headers of a file (everything but the bulk data the caller copies in at the layout's offsets): counts and lengths
extern "C" int zkc_pk_file_write_headers(uint8_t* file, size_t cap, uint32_t k, uint32_t extended_k, uint32_t num_fixed, uint32_t num_perm,
                                         uint32_t num_selectors, int be) {
  zkc_pk_file_layout_t L;
  if (!file || zkc_pk_file_layout(k, extended_k, num_fixed, num_perm, num_selectors, &L) != ZKC_OK || cap < L.total) return ZKC_ERR_BAD_ARG;
  const uint64_t n = 1ull << k, en = 1ull << extended_k;
  put_u32(file + L.k_off, k, be);
  put_u32(file + L.num_fixed_off, num_fixed, be);
  for (uint64_t off : {L.l0_off, L.l_last_off, L.l_active_row_off}) put_u32(file + off, (uint32_t)en, be);
  auto slice = [&](uint64_t off, uint32_t count, uint64_t len) {
    put_u32(file + off, count, be);
    for (uint32_t c = 0; c < count; ++c) put_u32(file + off + 4 + (uint64_t)c * (4 + 32 * len), (uint32_t)len, be);
  };
  slice(L.fixed_values_off, num_fixed, n); slice(L.fixed_polys_off, num_fixed, n); slice(L.fixed_cosets_off, num_fixed, en);
  slice(L.perm_values_off, num_perm, n); slice(L.perm_polys_off, num_perm, n); slice(L.perm_cosets_off, num_perm, en);
  return ZKC_OK;
}

// structural check of a file against the constraint system's shape: every count / length field must be what the layout says
// (endianness `be`; -1 = detect from the k field) and the size must match.  *be_out receives the endianness found.
extern "C" int zkc_pk_file_check(const uint8_t* file, size_t len, uint32_t k, uint32_t extended_k, uint32_t num_fixed, uint32_t num_perm,
                                 uint32_t num_selectors, int be, int* be_out) {
  zkc_pk_file_layout_t L;
  if (!file || zkc_pk_file_layout(k, extended_k, num_fixed, num_perm, num_selectors, &L) != ZKC_OK) return ZKC_ERR_BAD_ARG;
  if (len != L.total) return ZKC_ERR_BAD_ARG;
  if (be < 0) {
    if (get_u32(file + L.k_off, 1) == k) be = 1;
    else if (get_u32(file + L.k_off, 0) == k) be = 0;
    else return ZKC_ERR_BAD_ARG;
  }
  if (be_out) *be_out = be;
  const uint64_t n = 1ull << k, en = 1ull << extended_k;
  if (get_u32(file + L.k_off, be) != k || get_u32(file + L.num_fixed_off, be) != num_fixed) return ZKC_ERR_BAD_ARG;
  for (uint64_t off : {L.l0_off, L.l_last_off, L.l_active_row_off}) if (get_u32(file + off, be) != (uint32_t)en) return ZKC_ERR_BAD_ARG;
  auto slice = [&](uint64_t off, uint32_t count, uint64_t plen) {
    if (get_u32(file + off, be) != count) return false;
    for (uint32_t c = 0; c < count; ++c) if (get_u32(file + off + 4 + (uint64_t)c * (4 + 32 * plen), be) != (uint32_t)plen) return false;
    return true;
  };
  if (!slice(L.fixed_values_off, num_fixed, n) || !slice(L.fixed_polys_off, num_fixed, n) || !slice(L.fixed_cosets_off, num_fixed, en) ||
      !slice(L.perm_values_off, num_perm, n) || !slice(L.perm_polys_off, num_perm, n) || !slice(L.perm_cosets_off, num_perm, en))
    return ZKC_ERR_BAD_ARG;
  return ZKC_OK;
}

// SerdeFormat::RawBytes validation of one column of a file: every element canonical (< r).  Unchecked files skip this.
extern "C" int zkc_fr_column_is_canonical(const zkc_fr* col, size_t n) {
  const uint8_t* a = reinterpret_cast<const uint8_t*>(col);    // file sections are only 4-byte aligned
  for (size_t i = 0; i < n; ++i) { uint32_t v[8]; memcpy(v, a + 32 * i, 32); if (geq_mod<FrP>(v)) return 0; }
  return 1;
}
