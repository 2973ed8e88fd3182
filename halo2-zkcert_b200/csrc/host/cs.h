// Constraint-system wire format shared by the prover (zkc_pk_load) and the host verifier (zkc_verify): the blob written by
// halo2-zkcert_b200/circuit.py `serialize` — what a Rust host derives from `pk.vk.cs` (halo2_proofs 0.2.0 @4b42325
// src/plonk/circuit.rs ConstraintSystem: queries, permutation columns, gate / lookup expressions as postfix programs).
// The derived numbers follow SURVEY.md Appendix A.5 (blinding_factors, degree, permutation chunk length).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include "../ff.cuh"

#define PROG_STACK_HOST_MAX 12   // must not exceed PROG_STACK of the device interpreter (poly.cu)

namespace zkc {
namespace host {


struct HostProgram {
  std::vector<uint32_t> words;     // (op, arg) pairs
  std::vector<Fr> consts;          // Montgomery
  uint32_t nexprs = 0;
  std::vector<uint32_t> degrees;   // per expression
};

struct Cs {
  uint32_t k = 0, num_advice = 0, num_fixed = 0, num_instance = 0, min_degree = 0;
  std::vector<std::pair<uint32_t, int32_t>> aq, fq, iq;
  std::vector<std::pair<uint32_t, uint32_t>> perm;   // (kind, column)
  HostProgram gates;
  std::vector<std::pair<HostProgram, HostProgram>> lookups;
  uint32_t blinding_factors = 0, degree = 0, chunk_len = 0;
  uint64_t n() const { return 1ull << k; }
  uint64_t usable() const { return n() - (blinding_factors + 1); }
  uint32_t nsets() const { return perm.empty() ? 0 : (uint32_t)((perm.size() + chunk_len - 1) / chunk_len); }
};

struct Reader {
  const uint8_t* p; size_t len, pos = 0; bool ok = true;
  uint32_t u32() { if (pos + 4 > len) { ok = false; return 0; } uint32_t v; memcpy(&v, p + pos, 4); pos += 4; return v; }
  void bytes(void* out, size_t n) { if (pos + n > len) { ok = false; memset(out, 0, n); return; } memcpy(out, p + pos, n); pos += n; }
};

inline bool parse_program(Reader& r, uint32_t nexprs, HostProgram& out) {
  out.nexprs = nexprs;
  const uint32_t npairs = r.u32();
  if (!r.ok || (size_t)npairs * 8 > r.len) return false;
  out.words.resize((size_t)npairs * 2);
  for (auto& w : out.words) w = r.u32();
  const uint32_t nconsts = r.u32();
  if (!r.ok || (size_t)nconsts * 32 > r.len) return false;
  out.consts.resize(nconsts);
  for (auto& c : out.consts) { Fr raw; r.bytes(raw.v, 32); c = fe_from_canonical(raw); }
  // degrees (and a structural check) by abstract interpretation of the postfix stream
  std::vector<uint32_t> st;
  uint32_t ends = 0;
  for (uint32_t i = 0; i < npairs; ++i) {
    const uint32_t op = out.words[2 * i], arg = out.words[2 * i + 1];
    switch (op) {
      case 0: if (arg >= nconsts) return false; st.push_back(0); break;
      case 1: case 2: case 3: st.push_back(1); break;
      case 4: if (st.empty()) return false; break;
      case 5: if (st.size() < 2) return false; { uint32_t b = st.back(); st.pop_back(); st.back() = std::max(st.back(), b); } break;
      case 6: if (st.size() < 2) return false; { uint32_t b = st.back(); st.pop_back(); st.back() += b; } break;
      case 7: if (st.empty() || arg >= nconsts) return false; break;
      case 8: if (st.size() != 1) return false; out.degrees.push_back(st.back()); st.clear(); ++ends; break;
      default: return false;
    }
    if (st.size() > PROG_STACK_HOST_MAX) return false;
  }
  return r.ok && ends == nexprs && st.empty();
}


// Parses the whole blob and fills the derived numbers.  Returns false with a message on malformed input.
inline bool parse_cs(const uint8_t* cs_blob, size_t cs_len, Cs& cs, std::string& err) {
  Reader r{cs_blob, cs_len};
  uint8_t magic[4]; r.bytes(magic, 4);
  if (memcmp(magic, "ZKCS", 4) != 0 || r.u32() != 1) { err = "bad constraint-system blob"; return false; }
  cs.k = r.u32(); cs.num_advice = r.u32(); cs.num_fixed = r.u32(); cs.num_instance = r.u32(); cs.min_degree = r.u32();
  if (!r.ok || cs.k > 26) { err = "bad k"; return false; }
  auto read_queries = [&](std::vector<std::pair<uint32_t, int32_t>>& q, uint32_t ncols) {
    const uint32_t m = r.u32();
    if (!r.ok || m > (1u << 20)) { r.ok = false; return; }
    q.resize(m);
    for (auto& e : q) { e.first = r.u32(); e.second = (int32_t)r.u32(); if (e.first >= ncols) r.ok = false; }
  };
  read_queries(cs.aq, cs.num_advice); read_queries(cs.fq, cs.num_fixed); read_queries(cs.iq, cs.num_instance);
  const uint32_t nperm = r.u32();
  if (!r.ok || nperm > (1u << 16)) { err = "truncated blob"; return false; }
  cs.perm.resize(nperm);
  for (auto& e : cs.perm) {
    e.first = r.u32(); e.second = r.u32();
    const uint32_t lim = e.first == 0 ? cs.num_advice : (e.first == 1 ? cs.num_fixed : cs.num_instance);
    if (e.first > 2 || e.second >= lim) r.ok = false;
  }
  const uint32_t npolys = r.u32();
  if (!r.ok || !parse_program(r, npolys, cs.gates)) { err = "bad gate program"; return false; }
  const uint32_t nlk = r.u32();
  if (!r.ok || nlk > 4096) { err = "truncated blob"; return false; }
  cs.lookups.resize(nlk);
  for (auto& lk : cs.lookups) {
    const uint32_t ni = r.u32();
    if (!r.ok || !parse_program(r, ni, lk.first)) { err = "bad lookup input program"; return false; }
    const uint32_t nt = r.u32();
    if (!r.ok || !parse_program(r, nt, lk.second)) { err = "bad lookup table program"; return false; }
  }
  // query indices inside programs must be in range
  auto check_prog = [&](const HostProgram& h) {
    for (size_t i = 0; i < h.words.size(); i += 2) {
      const uint32_t op = h.words[i], arg = h.words[i + 1];
      if ((op == 1 && arg >= cs.aq.size()) || (op == 2 && arg >= cs.fq.size()) || (op == 3 && arg >= cs.iq.size())) return false;
    }
    return true;
  };
  bool okp = check_prog(cs.gates);
  for (auto& lk : cs.lookups) okp = okp && check_prog(lk.first) && check_prog(lk.second);
  if (!okp) { err = "query index out of range"; return false; }
  // A.5 numbers
  std::vector<uint32_t> per_col(cs.num_advice, 0);
  for (auto& q : cs.aq) per_col[q.first]++;
  uint32_t f = 3;
  for (uint32_t c : per_col) f = std::max(f, c);
  cs.blinding_factors = f + 2;
  uint32_t d = 3;
  for (auto& lk : cs.lookups) {
    uint32_t di = 1, dt = 1;
    for (uint32_t x : lk.first.degrees) di = std::max(di, x);
    for (uint32_t x : lk.second.degrees) dt = std::max(dt, x);
    d = std::max(d, std::max(4u, 2 + di + dt));
  }
  for (uint32_t x : cs.gates.degrees) d = std::max(d, x);
  cs.degree = std::max(d, std::max(cs.min_degree, 1u));
  cs.chunk_len = cs.degree - 2;
  return true;
}

}  // namespace host
}  // namespace zkc
