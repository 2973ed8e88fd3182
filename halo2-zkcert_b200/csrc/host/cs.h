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


// ---- device-side program optimisation: a shared leading factor is taken out of a run of expressions --------------------------------
// Gates come as `selector * constraint` (Constraints::with_selector, or q.clone() * expr by hand), and consecutive constraints
// usually share the selector.  The interpreter folds expressions with Horner in `mult`: acc = acc * mult + e_i.  For a run of
// L >= 3 expressions e_i = S * r_i with the same leaf S (a fixed / advice / instance query):
//      acc' = acc * mult^L + S * (sum_i r_i mult^(L - i))
// — exact field algebra, so the value is bit-identical, with L - 2 fewer products per row.  Emitted as
//      GROUP_BEGIN (9)           saved = acc; acc = 0
//      r_1 END r_2 END ... r_L END
//      S  GROUP_END (10, slot)   acc = saved * pow[slot] + S * acc        (pow[slot] = mult^L, filled per launch)
// Only the device copy is rewritten; the verifier and the degree analysis keep the original stream.
inline bool complete_expr(const std::vector<uint32_t>& w, size_t a, size_t b) {   // pairs [a, b) leave exactly one value, never underflow
  int d = 0;
  for (size_t i = a; i < b; ++i) {
    const uint32_t op = w[2 * i];
    if (op <= 3) ++d;
    else if (op == 4 || op == 7) { if (d < 1) return false; }
    else if (op == 5 || op == 6) { if (d < 2) return false; --d; }
    else return false;
  }
  return d == 1;
}
struct OptimizedProgram { std::vector<uint32_t> words; std::vector<uint32_t> pow_len; };
inline OptimizedProgram optimize_program(const HostProgram& h, uint32_t max_slots = 8) {
  struct Poly { size_t a, b; bool fact; uint32_t lop, larg; size_t ra, rb; };   // pairs [a, b) without the END; rest = [ra, rb)
  const std::vector<uint32_t>& w = h.words;
  std::vector<Poly> polys;
  size_t start = 0;
  for (size_t i = 0; i < w.size() / 2; ++i) {
    if (w[2 * i] != 8) continue;
    Poly p{start, i, false, 0, 0, 0, 0};
    if (i >= start + 3 && w[2 * (i - 1)] == 6) {
      const uint32_t f = w[2 * start], l = w[2 * (i - 2)];
      if (f >= 1 && f <= 3 && complete_expr(w, start + 1, i - 1)) { p.fact = true; p.lop = f; p.larg = w[2 * start + 1]; p.ra = start + 1; p.rb = i - 1; }
      else if (l >= 1 && l <= 3 && complete_expr(w, start, i - 2)) { p.fact = true; p.lop = l; p.larg = w[2 * (i - 2) + 1]; p.ra = start; p.rb = i - 2; }
    }
    polys.push_back(p);
    start = i + 1;
  }
  OptimizedProgram out;
  auto copy = [&](size_t a, size_t b) { out.words.insert(out.words.end(), w.begin() + 2 * a, w.begin() + 2 * b); };
  for (size_t i = 0; i < polys.size();) {
    size_t j = i + 1;
    if (polys[i].fact) while (j < polys.size() && polys[j].fact && polys[j].lop == polys[i].lop && polys[j].larg == polys[i].larg) ++j;
    const uint32_t L = (uint32_t)(j - i);
    int slot = -1;
    if (polys[i].fact && L >= 3) {
      for (size_t s = 0; s < out.pow_len.size(); ++s) if (out.pow_len[s] == L) slot = (int)s;
      if (slot < 0 && out.pow_len.size() < max_slots) { out.pow_len.push_back(L); slot = (int)out.pow_len.size() - 1; }
    }
    if (slot < 0) { for (size_t t = i; t < j; ++t) { copy(polys[t].a, polys[t].b); out.words.push_back(8); out.words.push_back(0); } i = j; continue; }
    out.words.push_back(9); out.words.push_back(0);
    for (size_t t = i; t < j; ++t) { copy(polys[t].ra, polys[t].rb); out.words.push_back(8); out.words.push_back(0); }
    out.words.push_back(polys[i].lop); out.words.push_back(polys[i].larg);
    out.words.push_back(10); out.words.push_back((uint32_t)slot);
    i = j;
  }
  if (max_slots == 0) return out;   // optimisation off: the stream as parsed
  // Peephole pass: an operand that is pushed only to be consumed by the next operation is folded into it, so the interpreter
  // (whose stack beyond the top value lives in local memory) neither stores nor reloads it:
  //   leaf MUL -> MUL_leaf (11 + kind)    leaf ADD -> ADD_leaf (14 + kind)    leaf NEG ADD -> SUB_leaf (17 + kind)
  //   NEG ADD  -> SUB (20)                const ADD -> ADD_CONST (21)         const MUL -> SCALE (7)
  // (kind = 0 advice, 1 fixed, 2 instance).  Same field operations in the same order: nothing changes but the bookkeeping.
  std::vector<uint32_t> pp;
  const std::vector<uint32_t>& v = out.words;
  const size_t np = v.size() / 2;
  auto op_at = [&](size_t i) { return i < np ? v[2 * i] : 0xffffffffu; };
  for (size_t i = 0; i < np;) {
    const uint32_t op = v[2 * i], arg = v[2 * i + 1];
    // a fused operand needs a value underneath it: never at the start of an expression (there the leaf is the left operand)
    const bool has_below = i > 0 && v[2 * (i - 1)] != 8 && v[2 * (i - 1)] != 9 && v[2 * (i - 1)] != 10;
    if (op >= 1 && op <= 3 && has_below) {
      if (op_at(i + 1) == 6) { pp.push_back(11 + (op - 1)); pp.push_back(arg); i += 2; continue; }
      if (op_at(i + 1) == 5) { pp.push_back(14 + (op - 1)); pp.push_back(arg); i += 2; continue; }
      if (op_at(i + 1) == 4 && op_at(i + 2) == 5) { pp.push_back(17 + (op - 1)); pp.push_back(arg); i += 3; continue; }
    }
    if (op == 0 && has_below) {
      if (op_at(i + 1) == 5) { pp.push_back(21); pp.push_back(arg); i += 2; continue; }
      if (op_at(i + 1) == 6) { pp.push_back(7); pp.push_back(arg); i += 2; continue; }
    }
    if (op == 4 && op_at(i + 1) == 5) { pp.push_back(20); pp.push_back(0); i += 2; continue; }
    pp.push_back(op); pp.push_back(arg); ++i;
  }
  out.words.swap(pp);
  return out;
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
