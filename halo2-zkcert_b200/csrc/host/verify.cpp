// zkc_verify: halo2 `plonk::verify_proof` with VerifierSHPLONK / VerifierGWC over KZG on BN254, on the host.
// (halo2_proofs 0.2.0 "halo2-axiom" @4b42325 src/plonk/verifier.rs, src/poly/kzg/multiopen/{shplonk,gwc}/verifier.rs,
// un-vendored — /root/reference/Cargo.lock:1320-1336.)  It is the self-check snark-verifier-sdk's gen_snark_shplonk
// runs right after create_proof in debug builds (/root/reference/src/helpers.rs:233,299) and what the reference's
// tests assert (src/tests/x509_aggregation.rs:64-105): the product can referee its own proofs without the oracle.
// Everything here is O(#queries) scalar / curve work plus one pairing product; no device is needed.
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>
#include "../../../include/zkcert_cuda.h"
#include "cs.h"
#include "hostutil.h"
#include "pairing.h"
#include "small_poly.h"

using namespace zkc;
using namespace zkc::host;

namespace {

const uint32_t ROOT_OF_UNITY_RAW[8] = {0x60c37c9cu, 0xd34f1ed9u, 0xd39329c8u, 0x3215cf6du, 0x3dd31f74u, 0x98865ea9u, 0x166d18b7u, 0x03ddb9f5u};
const uint32_t DELTA_RAW[8] = {0xe533e9a2u, 0x870e56bbu, 0x5e963f25u, 0x5b5f898eu, 0xd4c86e71u, 0x64ec26aau, 0x22c6f0cau, 0x09226b6eu};
Fr fr_words(const uint32_t w[8]) { Fr t; for (int i = 0; i < 8; ++i) t.v[i] = w[i]; return fe_from_canonical(t); }
Fr fr_u64(uint64_t x) { Fr t = fe_zero<FrP>(); t.v[0] = (uint32_t)x; t.v[1] = (uint32_t)(x >> 32); return fe_from_canonical(t); }

// [k]P on the host (XYZZ double-and-add); k in Montgomery form
G1Affine g1_mul(const G1Affine& p, const Fr& k_mont) {
  const Fr k = fe_to_canonical(k_mont);
  G1Xyzz acc = xyzz_identity();
  if (affine_is_identity(p) || fe_is_zero(k)) return xyzz_to_affine(acc);   // xyzz_madd's contract: the affine operand is not the identity
  for (int i = 255; i >= 0; --i) {
    acc = xyzz_dbl(acc);
    if ((k.v[i >> 5] >> (i & 31)) & 1) xyzz_madd(acc, p, false);
  }
  return xyzz_to_affine(acc);
}
G1Affine g1_add(const G1Affine& a, const G1Affine& b) {
  if (affine_is_identity(a)) return b;
  if (affine_is_identity(b)) return a;
  G1Xyzz acc = xyzz_from_affine(a);
  xyzz_add(acc, xyzz_from_affine(b));
  return xyzz_to_affine(acc);
}
G1Affine g1_neg(const G1Affine& a) { G1Affine r = a; if (!affine_is_identity(a)) r.y = fe_neg(a.y); return r; }
G1Affine g1_identity() { G1Affine r; r.x = fe_zero<FqP>(); r.y = fe_zero<FqP>(); return r; }
bool g1_on_curve(const G1Affine& p) {
  if (affine_is_identity(p)) return true;
  Fq three = fe_zero<FqP>(); three.v[0] = 3; three = fe_from_canonical(three);
  return fe_eq(fe_sqr(p.y), fe_add(fe_mul(fe_sqr(p.x), p.x), three));
}

// Lagrange basis polynomial of row i (mod n) at x:  (x^n - 1) / n * w^i / (x - w^i)
Fr l_i(const Fr& x, const Fr& xn, uint64_t n, const Fr& omega, const Fr& omega_inv, int64_t i) {
  const Fr wi = i >= 0 ? fe_pow_u64(omega, (uint64_t)i % n) : fe_pow_u64(omega_inv, (uint64_t)(-i) % n);
  const Fr num = fe_mul(fe_mul(fe_sub(xn, fe_one<FrP>()), fe_inv(fr_u64(n))), wi);
  return fe_mul(num, fe_inv(fe_sub(x, wi)));
}

struct Evals { const std::vector<Fr>*advice, *fixed, *instance; };
// postfix program over query evaluations; returns one value per expression
bool run_program(const HostProgram& h, const Evals& ev, std::vector<Fr>& out) {
  std::vector<Fr> st;
  for (size_t i = 0; i < h.words.size(); i += 2) {
    const uint32_t op = h.words[i], arg = h.words[i + 1];
    switch (op) {
      case 0: st.push_back(h.consts[arg]); break;
      case 1: st.push_back((*ev.advice)[arg]); break;
      case 2: st.push_back((*ev.fixed)[arg]); break;
      case 3: st.push_back((*ev.instance)[arg]); break;
      case 4: st.back() = fe_neg(st.back()); break;
      case 5: { Fr b = st.back(); st.pop_back(); st.back() = fe_add(st.back(), b); break; }
      case 6: { Fr b = st.back(); st.pop_back(); st.back() = fe_mul(st.back(), b); break; }
      case 7: st.back() = fe_mul(st.back(), h.consts[arg]); break;
      case 8: out.push_back(st.back()); st.pop_back(); break;
      default: return false;
    }
  }
  return st.empty();
}

struct VQuery { int id; Fr point; Fr eval; };
struct VSet { std::vector<Fr> points; std::vector<int> ids; std::vector<std::vector<Fr>> evals; };

int find_query(const std::vector<std::pair<uint32_t, int32_t>>& qs, uint32_t col) {
  for (size_t i = 0; i < qs.size(); ++i) if (qs[i].first == col && qs[i].second == 0) return (int)i;
  return -1;
}

}  // namespace

extern "C" int zkc_g2_generator(zkc_g2_affine* out) {
  if (!out) return ZKC_ERR_BAD_ARG;
  const G2Affine g = g2_generator();
  memcpy(out, &g, sizeof g);
  return ZKC_OK;
}
extern "C" int zkc_g2_mul(const zkc_g2_affine* p, const zkc_fr* scalar, zkc_g2_affine* out) {
  if (!p || !scalar || !out) return ZKC_ERR_BAD_ARG;
  G2Affine q; memcpy(&q, p, sizeof q);
  Fr s; memcpy(s.v, scalar, 32);
  if (!g2_on_curve(q)) return ZKC_ERR_BAD_ARG;
  const G2Affine r = g2_mul(q, s);
  memcpy(out, &r, sizeof r);
  return ZKC_OK;
}
extern "C" int zkc_pairing_check(const zkc_g1_affine* g1s, const zkc_g2_affine* g2s, size_t npairs, int* is_one) {
  if (!is_one || (npairs && (!g1s || !g2s))) return ZKC_ERR_BAD_ARG;
  std::vector<std::pair<G1Affine, G2Affine>> pairs(npairs);
  for (size_t i = 0; i < npairs; ++i) {
    memcpy(&pairs[i].first, &g1s[i], sizeof(G1Affine));
    memcpy(&pairs[i].second, &g2s[i], sizeof(G2Affine));
    if (!g1_on_curve(pairs[i].first) || !g2_on_curve(pairs[i].second)) return ZKC_ERR_BAD_ARG;
  }
  *is_one = pairing_product_is_one(pairs) ? 1 : 0;
  return ZKC_OK;
}

extern "C" int zkc_verify(const uint8_t* cs_blob, size_t cs_len, const zkc_g1_affine* fixed_comm, const zkc_g1_affine* sigma_comm,
                          const zkc_fr* transcript_repr, const zkc_g1_affine* g1_gen, const zkc_g2_affine* g2_abi, const zkc_g2_affine* s_g2_abi,
                          const zkc_fr* const* instances, const size_t* instance_lens, size_t num_instance_columns, const uint8_t* proof,
                          size_t proof_len, const zkc_prove_opts* opts, int* ok) {
  if (!cs_blob || !transcript_repr || !g1_gen || !g2_abi || !s_g2_abi || !proof || !opts || !ok) return ZKC_ERR_BAD_ARG;
  *ok = 0;
  if (opts->transcript < 0 || opts->transcript > 3 || opts->multiopen < 0 || opts->multiopen > 1) return ZKC_ERR_BAD_ARG;
  Cs cs;
  std::string perr;
  if (!parse_cs(cs_blob, cs_len, cs, perr)) return ZKC_ERR_BAD_ARG;
  if ((cs.num_fixed && !fixed_comm) || (!cs.perm.empty() && !sigma_comm) || (cs.num_instance && (!instances || !instance_lens))) return ZKC_ERR_BAD_ARG;
  if (num_instance_columns != cs.num_instance) return ZKC_ERR_INVALID_INSTANCES;   // plonk::Error::InvalidInstances upstream
  const uint64_t n = cs.n();
  const uint32_t bf = cs.blinding_factors, nsets = cs.nsets(), L = (uint32_t)cs.lookups.size();
  if (n < bf + 3) return ZKC_ERR_NOT_ENOUGH_ROWS;
  const Fr ONE = fe_one<FrP>(), ZERO = fe_zero<FrP>();
  Fr omega = fr_words(ROOT_OF_UNITY_RAW);
  for (uint32_t i = cs.k; i < 28; ++i) omega = fe_sqr(omega);
  const Fr omega_inv = fe_inv(omega);
  const Fr DELTA = fr_words(DELTA_RAW);
  G1Affine G; memcpy(&G, g1_gen, sizeof G);
  G2Affine g2, s_g2; memcpy(&g2, g2_abi, sizeof g2); memcpy(&s_g2, s_g2_abi, sizeof s_g2);
  if (!g1_on_curve(G) || !g2_on_curve(g2) || !g2_on_curve(s_g2)) return ZKC_ERR_BAD_ARG;
  auto load_g1 = [](const zkc_g1_affine* p, size_t i) { G1Affine a; memcpy(&a, &p[i], sizeof a); return a; };

  Transcript tr(opts->transcript, opts->point_format);
  tr.set_input(proof, proof_len);
  // a malformed proof is a rejected proof, not an error
#define RD_POINT(dst) do { if (tr.read_point(&(dst))) return ZKC_OK; } while (0)
#define RD_SCALAR(dst) do { if (tr.read_scalar(&(dst))) return ZKC_OK; } while (0)

  { Fr t; memcpy(t.v, transcript_repr, 32); tr.common_scalar(t); }
  std::vector<std::vector<Fr>> inst(cs.num_instance);
  for (uint32_t c = 0; c < cs.num_instance; ++c) {
    if (instance_lens[c] > cs.usable()) return ZKC_ERR_INVALID_INSTANCES;
    inst[c].resize(instance_lens[c]);
    for (size_t i = 0; i < instance_lens[c]; ++i) { memcpy(inst[c][i].v, &instances[c][i], 32); tr.common_scalar(inst[c][i]); }
  }
  std::vector<G1Affine> advice_comms(cs.num_advice);
  for (auto& p : advice_comms) RD_POINT(p);
  const Fr theta = tr.squeeze_challenge();
  std::vector<G1Affine> lk_a(L), lk_s(L), lk_z(L);
  for (uint32_t l = 0; l < L; ++l) { RD_POINT(lk_a[l]); RD_POINT(lk_s[l]); }
  const Fr beta = tr.squeeze_challenge();
  const Fr gamma = tr.squeeze_challenge();
  std::vector<G1Affine> perm_comms(nsets);
  for (auto& p : perm_comms) RD_POINT(p);
  for (uint32_t l = 0; l < L; ++l) RD_POINT(lk_z[l]);
  G1Affine random_comm;
  RD_POINT(random_comm);
  const Fr y = tr.squeeze_challenge();
  const uint32_t q = cs.degree - 1;
  std::vector<G1Affine> h_comms(q);
  for (auto& p : h_comms) RD_POINT(p);
  const Fr x = tr.squeeze_challenge();
  const Fr xn = fe_pow_u64(x, n);
  auto rot_point = [&](int32_t r) { return rotate_omega(x, omega, omega_inv, r); };

  // instance evaluations are the verifier's own (KZG: QUERY_INSTANCE = false)
  std::vector<Fr> instance_evals;
  for (auto& iq : cs.iq) {
    const Fr pt = rot_point(iq.second);
    const Fr ptn = fe_pow_u64(pt, n);
    Fr acc = ZERO;
    for (size_t i = 0; i < inst[iq.first].size(); ++i) acc = fe_add(acc, fe_mul(inst[iq.first][i], l_i(pt, ptn, n, omega, omega_inv, (int64_t)i)));
    instance_evals.push_back(acc);
  }
  std::vector<Fr> advice_evals(cs.aq.size()), fixed_evals(cs.fq.size()), sigma_evals(cs.perm.size());
  for (auto& e : advice_evals) RD_SCALAR(e);
  for (auto& e : fixed_evals) RD_SCALAR(e);
  Fr random_eval;
  RD_SCALAR(random_eval);
  for (auto& e : sigma_evals) RD_SCALAR(e);
  struct PermEv { Fr z, z_next, z_last; };
  std::vector<PermEv> perm_evals(nsets);
  for (uint32_t s = 0; s < nsets; ++s) {
    RD_SCALAR(perm_evals[s].z); RD_SCALAR(perm_evals[s].z_next);
    if (s + 1 != nsets) RD_SCALAR(perm_evals[s].z_last);
  }
  struct LkEv { Fr z, z_next, a, a_inv, s; };
  std::vector<LkEv> lk_evals(L);
  for (auto& e : lk_evals) { RD_SCALAR(e.z); RD_SCALAR(e.z_next); RD_SCALAR(e.a); RD_SCALAR(e.a_inv); RD_SCALAR(e.s); }

  // expected h(x)
  const Fr l_last = l_i(x, xn, n, omega, omega_inv, -(int64_t)(bf + 1));
  Fr l_blind = ZERO;
  for (uint32_t i = 1; i <= bf; ++i) l_blind = fe_add(l_blind, l_i(x, xn, n, omega, omega_inv, -(int64_t)i));
  const Fr l_0 = l_i(x, xn, n, omega, omega_inv, 0);
  const Fr active = fe_sub(ONE, fe_add(l_last, l_blind));
  const Evals evs{&advice_evals, &fixed_evals, &instance_evals};
  std::vector<Fr> exprs;
  if (!run_program(cs.gates, evs, exprs)) return ZKC_ERR_BAD_ARG;
  if (nsets) {
    auto col_eval = [&](uint32_t kind, uint32_t idx, Fr* out) -> bool {
      const int qi = find_query(kind == 0 ? cs.aq : (kind == 1 ? cs.fq : cs.iq), idx);
      if (qi < 0) return false;
      *out = kind == 0 ? advice_evals[qi] : (kind == 1 ? fixed_evals[qi] : instance_evals[qi]);
      return true;
    };
    exprs.push_back(fe_mul(l_0, fe_sub(ONE, perm_evals[0].z)));
    const Fr zl = perm_evals[nsets - 1].z;
    exprs.push_back(fe_mul(fe_sub(fe_sqr(zl), zl), l_last));
    for (uint32_t s = 1; s < nsets; ++s) exprs.push_back(fe_mul(fe_sub(perm_evals[s].z, perm_evals[s - 1].z_last), l_0));
    const uint32_t chunk = cs.chunk_len;
    for (uint32_t s = 0; s < nsets; ++s) {
      Fr left = perm_evals[s].z_next, right = perm_evals[s].z;
      Fr cur = fe_mul(fe_mul(beta, x), fe_pow_u64(DELTA, (uint64_t)s * chunk));
      for (uint32_t t = 0; t < chunk && (size_t)s * chunk + t < cs.perm.size(); ++t) {
        const size_t g = (size_t)s * chunk + t;
        Fr v;
        if (!col_eval(cs.perm[g].first, cs.perm[g].second, &v)) return ZKC_ERR_BAD_ARG;   // permutation column without a rotation-0 query
        left = fe_mul(left, fe_add(fe_add(v, fe_mul(beta, sigma_evals[g])), gamma));
        right = fe_mul(right, fe_add(fe_add(v, cur), gamma));
        cur = fe_mul(cur, DELTA);
      }
      exprs.push_back(fe_mul(fe_sub(left, right), active));
    }
  }
  for (uint32_t l = 0; l < L; ++l) {
    auto compress = [&](const HostProgram& h, Fr* out) -> bool {
      std::vector<Fr> vals;
      if (!run_program(h, evs, vals)) return false;
      Fr acc = ZERO;
      for (auto& v : vals) acc = fe_add(fe_mul(acc, theta), v);
      *out = acc;
      return true;
    };
    Fr cin, ctab;
    if (!compress(cs.lookups[l].first, &cin) || !compress(cs.lookups[l].second, &ctab)) return ZKC_ERR_BAD_ARG;
    const LkEv& e = lk_evals[l];
    exprs.push_back(fe_mul(l_0, fe_sub(ONE, e.z)));
    exprs.push_back(fe_mul(l_last, fe_sub(fe_sqr(e.z), e.z)));
    const Fr left = fe_mul(fe_mul(e.z_next, fe_add(e.a, beta)), fe_add(e.s, gamma));
    const Fr right = fe_mul(fe_mul(e.z, fe_add(cin, beta)), fe_add(ctab, gamma));
    exprs.push_back(fe_mul(fe_sub(left, right), active));
    exprs.push_back(fe_mul(l_0, fe_sub(e.a, e.s)));
    exprs.push_back(fe_mul(fe_mul(fe_sub(e.a, e.s), fe_sub(e.a, e.a_inv)), active));
  }
  Fr h_eval = ZERO;
  for (auto& v : exprs) h_eval = fe_add(fe_mul(h_eval, y), v);
  if (fe_is_zero(fe_sub(xn, ONE))) return ZKC_OK;   // x on the domain: negligible, rejected
  const Fr expected_h_eval = fe_mul(h_eval, fe_inv(fe_sub(xn, ONE)));
  G1Affine h_comm = g1_identity();
  for (size_t i = h_comms.size(); i-- > 0;) h_comm = g1_add(g1_mul(h_comm, xn), h_comms[i]);

  // queries in upstream order; commitment ids index `comms`
  std::vector<G1Affine> comms;
  std::vector<VQuery> queries;
  auto new_comm = [&](const G1Affine& c) { comms.push_back(c); return (int)comms.size() - 1; };
  std::vector<int> id_adv(cs.num_advice), id_pz(nsets), id_lz(L), id_la(L), id_ls(L), id_fix(cs.num_fixed), id_sig(cs.perm.size());
  for (uint32_t c = 0; c < cs.num_advice; ++c) id_adv[c] = new_comm(advice_comms[c]);
  for (uint32_t s = 0; s < nsets; ++s) id_pz[s] = new_comm(perm_comms[s]);
  for (uint32_t l = 0; l < L; ++l) { id_lz[l] = new_comm(lk_z[l]); id_la[l] = new_comm(lk_a[l]); id_ls[l] = new_comm(lk_s[l]); }
  for (uint32_t c = 0; c < cs.num_fixed; ++c) id_fix[c] = new_comm(load_g1(fixed_comm, c));
  for (size_t g = 0; g < cs.perm.size(); ++g) id_sig[g] = new_comm(load_g1(sigma_comm, g));
  const int id_h = new_comm(h_comm), id_rand = new_comm(random_comm);
  const Fr x_next = rot_point(1), x_last = rot_point(-(int32_t)(bf + 1)), x_inv = rot_point(-1);
  for (size_t i = 0; i < cs.aq.size(); ++i) queries.push_back({id_adv[cs.aq[i].first], rot_point(cs.aq[i].second), advice_evals[i]});
  for (uint32_t s = 0; s < nsets; ++s) {
    queries.push_back({id_pz[s], x, perm_evals[s].z});
    queries.push_back({id_pz[s], x_next, perm_evals[s].z_next});
  }
  for (int s = (int)nsets - 2; s >= 0; --s) queries.push_back({id_pz[s], x_last, perm_evals[s].z_last});
  for (uint32_t l = 0; l < L; ++l) {
    queries.push_back({id_lz[l], x, lk_evals[l].z});
    queries.push_back({id_la[l], x, lk_evals[l].a});
    queries.push_back({id_ls[l], x, lk_evals[l].s});
    queries.push_back({id_la[l], x_inv, lk_evals[l].a_inv});
    queries.push_back({id_lz[l], x_next, lk_evals[l].z_next});
  }
  for (size_t i = 0; i < cs.fq.size(); ++i) queries.push_back({id_fix[cs.fq[i].first], rot_point(cs.fq[i].second), fixed_evals[i]});
  for (size_t g = 0; g < cs.perm.size(); ++g) queries.push_back({id_sig[g], x, sigma_evals[g]});
  queries.push_back({id_h, x, expected_h_eval});
  queries.push_back({id_rand, x, random_eval});

  G1Affine left, right;   // accept iff e(left, [s]G2) == e(right, G2)
  if (opts->multiopen == 0) {
    const Fr yy = tr.squeeze_challenge();
    const Fr v = tr.squeeze_challenge();
    G1Affine h1, h2;
    RD_POINT(h1);
    const Fr u = tr.squeeze_challenge();
    RD_POINT(h2);
    // rotation sets: commitments grouped by their (sorted) point set, first-appearance order
    auto less = [](const Fr& a, const Fr& b) { return fr_cmp_canonical(a, b) < 0; };
    auto insert_sorted = [&](std::vector<Fr>& vec, const Fr& p) {
      for (auto& e : vec) if (fe_eq(e, p)) return;
      vec.insert(std::upper_bound(vec.begin(), vec.end(), p, less), p);
    };
    std::vector<Fr> super_points;
    std::vector<std::pair<int, std::vector<Fr>>> by_comm;
    for (auto& qq : queries) {
      insert_sorted(super_points, qq.point);
      bool found = false;
      for (auto& e : by_comm) if (e.first == qq.id) { insert_sorted(e.second, qq.point); found = true; break; }
      if (!found) by_comm.push_back({qq.id, std::vector<Fr>{qq.point}});
    }
    std::vector<VSet> sets;
    for (auto& bc : by_comm) {
      VSet* tgt = nullptr;
      for (auto& s : sets) {
        if (s.points.size() != bc.second.size()) continue;
        bool eq = true;
        for (size_t i = 0; i < s.points.size() && eq; ++i) eq = fe_eq(s.points[i], bc.second[i]);
        if (eq) { tgt = &s; break; }
      }
      if (!tgt) { sets.push_back(VSet()); tgt = &sets.back(); tgt->points = bc.second; }
      tgt->ids.push_back(bc.first);
      std::vector<Fr> ev;
      for (auto& pt : tgt->points)
        for (auto& qq : queries) if (qq.id == bc.first && fe_eq(qq.point, pt)) { ev.push_back(qq.eval); break; }
      tgt->evals.push_back(ev);
    }
    G1Affine outer = g1_identity();
    Fr r_outer = ZERO, z_0 = ZERO, z_0_diff_inv = ZERO, pv = ONE;
    for (size_t i = 0; i < sets.size(); ++i) {
      std::vector<Fr> diffs;
      for (auto& sp : super_points) {
        bool in = false;
        for (auto& p : sets[i].points) if (fe_eq(p, sp)) { in = true; break; }
        if (!in) diffs.push_back(sp);
      }
      Fr z_diff_i = vanishing_eval(diffs, u);
      if (i == 0) {
        z_0 = vanishing_eval(sets[i].points, u);
        if (fe_is_zero(z_diff_i)) return ZKC_OK;
        z_0_diff_inv = fe_inv(z_diff_i);
        z_diff_i = ONE;
      } else {
        z_diff_i = fe_mul(z_diff_i, z_0_diff_inv);
      }
      G1Affine inner = g1_identity();
      Fr r_inner = ZERO, py = ONE;
      const std::vector<std::vector<Fr>> basis = lagrange_basis(sets[i].points);
      for (size_t c = 0; c < sets[i].ids.size(); ++c) {
        const std::vector<Fr> r_x = interpolate_with_basis(basis, sets[i].evals[c]);
        r_inner = fe_add(r_inner, fe_mul(py, eval_small(r_x, u)));
        inner = g1_add(inner, g1_mul(comms[sets[i].ids[c]], py));
        py = fe_mul(py, yy);
      }
      const Fr w = fe_mul(pv, z_diff_i);
      outer = g1_add(outer, g1_mul(inner, w));
      r_outer = fe_add(r_outer, fe_mul(w, r_inner));
      pv = fe_mul(pv, v);
    }
    outer = g1_add(outer, g1_mul(G, fe_neg(r_outer)));
    outer = g1_add(outer, g1_mul(h1, fe_neg(z_0)));
    outer = g1_add(outer, g1_mul(h2, u));
    left = h2; right = outer;
  } else {
    const Fr v = tr.squeeze_challenge();
    std::vector<Fr> points;
    for (auto& qq : queries) {
      bool seen = false;
      for (auto& p : points) if (fe_eq(p, qq.point)) { seen = true; break; }
      if (!seen) points.push_back(qq.point);
    }
    std::vector<G1Affine> ws(points.size());
    for (auto& w : ws) RD_POINT(w);
    const Fr u = tr.squeeze_challenge();
    // sum_i u^i e(W_i, [s]G2) == sum_i u^i e(z_i W_i + C_i - [e_i]G, G2)
    left = g1_identity(); right = g1_identity();
    Fr pu = ONE;
    for (size_t i = 0; i < points.size(); ++i) {
      G1Affine cacc = g1_identity();
      Fr eacc = ZERO, pvv = ONE;
      for (auto& qq : queries) {
        if (!fe_eq(qq.point, points[i])) continue;
        cacc = g1_add(cacc, g1_mul(comms[qq.id], pvv));
        eacc = fe_add(eacc, fe_mul(qq.eval, pvv));
        pvv = fe_mul(pvv, v);
      }
      const G1Affine term = g1_add(g1_add(g1_mul(ws[i], points[i]), cacc), g1_mul(G, fe_neg(eacc)));
      left = g1_add(left, g1_mul(ws[i], pu));
      right = g1_add(right, g1_mul(term, pu));
      pu = fe_mul(pu, u);
    }
  }
  // Deliberately stricter than upstream: halo2's verify_proof never checks that the transcript is exhausted; a byte stream
  // with trailing garbage is rejected here (documented in DESIGN.md / INTEGRATION.md).
  if (tr.in_pos != proof_len) return ZKC_OK;
  std::vector<std::pair<G1Affine, G2Affine>> pairs;
  pairs.push_back({left, s_g2});
  pairs.push_back({g1_neg(right), g2});
  *ok = pairing_product_is_one(pairs) ? 1 : 0;
  return ZKC_OK;
#undef RD_POINT
#undef RD_SCALAR
}

// Host fold of the gate program exactly as the device interpreter folds it (acc = acc * mult + e, poly::eval_program), over
// given query values — with the stream as parsed (`factored` = 0) or with the factored stream the device runs (host/cs.h
// optimize_program).  Test hook for the program optimiser: both must give the same field element.  No device work.
extern "C" int zkc_host_fold_gates(const uint8_t* cs_blob, size_t cs_len, const zkc_fr* advice_q, const zkc_fr* fixed_q, const zkc_fr* instance_q,
                                   const zkc_fr* mult, int factored, zkc_fr* out, uint32_t* groups) {
  if (!cs_blob || !mult || !out) return ZKC_ERR_BAD_ARG;
  zkc::host::Cs cs;
  std::string err;
  if (!zkc::host::parse_cs(cs_blob, cs_len, cs, err)) return ZKC_ERR_BAD_ARG;
  if ((cs.aq.size() && !advice_q) || (cs.fq.size() && !fixed_q) || (cs.iq.size() && !instance_q)) return ZKC_ERR_BAD_ARG;
  std::vector<uint32_t> words = cs.gates.words;
  std::vector<Fr> pows;
  Fr m; memcpy(m.v, mult, 32);
  uint32_t ngroups = 0;
  if (factored) {
    const zkc::host::OptimizedProgram opt = zkc::host::optimize_program(cs.gates);
    words = opt.words;
    for (uint32_t L : opt.pow_len) pows.push_back(fe_pow_u64(m, L));
  }
  auto val = [](const zkc_fr* base, uint32_t i) { Fr f; memcpy(f.v, base + i, 32); return f; };
  std::vector<Fr> st;
  Fr acc = fe_zero<FrP>(), saved = fe_zero<FrP>();
  for (size_t i = 0; i < words.size(); i += 2) {
    const uint32_t op = words[i], arg = words[i + 1];
    switch (op) {
      case 0: st.push_back(cs.gates.consts[arg]); break;
      case 1: st.push_back(val(advice_q, arg)); break;
      case 2: st.push_back(val(fixed_q, arg)); break;
      case 3: st.push_back(val(instance_q, arg)); break;
      case 4: st.back() = fe_neg(st.back()); break;
      case 5: { Fr b = st.back(); st.pop_back(); st.back() = fe_add(st.back(), b); break; }
      case 6: { Fr b = st.back(); st.pop_back(); st.back() = fe_mul(st.back(), b); break; }
      case 7: st.back() = fe_mul(st.back(), cs.gates.consts[arg]); break;
      case 8: acc = fe_add(fe_mul(acc, m), st.back()); st.pop_back(); break;
      case 9: saved = acc; acc = fe_zero<FrP>(); ++ngroups; break;
      case 10: acc = fe_add(fe_mul(saved, pows[arg]), fe_mul(st.back(), acc)); st.pop_back(); break;
      case 11: case 12: case 13: st.back() = fe_mul(st.back(), val(op == 11 ? advice_q : (op == 12 ? fixed_q : instance_q), arg)); break;
      case 14: case 15: case 16: st.back() = fe_add(st.back(), val(op == 14 ? advice_q : (op == 15 ? fixed_q : instance_q), arg)); break;
      case 17: case 18: case 19: st.back() = fe_sub(st.back(), val(op == 17 ? advice_q : (op == 18 ? fixed_q : instance_q), arg)); break;
      case 20: { Fr b = st.back(); st.pop_back(); st.back() = fe_sub(st.back(), b); break; }
      case 21: st.back() = fe_add(st.back(), cs.gates.consts[arg]); break;
      default: return ZKC_ERR_BAD_ARG;
    }
  }
  if (!st.empty()) return ZKC_ERR_BAD_ARG;
  memcpy(out, acc.v, 32);
  if (groups) *groups = ngroups;
  return ZKC_OK;
}
