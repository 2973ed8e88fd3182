// ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.
//
// CPU restatement of the halo2-axiom operators on the create_proof hot path (SURVEY.md §8a rows
// a3-a6, Appendix A.3/A.4/A.13): best_multiexp / multiexp_serial, best_fft, EvaluationDomain
// conversions and ParamsKZG::setup.  The upstream source (halo2_proofs @4b42325:
// src/arithmetic.rs, src/poly/domain.rs, src/poly/kzg/commitment.rs) is not vendored under
// /root/reference (Cargo.lock:1320-1336 only pins it); reference call sites that reach these
// operators are /root/reference/src/helpers.rs:210,213,226,233,262,265,279,299.
// Validated against Python big-integer arithmetic in tests/ — never used by the product path.
#include "orc_field.hpp"
#include <vector>
#include <thread>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <functional>

using namespace orc;

namespace orc {

int default_threads() {
  const char* e = getenv("ZKC_ORACLE_THREADS");
  if (e && atoi(e) > 0) return atoi(e);
  unsigned h = std::thread::hardware_concurrency();
  return h ? (int)h : 1;
}

// Equivalent of halo2's `parallelize`: contiguous chunks, one per thread.
void parallel_chunks(size_t n, int threads, const std::function<void(size_t, size_t, int)>& f) {
  if (threads <= 1 || n < (size_t)threads * 4) { f(0, n, 0); return; }
  std::vector<std::thread> th;
  size_t chunk = (n + threads - 1) / threads;
  int t = 0;
  for (size_t s = 0; s < n; s += chunk, ++t) {
    size_t e = std::min(n, s + chunk);
    th.emplace_back([=, &f] { f(s, e, t); });
  }
  for (auto& x : th) x.join();
}

// ---- best_multiexp (A.3) ---------------------------------------------------------------------
// Bucket states None / Affine / Projective as upstream; the sum is algorithm-independent.
struct Bucket {
  int state = 0;  // 0 none, 1 affine, 2 projective
  G1Affine a; G1 p;
  void add_assign(const G1Affine& o) {
    if (state == 0) { a = o; state = 1; }
    else if (state == 1) { p = g1_add_mixed(to_jac(a), o); state = 2; }
    else { p = g1_add_mixed(p, o); }
  }
  G1 add_to(const G1& other) const {
    if (state == 0) return other;
    if (state == 1) return g1_add_mixed(other, a);
    return g1_add(other, p);
  }
};

static inline uint64_t get_at(size_t segment, size_t c, const uint8_t bytes[32]) {
  size_t skip_bits = segment * c;
  size_t skip_bytes = skip_bits / 8;
  if (skip_bytes >= 32) return 0;
  uint8_t v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (size_t i = 0; i < 8 && skip_bytes + i < 32; ++i) v[i] = bytes[skip_bytes + i];
  uint64_t tmp; memcpy(&tmp, v, 8);
  tmp >>= (skip_bits - skip_bytes * 8);
  tmp %= (1ULL << c);
  return tmp;
}

G1 multiexp_serial(const Fr* coeffs, const G1Affine* bases, size_t n) {
  G1 acc = g1_identity();
  if (n == 0) return acc;
  std::vector<uint8_t> reprs(n * 32);
  for (size_t i = 0; i < n; ++i) { uint64_t r[4]; coeffs[i].to_raw(r); memcpy(&reprs[i * 32], r, 32); }
  size_t c;
  if (n < 4) c = 1; else if (n < 32) c = 3; else c = (size_t)std::ceil(std::log((double)n));
  size_t segments = (256 / c) + 1;
  std::vector<Bucket> buckets((1ULL << c) - 1);
  for (size_t seg = segments; seg-- > 0;) {
    for (size_t d = 0; d < c; ++d) acc = g1_double(acc);
    for (auto& b : buckets) b.state = 0;
    for (size_t i = 0; i < n; ++i) {
      uint64_t v = get_at(seg, c, &reprs[i * 32]);
      if (v != 0) buckets[v - 1].add_assign(bases[i]);
    }
    // summation by parts
    G1 running = g1_identity();
    for (size_t b = buckets.size(); b-- > 0;) {
      running = buckets[b].add_to(running);
      acc = g1_add(acc, running);
    }
  }
  return acc;
}

G1 best_multiexp(const Fr* coeffs, const G1Affine* bases, size_t n, int threads) {
  if (threads < 1) threads = 1;
  if (n > (size_t)threads && threads > 1) {
    size_t chunk = n / threads;
    size_t nchunks = (n + chunk - 1) / chunk;   // `chunks(chunk)` may yield one extra tail chunk
    std::vector<G1> results(nchunks, g1_identity());
    std::vector<std::thread> th;
    for (size_t t = 0; t < nchunks; ++t) {
      size_t s = t * chunk, e = std::min(n, s + chunk);
      th.emplace_back([&, s, e, t] { results[t] = multiexp_serial(coeffs + s, bases + s, e - s); });
    }
    for (auto& x : th) x.join();
    G1 acc = g1_identity();
    for (auto& r : results) acc = g1_add(acc, r);
    return acc;
  }
  return multiexp_serial(coeffs, bases, n);
}

// ---- best_fft (A.4) --------------------------------------------------------------------------
static inline uint32_t bitreverse(uint32_t n, uint32_t l) {
  uint32_t r = 0;
  for (uint32_t i = 0; i < l; ++i) { r = (r << 1) | (n & 1); n >>= 1; }
  return r;
}

static void recursive_butterfly(Fr* a, size_t n, size_t twiddle_chunk, const Fr* twiddles, int par_depth) {
  if (n == 2) {
    Fr t = a[1];
    a[1] = a[0] - t;
    a[0] = a[0] + t;
    return;
  }
  Fr* left = a; Fr* right = a + n / 2;
  if (par_depth > 0) {
    std::thread t1([&] { recursive_butterfly(left, n / 2, twiddle_chunk * 2, twiddles, par_depth - 1); });
    recursive_butterfly(right, n / 2, twiddle_chunk * 2, twiddles, par_depth - 1);
    t1.join();
  } else {
    recursive_butterfly(left, n / 2, twiddle_chunk * 2, twiddles, 0);
    recursive_butterfly(right, n / 2, twiddle_chunk * 2, twiddles, 0);
  }
  // case when twiddle factor is one
  {
    Fr t = right[0];
    right[0] = left[0] - t;
    left[0] = left[0] + t;
  }
  for (size_t i = 1; i < n / 2; ++i) {
    Fr t = right[i] * twiddles[i * twiddle_chunk];
    right[i] = left[i] - t;
    left[i] = left[i] + t;
  }
}

void best_fft(Fr* a, const Fr& omega, uint32_t log_n, int threads) {
  size_t n = (size_t)1 << log_n;
  if (log_n == 0) return;
  for (size_t k = 0; k < n; ++k) {
    size_t rk = bitreverse((uint32_t)k, log_n);
    if (k < rk) std::swap(a[k], a[rk]);
  }
  if (n == 1) return;
  std::vector<Fr> tw(n / 2 > 0 ? n / 2 : 1);
  // twiddles omega^i, i < n/2 (upstream fills them chunk-parallel; values are what matter)
  parallel_chunks(n / 2, threads, [&](size_t s, size_t e, int) {
    uint64_t ex[4] = {s, 0, 0, 0};
    Fr w = omega.pow(ex);
    for (size_t i = s; i < e; ++i) { tw[i] = w; w = w * omega; }
  });
  int depth = 0;
  while ((1 << (depth + 1)) <= threads) ++depth;
  if (log_n <= (uint32_t)depth || n <= 2) depth = 0;
  recursive_butterfly(a, n, 1, tw.data(), depth);
}

// ---- EvaluationDomain (A.4) ------------------------------------------------------------------
static const uint64_t ROOT_OF_UNITY_RAW[4] = {0xd34f1ed960c37c9cULL, 0x3215cf6dd39329c8ULL, 0x98865ea93dd31f74ULL, 0x03ddb9f5166d18b7ULL};
static const uint64_t ZETA_RAW[4] = {0xb8ca0b2d36636f23ULL, 0xcc37a73fec2bc5e9ULL, 0x048b6e193fd84104ULL, 0x30644e72e131a029ULL};
static const uint64_t ZETA_ALT_RAW[4] = {0x8b17ea66b99c90ddULL, 0x5bfc41088d8daaa7ULL, 0xb3c4d79d41a91758ULL, 0x0ULL};
static const uint64_t DELTA_RAW[4] = {0x870e56bbe533e9a2ULL, 0x5b5f898e5e963f25ULL, 0x64ec26aad4c86e71ULL, 0x09226b6e22c6f0caULL};
static const uint32_t FR_S = 28;

struct Domain {
  uint32_t k, extended_k, j;
  size_t n, ext_n;
  Fr omega, omega_inv, extended_omega, extended_omega_inv, g_coset, g_coset_inv;
  Fr ifft_divisor, extended_ifft_divisor;
  std::vector<Fr> t_evaluations;  // inverted, as upstream stores them
  int zeta_choice;
};

Domain make_domain(uint32_t j, uint32_t k, int zeta_choice) {
  Domain d; d.k = k; d.j = j; d.zeta_choice = zeta_choice;
  uint32_t quotient_poly_degree = j - 1;
  d.n = (size_t)1 << k;
  d.extended_k = k;
  while (((size_t)1 << d.extended_k) < d.n * quotient_poly_degree) d.extended_k++;
  d.ext_n = (size_t)1 << d.extended_k;
  Fr eo = Fr::from_raw(ROOT_OF_UNITY_RAW);
  for (uint32_t i = d.extended_k; i < FR_S; ++i) eo = eo.sqr();
  d.extended_omega = eo;
  d.extended_omega_inv = eo.inv();
  Fr o = eo;
  for (uint32_t i = k; i < d.extended_k; ++i) o = o.sqr();
  d.omega = o; d.omega_inv = o.inv();
  d.g_coset = Fr::from_raw(zeta_choice == 0 ? ZETA_RAW : ZETA_ALT_RAW);
  d.g_coset_inv = d.g_coset.sqr();
  d.ifft_divisor = Fr::from_u64(1ULL << k).inv();
  d.extended_ifft_divisor = Fr::from_u64(1ULL << d.extended_k).inv();
  size_t tl = (size_t)1 << (d.extended_k - k);
  d.t_evaluations.resize(tl);
  Fr cur = d.g_coset;
  uint64_t nexp[4] = {d.n, 0, 0, 0};
  for (size_t i = 0; i < tl; ++i) {
    d.t_evaluations[i] = (cur.pow(nexp) - Fr::one()).inv();
    cur = cur * d.extended_omega;
  }
  return d;
}

static void distribute_powers_zeta(const Domain& d, Fr* a, size_t len, bool into_coset, int threads) {
  Fr z1 = into_coset ? d.g_coset : d.g_coset_inv;
  Fr z2 = z1.sqr();
  parallel_chunks(len, threads, [&](size_t s, size_t e, int) {
    for (size_t i = s; i < e; ++i) {
      size_t m = i % 3;
      if (m == 1) a[i] = a[i] * z1; else if (m == 2) a[i] = a[i] * z2;
    }
  });
}

}  // namespace orc

// ---- C exports -------------------------------------------------------------------------------
extern "C" {

int orc_default_threads() { return default_threads(); }

// field: which = 0 (Fr) / 1 (Fq); op: 0 add, 1 sub, 2 mul, 3 inv(a), 4 from_canonical(a), 5 to_canonical(a), 6 neg(a), 7 from_u512(a has 8 limbs)
void orc_field_op(int which, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
  auto run = [&](auto tag) {
    typedef decltype(tag) F;
    for (size_t i = 0; i < n; ++i) {
      F x, y = F::zero(), r;
      if (op == 7) { r = F::from_u512(a + 8 * i); memcpy(out + 4 * i, r.l, 32); continue; }
      memcpy(x.l, a + 4 * i, 32);
      if (b) memcpy(y.l, b + 4 * i, 32);
      switch (op) {
        case 0: r = x + y; break;
        case 1: r = x - y; break;
        case 2: r = x * y; break;
        case 3: r = x.inv(); break;
        case 4: r = F::from_raw(x.l); break;
        case 5: x.to_raw(r.l); break;
        case 6: r = x.neg(); break;
        default: r = F::zero();
      }
      memcpy(out + 4 * i, r.l, 32);
    }
  };
  if (which == 0) run(Fr()); else run(Fq());
}

void orc_g1_generator(uint64_t* out) { G1Affine g = g1_generator(); memcpy(out, &g, 64); }

// out = [s]P, all affine, Montgomery; s is an Fr in Montgomery form
void orc_g1_mul(const uint64_t* p, const uint64_t* s, uint64_t* out) {
  G1Affine a; memcpy(&a, p, 64); Fr f; memcpy(f.l, s, 32);
  G1Affine r = to_affine(g1_mul(to_jac(a), f)); memcpy(out, &r, 64);
}
void orc_g1_add(const uint64_t* p, const uint64_t* q, uint64_t* out) {
  G1Affine a, b; memcpy(&a, p, 64); memcpy(&b, q, 64);
  G1Affine r = to_affine(g1_add(to_jac(a), to_jac(b))); memcpy(out, &r, 64);
}
int orc_g1_on_curve(const uint64_t* p) {
  G1Affine a; memcpy(&a, p, 64);
  if (a.is_identity()) return 1;
  return (a.y.sqr() == a.x.sqr() * a.x + Fq::from_u64(3)) ? 1 : 0;
}
// Jacobian (96 B) -> affine (64 B), n points
void orc_g1_to_affine(const uint64_t* jac, uint64_t* out, size_t n) {
  for (size_t i = 0; i < n; ++i) { G1 p; memcpy(&p, jac + 12 * i, 96); G1Affine a = to_affine(p); memcpy(out + 8 * i, &a, 64); }
}

// sum_i [s_i] P_i by plain double-and-add (the uniqueness check for MSM)
void orc_msm_naive(const uint64_t* scalars, const uint64_t* bases, size_t n, uint64_t* out_affine) {
  int threads = default_threads();
  std::vector<G1> part(threads, g1_identity());
  parallel_chunks(n, threads, [&](size_t s, size_t e, int t) {
    G1 acc = g1_identity();
    for (size_t i = s; i < e; ++i) {
      G1Affine a; memcpy(&a, bases + 8 * i, 64); Fr f; memcpy(f.l, scalars + 4 * i, 32);
      acc = g1_add(acc, g1_mul(to_jac(a), f));
    }
    part[t] = acc;
  });
  G1 acc = g1_identity();
  for (auto& p : part) acc = g1_add(acc, p);
  G1Affine r = to_affine(acc); memcpy(out_affine, &r, 64);
}

// best_multiexp restatement; out = affine (64 B)
void orc_best_multiexp(const uint64_t* scalars, const uint64_t* bases, size_t n, int threads, uint64_t* out_affine) {
  if (threads <= 0) threads = default_threads();
  G1 r = best_multiexp((const Fr*)scalars, (const G1Affine*)bases, n, threads);
  G1Affine a = to_affine(r); memcpy(out_affine, &a, 64);
}

void orc_best_fft(uint64_t* a, const uint64_t* omega, uint32_t log_n, int threads) {
  if (threads <= 0) threads = default_threads();
  Fr w; memcpy(w.l, omega, 32);
  best_fft((Fr*)a, w, log_n, threads);
}

// Domain constants: out[0]=omega, [1]=omega_inv, [2]=extended_omega, [3]=extended_omega_inv,
// [4]=g_coset, [5]=g_coset_inv, [6]=ifft_divisor, [7]=extended_ifft_divisor; returns extended_k.
uint32_t orc_domain_constants(uint32_t j, uint32_t k, int zeta_choice, uint64_t* out) {
  Domain d = make_domain(j, k, zeta_choice);
  const Fr* v[8] = {&d.omega, &d.omega_inv, &d.extended_omega, &d.extended_omega_inv, &d.g_coset, &d.g_coset_inv, &d.ifft_divisor, &d.extended_ifft_divisor};
  for (int i = 0; i < 8; ++i) memcpy(out + 4 * i, v[i]->l, 32);
  return d.extended_k;
}

// EvaluationDomain::lagrange_to_coeff: a (n) in place
void orc_lagrange_to_coeff(uint32_t j, uint32_t k, uint64_t* a, int threads) {
  if (threads <= 0) threads = default_threads();
  Domain d = make_domain(j, k, 0);
  Fr* p = (Fr*)a;
  best_fft(p, d.omega_inv, k, threads);
  parallel_chunks(d.n, threads, [&](size_t s, size_t e, int) { for (size_t i = s; i < e; ++i) p[i] = p[i] * d.ifft_divisor; });
}
// EvaluationDomain::coeff_to_lagrange (fft with omega) — used by keygen-side code
void orc_coeff_to_lagrange(uint32_t j, uint32_t k, uint64_t* a, int threads) {
  if (threads <= 0) threads = default_threads();
  Domain d = make_domain(j, k, 0);
  best_fft((Fr*)a, d.omega, k, threads);
}
// EvaluationDomain::coeff_to_extended: in (n coeffs) -> out (2^extended_k evaluations on the zeta coset)
void orc_coeff_to_extended(uint32_t j, uint32_t k, int zeta_choice, const uint64_t* in, uint64_t* out, int threads) {
  if (threads <= 0) threads = default_threads();
  Domain d = make_domain(j, k, zeta_choice);
  Fr* o = (Fr*)out;
  memcpy(o, in, d.n * 32);
  for (size_t i = d.n; i < d.ext_n; ++i) o[i] = Fr::zero();
  distribute_powers_zeta(d, o, d.n, true, threads);
  best_fft(o, d.extended_omega, d.extended_k, threads);
}
// EvaluationDomain::extended_to_coeff: a (2^extended_k) in place; first n*(j-1) entries are the result
void orc_extended_to_coeff(uint32_t j, uint32_t k, int zeta_choice, uint64_t* a, int threads) {
  if (threads <= 0) threads = default_threads();
  Domain d = make_domain(j, k, zeta_choice);
  Fr* p = (Fr*)a;
  best_fft(p, d.extended_omega_inv, d.extended_k, threads);
  parallel_chunks(d.ext_n, threads, [&](size_t s, size_t e, int) { for (size_t i = s; i < e; ++i) p[i] = p[i] * d.extended_ifft_divisor; });
  distribute_powers_zeta(d, p, d.ext_n, false, threads);
  // upstream truncates to n * quotient_poly_degree; zero the tail so callers can compare whole buffers
  for (size_t i = d.n * (j - 1); i < d.ext_n; ++i) p[i] = Fr::zero();
}
// EvaluationDomain::divide_by_vanishing_poly: a[i] *= t_inv[i mod 2^(ext_k-k)]
void orc_divide_by_vanishing(uint32_t j, uint32_t k, int zeta_choice, uint64_t* a, int threads) {
  if (threads <= 0) threads = default_threads();
  Domain d = make_domain(j, k, zeta_choice);
  Fr* p = (Fr*)a; size_t tl = d.t_evaluations.size();
  parallel_chunks(d.ext_n, threads, [&](size_t s, size_t e, int) { for (size_t i = s; i < e; ++i) p[i] = p[i] * d.t_evaluations[i % tl]; });
}

// [s_i]G for a batch (fixed-base, 8-bit windows) -> affine outputs; used by the SRS setup
void orc_fixed_base_batch(const uint64_t* scalars, size_t n, uint64_t* out_affine) {
  int threads = default_threads();
  // table[w][d] = [d * 256^w] G, d in 1..255
  std::vector<G1Affine> table(32 * 256);
  {
    G1 base = to_jac(g1_generator());
    for (int w = 0; w < 32; ++w) {
      G1 acc = g1_identity();
      std::vector<G1> row(256);
      for (int d = 1; d < 256; ++d) { acc = g1_add(acc, base); row[d] = acc; }
      for (int d = 1; d < 256; ++d) table[w * 256 + d] = to_affine(row[d]);
      table[w * 256] = g1a_identity();
      base = g1_add(acc, base);  // 256 * base
    }
  }
  parallel_chunks(n, threads, [&](size_t s, size_t e, int) {
    for (size_t i = s; i < e; ++i) {
      Fr f; memcpy(f.l, scalars + 4 * i, 32);
      uint64_t r[4]; f.to_raw(r);
      const uint8_t* b = (const uint8_t*)r;
      G1 acc = g1_identity();
      for (int w = 0; w < 32; ++w) if (b[w]) acc = g1_add_mixed(acc, table[w * 256 + b[w]]);
      G1Affine a = to_affine(acc);
      memcpy(out_affine + 8 * i, &a, 64);
    }
  });
}

// ParamsKZG::setup restatement (A.13): g[i] = [s^i]G, g_lagrange[i] = [l_i(s)]G, l_i(s) = (s^n-1)/n * w^i/(s-w^i)
void orc_srs_setup(uint32_t k, const uint64_t* s_mont, uint64_t* g_out, uint64_t* g_lagrange_out) {
  size_t n = (size_t)1 << k;
  Fr s; memcpy(s.l, s_mont, 32);
  std::vector<Fr> sc(n);
  Fr cur = Fr::one();
  for (size_t i = 0; i < n; ++i) { sc[i] = cur; cur = cur * s; }
  orc_fixed_base_batch((const uint64_t*)sc.data(), n, g_out);
  if (!g_lagrange_out) return;
  Domain d = make_domain(2, k, 0);
  uint64_t nexp[4] = {n, 0, 0, 0};
  Fr zn = (s.pow(nexp) - Fr::one()) * d.ifft_divisor;
  Fr w = Fr::one();
  for (size_t i = 0; i < n; ++i) { sc[i] = zn * w * (s - w).inv(); w = w * d.omega; }
  orc_fixed_base_batch((const uint64_t*)sc.data(), n, g_lagrange_out);
}

}  // extern "C"
