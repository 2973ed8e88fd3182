// ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.
//
// CPU restatement of the polynomial helpers create_proof uses (SURVEY.md §8a rows a8-a11,
// Appendix A.6-A.11): eval_polynomial, kate_division, prefix products, scalar/vector arithmetic.
// Upstream: halo2_proofs @4b42325 src/arithmetic.rs, src/poly.rs (un-vendored; Cargo.lock:1320-1336).
#include "orc_field.hpp"
#include <vector>
#include <thread>
#include <functional>

namespace orc {
int default_threads();
void parallel_chunks(size_t n, int threads, const std::function<void(size_t, size_t, int)>& f);
}
using namespace orc;

extern "C" {

// out[i] = a[i] (op) s   op: 0 add, 1 sub, 2 mul, 3 rsub (s - a[i])
void orc_fr_vec_scalar(int op, const uint64_t* a, const uint64_t* s, uint64_t* out, size_t n) {
  Fr sc; memcpy(sc.l, s, 32);
  const Fr* A = (const Fr*)a; Fr* O = (Fr*)out;
  parallel_chunks(n, default_threads(), [&](size_t lo, size_t hi, int) {
    for (size_t i = lo; i < hi; ++i) {
      switch (op) {
        case 0: O[i] = A[i] + sc; break;
        case 1: O[i] = A[i] - sc; break;
        case 2: O[i] = A[i] * sc; break;
        default: O[i] = sc - A[i];
      }
    }
  });
}
// threaded element-wise ops: 0 add, 1 sub, 2 mul
void orc_fr_vec_vec(int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
  const Fr* A = (const Fr*)a; const Fr* B = (const Fr*)b; Fr* O = (Fr*)out;
  parallel_chunks(n, default_threads(), [&](size_t lo, size_t hi, int) {
    for (size_t i = lo; i < hi; ++i) O[i] = op == 0 ? A[i] + B[i] : (op == 1 ? A[i] - B[i] : A[i] * B[i]);
  });
}
// batch inversion with zeros passed through (halo2 `batch_invert`)
void orc_fr_batch_invert(const uint64_t* a, uint64_t* out, size_t n) {
  const Fr* A = (const Fr*)a; Fr* O = (Fr*)out;
  parallel_chunks(n, default_threads(), [&](size_t lo, size_t hi, int) {
    std::vector<Fr> pre(hi - lo);
    Fr acc = Fr::one();
    for (size_t i = lo; i < hi; ++i) { pre[i - lo] = acc; if (!A[i].is_zero()) acc = acc * A[i]; }
    acc = acc.inv();
    for (size_t i = hi; i-- > lo;) {
      if (A[i].is_zero()) { O[i] = A[i]; continue; }
      Fr v = A[i];
      O[i] = acc * pre[i - lo];
      acc = acc * v;
    }
  });
}
// eval_polynomial: Horner
void orc_eval_poly(const uint64_t* coeffs, size_t n, const uint64_t* x, uint64_t* out) {
  Fr X; memcpy(X.l, x, 32);
  const Fr* c = (const Fr*)coeffs;
  int threads = default_threads();
  if (n < 4096) threads = 1;
  std::vector<Fr> part(threads, Fr::zero());
  std::vector<size_t> starts(threads, 0);
  size_t chunk = (n + threads - 1) / threads;
  parallel_chunks(n, threads, [&](size_t lo, size_t hi, int t) {
    Fr acc = Fr::zero();
    for (size_t i = hi; i-- > lo;) acc = acc * X + c[i];
    part[t] = acc; starts[t] = lo;
  });
  Fr res = Fr::zero();
  for (int t = 0; t < threads; ++t) {
    if (t > 0 && starts[t] == 0) continue;  // unused slot
    uint64_t e[4] = {starts[t], 0, 0, 0};
    res = res + part[t] * X.pow(e);
  }
  (void)chunk;
  memcpy(out, res.l, 32);
}
// kate_division: a(X) / (X - z), remainder dropped; out has n-1 coefficients
void orc_kate_division(const uint64_t* a, size_t n, const uint64_t* z, uint64_t* out) {
  Fr Z; memcpy(Z.l, z, 32);
  const Fr* A = (const Fr*)a; Fr* Q = (Fr*)out;
  if (n < 2) return;
  Fr tmp = Fr::zero();
  for (size_t i = n - 1; i >= 1; --i) {
    Fr lead = A[i] + tmp;      // upstream: b = -z; tmp = q[i]*b; lead_coeff -= tmp
    Q[i - 1] = lead;
    tmp = lead * Z;
  }
}
// out[0] = start; out[i+1] = out[i] * r[i] for i < n-1   (grand product column)
void orc_prefix_product(const uint64_t* r, size_t n, const uint64_t* start, uint64_t* out) {
  Fr acc; memcpy(acc.l, start, 32);
  const Fr* R = (const Fr*)r; Fr* O = (Fr*)out;
  for (size_t i = 0; i < n; ++i) { O[i] = acc; acc = acc * R[i]; }
}
// out[i] = base^i * first
void orc_powers(const uint64_t* base, const uint64_t* first, size_t n, uint64_t* out) {
  Fr b, f; memcpy(b.l, base, 32); memcpy(f.l, first, 32);
  Fr* O = (Fr*)out;
  parallel_chunks(n, default_threads(), [&](size_t lo, size_t hi, int) {
    uint64_t e[4] = {lo, 0, 0, 0};
    Fr w = f * b.pow(e);
    for (size_t i = lo; i < hi; ++i) { O[i] = w; w = w * b; }
  });
}

}  // extern "C"
