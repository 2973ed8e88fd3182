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

// ---- ChaCha20Rng / Fr::random stream (rand_chacha 0.3.1 + halo2curves from_u512) -------------
// draw #i = keystream block (skip + i) of ChaCha20(key = seed, 64-bit block counter, stream 0), read as a
// 512-bit little-endian integer and reduced mod r.  Pinned against tests/pyref.py's pure-Python restatement
// and the known answers of SURVEY §8c-4.
static inline uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
extern "C" void orc_chacha_fr_random(const uint8_t* seed, int double_rounds, uint64_t skip, size_t count, uint64_t* out) {
  uint32_t key[8]; memcpy(key, seed, 32);
  parallel_chunks(count, default_threads(), [&](size_t lo, size_t hi, int) {
    for (size_t i = lo; i < hi; ++i) {
      const uint64_t ctr = skip + i;
      uint32_t s[16] = {0x61707865, 0x3320646e, 0x79622d32, 0x6b206574, key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                        (uint32_t)ctr, (uint32_t)(ctr >> 32), 0, 0};
      uint32_t x[16]; memcpy(x, s, sizeof x);
#define ORC_QR(a, b, c, d) \
  x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 16); x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 12); \
  x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 8);  x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 7);
      for (int r = 0; r < double_rounds; ++r) {
        ORC_QR(0, 4, 8, 12) ORC_QR(1, 5, 9, 13) ORC_QR(2, 6, 10, 14) ORC_QR(3, 7, 11, 15)
        ORC_QR(0, 5, 10, 15) ORC_QR(1, 6, 11, 12) ORC_QR(2, 7, 8, 13) ORC_QR(3, 4, 9, 14)
      }
#undef ORC_QR
      uint64_t wide[8];
      for (int j = 0; j < 8; ++j) wide[j] = (uint64_t)(x[2 * j] + s[2 * j]) | ((uint64_t)(x[2 * j + 1] + s[2 * j + 1]) << 32);
      Fr f = Fr::from_u512(wide);
      memcpy(out + 4 * i, f.l, 32);
    }
  });
}

// raw keystream words [first_word, first_word + count) of the same stream (BlockRng hands the stream out word by word:
// next_u64 = two consecutive words, fill_bytes(32) = eight words) — used where draws are not 16-word aligned (SURVEY OPEN-3)
extern "C" void orc_chacha_words(const uint8_t* seed, int double_rounds, uint64_t first_word, size_t count, uint32_t* out) {
  uint32_t key[8]; memcpy(key, seed, 32);
  size_t done = 0;
  while (done < count) {
    const uint64_t w = first_word + done, ctr = w / 16;
    uint32_t s[16] = {0x61707865, 0x3320646e, 0x79622d32, 0x6b206574, key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                      (uint32_t)ctr, (uint32_t)(ctr >> 32), 0, 0};
    uint32_t x[16]; memcpy(x, s, sizeof x);
#define ORC_QR(a, b, c, d) \
  x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 16); x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 12); \
  x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 8);  x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 7);
    for (int r = 0; r < double_rounds; ++r) {
      ORC_QR(0, 4, 8, 12) ORC_QR(1, 5, 9, 13) ORC_QR(2, 6, 10, 14) ORC_QR(3, 7, 11, 15)
      ORC_QR(0, 5, 10, 15) ORC_QR(1, 6, 11, 12) ORC_QR(2, 7, 8, 13) ORC_QR(3, 4, 9, 14)
    }
#undef ORC_QR
    for (size_t j = (size_t)(w % 16); j < 16 && done < count; ++j) out[done++] = x[j] + s[j];
  }
}
