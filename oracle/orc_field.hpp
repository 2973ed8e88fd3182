// ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.
//
// CPU restatement of the BN254 field / curve arithmetic that halo2-zkcert reaches through
// halo2curves 0.4.0 (axiom fork @ e185711, /root/reference/Cargo.lock:1359-1380) and
// halo2_proofs 0.2.0 "halo2-axiom" (@ 4b42325, /root/reference/Cargo.lock:1320-1336).
// Neither crate is vendored under /root/reference and there is no Rust toolchain in this image,
// so this file follows the published algorithms and is
// validated by Python big-integer arithmetic (tests/test_oracle_*.py), not by upstream code.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
//
// Memory layout matches halo2curves: a field element is 4 x u64 little-endian limbs holding the
// Montgomery residue a*2^256 mod m; a G1Affine is (x, y) = 64 bytes with identity = (0, 0).
#pragma once
#include <cstdint>
#include <cstring>
#include <cstddef>

namespace orc {

typedef unsigned __int128 u128;

struct FrParams {
  static constexpr uint64_t M[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
  static constexpr uint64_t R[4] = {0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL};
  static constexpr uint64_t R2[4] = {0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL};
  static constexpr uint64_t INV = 0xc2e1f593efffffffULL;
};
struct FqParams {
  static constexpr uint64_t M[4] = {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
  static constexpr uint64_t R[4] = {0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL};
  static constexpr uint64_t R2[4] = {0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL};
  static constexpr uint64_t INV = 0x87d20782e4866389ULL;
};

template <class P>
struct Fp {
  uint64_t l[4];

  static Fp zero() { Fp r; r.l[0] = r.l[1] = r.l[2] = r.l[3] = 0; return r; }
  static Fp one() { Fp r; memcpy(r.l, P::R, 32); return r; }
  static Fp from_raw(const uint64_t v[4]) {  // canonical integer -> Montgomery
    Fp t; memcpy(t.l, v, 32);
    Fp r2; memcpy(r2.l, P::R2, 32);
    return t * r2;
  }
  static Fp from_u64(uint64_t v) { uint64_t t[4] = {v, 0, 0, 0}; return from_raw(t); }
  // halo2curves `from_u512` (Fr::random / from_uniform_bytes): value = lo + hi*2^256 mod m.
  static Fp from_u512(const uint64_t v[8]) {
    Fp lo, hi, r2;
    memcpy(lo.l, v, 32); memcpy(hi.l, v + 4, 32); memcpy(r2.l, P::R2, 32);
    // lo, hi may exceed m; the Montgomery product still reduces correctly for inputs < 2^256
    // provided the final conditional subtraction loop handles up to a few multiples of m.
    Fp r3 = r2 * r2;   // R^3 in Montgomery terms: mont(R2,R2) = R^2*R^2/R = R^3
    return mont_wide(lo, r2) + mont_wide(hi, r3);
  }
  void to_raw(uint64_t out[4]) const {  // Montgomery -> canonical integer
    Fp o; o.l[0] = 1; o.l[1] = o.l[2] = o.l[3] = 0;
    Fp r = (*this) * o;
    memcpy(out, r.l, 32);
  }
  bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
  bool operator==(const Fp& o) const { return l[0] == o.l[0] && l[1] == o.l[1] && l[2] == o.l[2] && l[3] == o.l[3]; }
  bool operator!=(const Fp& o) const { return !(*this == o); }

  static bool geq_m(const uint64_t a[4]) {
    for (int i = 3; i >= 0; --i) {
      if (a[i] > P::M[i]) return true;
      if (a[i] < P::M[i]) return false;
    }
    return true;
  }
  static void sub_m(uint64_t a[4]) {
    u128 b = 0;
    for (int i = 0; i < 4; ++i) {
      u128 d = (u128)a[i] - P::M[i] - (uint64_t)b;
      a[i] = (uint64_t)d;
      b = (d >> 64) & 1;
    }
  }
  Fp operator+(const Fp& o) const {
    Fp r; u128 c = 0;
    for (int i = 0; i < 4; ++i) { c += (u128)l[i] + o.l[i]; r.l[i] = (uint64_t)c; c >>= 64; }
    if (c || geq_m(r.l)) sub_m(r.l);
    return r;
  }
  Fp operator-(const Fp& o) const {
    Fp r; u128 b = 0;
    for (int i = 0; i < 4; ++i) {
      u128 d = (u128)l[i] - o.l[i] - (uint64_t)b;
      r.l[i] = (uint64_t)d; b = (d >> 64) & 1;
    }
    if (b) { u128 c = 0; for (int i = 0; i < 4; ++i) { c += (u128)r.l[i] + P::M[i]; r.l[i] = (uint64_t)c; c >>= 64; } }
    return r;
  }
  Fp neg() const { return zero() - *this; }
  Fp dbl() const { return *this + *this; }

  // CIOS Montgomery product; inputs may be any 256-bit values, output fully reduced.
  static Fp mont_wide(const Fp& a, const Fp& b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
      u128 c = 0;
      for (int j = 0; j < 4; ++j) { c += (u128)a.l[j] * b.l[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
      c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
      uint64_t m = t[0] * P::INV;
      c = (u128)m * P::M[0] + t[0]; c >>= 64;
      for (int j = 1; j < 4; ++j) { c += (u128)m * P::M[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
      c += t[4]; t[3] = (uint64_t)c; c >>= 64;
      t[4] = t[5] + (uint64_t)c; t[5] = 0;
    }
    Fp r; memcpy(r.l, t, 32);
    uint64_t hi = t[4];
    while (hi || geq_m(r.l)) {
      u128 bb = 0;
      for (int i = 0; i < 4; ++i) { u128 d = (u128)r.l[i] - P::M[i] - (uint64_t)bb; r.l[i] = (uint64_t)d; bb = (d >> 64) & 1; }
      hi -= (uint64_t)bb;
    }
    return r;
  }
  Fp operator*(const Fp& o) const { return mont_wide(*this, o); }
  Fp sqr() const { return mont_wide(*this, *this); }

  Fp pow(const uint64_t e[4]) const {
    Fp acc = one();
    for (int i = 255; i >= 0; --i) {
      acc = acc.sqr();
      if ((e[i / 64] >> (i % 64)) & 1) acc = acc * (*this);
    }
    return acc;
  }
  Fp pow_u64(uint64_t e) const { uint64_t ee[4] = {e, 0, 0, 0}; return pow(ee); }
  // Fermat inversion; 0 -> 0 (halo2curves' invert() returns CtOption; callers here treat 0 specially).
  Fp inv() const {
    uint64_t e[4]; memcpy(e, P::M, 32);
    e[0] -= 2;  // M[0] >= 2, no borrow
    return pow(e);
  }
  // compare canonical integers: -1, 0, 1  (halo2curves `Ord for Fr`)
  static int cmp_canonical(const Fp& a, const Fp& b) {
    uint64_t x[4], y[4]; a.to_raw(x); b.to_raw(y);
    for (int i = 3; i >= 0; --i) { if (x[i] < y[i]) return -1; if (x[i] > y[i]) return 1; }
    return 0;
  }
};

typedef Fp<FrParams> Fr;
typedef Fp<FqParams> Fq;

// ---- G1: y^2 = x^3 + 3 over Fq --------------------------------------------------------------
struct G1Affine { Fq x, y; bool is_identity() const { return x.is_zero() && y.is_zero(); } };
struct G1 { Fq x, y, z; bool is_identity() const { return z.is_zero(); } };

inline G1 g1_identity() { G1 r; r.x = Fq::zero(); r.y = Fq::one(); r.z = Fq::zero(); return r; }
inline G1Affine g1a_identity() { G1Affine r; r.x = Fq::zero(); r.y = Fq::zero(); return r; }
inline G1Affine g1_generator() { G1Affine g; g.x = Fq::from_u64(1); g.y = Fq::from_u64(2); return g; }
inline G1 to_jac(const G1Affine& a) {
  if (a.is_identity()) return g1_identity();
  G1 r; r.x = a.x; r.y = a.y; r.z = Fq::one(); return r;
}
inline G1 g1_double(const G1& p) {
  if (p.is_identity()) return p;
  // dbl-2009-l (a = 0)
  Fq A = p.x.sqr(), B = p.y.sqr(), C = B.sqr();
  Fq D = ((p.x + B).sqr() - A - C).dbl();
  Fq E = A.dbl() + A, F = E.sqr();
  G1 r;
  r.x = F - D.dbl();
  r.y = E * (D - r.x) - C.dbl().dbl().dbl();
  r.z = (p.y * p.z).dbl();
  return r;
}
inline G1 g1_add(const G1& p, const G1& q) {
  if (p.is_identity()) return q;
  if (q.is_identity()) return p;
  Fq z1z1 = p.z.sqr(), z2z2 = q.z.sqr();
  Fq u1 = p.x * z2z2, u2 = q.x * z1z1;
  Fq s1 = p.y * q.z * z2z2, s2 = q.y * p.z * z1z1;
  if (u1 == u2) {
    if (s1 == s2) return g1_double(p);
    return g1_identity();
  }
  Fq h = u2 - u1, r = s2 - s1;
  Fq hh = h.sqr(), hhh = h * hh, v = u1 * hh;
  G1 o;
  o.x = r.sqr() - hhh - v.dbl();
  o.y = r * (v - o.x) - s1 * hhh;
  o.z = p.z * q.z * h;
  return o;
}
inline G1 g1_add_mixed(const G1& p, const G1Affine& q) {
  if (q.is_identity()) return p;
  if (p.is_identity()) return to_jac(q);
  Fq z1z1 = p.z.sqr();
  Fq u2 = q.x * z1z1, s2 = q.y * p.z * z1z1;
  if (p.x == u2) {
    if (p.y == s2) return g1_double(p);
    return g1_identity();
  }
  Fq h = u2 - p.x, r = s2 - p.y;
  Fq hh = h.sqr(), hhh = h * hh, v = p.x * hh;
  G1 o;
  o.x = r.sqr() - hhh - v.dbl();
  o.y = r * (v - o.x) - p.y * hhh;
  o.z = p.z * h;
  return o;
}
inline G1 g1_neg(const G1& p) { G1 r = p; r.y = p.y.neg(); return r; }
inline G1Affine g1a_neg(const G1Affine& p) { G1Affine r = p; if (!p.is_identity()) r.y = p.y.neg(); return r; }
inline G1Affine to_affine(const G1& p) {
  if (p.is_identity()) return g1a_identity();
  Fq zi = p.z.inv(), zi2 = zi.sqr();
  G1Affine a; a.x = p.x * zi2; a.y = p.y * zi2 * zi; return a;
}
// scalar given as canonical 256-bit integer
inline G1 g1_mul_raw(const G1& p, const uint64_t e[4]) {
  G1 acc = g1_identity();
  for (int i = 255; i >= 0; --i) {
    acc = g1_double(acc);
    if ((e[i / 64] >> (i % 64)) & 1) acc = g1_add(acc, p);
  }
  return acc;
}
inline G1 g1_mul(const G1& p, const Fr& s) { uint64_t e[4]; s.to_raw(e); return g1_mul_raw(p, e); }

}  // namespace orc
