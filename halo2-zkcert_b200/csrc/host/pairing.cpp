// Host BN254 pairing (see pairing.h).  Textbook construction, kept small rather than fast: the product calls it once
// per verified proof.
#include "pairing.h"
#include <cstring>

namespace zkc {
namespace host {

namespace {
typedef unsigned long long u64w;

Fq fq_small(uint32_t v) { Fq t = fe_zero<FqP>(); t.v[0] = v; return fe_from_canonical(t); }
Fq fq_from_words(const uint32_t w[8]) { Fq t; for (int i = 0; i < 8; ++i) t.v[i] = w[i]; return fe_from_canonical(t); }

// generator of G2 (EIP-197 / halo2curves G2Affine::generator), canonical little-endian 32-bit words
const uint32_t G2_X0[8] = {0xd992f6edu, 0x46debd5cu, 0xf75edaddu, 0x674322d4u, 0x5e5c4479u, 0x426a0066u, 0x121f1e76u, 0x1800deefu};
const uint32_t G2_X1[8] = {0xaef312c2u, 0x97e485b7u, 0x35a9e712u, 0xf1aa4933u, 0x31fb5d25u, 0x7260bfb7u, 0x920d483au, 0x198e9393u};
const uint32_t G2_Y0[8] = {0x66fa7daau, 0x4ce6cc01u, 0x0c43d37bu, 0xe3d1e769u, 0x8dcb408fu, 0x4aab7180u, 0xdb8c6debu, 0x12c85ea5u};
const uint32_t G2_Y1[8] = {0xd122975bu, 0x55acdadcu, 0x70b38ef3u, 0xbc4b3133u, 0x690c3395u, 0xec9e99adu, 0x585ff075u, 0x090689d0u};

// 6x + 2 for the BN parameter x = 4965661367192848881 (65 bits; the loop starts below the top bit)
const u64w ATE_LOOP_LOW = 0x9d797039be763ba8ull;   // low 64 bits; bit 64 is set
// (p - 1) / 6
const u64w XI_EXP[4] = {0x34b017592414d4e1ull, 0xee9591c2e6bda1c2ull, 0xf40d60f3c0403964ull, 0x0810b7bdd032f006ull};
// (p^12 - 1) / r
const u64w FINAL_EXP[44] = {
    0x86964b64ca86f120ull, 0x40a4efb7e54523a4ull, 0x837fa97896e84abbull, 0x361102b6b9b2b918ull, 0xc0de81def35692daull, 0xbe04c7e8a6c3c760ull,
    0xd766f9c9d570bb7full, 0xc230974d83561841ull, 0x5bba1668c3be69a3ull, 0x7f3811c410526294ull, 0x29baee7ddadda71cull, 0xbf813b8d145da900ull,
    0x641bbadf423f9a2cull, 0xa80bb4ea44eacc5eull, 0xcd65664814fde37cull, 0x4a0364b9580291d2ull, 0xee93dfb10826f0ddull, 0x6b42db8dc5514724ull,
    0xbb10cf430b0f3785ull, 0x40494e406f804216ull, 0x55cfe107acf3aafbull, 0x2088ec80e0ebae87ull, 0x846a3ed011a337a0ull, 0x48a45a4a1e3a5195ull,
    0xe5664568dfc50e16ull, 0xab6a41294c0cc4ebull, 0x82d0d602d268c7daull, 0x6668449aed3cc48aull, 0x5062cd0fb2015dfcull, 0x7f2940a8b1ddb3d1ull,
    0x77f5b63a2a226448ull, 0xfef0781361e443aeull, 0xf977870e88d5c6c8ull, 0x790364a61f676baaull, 0x5887e72eceaddea3ull, 0x1377e563a09a1b70ull,
    0x0c54efee1bd8c3b2ull, 0x3ec3d15ad524d8f7ull, 0xdaf15466b2383a5dull, 0xe1e30a73bb94fec0ull, 0x6a1c71015f3f7be2ull, 0x842d43bf6369b1ffull,
    0x20fddadf107d20bcull, 0x0000002f4b6dc970ull};

// ---- Fq12 = Fq[w] / (w^12 - 18 w^6 + 82), coefficients low to high ----------------------------------------------------
struct Fq12 { Fq c[12]; };
Fq12 fq12_one() { Fq12 r; for (auto& x : r.c) x = fe_zero<FqP>(); r.c[0] = fe_one<FqP>(); return r; }
bool fq12_is_one(const Fq12& a) {
  if (!fe_eq(a.c[0], fe_one<FqP>())) return false;
  for (int i = 1; i < 12; ++i) if (!fe_is_zero(a.c[i])) return false;
  return true;
}
Fq12 fq12_mul(const Fq12& a, const Fq12& b) {
  static const Fq C18 = fq_small(18), C82 = fq_small(82);
  Fq t[23];
  for (auto& x : t) x = fe_zero<FqP>();
  for (int i = 0; i < 12; ++i) {
    if (fe_is_zero(a.c[i])) continue;
    for (int j = 0; j < 12; ++j) {
      if (fe_is_zero(b.c[j])) continue;
      t[i + j] = fe_add(t[i + j], fe_mul(a.c[i], b.c[j]));
    }
  }
  for (int top = 22; top >= 12; --top) {   // w^top = 18 w^(top-6) - 82 w^(top-12)
    if (fe_is_zero(t[top])) continue;
    t[top - 6] = fe_add(t[top - 6], fe_mul(t[top], C18));
    t[top - 12] = fe_sub(t[top - 12], fe_mul(t[top], C82));
  }
  Fq12 r;
  for (int i = 0; i < 12; ++i) r.c[i] = t[i];
  return r;
}
Fq12 fq12_pow(const Fq12& x, const u64w* e, int nwords) {
  Fq12 acc = fq12_one();
  bool started = false;
  for (int w = nwords - 1; w >= 0; --w)
    for (int b = 63; b >= 0; --b) {
      if (started) acc = fq12_mul(acc, acc);
      if ((e[w] >> b) & 1) { acc = started ? fq12_mul(acc, x) : x; started = true; }
    }
  return acc;
}

// Fq2 element e placed at w^d: e = a + b u with u = w^6 - 9  ->  (a - 9 b) w^d + b w^(d+6)
void fq12_put(Fq12& f, int d, const Fq2& e) {
  static const Fq C9 = fq_small(9);
  f.c[d] = fe_add(f.c[d], fe_sub(e.c0, fe_mul(e.c1, C9)));
  f.c[d + 6] = fe_add(f.c[d + 6], e.c1);
}
Fq2 fq2_scale(const Fq2& a, const Fq& k) { return Fq2{fe_mul(a.c0, k), fe_mul(a.c1, k)}; }
Fq2 fq2_conj(const Fq2& a) { return Fq2{a.c0, fe_neg(a.c1)}; }
Fq2 fq2_pow(const Fq2& x, const u64w* e, int nwords) {
  Fq2 acc = fq2_one();
  for (int w = nwords - 1; w >= 0; --w)
    for (int b = 63; b >= 0; --b) {
      acc = fq2_mul(acc, acc);
      if ((e[w] >> b) & 1) acc = fq2_mul(acc, x);
    }
  return acc;
}

// Line through the twist points t1, t2 (their images x w^2, y w^3 on the curve over Fq12), evaluated at the G1 point P:
//   slope m (over Fq2) != vertical:  l = -y_P + (m x_P) w + (y_1 - m x_1) w^3;   vertical:  l = x_P - x_1 w^2
Fq12 line(const G2Affine& t1, const G2Affine& t2, const G1Affine& p) {
  Fq12 f;
  for (auto& x : f.c) x = fe_zero<FqP>();
  Fq2 m;
  if (!fq2_eq(t1.x, t2.x)) {
    m = fq2_mul(fq2_sub(t2.y, t1.y), fq2_inv(fq2_sub(t2.x, t1.x)));
  } else if (fq2_eq(t1.y, t2.y)) {
    const Fq2 x2 = fq2_mul(t1.x, t1.x);
    m = fq2_mul(fq2_add(fq2_add(x2, x2), x2), fq2_inv(fq2_add(t1.y, t1.y)));
  } else {
    f.c[0] = p.x;
    fq12_put(f, 2, fq2_neg(t1.x));
    return f;
  }
  f.c[0] = fe_neg(p.y);
  fq12_put(f, 1, fq2_scale(m, p.x));
  fq12_put(f, 3, fq2_sub(t1.y, fq2_mul(m, t1.x)));
  return f;
}

Fq12 miller_loop(const G2Affine& q, const G1Affine& p) {
  Fq12 f = fq12_one();
  G2Affine r = q;
  for (int i = 63; i >= 0; --i) {
    f = fq12_mul(fq12_mul(f, f), line(r, r, p));
    r = g2_add(r, r);
    if ((ATE_LOOP_LOW >> i) & 1) {
      f = fq12_mul(f, line(r, q, p));
      r = g2_add(r, q);
    }
  }
  // Frobenius images of Q on the twist: (conj(x) g^2, conj(y) g^3) with g = (9 + u)^((p - 1) / 6)
  static const Fq2 G1C = fq2_pow(Fq2{fq_small(9), fe_one<FqP>()}, XI_EXP, 4);
  static const Fq2 G2C = fq2_mul(G1C, G1C), G3C = fq2_mul(G2C, G1C);
  const G2Affine q1{fq2_mul(fq2_conj(q.x), G2C), fq2_mul(fq2_conj(q.y), G3C)};
  const G2Affine nq2{fq2_mul(fq2_conj(q1.x), G2C), fq2_neg(fq2_mul(fq2_conj(q1.y), G3C))};
  f = fq12_mul(f, line(r, q1, p));
  r = g2_add(r, q1);
  f = fq12_mul(f, line(r, nq2, p));
  return f;
}

}  // namespace

Fq fq_pow_words(const Fq& x, const unsigned long long* e, int nwords) {
  Fq acc = fe_one<FqP>();
  for (int w = nwords - 1; w >= 0; --w)
    for (int b = 63; b >= 0; --b) {
      acc = fe_sqr(acc);
      if ((e[w] >> b) & 1) acc = fe_mul(acc, x);
    }
  return acc;
}

Fq2 fq2_zero() { return Fq2{fe_zero<FqP>(), fe_zero<FqP>()}; }
Fq2 fq2_one() { return Fq2{fe_one<FqP>(), fe_zero<FqP>()}; }
bool fq2_is_zero(const Fq2& a) { return fe_is_zero(a.c0) && fe_is_zero(a.c1); }
bool fq2_eq(const Fq2& a, const Fq2& b) { return fe_eq(a.c0, b.c0) && fe_eq(a.c1, b.c1); }
Fq2 fq2_add(const Fq2& a, const Fq2& b) { return Fq2{fe_add(a.c0, b.c0), fe_add(a.c1, b.c1)}; }
Fq2 fq2_sub(const Fq2& a, const Fq2& b) { return Fq2{fe_sub(a.c0, b.c0), fe_sub(a.c1, b.c1)}; }
Fq2 fq2_neg(const Fq2& a) { return Fq2{fe_neg(a.c0), fe_neg(a.c1)}; }
Fq2 fq2_mul(const Fq2& a, const Fq2& b) {
  const Fq t0 = fe_mul(a.c0, b.c0), t1 = fe_mul(a.c1, b.c1);
  return Fq2{fe_sub(t0, t1), fe_sub(fe_sub(fe_mul(fe_add(a.c0, a.c1), fe_add(b.c0, b.c1)), t0), t1)};
}
Fq2 fq2_inv(const Fq2& a) {
  const Fq ninv = fe_inv(fe_add(fe_sqr(a.c0), fe_sqr(a.c1)));
  return Fq2{fe_mul(a.c0, ninv), fe_neg(fe_mul(a.c1, ninv))};
}

bool g2_is_identity(const G2Affine& p) { return fq2_is_zero(p.x) && fq2_is_zero(p.y); }
bool g2_on_curve(const G2Affine& p) {
  if (g2_is_identity(p)) return true;
  static const Fq2 B2 = fq2_mul(Fq2{fq_small(3), fe_zero<FqP>()}, fq2_inv(Fq2{fq_small(9), fe_one<FqP>()}));
  return fq2_eq(fq2_mul(p.y, p.y), fq2_add(fq2_mul(fq2_mul(p.x, p.x), p.x), B2));
}
G2Affine g2_generator() { return G2Affine{Fq2{fq_from_words(G2_X0), fq_from_words(G2_X1)}, Fq2{fq_from_words(G2_Y0), fq_from_words(G2_Y1)}}; }
G2Affine g2_add(const G2Affine& p, const G2Affine& q) {
  if (g2_is_identity(p)) return q;
  if (g2_is_identity(q)) return p;
  Fq2 m;
  if (fq2_eq(p.x, q.x)) {
    if (!fq2_eq(p.y, q.y) || fq2_is_zero(p.y)) return G2Affine{fq2_zero(), fq2_zero()};
    const Fq2 x2 = fq2_mul(p.x, p.x);
    m = fq2_mul(fq2_add(fq2_add(x2, x2), x2), fq2_inv(fq2_add(p.y, p.y)));
  } else {
    m = fq2_mul(fq2_sub(q.y, p.y), fq2_inv(fq2_sub(q.x, p.x)));
  }
  const Fq2 x3 = fq2_sub(fq2_sub(fq2_mul(m, m), p.x), q.x);
  return G2Affine{x3, fq2_sub(fq2_mul(m, fq2_sub(p.x, x3)), p.y)};
}
G2Affine g2_mul(const G2Affine& p, const Fr& scalar) {
  const Fr k = fe_to_canonical(scalar);
  G2Affine acc{fq2_zero(), fq2_zero()};
  for (int i = 255; i >= 0; --i) {
    acc = g2_add(acc, acc);
    if ((k.v[i >> 5] >> (i & 31)) & 1) acc = g2_add(acc, p);
  }
  return acc;
}

bool pairing_product_is_one(const std::vector<std::pair<G1Affine, G2Affine>>& pairs) {
  Fq12 f = fq12_one();
  for (const auto& pr : pairs) {
    if (affine_is_identity(pr.first) || g2_is_identity(pr.second)) continue;
    f = fq12_mul(f, miller_loop(pr.second, pr.first));
  }
  return fq12_is_one(fq12_pow(f, FINAL_EXP, 44));
}

}  // namespace host
}  // namespace zkc
