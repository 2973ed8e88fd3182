// BN254 G1 (y^2 = x^3 + 3 over Fq) point arithmetic for sm_100a.
//
// Replaces halo2curves::bn256::{G1Affine, G1} group operations used by best_multiexp
// (halo2curves 0.4.0 @ e185711, /root/reference/Cargo.lock:1359-1380).
// Interface types keep the reference's memory layout (affine (x, y) with identity = (0, 0);
// Jacobian (x, y, z)); internally buckets use XYZZ coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2),
// whose mixed addition costs 8M + 2S with no inversion.  Group elements are unique, so any
// coordinate system yields the reference's bytes after normalisation.
#pragma once
#include "ff.cuh"

namespace zkc {

struct alignas(16) G1Affine { Fq x, y; };
struct alignas(16) G1Xyzz { Fq x, y, zz, zzz; };
struct alignas(16) G1Jac { Fq x, y, z; };

ZKC_HD bool affine_is_identity(const G1Affine& p) { return fe_is_zero(p.x) && fe_is_zero(p.y); }
ZKC_HD bool xyzz_is_identity(const G1Xyzz& p) { return fe_is_zero(p.zz); }
ZKC_HD G1Xyzz xyzz_identity() { G1Xyzz r; r.x = fe_zero<FqP>(); r.y = fe_zero<FqP>(); r.zz = fe_zero<FqP>(); r.zzz = fe_zero<FqP>(); return r; }
ZKC_HD G1Xyzz xyzz_from_affine(const G1Affine& p) {
  if (affine_is_identity(p)) return xyzz_identity();
  G1Xyzz r; r.x = p.x; r.y = p.y; r.zz = fe_one<FqP>(); r.zzz = fe_one<FqP>(); return r;
}
ZKC_HD G1Xyzz xyzz_neg(const G1Xyzz& p) { G1Xyzz r = p; r.y = fe_neg(p.y); return r; }

// 2 * (affine q), q != identity.   mdbl-2008-s-1
ZKC_HD G1Xyzz xyzz_dbl_affine(const G1Affine& q) {
  G1Xyzz r;
  if (fe_is_zero(q.y)) return xyzz_identity();  // order-2 point (none on BN254 G1, kept for safety)
  Fq u = fe_dbl(q.y);
  Fq v = fe_sqr(u);
  Fq w = fe_mul(u, v);
  Fq s = fe_mul(q.x, v);
  Fq xx = fe_sqr(q.x);
  Fq m = fe_add(fe_dbl(xx), xx);
  r.x = fe_sub(fe_sqr(m), fe_dbl(s));
  r.y = fe_sub(fe_mul(m, fe_sub(s, r.x)), fe_mul(w, q.y));
  r.zz = v;
  r.zzz = w;
  return r;
}

// 2 * p.   dbl-2008-s-1 (a = 0)
ZKC_HD G1Xyzz xyzz_dbl(const G1Xyzz& p) {
  if (xyzz_is_identity(p)) return p;
  G1Xyzz r;
  Fq u = fe_dbl(p.y);
  Fq v = fe_sqr(u);
  Fq w = fe_mul(u, v);
  Fq s = fe_mul(p.x, v);
  Fq xx = fe_sqr(p.x);
  Fq m = fe_add(fe_dbl(xx), xx);
  r.x = fe_sub(fe_sqr(m), fe_dbl(s));
  r.y = fe_sub(fe_mul(m, fe_sub(s, r.x)), fe_mul(w, p.y));
  r.zz = fe_mul(v, p.zz);
  r.zzz = fe_mul(w, p.zzz);
  return r;
}

// p += q (affine, q != identity; `neg` adds -q).   madd-2008-s, with the exceptional cases.
ZKC_HD void xyzz_madd(G1Xyzz& p, const G1Affine& q, bool neg) {
  Fq qy = neg ? fe_neg(q.y) : q.y;
  if (xyzz_is_identity(p)) { p.x = q.x; p.y = qy; p.zz = fe_one<FqP>(); p.zzz = fe_one<FqP>(); return; }
  Fq u2 = fe_mul(q.x, p.zz);
  Fq s2 = fe_mul(qy, p.zzz);
  Fq pp_ = fe_sub(u2, p.x);
  Fq r = fe_sub(s2, p.y);
  if (fe_is_zero(pp_)) {
    if (fe_is_zero(r)) { G1Affine t; t.x = q.x; t.y = qy; p = xyzz_dbl_affine(t); }
    else p = xyzz_identity();
    return;
  }
  Fq pp = fe_sqr(pp_);
  Fq ppp = fe_mul(pp_, pp);
  Fq qq = fe_mul(p.x, pp);
  Fq x3 = fe_sub(fe_sub(fe_sqr(r), ppp), fe_dbl(qq));
  p.y = fe_sub(fe_mul(r, fe_sub(qq, x3)), fe_mul(p.y, ppp));
  p.x = x3;
  p.zz = fe_mul(p.zz, pp);
  p.zzz = fe_mul(p.zzz, ppp);
}

// p += q (both XYZZ).   add-2008-s, with the exceptional cases.
ZKC_HD void xyzz_add(G1Xyzz& p, const G1Xyzz& q) {
  if (xyzz_is_identity(q)) return;
  if (xyzz_is_identity(p)) { p = q; return; }
  Fq u1 = fe_mul(p.x, q.zz);
  Fq u2 = fe_mul(q.x, p.zz);
  Fq s1 = fe_mul(p.y, q.zzz);
  Fq s2 = fe_mul(q.y, p.zzz);
  Fq pp_ = fe_sub(u2, u1);
  Fq r = fe_sub(s2, s1);
  if (fe_is_zero(pp_)) {
    if (fe_is_zero(r)) p = xyzz_dbl(p);
    else p = xyzz_identity();
    return;
  }
  Fq pp = fe_sqr(pp_);
  Fq ppp = fe_mul(pp_, pp);
  Fq qq = fe_mul(u1, pp);
  Fq x3 = fe_sub(fe_sub(fe_sqr(r), ppp), fe_dbl(qq));
  p.y = fe_sub(fe_mul(r, fe_sub(qq, x3)), fe_mul(s1, ppp));
  p.x = x3;
  p.zz = fe_mul(fe_mul(p.zz, q.zz), pp);
  p.zzz = fe_mul(fe_mul(p.zzz, q.zzz), ppp);
}

// XYZZ -> affine (one inversion): x = X / ZZ, y = Y / ZZZ
ZKC_HD G1Affine xyzz_to_affine(const G1Xyzz& p) {
  G1Affine a;
  if (xyzz_is_identity(p)) { a.x = fe_zero<FqP>(); a.y = fe_zero<FqP>(); return a; }
  // 1/ZZZ, then 1/ZZ = ZZZ^-1 * ZZZ / ZZ ... simpler: invert the product and split
  Fq prod = fe_mul(p.zz, p.zzz);
  Fq inv = fe_inv(prod);
  Fq izz = fe_mul(inv, p.zzz);
  Fq izzz = fe_mul(inv, p.zz);
  a.x = fe_mul(p.x, izz);
  a.y = fe_mul(p.y, izzz);
  return a;
}

#if defined(__CUDACC__)
ZKC_D G1Affine affine_load_nc(const G1Affine* p) { G1Affine r; r.x = fe_load_nc(&p->x); r.y = fe_load_nc(&p->y); return r; }
ZKC_D G1Xyzz xyzz_load(const G1Xyzz* p) { G1Xyzz r; r.x = fe_load(&p->x); r.y = fe_load(&p->y); r.zz = fe_load(&p->zz); r.zzz = fe_load(&p->zzz); return r; }
ZKC_D void xyzz_store(G1Xyzz* p, const G1Xyzz& r) { fe_store(&p->x, r.x); fe_store(&p->y, r.y); fe_store(&p->zz, r.zz); fe_store(&p->zzz, r.zzz); }
#endif

}  // namespace zkc
