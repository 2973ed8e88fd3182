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
// ---- cooperative addition: FOUR lanes of a warp (an aligned quad) share one XYZZ addition ----------------------------------
// The back end of an MSM (bucket sums, the bit-plane tree) is a chain of dependent additions on a machine that is mostly idle:
// what counts there is the latency of ONE addition (14 Montgomery products in a row for a lone thread), not throughput.  The
// products of add-2008-s fall into four independent groups { u1, u2, s1, s2 } -> { pp, r^2, zz1 zz2, zzz1 zzz2 } ->
// { ppp, qq, zz } -> { r (qq - x3), s1 ppp, zzz }; a quad computes one group per step, one product per lane, and exchanges the
// results with shuffles: 4 product latencies instead of 14.  All four lanes hold the same p and q on entry and the same sum
// on exit.
ZKC_D Fq fq_quad_bcast(const Fq& v, int src, unsigned qmask) {
  Fq r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = __shfl_sync(qmask, v.v[i], src, 4);
  return r;
}
// a_role, selected with bit masks: a chain of ?: on the role compiles to a divergent branch region PER WORD, which serialises
// the four roles and costs more than the products it feeds
ZKC_D Fq fq_sel4(int role, const Fq& a0, const Fq& a1, const Fq& a2, const Fq& a3) {
  const uint32_t m0 = role == 0 ? 0xffffffffu : 0u, m1 = role == 1 ? 0xffffffffu : 0u, m2 = role == 2 ? 0xffffffffu : 0u,
                 m3 = role == 3 ? 0xffffffffu : 0u;
  Fq r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = (a0.v[i] & m0) | (a1.v[i] & m1) | (a2.v[i] & m2) | (a3.v[i] & m3);
  return r;
}
// Every lane of the WARP must call this together (the shuffles name the full warp: a shuffle with a run-time sub-mask costs a
// WARPSYNC each, which is what made a quad-masked variant no faster than a lone thread); quads with nothing to add pass the
// identity for q.  The exceptional cases are resolved by selection after the common path, so the warp never diverges around
// a shuffle.
ZKC_D void xyzz_add_quad(G1Xyzz& p, const G1Xyzz& q) {
  const unsigned FULL = 0xffffffffu;
  const int role = threadIdx.x & 3;
  const bool q_id = xyzz_is_identity(q), p_id = xyzz_is_identity(p);
  Fq m = fe_mul(fq_sel4(role, p.x, q.x, p.y, q.y), fq_sel4(role, q.zz, p.zz, q.zzz, p.zzz));
  const Fq u1 = fq_quad_bcast(m, 0, FULL), u2 = fq_quad_bcast(m, 1, FULL), s1 = fq_quad_bcast(m, 2, FULL), s2 = fq_quad_bcast(m, 3, FULL);
  const Fq pp_ = fe_sub(u2, u1), r = fe_sub(s2, s1);
  m = fe_mul(fq_sel4(role, pp_, r, p.zz, p.zzz), fq_sel4(role, pp_, r, q.zz, q.zzz));
  const Fq pp = fq_quad_bcast(m, 0, FULL), rr = fq_quad_bcast(m, 1, FULL), zz12 = fq_quad_bcast(m, 2, FULL), zzz12 = fq_quad_bcast(m, 3, FULL);
  m = fe_mul(fq_sel4(role, pp_, u1, zz12, zz12), pp);
  const Fq ppp = fq_quad_bcast(m, 0, FULL), qq = fq_quad_bcast(m, 1, FULL), zz3 = fq_quad_bcast(m, 2, FULL);
  const Fq x3 = fe_sub(fe_sub(rr, ppp), fe_dbl(qq));
  m = fe_mul(fq_sel4(role, r, s1, zzz12, zzz12), fq_sel4(role, fe_sub(qq, x3), ppp, ppp, ppp));
  const Fq y1 = fq_quad_bcast(m, 0, FULL), y2 = fq_quad_bcast(m, 1, FULL), zzz3 = fq_quad_bcast(m, 2, FULL);
  if (q_id) return;                       // p + 0
  if (p_id) { p = q; return; }            // 0 + q
  if (fe_is_zero(pp_)) {                  // same x: P + P or P - P (no shuffles below: lanes may diverge freely)
    if (fe_is_zero(r)) p = xyzz_dbl(p);
    else p = xyzz_identity();
    return;
  }
  p.x = x3; p.y = fe_sub(y1, y2); p.zz = zz3; p.zzz = zzz3;
}

ZKC_D G1Affine affine_load_nc(const G1Affine* p) { G1Affine r; r.x = fe_load_nc(&p->x); r.y = fe_load_nc(&p->y); return r; }
ZKC_D G1Xyzz xyzz_load(const G1Xyzz* p) { G1Xyzz r; r.x = fe_load(&p->x); r.y = fe_load(&p->y); r.zz = fe_load(&p->zz); r.zzz = fe_load(&p->zzz); return r; }
ZKC_D void xyzz_store(G1Xyzz* p, const G1Xyzz& r) { fe_store(&p->x, r.x); fe_store(&p->y, r.y); fe_store(&p->zz, r.zz); fe_store(&p->zzz, r.zzz); }
#endif

}  // namespace zkc
