// BN254 Fr / Fq Montgomery arithmetic for sm_100a, 8 x 32-bit limbs in registers.
//
// Replaces halo2curves::bn256::{Fr,Fq} (halo2curves 0.4.0 @ e185711, pinned at
// /root/reference/Cargo.lock:1359-1380).  Bytes are identical to the
// reference's `[u64; 4]` little-endian Montgomery limbs (R = 2^256), so host buffers pass through
// the C ABI with zero conversion.  All results are canonical (< modulus).
//
// Multiplication: row-interleaved Montgomery (CIOS) on *64-bit column accumulators*.  Products of
// even limbs of `a` land on even 64-bit columns (E), odd limbs on odd columns (O), so each
// 32x32->64 product + 64-bit accumulate + carry is ONE `IMAD.WIDE.U32(.X)` (PTX: mul.wide.u32 +
// add.cc.u64 / addc.cc.u64, which ptxas fuses).  After each row the low limb is cancelled with
// m = t0 * (-p^-1) and the state shifts one limb by swapping the roles of E and O; the limb that
// falls between the two alignments is carried as a 32-bit `pend` word whose only effect is a
// carry-in bit into the next odd chain.  16 IMAD.WIDE + ~5 other instructions per row.
#pragma once
#include <cstring>
#include <cstdint>

#if defined(__CUDACC__)
#define ZKC_HD __host__ __device__ __forceinline__
#define ZKC_D __device__ __forceinline__
#else
#define ZKC_HD inline
#define ZKC_D inline
#endif

namespace zkc {

typedef unsigned long long u64;

struct FrP {
  static constexpr uint32_t INV = 0xefffffffu;
  ZKC_HD static constexpr uint32_t M(int i) {
    return i == 0 ? 0xf0000001u : i == 1 ? 0x43e1f593u : i == 2 ? 0x79b97091u : i == 3 ? 0x2833e848u
         : i == 4 ? 0x8181585du : i == 5 ? 0xb85045b6u : i == 6 ? 0xe131a029u : 0x30644e72u;
  }
  ZKC_HD static constexpr uint32_t ONE(int i) {  // R mod r
    return i == 0 ? 0x4ffffffbu : i == 1 ? 0xac96341cu : i == 2 ? 0x9f60cd29u : i == 3 ? 0x36fc7695u
         : i == 4 ? 0x7879462eu : i == 5 ? 0x666ea36fu : i == 6 ? 0x9a07df2fu : 0x0e0a77c1u;
  }
  ZKC_HD static constexpr uint32_t R2(int i) {  // R^2 mod r
    return i == 0 ? 0xae216da7u : i == 1 ? 0x1bb8e645u : i == 2 ? 0xe35c59e3u : i == 3 ? 0x53fe3ab1u
         : i == 4 ? 0x53bb8085u : i == 5 ? 0x8c49833du : i == 6 ? 0x7f4e44a5u : 0x0216d0b1u;
  }
};
struct FqP {
  static constexpr uint32_t INV = 0xe4866389u;
  ZKC_HD static constexpr uint32_t M(int i) {
    return i == 0 ? 0xd87cfd47u : i == 1 ? 0x3c208c16u : i == 2 ? 0x6871ca8du : i == 3 ? 0x97816a91u
         : i == 4 ? 0x8181585du : i == 5 ? 0xb85045b6u : i == 6 ? 0xe131a029u : 0x30644e72u;
  }
  ZKC_HD static constexpr uint32_t ONE(int i) {  // R mod p
    return i == 0 ? 0xc58f0d9du : i == 1 ? 0xd35d438du : i == 2 ? 0xf5c70b3du : i == 3 ? 0x0a78eb28u
         : i == 4 ? 0x7879462cu : i == 5 ? 0x666ea36fu : i == 6 ? 0x9a07df2fu : 0x0e0a77c1u;
  }
  ZKC_HD static constexpr uint32_t R2(int i) {  // R^2 mod p
    return i == 0 ? 0x538afa89u : i == 1 ? 0xf32cfc5bu : i == 2 ? 0xd44501fbu : i == 3 ? 0xb5e71911u
         : i == 4 ? 0x0a417ff6u : i == 5 ? 0x47ab1effu : i == 6 ? 0xcab8351fu : 0x06d89f71u;
  }
};

template <class P>
struct alignas(16) Fe {
  uint32_t v[8];
};
typedef Fe<FrP> Fr;
typedef Fe<FqP> Fq;

template <class P> ZKC_HD Fe<P> fe_zero() { Fe<P> r; for (int i = 0; i < 8; ++i) r.v[i] = 0; return r; }
template <class P> ZKC_HD Fe<P> fe_one() { Fe<P> r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = P::ONE(i); return r; }
template <class P> ZKC_HD Fe<P> fe_r2() { Fe<P> r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = P::R2(i); return r; }
template <class P> ZKC_HD bool fe_is_zero(const Fe<P>& a) {
  return (a.v[0] | a.v[1] | a.v[2] | a.v[3] | a.v[4] | a.v[5] | a.v[6] | a.v[7]) == 0;
}
template <class P> ZKC_HD bool fe_eq(const Fe<P>& a, const Fe<P>& b) {
  uint32_t d = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) d |= a.v[i] ^ b.v[i];
  return d == 0;
}

// ------------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)

// r = a - M with borrow-out; returns borrow (1 if a < M)
template <class P> ZKC_D uint32_t sub_mod_raw(uint32_t* r, const uint32_t* a) {
  uint32_t bw;
  asm("sub.cc.u32 %0,%9,%17;\n\t"
      "subc.cc.u32 %1,%10,%18;\n\t"
      "subc.cc.u32 %2,%11,%19;\n\t"
      "subc.cc.u32 %3,%12,%20;\n\t"
      "subc.cc.u32 %4,%13,%21;\n\t"
      "subc.cc.u32 %5,%14,%22;\n\t"
      "subc.cc.u32 %6,%15,%23;\n\t"
      "subc.cc.u32 %7,%16,%24;\n\t"
      "subc.u32 %8,0,0;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(bw)
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "r"(P::M(0)), "r"(P::M(1)), "r"(P::M(2)), "r"(P::M(3)), "r"(P::M(4)), "r"(P::M(5)), "r"(P::M(6)), "r"(P::M(7)));
  return bw;  // 0 or 0xffffffff
}

// final reduction of a value in [0, 2M)
template <class P> ZKC_D void reduce_once(uint32_t* t) {
  uint32_t s[8];
  uint32_t bw = sub_mod_raw<P>(s, t);
#pragma unroll
  for (int i = 0; i < 8; ++i) t[i] = bw ? t[i] : s[i];
}

template <class P> ZKC_D Fe<P> fe_add(const Fe<P>& a, const Fe<P>& b) {
  Fe<P> r;
  asm("add.cc.u32 %0,%8,%16;\n\t"
      "addc.cc.u32 %1,%9,%17;\n\t"
      "addc.cc.u32 %2,%10,%18;\n\t"
      "addc.cc.u32 %3,%11,%19;\n\t"
      "addc.cc.u32 %4,%12,%20;\n\t"
      "addc.cc.u32 %5,%13,%21;\n\t"
      "addc.cc.u32 %6,%14,%22;\n\t"
      "addc.u32 %7,%15,%23;"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
        "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
  reduce_once<P>(r.v);  // a + b < 2M < 2^255: no carry out of limb 7
  return r;
}

template <class P> ZKC_D Fe<P> fe_sub(const Fe<P>& a, const Fe<P>& b) {
  Fe<P> r;
  uint32_t bw;
  asm("sub.cc.u32 %0,%9,%17;\n\t"
      "subc.cc.u32 %1,%10,%18;\n\t"
      "subc.cc.u32 %2,%11,%19;\n\t"
      "subc.cc.u32 %3,%12,%20;\n\t"
      "subc.cc.u32 %4,%13,%21;\n\t"
      "subc.cc.u32 %5,%14,%22;\n\t"
      "subc.cc.u32 %6,%15,%23;\n\t"
      "subc.cc.u32 %7,%16,%24;\n\t"
      "subc.u32 %8,0,0;"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]), "=r"(bw)
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
        "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
  // add back M masked by the borrow
  asm("add.cc.u32 %0,%0,%8;\n\t"
      "addc.cc.u32 %1,%1,%9;\n\t"
      "addc.cc.u32 %2,%2,%10;\n\t"
      "addc.cc.u32 %3,%3,%11;\n\t"
      "addc.cc.u32 %4,%4,%12;\n\t"
      "addc.cc.u32 %5,%5,%13;\n\t"
      "addc.cc.u32 %6,%6,%14;\n\t"
      "addc.u32 %7,%7,%15;"
      : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]), "+r"(r.v[7])
      : "r"(P::M(0) & bw), "r"(P::M(1) & bw), "r"(P::M(2) & bw), "r"(P::M(3) & bw), "r"(P::M(4) & bw), "r"(P::M(5) & bw),
        "r"(P::M(6) & bw), "r"(P::M(7) & bw));
  return r;
}

// acc(4 x u64) += x[0..3] * m, carry-out accumulated into the 32-bit `top`
#define ZKC_CHAIN4_CO(e0, e1, e2, e3, top, x0, x1, x2, x3, m)                               \
  asm("{\n\t.reg .u64 t;\n\t"                                                                \
      "mul.wide.u32 t,%5,%9;\n\tadd.cc.u64 %0,%0,t;\n\t"                                     \
      "mul.wide.u32 t,%6,%9;\n\taddc.cc.u64 %1,%1,t;\n\t"                                    \
      "mul.wide.u32 t,%7,%9;\n\taddc.cc.u64 %2,%2,t;\n\t"                                    \
      "mul.wide.u32 t,%8,%9;\n\taddc.cc.u64 %3,%3,t;\n\t"                                    \
      "addc.u32 %4,%4,0;\n\t}"                                                               \
      : "+l"(e0), "+l"(e1), "+l"(e2), "+l"(e3), "+r"(top)                                    \
      : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(m))
// acc(4 x u64) += x[0..3] * m   (provably no carry-out: see header comment / DESIGN.md)
#define ZKC_CHAIN4(o0, o1, o2, o3, x0, x1, x2, x3, m)                                        \
  asm("{\n\t.reg .u64 t;\n\t"                                                                \
      "mul.wide.u32 t,%4,%8;\n\tadd.cc.u64 %0,%0,t;\n\t"                                     \
      "mul.wide.u32 t,%5,%8;\n\taddc.cc.u64 %1,%1,t;\n\t"                                    \
      "mul.wide.u32 t,%6,%8;\n\taddc.cc.u64 %2,%2,t;\n\t"                                    \
      "mul.wide.u32 t,%7,%8;\n\taddc.u64 %3,%3,t;\n\t}"                                      \
      : "+l"(o0), "+l"(o1), "+l"(o2), "+l"(o3)                                               \
      : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(m))
// same, with carry-in = (pend != 0)
#define ZKC_CHAIN4_CI(o0, o1, o2, o3, x0, x1, x2, x3, m, pend)                               \
  asm("{\n\t.reg .u64 t;\n\t.reg .u32 d;\n\t"                                                \
      "add.cc.u32 d,%9,0xffffffff;\n\t"                                                      \
      "mul.wide.u32 t,%4,%8;\n\taddc.cc.u64 %0,%0,t;\n\t"                                    \
      "mul.wide.u32 t,%5,%8;\n\taddc.cc.u64 %1,%1,t;\n\t"                                    \
      "mul.wide.u32 t,%6,%8;\n\taddc.cc.u64 %2,%2,t;\n\t"                                    \
      "mul.wide.u32 t,%7,%8;\n\taddc.u64 %3,%3,t;\n\t}"                                      \
      : "+l"(o0), "+l"(o1), "+l"(o2), "+l"(o3)                                               \
      : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(m), "r"(pend))

template <class P> ZKC_D Fe<P> fe_mul(const Fe<P>& a, const Fe<P>& b) {
  u64 e0 = 0, e1 = 0, e2 = 0, e3 = 0, o0 = 0, o1 = 0, o2 = 0, o3 = 0;
  uint32_t pend = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t bi = b.v[i];
    uint32_t top = 0;
    ZKC_CHAIN4_CO(e0, e1, e2, e3, top, a.v[0], a.v[2], a.v[4], a.v[6], bi);
    ZKC_CHAIN4(o0, o1, o2, o3, a.v[1], a.v[3], a.v[5], a.v[7], bi);
    const uint32_t m = ((uint32_t)e0 + pend) * P::INV;
    ZKC_CHAIN4_CO(e0, e1, e2, e3, top, P::M(0), P::M(2), P::M(4), P::M(6), m);
    ZKC_CHAIN4_CI(o0, o1, o2, o3, P::M(1), P::M(3), P::M(5), P::M(7), m, pend);
    // shift one limb: E <- O, O <- E >> 64, pend <- hi32(E0)
    const uint32_t np = (uint32_t)(e0 >> 32);
    const u64 n0 = o0, n1 = o1, n2 = o2, n3 = o3;
    o0 = e1; o1 = e2; o2 = e3; o3 = (u64)top;
    e0 = n0; e1 = n1; e2 = n2; e3 = n3;
    pend = np;
  }
  Fe<P> r;
  asm("add.cc.u32 %0,%8,%16;\n\t"
      "addc.cc.u32 %1,%9,%17;\n\t"
      "addc.cc.u32 %2,%10,%18;\n\t"
      "addc.cc.u32 %3,%11,%19;\n\t"
      "addc.cc.u32 %4,%12,%20;\n\t"
      "addc.cc.u32 %5,%13,%21;\n\t"
      "addc.cc.u32 %6,%14,%22;\n\t"
      "addc.u32 %7,%15,%23;"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
      : "r"((uint32_t)e0), "r"((uint32_t)(e0 >> 32)), "r"((uint32_t)e1), "r"((uint32_t)(e1 >> 32)), "r"((uint32_t)e2),
        "r"((uint32_t)(e2 >> 32)), "r"((uint32_t)e3), "r"((uint32_t)(e3 >> 32)),
        "r"(pend), "r"((uint32_t)o0), "r"((uint32_t)(o0 >> 32)), "r"((uint32_t)o1), "r"((uint32_t)(o1 >> 32)),
        "r"((uint32_t)o2), "r"((uint32_t)(o2 >> 32)), "r"((uint32_t)o3));
  reduce_once<P>(r.v);
  return r;
}


// ---- dedicated squaring: 100 instead of 128 wide multiplies --------------------------------------------------------------------
// a^2 = 2 * sum_{i<j} a_i a_j B^(i+j) + sum_i a_i^2 B^(2i)  (B = 2^32): the 28 cross products are accumulated once on the same
// even / odd aligned 64-bit columns as fe_mul (a_i a_j lands on E[(i+j)/2] or O[(i+j-1)/2]), merged and doubled with plain
// adds, the 8 squares sit on the even columns without overlapping; the low half of the 512-bit square is then cancelled by
// the same eight rows of m * p as in fe_mul (64 multiplies) and the high half is added at the end.  Where a chain's carry can
// reach one accumulator further than its last product a carry-only step follows; the chain extents are those of a worst-case
// (all-ones limbs) simulation in which no carry is lost (the partial sums are monotone in the limbs).
#define ZKC_SQ_MAC1(a0, x0, m) asm("mad.wide.u32 %0,%1,%2,%0;" : "+l"(a0) : "r"(x0), "r"(m))
#define ZKC_SQ_MAC1C(a0, c1, x0, m)                                                         \
  asm("{\n\t.reg .u64 t;\n\t"                                                                \
      "mul.wide.u32 t,%2,%3;\n\tadd.cc.u64 %0,%0,t;\n\taddc.u64 %1,%1,0;\n\t}"               \
      : "+l"(a0), "+l"(c1) : "r"(x0), "r"(m))
#define ZKC_SQ_MAC2(a0, a1, x0, x1, m)                                                      \
  asm("{\n\t.reg .u64 t;\n\t"                                                                \
      "mul.wide.u32 t,%2,%4;\n\tadd.cc.u64 %0,%0,t;\n\t"                                     \
      "mul.wide.u32 t,%3,%4;\n\taddc.u64 %1,%1,t;\n\t}"                                      \
      : "+l"(a0), "+l"(a1) : "r"(x0), "r"(x1), "r"(m))
#define ZKC_SQ_MAC2C(a0, a1, c2, x0, x1, m)                                                 \
  asm("{\n\t.reg .u64 t;\n\t"                                                                \
      "mul.wide.u32 t,%3,%5;\n\tadd.cc.u64 %0,%0,t;\n\t"                                     \
      "mul.wide.u32 t,%4,%5;\n\taddc.cc.u64 %1,%1,t;\n\taddc.u64 %2,%2,0;\n\t}"              \
      : "+l"(a0), "+l"(a1), "+l"(c2) : "r"(x0), "r"(x1), "r"(m))
#define ZKC_SQ_MAC3(a0, a1, a2, x0, x1, x2, m)                                              \
  asm("{\n\t.reg .u64 t;\n\t"                                                                \
      "mul.wide.u32 t,%3,%6;\n\tadd.cc.u64 %0,%0,t;\n\t"                                     \
      "mul.wide.u32 t,%4,%6;\n\taddc.cc.u64 %1,%1,t;\n\t"                                    \
      "mul.wide.u32 t,%5,%6;\n\taddc.u64 %2,%2,t;\n\t}"                                      \
      : "+l"(a0), "+l"(a1), "+l"(a2) : "r"(x0), "r"(x1), "r"(x2), "r"(m))
#define ZKC_SQ_MAC3C(a0, a1, a2, c3, x0, x1, x2, m)                                         \
  asm("{\n\t.reg .u64 t;\n\t"                                                                \
      "mul.wide.u32 t,%4,%7;\n\tadd.cc.u64 %0,%0,t;\n\t"                                     \
      "mul.wide.u32 t,%5,%7;\n\taddc.cc.u64 %1,%1,t;\n\t"                                    \
      "mul.wide.u32 t,%6,%7;\n\taddc.cc.u64 %2,%2,t;\n\taddc.u64 %3,%3,0;\n\t}"              \
      : "+l"(a0), "+l"(a1), "+l"(a2), "+l"(c3) : "r"(x0), "r"(x1), "r"(x2), "r"(m))

template <class P> ZKC_D Fe<P> fe_sqr(const Fe<P>& a) {
  const uint32_t a0 = a.v[0], a1 = a.v[1], a2 = a.v[2], a3 = a.v[3], a4 = a.v[4], a5 = a.v[5], a6 = a.v[6], a7 = a.v[7];
  // cross products, row i = a_i * (a_j, j > i): even columns E1..E6, odd columns O0..O6
  u64 E1 = (u64)a0 * a2, E2 = (u64)a0 * a4, E3 = (u64)a0 * a6, E4 = 0, E5 = 0, E6 = 0;
  u64 O0 = (u64)a0 * a1, O1 = (u64)a0 * a3, O2 = (u64)a0 * a5, O3 = (u64)a0 * a7, O4 = 0, O5 = 0, O6 = 0;
  ZKC_SQ_MAC3(E2, E3, E4, a3, a5, a7, a1);       ZKC_SQ_MAC3C(O1, O2, O3, O4, a2, a4, a6, a1);
  ZKC_SQ_MAC2C(E3, E4, E5, a4, a6, a2);          ZKC_SQ_MAC3(O2, O3, O4, a3, a5, a7, a2);
  ZKC_SQ_MAC2(E4, E5, a5, a7, a3);               ZKC_SQ_MAC2C(O3, O4, O5, a4, a6, a3);
  ZKC_SQ_MAC1C(E5, E6, a6, a4);                  ZKC_SQ_MAC2(O4, O5, a5, a7, a4);
  ZKC_SQ_MAC1(E6, a7, a5);                       ZKC_SQ_MAC1C(O5, O6, a6, a5);
  ZKC_SQ_MAC1(O6, a7, a6);
  // c = E + (O << 32) as 16 limbs (c0 = 0): c[k] = E-limb k + O-limb k
  uint32_t c[16];
  c[0] = 0; c[1] = (uint32_t)O0;
  asm("add.cc.u32 %0,%14,%15;\n\t"
      "addc.cc.u32 %1,%16,%17;\n\t"
      "addc.cc.u32 %2,%18,%19;\n\t"
      "addc.cc.u32 %3,%20,%21;\n\t"
      "addc.cc.u32 %4,%22,%23;\n\t"
      "addc.cc.u32 %5,%24,%25;\n\t"
      "addc.cc.u32 %6,%26,%27;\n\t"
      "addc.cc.u32 %7,%28,%29;\n\t"
      "addc.cc.u32 %8,%30,%31;\n\t"
      "addc.cc.u32 %9,%32,%33;\n\t"
      "addc.cc.u32 %10,%34,%35;\n\t"
      "addc.cc.u32 %11,%36,%37;\n\t"
      "addc.cc.u32 %12,%38,0;\n\t"
      "addc.u32 %13,0,0;"
      : "=r"(c[2]), "=r"(c[3]), "=r"(c[4]), "=r"(c[5]), "=r"(c[6]), "=r"(c[7]), "=r"(c[8]), "=r"(c[9]), "=r"(c[10]), "=r"(c[11]),
        "=r"(c[12]), "=r"(c[13]), "=r"(c[14]), "=r"(c[15])
      : "r"((uint32_t)E1), "r"((uint32_t)(O0 >> 32)), "r"((uint32_t)(E1 >> 32)), "r"((uint32_t)O1),
        "r"((uint32_t)E2), "r"((uint32_t)(O1 >> 32)), "r"((uint32_t)(E2 >> 32)), "r"((uint32_t)O2),
        "r"((uint32_t)E3), "r"((uint32_t)(O2 >> 32)), "r"((uint32_t)(E3 >> 32)), "r"((uint32_t)O3),
        "r"((uint32_t)E4), "r"((uint32_t)(O3 >> 32)), "r"((uint32_t)(E4 >> 32)), "r"((uint32_t)O4),
        "r"((uint32_t)E5), "r"((uint32_t)(O4 >> 32)), "r"((uint32_t)(E5 >> 32)), "r"((uint32_t)O5),
        "r"((uint32_t)E6), "r"((uint32_t)(O5 >> 32)), "r"((uint32_t)(E6 >> 32)), "r"((uint32_t)O6),
        "r"((uint32_t)(O6 >> 32)));
  // t = 2 c + squares (a_i^2 on limbs 2i, 2i + 1)
  const u64 S0 = (u64)a0 * a0, S1 = (u64)a1 * a1, S2 = (u64)a2 * a2, S3 = (u64)a3 * a3, S4 = (u64)a4 * a4, S5 = (u64)a5 * a5,
            S6 = (u64)a6 * a6, S7 = (u64)a7 * a7;
  uint32_t t[16];
  asm("add.cc.u32 %0,%15,%15;\n\t"
      "addc.cc.u32 %1,%16,%16;\n\t"
      "addc.cc.u32 %2,%17,%17;\n\t"
      "addc.cc.u32 %3,%18,%18;\n\t"
      "addc.cc.u32 %4,%19,%19;\n\t"
      "addc.cc.u32 %5,%20,%20;\n\t"
      "addc.cc.u32 %6,%21,%21;\n\t"
      "addc.cc.u32 %7,%22,%22;\n\t"
      "addc.cc.u32 %8,%23,%23;\n\t"
      "addc.cc.u32 %9,%24,%24;\n\t"
      "addc.cc.u32 %10,%25,%25;\n\t"
      "addc.cc.u32 %11,%26,%26;\n\t"
      "addc.cc.u32 %12,%27,%27;\n\t"
      "addc.cc.u32 %13,%28,%28;\n\t"
      "addc.u32 %14,%29,%29;"
      : "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(t[8]), "=r"(t[9]), "=r"(t[10]),
        "=r"(t[11]), "=r"(t[12]), "=r"(t[13]), "=r"(t[14]), "=r"(t[15])
      : "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]), "r"(c[5]), "r"(c[6]), "r"(c[7]), "r"(c[8]), "r"(c[9]), "r"(c[10]), "r"(c[11]),
        "r"(c[12]), "r"(c[13]), "r"(c[14]), "r"(c[15]));
  t[0] = (uint32_t)S0;
  asm("add.cc.u32 %0,%0,%15;\n\t"
      "addc.cc.u32 %1,%1,%16;\n\t"
      "addc.cc.u32 %2,%2,%17;\n\t"
      "addc.cc.u32 %3,%3,%18;\n\t"
      "addc.cc.u32 %4,%4,%19;\n\t"
      "addc.cc.u32 %5,%5,%20;\n\t"
      "addc.cc.u32 %6,%6,%21;\n\t"
      "addc.cc.u32 %7,%7,%22;\n\t"
      "addc.cc.u32 %8,%8,%23;\n\t"
      "addc.cc.u32 %9,%9,%24;\n\t"
      "addc.cc.u32 %10,%10,%25;\n\t"
      "addc.cc.u32 %11,%11,%26;\n\t"
      "addc.cc.u32 %12,%12,%27;\n\t"
      "addc.cc.u32 %13,%13,%28;\n\t"
      "addc.u32 %14,%14,%29;"
      : "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "+r"(t[8]), "+r"(t[9]), "+r"(t[10]),
        "+r"(t[11]), "+r"(t[12]), "+r"(t[13]), "+r"(t[14]), "+r"(t[15])
      : "r"((uint32_t)(S0 >> 32)), "r"((uint32_t)S1), "r"((uint32_t)(S1 >> 32)), "r"((uint32_t)S2), "r"((uint32_t)(S2 >> 32)),
        "r"((uint32_t)S3), "r"((uint32_t)(S3 >> 32)), "r"((uint32_t)S4), "r"((uint32_t)(S4 >> 32)), "r"((uint32_t)S5),
        "r"((uint32_t)(S5 >> 32)), "r"((uint32_t)S6), "r"((uint32_t)(S6 >> 32)), "r"((uint32_t)S7), "r"((uint32_t)(S7 >> 32)));
  // Montgomery reduction of the low half: fe_mul's rows without their a * b_i part
  u64 e0 = ((u64)t[1] << 32) | t[0], e1 = ((u64)t[3] << 32) | t[2], e2 = ((u64)t[5] << 32) | t[4], e3 = ((u64)t[7] << 32) | t[6];
  u64 o0 = 0, o1 = 0, o2 = 0, o3 = 0;
  uint32_t pend = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint32_t top = 0;
    const uint32_t m = ((uint32_t)e0 + pend) * P::INV;
    ZKC_CHAIN4_CO(e0, e1, e2, e3, top, P::M(0), P::M(2), P::M(4), P::M(6), m);
    ZKC_CHAIN4_CI(o0, o1, o2, o3, P::M(1), P::M(3), P::M(5), P::M(7), m, pend);
    const uint32_t np = (uint32_t)(e0 >> 32);
    const u64 n0 = o0, n1 = o1, n2 = o2, n3 = o3;
    o0 = e1; o1 = e2; o2 = e3; o3 = (u64)top;
    e0 = n0; e1 = n1; e2 = n2; e3 = n3;
    pend = np;
  }
  Fe<P> r;
  asm("add.cc.u32 %0,%8,%16;\n\t"
      "addc.cc.u32 %1,%9,%17;\n\t"
      "addc.cc.u32 %2,%10,%18;\n\t"
      "addc.cc.u32 %3,%11,%19;\n\t"
      "addc.cc.u32 %4,%12,%20;\n\t"
      "addc.cc.u32 %5,%13,%21;\n\t"
      "addc.cc.u32 %6,%14,%22;\n\t"
      "addc.u32 %7,%15,%23;"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
      : "r"((uint32_t)e0), "r"((uint32_t)(e0 >> 32)), "r"((uint32_t)e1), "r"((uint32_t)(e1 >> 32)), "r"((uint32_t)e2),
        "r"((uint32_t)(e2 >> 32)), "r"((uint32_t)e3), "r"((uint32_t)(e3 >> 32)),
        "r"(pend), "r"((uint32_t)o0), "r"((uint32_t)(o0 >> 32)), "r"((uint32_t)o1), "r"((uint32_t)(o1 >> 32)),
        "r"((uint32_t)o2), "r"((uint32_t)(o2 >> 32)), "r"((uint32_t)o3));
  // + the high half of the square: the sum stays below 2p (the reduced low half is <= p, the high half < p^2 / 2^256 < p / 5)
  asm("add.cc.u32 %0,%0,%8;\n\t"
      "addc.cc.u32 %1,%1,%9;\n\t"
      "addc.cc.u32 %2,%2,%10;\n\t"
      "addc.cc.u32 %3,%3,%11;\n\t"
      "addc.cc.u32 %4,%4,%12;\n\t"
      "addc.cc.u32 %5,%5,%13;\n\t"
      "addc.cc.u32 %6,%6,%14;\n\t"
      "addc.u32 %7,%7,%15;"
      : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]), "+r"(r.v[7])
      : "r"(t[8]), "r"(t[9]), "r"(t[10]), "r"(t[11]), "r"(t[12]), "r"(t[13]), "r"(t[14]), "r"(t[15]));
  reduce_once<P>(r.v);
  return r;
}

#else  // ---- host path: 4 x 64-bit limbs (the 8 x u32 limbs are the same bytes on a little-endian host) ----
// Used by everything the host keeps: MSM epilogues (Horner over the bit-plane sums, normalisation), the driver's transcript
// scalars, keygen pieces, zkc_verify.  Never a fallback for device work.

template <class P> struct HostMod {
  static constexpr uint64_t m(int i) { return (uint64_t)P::M(2 * i) | ((uint64_t)P::M(2 * i + 1) << 32); }
  static constexpr uint64_t inv64() {   // -M^-1 mod 2^64 from the 32-bit constant by one Newton step
    uint64_t ninv = (uint64_t)0 - (uint64_t)P::INV;      // M^-1 mod 2^32
    ninv = ninv * (2 - m(0) * ninv);                      // M^-1 mod 2^64
    return (uint64_t)0 - ninv;
  }
};
template <class P> inline void fe_to_u64(const Fe<P>& a, uint64_t t[4]) { memcpy(t, a.v, 32); }
template <class P> inline Fe<P> fe_from_u64(const uint64_t t[4]) { Fe<P> r; memcpy(r.v, t, 32); return r; }
// t (4 limbs + carry word `hi`) -> t - M when t >= M
template <class P> inline void host_reduce_once(uint64_t t[4], uint64_t hi) {
  typedef unsigned __int128 u128;
  uint64_t d[4];
  u128 bw = 0;
  for (int i = 0; i < 4; ++i) { const u128 x = (u128)t[i] - HostMod<P>::m(i) - (uint64_t)bw; d[i] = (uint64_t)x; bw = (x >> 64) & 1; }
  if (hi || !bw) { t[0] = d[0]; t[1] = d[1]; t[2] = d[2]; t[3] = d[3]; }
}
template <class P> inline bool geq_mod(const uint32_t* a) {
  for (int i = 7; i >= 0; --i) { if (a[i] > P::M(i)) return true; if (a[i] < P::M(i)) return false; }
  return true;
}
template <class P> inline void sub_mod_inplace(uint32_t* a) {
  int64_t bw = 0;
  for (int i = 0; i < 8; ++i) { int64_t d = (int64_t)a[i] - P::M(i) + bw; a[i] = (uint32_t)d; bw = d >> 32; }
}
template <class P> inline Fe<P> fe_add(const Fe<P>& a, const Fe<P>& b) {
  typedef unsigned __int128 u128;
  uint64_t A[4], B[4], t[4];
  fe_to_u64(a, A); fe_to_u64(b, B);
  u128 c = 0;
  for (int i = 0; i < 4; ++i) { c += (u128)A[i] + B[i]; t[i] = (uint64_t)c; c >>= 64; }
  host_reduce_once<P>(t, (uint64_t)c);
  return fe_from_u64<P>(t);
}
template <class P> inline Fe<P> fe_sub(const Fe<P>& a, const Fe<P>& b) {
  typedef unsigned __int128 u128;
  uint64_t A[4], B[4], t[4];
  fe_to_u64(a, A); fe_to_u64(b, B);
  u128 bw = 0;
  for (int i = 0; i < 4; ++i) { const u128 x = (u128)A[i] - B[i] - (uint64_t)bw; t[i] = (uint64_t)x; bw = (x >> 64) & 1; }
  if (bw) { u128 c = 0; for (int i = 0; i < 4; ++i) { c += (u128)t[i] + HostMod<P>::m(i); t[i] = (uint64_t)c; c >>= 64; } }
  return fe_from_u64<P>(t);
}
// host product: 4 x 64-bit CIOS with 128-bit accumulators
template <class P> inline Fe<P> fe_mul(const Fe<P>& a, const Fe<P>& b) {
  typedef unsigned __int128 u128;
  uint64_t A[4], B[4], t[6] = {0, 0, 0, 0, 0, 0};
  fe_to_u64(a, A); fe_to_u64(b, B);
  constexpr uint64_t M0 = HostMod<P>::m(0), M1 = HostMod<P>::m(1), M2 = HostMod<P>::m(2), M3 = HostMod<P>::m(3), inv64 = HostMod<P>::inv64();
  const uint64_t M[4] = {M0, M1, M2, M3};
  for (int i = 0; i < 4; ++i) {
    u128 c = 0;
    for (int j = 0; j < 4; ++j) { c += (u128)A[j] * B[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
    c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
    const uint64_t m = t[0] * inv64;
    c = ((u128)m * M[0] + t[0]) >> 64;
    for (int j = 1; j < 4; ++j) { c += (u128)m * M[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
    c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64); t[5] = 0;
  }
  host_reduce_once<P>(t, t[4]);
  return fe_from_u64<P>(t);
}
template <class P> inline Fe<P> fe_sqr(const Fe<P>& a) { return fe_mul(a, a); }
#endif

template <class P> ZKC_HD Fe<P> fe_neg(const Fe<P>& a) { return fe_sub(fe_zero<P>(), a); }
template <class P> ZKC_HD Fe<P> fe_dbl(const Fe<P>& a) { return fe_add(a, a); }
template <class P> ZKC_HD Fe<P> fe_from_canonical(const Fe<P>& a) { return fe_mul(a, fe_r2<P>()); }
template <class P> ZKC_HD Fe<P> fe_to_canonical(const Fe<P>& a) {
  Fe<P> o = fe_zero<P>(); o.v[0] = 1; return fe_mul(a, o);
}
// x^e for a small exponent
template <class P> ZKC_HD Fe<P> fe_pow_u64(Fe<P> x, u64 e) {
  Fe<P> acc = fe_one<P>();
  while (e) { if (e & 1) acc = fe_mul(acc, x); x = fe_sqr(x); e >>= 1; }
  return acc;
}
#if !defined(__CUDA_ARCH__)
// Host inversion by the binary extended Euclidean algorithm on 4 x 64-bit limbs (0 -> 0): about 3x faster than the Fermat
// ladder below, and the host driver inverts on the Fiat-Shamir critical path (MSM epilogues, Lagrange bases of the opening
// sets).  Works on the Montgomery residue as a plain integer: (aR)^-1 = a^-1 R^-1, then two products by R^2 give a^-1 R.
template <class P> inline Fe<P> fe_inv_host(const Fe<P>& a) {
  if (fe_is_zero(a)) return a;
  typedef unsigned __int128 u128;
  struct U { uint64_t l[4]; };
  auto load = [](const uint32_t* v) { U r; for (int i = 0; i < 4; ++i) r.l[i] = (uint64_t)v[2 * i] | ((uint64_t)v[2 * i + 1] << 32); return r; };
  auto is_one = [](const U& x) { return x.l[0] == 1 && !(x.l[1] | x.l[2] | x.l[3]); };
  auto geq = [](const U& x, const U& y) { for (int i = 3; i >= 0; --i) { if (x.l[i] != y.l[i]) return x.l[i] > y.l[i]; } return true; };
  auto sub = [](U& x, const U& y) { uint64_t bw = 0; for (int i = 0; i < 4; ++i) { const u128 d = (u128)x.l[i] - y.l[i] - bw; x.l[i] = (uint64_t)d; bw = (uint64_t)(d >> 64) & 1; } };
  auto add = [](U& x, const U& y) { u128 c = 0; for (int i = 0; i < 4; ++i) { c += (u128)x.l[i] + y.l[i]; x.l[i] = (uint64_t)c; c >>= 64; } };
  uint32_t mw[8];
  for (int i = 0; i < 8; ++i) mw[i] = P::M(i);
  const U m = load(mw);
  // -m^-1 mod 2^64 (one Newton step from the 32-bit Montgomery constant)
  uint64_t minv = (uint64_t)0 - (uint64_t)P::INV;
  minv = minv * (2 - m.l[0] * minv);
  const uint64_t neg_minv = (uint64_t)0 - minv;
  // x >>= t for 1 <= t <= 63
  auto shr = [](U& x, int t) { for (int i = 0; i < 3; ++i) x.l[i] = (x.l[i] >> t) | (x.l[i + 1] << (64 - t)); x.l[3] >>= t; };
  // x / 2^t mod m for x < m: add the multiple k*m that clears the low t bits, then shift (the sum stays below 2^t * m)
  auto div2t = [&](U& x, int t) {
    const uint64_t k = (x.l[0] * neg_minv) & (((uint64_t)1 << t) - 1);
    uint64_t w[5];
    u128 c = 0;
    for (int i = 0; i < 4; ++i) { c += (u128)k * m.l[i] + x.l[i]; w[i] = (uint64_t)c; c >>= 64; }
    w[4] = (uint64_t)c;
    for (int i = 0; i < 4; ++i) x.l[i] = (w[i] >> t) | (w[i + 1] << (64 - t));
  };
  // strip the factors of two of an even, non-zero u, keeping x * a = u (mod m)
  auto make_odd = [&](U& u, U& x) {
    while (!(u.l[0] & 1)) {
      int t = u.l[0] ? __builtin_ctzll(u.l[0]) : 63;
      if (t > 63) t = 63;
      shr(u, t);
      div2t(x, t);
    }
  };
  auto sub_mod = [&](U& x, const U& y) { if (!geq(x, y)) add(x, m); sub(x, y); };   // x - y mod m for x, y < m
  U u = load(a.v), v = m, x1{{1, 0, 0, 0}}, x2{{0, 0, 0, 0}};
  make_odd(u, x1);
  for (;;) {
    if (is_one(u)) break;
    if (is_one(v)) { x1 = x2; break; }
    if (geq(u, v)) { sub(u, v); sub_mod(x1, x2); make_odd(u, x1); }
    else { sub(v, u); sub_mod(x2, x1); make_odd(v, x2); }
  }
  Fe<P> t;
  for (int i = 0; i < 4; ++i) { t.v[2 * i] = (uint32_t)x1.l[i]; t.v[2 * i + 1] = (uint32_t)(x1.l[i] >> 32); }
  const Fe<P> r2 = fe_r2<P>();
  return fe_mul(fe_mul(t, r2), r2);
}
#endif
// Fermat inversion (0 -> 0).  ~380 multiplications; batch inversion is preferred on hot paths.
template <class P> ZKC_HD Fe<P> fe_inv(const Fe<P>& a) {
#if !defined(__CUDA_ARCH__)
  return fe_inv_host(a);
#else
  Fe<P> acc = fe_one<P>();
  for (int i = 7; i >= 0; --i) {
    uint32_t w = P::M(i) - (i == 0 ? 2u : 0u);
    for (int bit = 31; bit >= 0; --bit) {
      acc = fe_sqr(acc);
      if ((w >> bit) & 1) acc = fe_mul(acc, a);
    }
  }
  return acc;
#endif
}
// the Fermat ladder on the host too (cross-check of fe_inv_host in tests)
template <class P> inline Fe<P> fe_inv_fermat_host(const Fe<P>& a) {
  Fe<P> acc = fe_one<P>();
  for (int i = 7; i >= 0; --i) {
    uint32_t w = P::M(i) - (i == 0 ? 2u : 0u);
    for (int bit = 31; bit >= 0; --bit) {
      acc = fe_mul(acc, acc);
      if ((w >> bit) & 1) acc = fe_mul(acc, a);
    }
  }
  return acc;
}

// ---- 16-byte vector load/store helpers -----------------------------------------------------------
#if defined(__CUDACC__)
template <class P> ZKC_D Fe<P> fe_load(const Fe<P>* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1];
  Fe<P> r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
template <class P> ZKC_D Fe<P> fe_load_nc(const Fe<P>* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  Fe<P> r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
template <class P> ZKC_D void fe_store(Fe<P>* p, const Fe<P>& r) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
  q[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
template <class P> ZKC_D Fe<P> fe_from_halves(const uint4& a, const uint4& b) {
  Fe<P> r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
template <class P> ZKC_D uint4 fe_lo(const Fe<P>& r) { return make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]); }
template <class P> ZKC_D uint4 fe_hi(const Fe<P>& r) { return make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]); }
#endif

}  // namespace zkc
