// ParamsKZG<Bn256> on disk — the `kzg_bn254_{k}.srs` files halo2-base's gen_srs writes and every prove command of the
// reference reads (/root/reference/src/helpers.rs:210,262; src/bin/cli.rs:221-224, 234, 256 ...).  Layout of
// `ParamsKZG::write_custom` (halo2_proofs 0.2.0 @4b42325 src/poly/kzg/commitment.rs, un-vendored; SURVEY OPEN-8):
//     k : u32 little-endian | g[0..n) | g_lagrange[0..n) | g2 | s_g2
// with every point in `SerdeFormat::RawBytes` / `RawBytesUnchecked` form: the coordinates' Montgomery limbs as they sit in
// memory (G1Affine 64 B, G2Affine 128 B).  RawBytes validates what it reads (canonical coordinates, on the curve);
// Unchecked does not.  `SerdeFormat::Processed` (compressed points) is not handled: its G2 sign convention is not
// recoverable from what is on disk here.  Host only: the caller hands the arrays to zkc_srs_load.
#include <cstring>
#include "../../../include/zkcert_cuda.h"
#include "pairing.h"

using namespace zkc;
using namespace zkc::host;

namespace {
bool fq_canonical_limbs(const Fq& a) { return !geq_mod<FqP>(a.v); }   // Montgomery residues are stored reduced
bool g1_valid(const G1Affine& p) {
  if (!fq_canonical_limbs(p.x) || !fq_canonical_limbs(p.y)) return false;
  if (affine_is_identity(p)) return true;
  Fq three = fe_zero<FqP>(); three.v[0] = 3; three = fe_from_canonical(three);
  return fe_eq(fe_sqr(p.y), fe_add(fe_mul(fe_sqr(p.x), p.x), three));
}
bool g2_valid(const G2Affine& p) {
  return fq_canonical_limbs(p.x.c0) && fq_canonical_limbs(p.x.c1) && fq_canonical_limbs(p.y.c0) && fq_canonical_limbs(p.y.c1) && g2_on_curve(p);
}
}  // namespace

extern "C" size_t zkc_params_size(uint32_t k) { return k > 28 ? 0 : 4 + 2 * ((size_t)64 << k) + 2 * 128; }

extern "C" int zkc_params_write(uint32_t k, const zkc_g1_affine* g, const zkc_g1_affine* g_lagrange, const zkc_g2_affine* g2,
                                const zkc_g2_affine* s_g2, uint8_t* out, size_t cap) {
  if (!g || !g_lagrange || !g2 || !s_g2 || !out || k > 28 || cap < zkc_params_size(k)) return ZKC_ERR_BAD_ARG;
  const size_t nb = (size_t)64 << k;
  memcpy(out, &k, 4);
  memcpy(out + 4, g, nb);
  memcpy(out + 4 + nb, g_lagrange, nb);
  memcpy(out + 4 + 2 * nb, g2, 128);
  memcpy(out + 4 + 2 * nb + 128, s_g2, 128);
  return ZKC_OK;
}

// `checked` != 0: SerdeFormat::RawBytes (every point validated); 0: RawBytesUnchecked.  g / g_lagrange may be NULL to read k only.
extern "C" int zkc_params_read(const uint8_t* in, size_t len, int checked, uint32_t* k_out, zkc_g1_affine* g, zkc_g1_affine* g_lagrange,
                               zkc_g2_affine* g2, zkc_g2_affine* s_g2) {
  if (!in || len < 4 || !k_out) return ZKC_ERR_BAD_ARG;
  uint32_t k;
  memcpy(&k, in, 4);
  if (k > 28 || len != zkc_params_size(k)) return ZKC_ERR_BAD_ARG;
  *k_out = k;
  if (!g && !g_lagrange && !g2 && !s_g2) return ZKC_OK;
  if (!g || !g_lagrange || !g2 || !s_g2) return ZKC_ERR_BAD_ARG;
  const size_t nb = (size_t)64 << k, n = (size_t)1 << k;
  memcpy(g, in + 4, nb);
  memcpy(g_lagrange, in + 4 + nb, nb);
  memcpy(g2, in + 4 + 2 * nb, 128);
  memcpy(s_g2, in + 4 + 2 * nb + 128, 128);
  if (checked) {
    const G1Affine* a = reinterpret_cast<const G1Affine*>(g);
    const G1Affine* b = reinterpret_cast<const G1Affine*>(g_lagrange);
    for (size_t i = 0; i < n; ++i) if (!g1_valid(a[i]) || !g1_valid(b[i])) return ZKC_ERR_BAD_ARG;
    G2Affine p, q;
    memcpy(&p, g2, 128); memcpy(&q, s_g2, 128);
    if (!g2_valid(p) || !g2_valid(q)) return ZKC_ERR_BAD_ARG;
  }
  return ZKC_OK;
}
