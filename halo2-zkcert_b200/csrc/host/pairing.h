// BN254 G2 arithmetic and the optimal-ate pairing on the host — what `verify_proof` needs for the final KZG check
// e(W, [s]G2) = e(P, G2) (halo2_proofs 0.2.0 @4b42325 src/poly/kzg/strategy.rs / msm.rs DualMSM::check through
// halo2curves 0.4.0 @e185711 `multi_miller_loop` + `final_exponentiation`; un-vendored, /root/reference/Cargo.lock:1320-1380).
// Host-only, a few tens of milliseconds per check: Fq2 affine G2 arithmetic, line functions embedded in
// Fq12 = Fq[w] / (w^12 - 18 w^6 + 82) (w^6 = 9 + u), Miller loop over 6x+2 with the two Frobenius corrections, final
// exponentiation by (p^12 - 1) / r as a plain square-and-multiply.  Checked against oracle/pairing.py in tests/.
#pragma once
#include <vector>
#include "../ff.cuh"
#include "../ec.cuh"

namespace zkc {
namespace host {

struct Fq2 { Fq c0, c1; };              // c0 + c1 u, u^2 = -1 (halo2curves layout)
struct G2Affine { Fq2 x, y; };          // identity = (0, 0)

Fq2 fq2_zero();
Fq2 fq2_one();
bool fq2_is_zero(const Fq2& a);
bool fq2_eq(const Fq2& a, const Fq2& b);
Fq2 fq2_add(const Fq2& a, const Fq2& b);
Fq2 fq2_sub(const Fq2& a, const Fq2& b);
Fq2 fq2_neg(const Fq2& a);
Fq2 fq2_mul(const Fq2& a, const Fq2& b);
Fq2 fq2_inv(const Fq2& a);

bool g2_is_identity(const G2Affine& p);
bool g2_on_curve(const G2Affine& p);                     // y^2 = x^3 + 3 / (9 + u)
G2Affine g2_generator();
G2Affine g2_add(const G2Affine& p, const G2Affine& q);
G2Affine g2_mul(const G2Affine& p, const Fr& scalar);    // scalar in Montgomery form

// prod_i e(P_i, Q_i) == 1 ?   (pairs with an identity on either side contribute 1)
bool pairing_product_is_one(const std::vector<std::pair<G1Affine, G2Affine>>& pairs);

// x^e for a little-endian multi-word exponent
Fq fq_pow_words(const Fq& x, const unsigned long long* e, int nwords);

}  // namespace host
}  // namespace zkc
