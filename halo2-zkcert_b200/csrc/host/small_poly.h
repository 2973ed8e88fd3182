// Small host-side polynomial helpers shared by the create_proof driver and the verifier: O(#queries) scalar work that
// halo2 also keeps on the CPU (arithmetic::lagrange_interpolate, eval_polynomial, evaluate_vanishing_polynomial,
// EvaluationDomain::rotate_omega; halo2_proofs 0.2.0 @4b42325, un-vendored).
#pragma once
#include <cstdint>
#include <vector>
#include "../ff.cuh"

namespace zkc {
namespace host {

inline Fr rotate_omega(const Fr& x, const Fr& omega, const Fr& omega_inv, int32_t rot) {
  return rot >= 0 ? fe_mul(x, fe_pow_u64(omega, (u64)rot)) : fe_mul(x, fe_pow_u64(omega_inv, (u64)(-(int64_t)rot)));
}

// coefficients of the interpolation polynomial through (points[i], evals[i])
inline std::vector<Fr> lagrange_interpolate(const std::vector<Fr>& pts, const std::vector<Fr>& evals) {
  const size_t m = pts.size();
  std::vector<Fr> coeffs(m, fe_zero<FrP>());
  if (m == 1) { coeffs[0] = evals[0]; return coeffs; }
  for (size_t j = 0; j < m; ++j) {
    std::vector<Fr> num(1, fe_one<FrP>());
    Fr den = fe_one<FrP>();
    for (size_t kx = 0; kx < m; ++kx) {
      if (kx == j) continue;
      std::vector<Fr> nxt(num.size() + 1, fe_zero<FrP>());
      for (size_t i = 0; i < num.size(); ++i) {
        nxt[i + 1] = fe_add(nxt[i + 1], num[i]);
        nxt[i] = fe_sub(nxt[i], fe_mul(pts[kx], num[i]));
      }
      num.swap(nxt);
      den = fe_mul(den, fe_sub(pts[j], pts[kx]));
    }
    const Fr scale = fe_mul(evals[j], fe_inv(den));
    for (size_t i = 0; i < m; ++i) coeffs[i] = fe_add(coeffs[i], fe_mul(num[i], scale));
  }
  return coeffs;
}
// Lagrange basis of a point set: basis[j] = coefficients of L_j(X) (L_j(points[i]) = [i == j]).  One inversion per point,
// shared by every polynomial opened at this set (r(X) = sum_j evals[j] * L_j(X)).
inline std::vector<std::vector<Fr>> lagrange_basis(const std::vector<Fr>& pts) {
  const size_t m = pts.size();
  std::vector<std::vector<Fr>> basis(m);
  for (size_t j = 0; j < m; ++j) {
    std::vector<Fr> num(1, fe_one<FrP>());
    Fr den = fe_one<FrP>();
    for (size_t kx = 0; kx < m; ++kx) {
      if (kx == j) continue;
      std::vector<Fr> nxt(num.size() + 1, fe_zero<FrP>());
      for (size_t i = 0; i < num.size(); ++i) {
        nxt[i + 1] = fe_add(nxt[i + 1], num[i]);
        nxt[i] = fe_sub(nxt[i], fe_mul(pts[kx], num[i]));
      }
      num.swap(nxt);
      den = fe_mul(den, fe_sub(pts[j], pts[kx]));
    }
    const Fr scale = m == 1 ? fe_one<FrP>() : fe_inv(den);
    basis[j].resize(m);
    for (size_t i = 0; i < m; ++i) basis[j][i] = fe_mul(num[i], scale);
  }
  return basis;
}
inline std::vector<Fr> interpolate_with_basis(const std::vector<std::vector<Fr>>& basis, const std::vector<Fr>& evals) {
  const size_t m = basis.size();
  std::vector<Fr> coeffs(m, fe_zero<FrP>());
  for (size_t j = 0; j < m; ++j)
    for (size_t i = 0; i < m; ++i) coeffs[i] = fe_add(coeffs[i], fe_mul(basis[j][i], evals[j]));
  return coeffs;
}

inline Fr eval_small(const std::vector<Fr>& c, const Fr& x) {
  Fr acc = fe_zero<FrP>();
  for (size_t i = c.size(); i-- > 0;) acc = fe_add(fe_mul(acc, x), c[i]);
  return acc;
}
inline Fr vanishing_eval(const std::vector<Fr>& roots, const Fr& z) {
  Fr acc = fe_one<FrP>();
  for (auto& r : roots) acc = fe_mul(acc, fe_sub(z, r));
  return acc;
}


}  // namespace host
}  // namespace zkc
