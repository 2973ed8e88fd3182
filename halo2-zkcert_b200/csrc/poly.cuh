// Device-side polynomial / column primitives used by the create_proof pipeline (prover.cu).
// Each replaces a serial or rayon loop of halo2_proofs; see poly.cu.
#pragma once
#include "common.cuh"

namespace zkc {

// compiled constraint-system program resident on the device (see circuit.py for the opcodes)
struct DevProgram {
  uint32_t* words = nullptr;   // (op, arg) pairs
  uint32_t npairs = 0;
  Fr* consts = nullptr;        // Montgomery
  uint32_t nconsts = 0;
  uint32_t nexprs = 0;
  // factored groups (host/cs.h optimize_program): run lengths whose power of `mult` the GROUP_END op needs, by slot
  uint32_t npows = 0;
  uint32_t pow_len[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

// query tables + column pointer tables for one evaluation domain (Lagrange rows or extended coset)
struct DevQueries {
  const uint32_t* aq_col; const int32_t* aq_rot;
  const uint32_t* fq_col; const int32_t* fq_rot;
  const uint32_t* iq_col; const int32_t* iq_rot;
  const Fr* const* advice; const Fr* const* fixed; const Fr* const* instance;   // device arrays of column pointers
};

enum { SCAN_ADD = 0, SCAN_MUL = 1 };

// out[i] = start (+|*) fold_{j<i} in[j]            (reverse = 0)
// out[i] = start (+|*) fold_{j>i} in[j]            (reverse = 1)        in-place allowed
int fr_scan(zkc_ctx* ctx, const Fr* in, Fr* out, uint64_t n, int op, int reverse, const Fr& start);
// exclusive u32 prefix sum; total written to *total_dev (device) if non-null.  in-place allowed
int u32_scan(zkc_ctx* ctx, const uint32_t* in, uint32_t* out, uint64_t n, uint32_t* total_dev);
// out[i] = first * base^i
int fr_powers(zkc_ctx* ctx, Fr* out, uint64_t n, const Fr& base, const Fr& first);
// out = sum_j coef[j] * polys[j][0..len_j)  (zero-extended to n).  `polys`/`lens`/`coefs` are host arrays.
int fr_lincomb(zkc_ctx* ctx, Fr* out, uint64_t n, const std::vector<const Fr*>& polys, const std::vector<Fr>& coefs);
// a[i] -= low[i] for i < m (m small; `low` host values)
int fr_sub_low(zkc_ctx* ctx, Fr* a, const std::vector<Fr>& low);
// a[i] *= s
int fr_scale(zkc_ctx* ctx, Fr* a, uint64_t n, const Fr& s);
// a[i] = a[i] * s + b[i]
int fr_mul_add(zkc_ctx* ctx, Fr* a, const Fr* b, uint64_t n, const Fr& s);
// kate_division: q(X) = (a(X) - a(z)) / (X - z); writes n coefficients (q[n-1] = 0).  Uses two scratch columns.
int fr_kate_division(zkc_ctx* ctx, const Fr* a, Fr* q, uint64_t n, const Fr& z, Fr* tmp1, Fr* tmp2);
// in-place batch: polys[j] <- polys[j] / (X - roots[j]) for independent jobs, one set of launches (tmp: n elements)
#define KD_MAX_JOBS 16
int fr_kate_division_batch(zkc_ctx* ctx, const std::vector<Fr*>& polys, const std::vector<Fr>& roots, uint64_t n, Fr* tmp);
// q[j] += C * z^(len-1-j), j < len
int fr_add_geometric(zkc_ctx* ctx, Fr* q, uint64_t len, const Fr& C, const Fr& z);
// evaluate polys[j] (n coefficients each) at points[j]; results to host `out`
int fr_eval_batch(zkc_ctx* ctx, const std::vector<const Fr*>& polys, uint64_t n, const std::vector<Fr>& points, std::vector<Fr>& out);
// batch inversion (zeros pass through)
int fr_batch_invert(zkc_ctx* ctx, const Fr* a, Fr* out, size_t n);

// postfix-program evaluation over `rows` rows: acc = 0; for each expression: acc = acc * mult + expr
// (mult = theta for lookup compression, y for the custom-gate part of h(X)).  If `accumulate`, the
// running value starts from out[i] instead of 0.  `rows` is the domain size (rotations wrap modulo it); only rows
// [row0, row0 + cnt) are evaluated (default: all).
int eval_program(zkc_ctx* ctx, const DevProgram& prog, const DevQueries& q, Fr* out, uint64_t rows, uint32_t rot_scale, const Fr& mult,
                 int accumulate, uint64_t row0 = 0, uint64_t cnt = UINT64_MAX);

// 256-bit ascending sort of canonical values (n a power of two)
int sort_u256(zkc_ctx* ctx, Fr* keys, uint64_t n);
// the same for keys[0..U) canonical + keys[U..n) all-ones sentinel; counting sort when every value is < 2^24
int sort_u256_padded(zkc_ctx* ctx, Fr* keys, uint64_t n, uint64_t U);

}  // namespace zkc
