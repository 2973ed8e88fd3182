// create_proof on the B200: the host-side driver that mirrors halo2_proofs::plonk::create_proof
// (halo2_proofs 0.2.0 "halo2-axiom" @4b42325 src/plonk/prover.rs, with ProverSHPLONK / ProverGWC and
// the Blake2b / Keccak256 transcripts) — un-vendored, pinned at /root/reference/Cargo.lock:1320-1336;
// reached from /root/reference/src/helpers.rs:233,299 and src/bin/cli.rs:320,369,462 through
// snark-verifier-sdk's gen_snark_shplonk.  SURVEY.md §3.2 is the step list this file follows.
//
// The host keeps what the reference keeps on the CPU: transcript hashing, challenge derivation,
// the RNG cursor, rotation-set bookkeeping and O(#queries) scalar work.  Every O(n) operation runs
// on the device; columns stay resident in HBM from the advice upload to the last opening proof,
// and only commitments (64 B), evaluations (32 B) and challenges cross PCIe.
#include "prover_kernels.cuh"
#include "msm.cuh"
#include "dist.cuh"
#include "host/hostutil.h"
#include "host/cs.h"
#include "host/small_poly.h"
#include <algorithm>
#include <functional>
#include <memory>

namespace zkc {
// implemented in ntt.cu / msm.cu
int ntt_twiddles(zkc_ctx* ctx, uint32_t log_n, const Fr** out);
int dom_lagrange_to_coeff(zkc_ctx* ctx, const zkc_domain* d, Fr* a, uint32_t ncols);
int dom_coeff_to_extended(zkc_ctx* ctx, const zkc_domain* d, const Fr* in, uint64_t in_stride, Fr* out, uint32_t ncols);
int dom_extended_to_coeff(zkc_ctx* ctx, const zkc_domain* d, Fr* a, uint32_t ncols);
int dom_divide_by_vanishing(zkc_ctx* ctx, const zkc_domain* d, Fr* a, uint64_t row0, uint64_t cnt);
int dom_coeff_to_classes(zkc_ctx* ctx, const zkc_domain* d, const Fr* in, uint64_t in_stride, Fr* out, uint32_t ncols, uint32_t c0, uint32_t c1);
int dom_classes_to_natural(zkc_ctx* ctx, const zkc_domain* d, const Fr* cm, Fr* nat);
int dom_classes_to_pieces(zkc_ctx* ctx, const zkc_domain* d, Fr* vals, Fr* out);
int dom_classes_inverse(zkc_ctx* ctx, const zkc_domain* d, Fr* vals, uint32_t c0, uint32_t c1);
int dom_classes_mix(zkc_ctx* ctx, const zkc_domain* d, const Fr* g, Fr* out);
int dom_divide_by_vanishing_classes(zkc_ctx* ctx, const zkc_domain* d, Fr* a, uint64_t row0, uint64_t cnt);
int srs_commit_dev(zkc_ctx* ctx, const zkc_srs* s, int basis, const Fr* polys, uint64_t len, uint32_t ncols, zkc_g1* out);
}  // namespace zkc

using namespace zkc;
using zkc::host::Transcript;

extern "C" uint32_t zkc_srs_k(const zkc_srs* srs);

using zkc::host::Cs;
using zkc::host::HostProgram;
using zkc::host::rotate_omega;
using zkc::host::lagrange_interpolate;
using zkc::host::eval_small;
using zkc::host::vanishing_eval;

// ---- proving key -------------------------------------------------------------------------------------------
struct zkc_pk {
  zkc_ctx* ctx = nullptr;
  const zkc_srs* srs = nullptr;
  Cs cs;
  zkc_domain* dom = nullptr;
  uint32_t ext_k = 0;
  Fr transcript_repr;
  // device-resident columns
  Fr *fixed_values = nullptr, *fixed_polys = nullptr, *fixed_cosets = nullptr;
  Fr *sigma_values = nullptr, *sigma_polys = nullptr, *sigma_cosets = nullptr;
  Fr *l0 = nullptr, *l_last = nullptr, *l_active = nullptr, *omega_pows = nullptr;
  Fr* l_polys = nullptr;   // l_0, l_last, l_blind in coefficient form (3 n): ProvingKey::write re-derives the natural-order cosets from them
  // programs + query tables
  DevProgram gates;
  std::vector<std::pair<DevProgram, DevProgram>> lookups;
  uint32_t* qtab = nullptr;          // aq_col, aq_rot, fq_col, fq_rot, iq_col, iq_rot packed
  const Fr** fixed_val_ptrs = nullptr;   // device arrays of column pointers
  const Fr** fixed_coset_ptrs = nullptr;
  std::vector<void*> owned;          // everything to cudaFree
  std::vector<G1Affine> fixed_comm, sigma_comm;
};

namespace {

int upload_program(zkc_ctx* ctx, zkc_pk* pk, const HostProgram& h, DevProgram& d) {
  // the device runs the factored stream (host/cs.h optimize_program: shared selectors taken out of runs of constraints)
  const zkc::host::OptimizedProgram opt = zkc::host::optimize_program(h, ctx->tune.no_program_factoring ? 0 : 8);
  d.npairs = (uint32_t)(opt.words.size() / 2); d.nconsts = (uint32_t)h.consts.size(); d.nexprs = h.nexprs;
  d.npows = (uint32_t)opt.pow_len.size();
  for (uint32_t s = 0; s < d.npows; ++s) d.pow_len[s] = opt.pow_len[s];
  if (d.npairs) {
    ZKC_CUDA_TRY(ctx, cudaMalloc(&d.words, opt.words.size() * 4)); pk->owned.push_back(d.words);
    ZKC_CUDA_TRY(ctx, cudaMemcpy(d.words, opt.words.data(), opt.words.size() * 4, cudaMemcpyHostToDevice));
  }
  if (d.nconsts) {
    ZKC_CUDA_TRY(ctx, cudaMalloc(&d.consts, h.consts.size() * sizeof(Fr))); pk->owned.push_back(d.consts);
    ZKC_CUDA_TRY(ctx, cudaMemcpy(d.consts, h.consts.data(), h.consts.size() * sizeof(Fr), cudaMemcpyHostToDevice));
  }
  return ZKC_OK;
}

template <class T> int dev_alloc(zkc_ctx* ctx, zkc_pk* pk, T** p, size_t count) {
  ZKC_CUDA_TRY(ctx, cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T)));
  pk->owned.push_back((void*)*p);
  return ZKC_OK;
}

void g1_from_abi(const zkc_g1& j, G1Affine& a) {
  Fq z; memcpy(z.v, &j.z, 32);
  if (fe_is_zero(z)) { a.x = fe_zero<FqP>(); a.y = fe_zero<FqP>(); return; }
  memcpy(a.x.v, &j.x, 32); memcpy(a.y.v, &j.y, 32);
}

int commit_points(zkc_ctx* ctx, const zkc_srs* srs, int basis, const Fr* polys, uint64_t len, uint32_t ncols, std::vector<G1Affine>& out) {
  std::vector<zkc_g1> tmp(ncols);
  ZKC_TRY(srs_commit_dev(ctx, srs, basis, polys, len, ncols, tmp.data()));
  out.resize(ncols);
  for (uint32_t i = 0; i < ncols; ++i) g1_from_abi(tmp[i], out[i]);
  return ZKC_OK;
}

}  // namespace

extern "C" void zkc_pk_free(zkc_pk* pk) {
  if (!pk) return;
  cudaSetDevice(pk->ctx->dev);
  cudaStreamSynchronize(pk->ctx->stream);
  for (void* p : pk->owned) cudaFree(p);
  if (pk->dom) zkc_domain_free(pk->dom);
  delete pk;
}

namespace {
// where the Lagrange columns of a key come from: one contiguous host / device array, or per-column host pointers (a `.pk`
// file: every column sits behind its own length prefix), or — sigma only — a permutation mapping to expand on the device
struct ColumnSource {
  const zkc_fr* contiguous = nullptr; bool on_device = false;
  const uint8_t* file = nullptr; uint64_t first_off = 0, col_stride = 0;   // column c at file + first_off + c * col_stride
  const uint64_t* mapping = nullptr;                                        // sigma: host mapping[col * n + row] = col' * n + row'
  bool present() const { return contiguous || file || mapping; }
};
// sigma_col[row] = DELTA^col' * omega^row' through the permutation mapping (permutation::keygen::Assembly::build_pk)
__global__ void k_sigma_from_mapping(const uint64_t* mapping, const Fr* delta_pows, const Fr* omega_pows, Fr* sigma, uint64_t n, uint64_t total) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const uint64_t m = mapping[i];
  fe_store(sigma + i, fe_mul(fe_load(delta_pows + m / n), fe_load_nc(omega_pows + m % n)));
}
int upload_columns(zkc_ctx* ctx, Fr* dst, const ColumnSource& src, uint32_t ncols, uint64_t n) {
  cudaStream_t st = ctx->stream;
  if (src.contiguous) {
    ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(dst, src.contiguous, (size_t)ncols * n * sizeof(Fr), src.on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
  } else {
    for (uint32_t c = 0; c < ncols; ++c)
      ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(dst + (size_t)c * n, src.file + src.first_off + (uint64_t)c * src.col_stride, n * sizeof(Fr), cudaMemcpyHostToDevice, st));
  }
  return ZKC_OK;
}
int pk_build(zkc_ctx* ctx, const zkc_srs* srs, const uint8_t* cs_blob, size_t cs_len, const ColumnSource& fixed, const ColumnSource& sigma,
             const zkc_fr* transcript_repr, int zeta_choice, zkc_pk** out);
}  // namespace

extern "C" int zkc_pk_load(zkc_ctx* ctx, const zkc_srs* srs, const uint8_t* cs_blob, size_t cs_len, const zkc_fr* fixed,
                           const zkc_fr* sigma, const zkc_fr* transcript_repr, int zeta_choice, zkc_pk** out) {
  if (!ctx || !srs || !cs_blob || !transcript_repr || !out) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_pk_load: null argument");
  CtxLock lock(ctx);
  ColumnSource f, s;
  f.contiguous = fixed; s.contiguous = sigma;
  return pk_build(ctx, srs, cs_blob, cs_len, f, s, transcript_repr, zeta_choice, out);
}

namespace {
int pk_build(zkc_ctx* ctx, const zkc_srs* srs, const uint8_t* cs_blob, size_t cs_len, const ColumnSource& fixed, const ColumnSource& sigma,
             const zkc_fr* transcript_repr, int zeta_choice, zkc_pk** out) {
  std::unique_ptr<zkc_pk, void (*)(zkc_pk*)> pk(new zkc_pk(), zkc_pk_free);
  pk->ctx = ctx; pk->srs = srs;
  Cs& cs = pk->cs;
  {
    std::string perr;
    if (!zkc::host::parse_cs(cs_blob, cs_len, cs, perr)) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_pk_load: " + perr);
    if (cs.k != zkc_srs_k(srs)) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_pk_load: k does not match the SRS");
  }
  const uint32_t nperm = (uint32_t)cs.perm.size();
  const uint64_t n = cs.n();
  if (n < cs.blinding_factors + 3) return set_err(ctx, ZKC_ERR_NOT_ENOUGH_ROWS, "zkc_pk_load: not enough rows for the blinding factors");
  if (cs.chunk_len > PERM_MAX_CHUNK) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_pk_load: permutation chunk longer than PERM_MAX_CHUNK");
  if (cs.nsets() > PERM_MAX_SETS) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_pk_load: too many permutation sets");
  if ((cs.num_fixed && !fixed.present()) || (nperm && !sigma.present())) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_pk_load: missing fixed / sigma columns");
  ZKC_TRY(zkc_domain_create(ctx, cs.degree, cs.k, zeta_choice, &pk->dom));
  zkc_domain_info di;
  zkc_domain_get_info(pk->dom, &di);
  pk->ext_k = di.extended_k;
  const uint64_t en = 1ull << pk->ext_k;
  // every extended coset of the key is kept CLASS-MAJOR (ntt.cu, dom_coeff_to_classes), and only the degree - 1 residue classes
  // that determine h(X) are ever evaluated (dom_classes_to_pieces): the remaining 2^e - (degree - 1) class blocks stay unused
  const uint32_t ncls = cs.degree - 1;
  memcpy(pk->transcript_repr.v, transcript_repr, 32);
  zkc_pk* P = pk.get();
  const uint32_t F = cs.num_fixed, S = nperm;
  // fixed + sigma: values -> polys -> cosets
  ZKC_TRY(dev_alloc(ctx, P, &P->fixed_values, (size_t)F * n)); ZKC_TRY(dev_alloc(ctx, P, &P->fixed_polys, (size_t)F * n));
  ZKC_TRY(dev_alloc(ctx, P, &P->fixed_cosets, (size_t)F * en));
  ZKC_TRY(dev_alloc(ctx, P, &P->sigma_values, (size_t)S * n)); ZKC_TRY(dev_alloc(ctx, P, &P->sigma_polys, (size_t)S * n));
  ZKC_TRY(dev_alloc(ctx, P, &P->sigma_cosets, (size_t)S * en));
  cudaStream_t st = ctx->stream;
  // omega^i (needed first when the sigma columns are expanded from a permutation mapping)
  ZKC_TRY(dev_alloc(ctx, P, &P->omega_pows, n));
  {
    Fr omega; memcpy(omega.v, &di.omega, 32);
    ZKC_TRY(fr_powers(ctx, P->omega_pows, n, omega, fe_one<FrP>()));
  }
  if (F) {
    ZKC_TRY(upload_columns(ctx, P->fixed_values, fixed, F, n));
    ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(P->fixed_polys, P->fixed_values, (size_t)F * n * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
    ZKC_TRY(dom_lagrange_to_coeff(ctx, P->dom, P->fixed_polys, F));
    ZKC_TRY(dom_coeff_to_classes(ctx, P->dom, P->fixed_polys, n, P->fixed_cosets, F, 0, ncls));
  }
  if (S && sigma.mapping) {
    // DELTA^c on the host (a handful of products), the n * S gathers on the device
    std::vector<Fr> dp(S);
    Fr d = fe_one<FrP>();
    const Fr DELTA = fr_from_raw_words(FR_DELTA_RAW);
    for (uint32_t c = 0; c < S; ++c) { dp[c] = d; d = fe_mul(d, DELTA); }
    struct DevTmp { void* p = nullptr; ~DevTmp() { if (p) cudaFree(p); } } map_h, dp_h;
    ZKC_CUDA_TRY(ctx, cudaMalloc(&map_h.p, (size_t)S * n * sizeof(uint64_t)));
    ZKC_CUDA_TRY(ctx, cudaMalloc(&dp_h.p, (size_t)S * sizeof(Fr)));
    ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(map_h.p, sigma.mapping, (size_t)S * n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(dp_h.p, dp.data(), (size_t)S * sizeof(Fr), cudaMemcpyHostToDevice, st));
    k_sigma_from_mapping<<<(unsigned)(((uint64_t)S * n + 255) / 256), 256, 0, st>>>((const uint64_t*)map_h.p, (const Fr*)dp_h.p, P->omega_pows, P->sigma_values, n, (uint64_t)S * n);
    ctx->launches++;
    ZKC_CUDA_TRY(ctx, cudaGetLastError());
    ZKC_CUDA_TRY(ctx, cudaStreamSynchronize(st));   // the temporaries are released when this block ends
  }
  if (S) {
    if (!sigma.mapping) ZKC_TRY(upload_columns(ctx, P->sigma_values, sigma, S, n));
    ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(P->sigma_polys, P->sigma_values, (size_t)S * n * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
    ZKC_TRY(dom_lagrange_to_coeff(ctx, P->dom, P->sigma_polys, S));
    ZKC_TRY(dom_coeff_to_classes(ctx, P->dom, P->sigma_polys, n, P->sigma_cosets, S, 0, ncls));
  }
  // l0, l_last, l_blind -> cosets; l_active = 1 - l_last - l_blind
  {
    ZKC_TRY(dev_alloc(ctx, P, &P->l0, en)); ZKC_TRY(dev_alloc(ctx, P, &P->l_last, en)); ZKC_TRY(dev_alloc(ctx, P, &P->l_active, en));
    struct DevTmp { void* p = nullptr; ~DevTmp() { if (p) cudaFree(p); } } lb_h;   // freed on every exit path
    ZKC_TRY(dev_alloc(ctx, P, &P->l_polys, 3 * n));
    ZKC_CUDA_TRY(ctx, cudaMalloc(&lb_h.p, en * sizeof(Fr)));
    Fr* tmp = P->l_polys;   // three Lagrange columns: l0, l_last, l_blind
    Fr* lb = (Fr*)lb_h.p;
    ZKC_CUDA_TRY(ctx, cudaMemsetAsync(tmp, 0, 3 * n * sizeof(Fr), st));
    const Fr one = fe_one<FrP>();
    const uint32_t bf = cs.blinding_factors;
    std::vector<Fr> ones(bf, one);
    ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(tmp, &one, sizeof(Fr), cudaMemcpyHostToDevice, st));
    ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(tmp + n + (n - bf - 1), &one, sizeof(Fr), cudaMemcpyHostToDevice, st));
    ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(tmp + 2 * n + (n - bf), ones.data(), bf * sizeof(Fr), cudaMemcpyHostToDevice, st));
    ZKC_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    int s1 = dom_lagrange_to_coeff(ctx, P->dom, tmp, 3);
    if (s1 == ZKC_OK) s1 = dom_coeff_to_classes(ctx, P->dom, tmp, n, P->l0, 1, 0, ncls);
    if (s1 == ZKC_OK) s1 = dom_coeff_to_classes(ctx, P->dom, tmp + n, n, P->l_last, 1, 0, ncls);
    if (s1 == ZKC_OK) s1 = dom_coeff_to_classes(ctx, P->dom, tmp + 2 * n, n, lb, 1, 0, ncls);
    if (s1 == ZKC_OK) { k_l_active<<<(unsigned)(((uint64_t)ncls * n + 255) / 256), 256, 0, st>>>(P->l_last, lb, P->l_active, (uint64_t)ncls * n); ctx->launches++; }
    cudaStreamSynchronize(st);   // the temporaries are released when this block ends
    ZKC_TRY(s1);
  }
  // programs, query tables, pointer tables
  ZKC_TRY(upload_program(ctx, P, cs.gates, P->gates));
  P->lookups.resize(cs.lookups.size());
  for (size_t i = 0; i < cs.lookups.size(); ++i) {
    ZKC_TRY(upload_program(ctx, P, cs.lookups[i].first, P->lookups[i].first));
    ZKC_TRY(upload_program(ctx, P, cs.lookups[i].second, P->lookups[i].second));
  }
  {
    std::vector<uint32_t> q;
    for (auto* v : {&cs.aq, &cs.fq, &cs.iq}) {
      for (auto& e : *v) q.push_back(e.first);
      for (auto& e : *v) q.push_back((uint32_t)e.second);
    }
    ZKC_TRY(dev_alloc(ctx, P, &P->qtab, q.size()));
    if (!q.empty()) ZKC_CUDA_TRY(ctx, cudaMemcpy(P->qtab, q.data(), q.size() * 4, cudaMemcpyHostToDevice));
    std::vector<const Fr*> pv(F), pc(F);
    for (uint32_t c = 0; c < F; ++c) { pv[c] = P->fixed_values + (size_t)c * n; pc[c] = P->fixed_cosets + (size_t)c * en; }
    ZKC_TRY(dev_alloc(ctx, P, &P->fixed_val_ptrs, F)); ZKC_TRY(dev_alloc(ctx, P, &P->fixed_coset_ptrs, F));
    if (F) {
      ZKC_CUDA_TRY(ctx, cudaMemcpy(P->fixed_val_ptrs, pv.data(), F * sizeof(void*), cudaMemcpyHostToDevice));
      ZKC_CUDA_TRY(ctx, cudaMemcpy(P->fixed_coset_ptrs, pc.data(), F * sizeof(void*), cudaMemcpyHostToDevice));
    }
  }
  // vk commitments (keygen_vk: fixed and permutation columns are committed in the Lagrange basis)
  if (F) ZKC_TRY(commit_points(ctx, srs, 1, P->fixed_values, n, F, P->fixed_comm));
  if (S) ZKC_TRY(commit_points(ctx, srs, 1, P->sigma_values, n, S, P->sigma_comm));
  ZKC_CUDA_TRY(ctx, cudaStreamSynchronize(st));
  *out = pk.release();
  return ZKC_OK;
}
}  // namespace

// keygen_pk from the fixed columns and the copy constraints synthesis recorded (gen_pk: /root/reference/src/helpers.rs:213,265)
extern "C" int zkc_keygen_pk(zkc_ctx* ctx, const zkc_srs* srs, const uint8_t* cs_blob, size_t cs_len, const zkc_fr* fixed, const uint32_t* copies,
                             size_t num_copies, const zkc_fr* transcript_repr, int zeta_choice, zkc_pk** out) {
  if (!ctx || !srs || !cs_blob || !transcript_repr || !out || (num_copies && !copies)) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_keygen_pk: null argument");
  CtxLock lock(ctx);
  Cs cs;
  std::string perr;
  if (!zkc::host::parse_cs(cs_blob, cs_len, cs, perr)) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_keygen_pk: " + perr);
  std::vector<uint64_t> mapping(cs.perm.size() * cs.n());
  const int st = zkc_keygen_permutation_mapping(cs.k, (uint32_t)cs.perm.size(), copies, num_copies, mapping.data());
  if (st != ZKC_OK) return set_err(ctx, st, "zkc_keygen_pk: copy constraint outside the permutation columns / rows");
  ColumnSource f, s;
  f.contiguous = fixed; s.mapping = mapping.data();
  return pk_build(ctx, srs, cs_blob, cs_len, f, s, transcript_repr, zeta_choice, out);
}

extern "C" int zkc_pk_get_sigma(zkc_ctx* ctx, const zkc_pk* pk, zkc_fr* sigma_out) {
  if (!ctx || !pk || (!pk->cs.perm.empty() && !sigma_out)) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_pk_get_sigma: null argument");
  CtxLock lock(ctx);
  if (!pk->cs.perm.empty()) {
    ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(sigma_out, pk->sigma_values, pk->cs.perm.size() * pk->cs.n() * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    ZKC_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return ZKC_OK;
}

// ---- ProvingKey files (layout: csrc/host/keygen.cpp) --------------------------------------------------------------------------
extern "C" size_t zkc_pk_file_size(const zkc_pk* pk, uint32_t num_selectors) {
  zkc_pk_file_layout_t L;
  if (!pk || zkc_pk_file_layout(pk->cs.k, pk->ext_k, pk->cs.num_fixed, (uint32_t)pk->cs.perm.size(), num_selectors, &L) != ZKC_OK) return 0;
  return (size_t)L.total;
}

extern "C" int zkc_pk_write(zkc_ctx* ctx, const zkc_pk* pk, const uint8_t* selectors, uint32_t num_selectors, int be, uint8_t* out, size_t cap) {
  if (!ctx || !pk || !out || (num_selectors && !selectors)) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_pk_write: null argument");
  CtxLock lock(ctx);
  const Cs& cs = pk->cs;
  const uint32_t F = cs.num_fixed, S = (uint32_t)cs.perm.size();
  const uint64_t n = cs.n(), en = 1ull << pk->ext_k;
  zkc_pk_file_layout_t L;
  if (zkc_pk_file_layout(cs.k, pk->ext_k, F, S, num_selectors, &L) != ZKC_OK || cap < L.total) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_pk_write: buffer too small");
  ZKC_TRY(zkc_pk_file_write_headers(out, cap, cs.k, pk->ext_k, F, S, num_selectors, be));
  if (F) memcpy(out + L.fixed_commitments_off, pk->fixed_comm.data(), (size_t)F * 64);
  if (S) memcpy(out + L.perm_commitments_off, pk->sigma_comm.data(), (size_t)S * 64);
  if (num_selectors) memcpy(out + L.selectors_off, selectors, (size_t)num_selectors * ((n + 7) / 8));
  cudaStream_t st = ctx->stream;
  auto d2h = [&](uint64_t off, const Fr* src, uint64_t len) { return cudaMemcpyAsync(out + off, src, len * sizeof(Fr), cudaMemcpyDeviceToHost, st); };
  // the key keeps only the residue classes of its extended cosets that h(X) needs, class-major; the file holds whole cosets in
  // upstream's natural order, so they are re-derived here from the coefficient forms (off the proving path)
  Fr* nat;
  ZKC_TRY(scratch_reserve(ctx, SCR_MISC3, 2 * en * sizeof(Fr), (void**)&nat));
  auto coset_out = [&](uint64_t off, const Fr* poly) -> int {
    ZKC_TRY(dom_coeff_to_extended(ctx, pk->dom, poly, n, nat, 1));
    ZKC_CUDA_TRY(ctx, d2h(off, nat, en));
    ZKC_CUDA_TRY(ctx, cudaStreamSynchronize(st));   // `nat` is reused by the next column
    return ZKC_OK;
  };
  ZKC_TRY(coset_out(L.l0_off + 4, pk->l_polys)); ZKC_TRY(coset_out(L.l_last_off + 4, pk->l_polys + n));
  {
    ZKC_TRY(dom_coeff_to_extended(ctx, pk->dom, pk->l_polys + n, n, nat, 1));
    ZKC_TRY(dom_coeff_to_extended(ctx, pk->dom, pk->l_polys + 2 * n, n, nat + en, 1));
    k_l_active<<<(unsigned)((en + 255) / 256), 256, 0, st>>>(nat, nat + en, nat, en); ZKC_LAUNCH_CHECK(ctx);
    ZKC_CUDA_TRY(ctx, d2h(L.l_active_row_off + 4, nat, en));
    ZKC_CUDA_TRY(ctx, cudaStreamSynchronize(st));
  }
  auto slice = [&](uint64_t off, const Fr* base, uint32_t count, uint64_t len) -> int {
    for (uint32_t c = 0; c < count; ++c) ZKC_CUDA_TRY(ctx, d2h(off + 4 + (uint64_t)c * (4 + 32 * len) + 4, base + (size_t)c * len, len));
    return ZKC_OK;
  };
  auto cosets = [&](uint64_t off, const Fr* polys, uint32_t count) -> int {
    for (uint32_t c = 0; c < count; ++c) ZKC_TRY(coset_out(off + 4 + (uint64_t)c * (4 + 32 * en) + 4, polys + (size_t)c * n));
    return ZKC_OK;
  };
  ZKC_TRY(slice(L.fixed_values_off, pk->fixed_values, F, n)); ZKC_TRY(slice(L.fixed_polys_off, pk->fixed_polys, F, n)); ZKC_TRY(cosets(L.fixed_cosets_off, pk->fixed_polys, F));
  ZKC_TRY(slice(L.perm_values_off, pk->sigma_values, S, n)); ZKC_TRY(slice(L.perm_polys_off, pk->sigma_polys, S, n)); ZKC_TRY(cosets(L.perm_cosets_off, pk->sigma_polys, S));
  ZKC_CUDA_TRY(ctx, cudaStreamSynchronize(st));
  return ZKC_OK;
}

extern "C" int zkc_pk_read(zkc_ctx* ctx, const zkc_srs* srs, const uint8_t* cs_blob, size_t cs_len, const uint8_t* file, size_t file_len,
                           uint32_t num_selectors, int format, const zkc_fr* transcript_repr, int zeta_choice, zkc_pk** out) {
  if (!ctx || !srs || !cs_blob || !file || !transcript_repr || !out || format < 0 || format > 1) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_pk_read: bad argument");
  CtxLock lock(ctx);
  Cs cs;
  std::string perr;
  if (!zkc::host::parse_cs(cs_blob, cs_len, cs, perr)) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_pk_read: " + perr);
  const uint32_t F = cs.num_fixed, S = (uint32_t)cs.perm.size();
  const uint64_t n = cs.n();
  // extended_k follows from the constraint system's degree (EvaluationDomain::new)
  uint32_t ext_k = cs.k;
  while ((1ull << ext_k) < n * (uint64_t)(cs.degree - 1)) ++ext_k;
  int be = 1;
  if (zkc_pk_file_check(file, file_len, cs.k, ext_k, F, S, num_selectors, -1, &be) != ZKC_OK)
    return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_pk_read: the file does not have the shape of a ProvingKey for this constraint system");
  zkc_pk_file_layout_t L;
  zkc_pk_file_layout(cs.k, ext_k, F, S, num_selectors, &L);
  ColumnSource f, s;
  f.file = file; f.first_off = L.fixed_values_off + 8; f.col_stride = 4 + 32 * n;
  s.file = file; s.first_off = L.perm_values_off + 8; s.col_stride = 4 + 32 * n;
  if (format == 0) {   // SerdeFormat::RawBytes: every scalar canonical
    for (uint32_t c = 0; c < F; ++c) if (!zkc_fr_column_is_canonical((const zkc_fr*)(file + f.first_off + (uint64_t)c * f.col_stride), n)) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_pk_read: non-canonical field element in fixed_values");
    for (uint32_t c = 0; c < S; ++c) if (!zkc_fr_column_is_canonical((const zkc_fr*)(file + s.first_off + (uint64_t)c * s.col_stride), n)) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_pk_read: non-canonical field element in the permutation columns");
  }
  std::unique_ptr<zkc_pk, void (*)(zkc_pk*)> pk(nullptr, zkc_pk_free);
  {
    zkc_pk* raw = nullptr;
    ZKC_TRY(pk_build(ctx, srs, cs_blob, cs_len, f, s, transcript_repr, zeta_choice, &raw));
    pk.reset(raw);
  }
  if (format == 0) {   // the commitments in the file must be the commitments of the columns in the file
    if ((F && memcmp(file + L.fixed_commitments_off, pk->fixed_comm.data(), (size_t)F * 64)) || (S && memcmp(file + L.perm_commitments_off, pk->sigma_comm.data(), (size_t)S * 64)))
      return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_pk_read: the verifying key's commitments do not match the columns");
  }
  *out = pk.release();
  return ZKC_OK;
}

extern "C" int zkc_pk_get_commitments(zkc_ctx* ctx, const zkc_pk* pk, zkc_g1_affine* fixed_out, zkc_g1_affine* sigma_out) {
  if (!ctx || !pk) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_pk_get_commitments: null argument");
  if (fixed_out && !pk->fixed_comm.empty()) memcpy(fixed_out, pk->fixed_comm.data(), pk->fixed_comm.size() * sizeof(G1Affine));
  if (sigma_out && !pk->sigma_comm.empty()) memcpy(sigma_out, pk->sigma_comm.data(), pk->sigma_comm.size() * sizeof(G1Affine));
  return ZKC_OK;
}

extern "C" int zkc_pk_info(const zkc_pk* pk, uint32_t* out /* k, extended_k, degree, blinding_factors, nsets, nlookups, nfixed, nperm */) {
  if (!pk || !out) return ZKC_ERR_BAD_ARG;
  out[0] = pk->cs.k; out[1] = pk->ext_k; out[2] = pk->cs.degree; out[3] = pk->cs.blinding_factors; out[4] = pk->cs.nsets();
  out[5] = (uint32_t)pk->cs.lookups.size(); out[6] = pk->cs.num_fixed; out[7] = (uint32_t)pk->cs.perm.size();
  return ZKC_OK;
}

// ---- create_proof ----------------------------------------------------------------------------------------------
namespace {

// stream-ordered temporary allocations for one proof
struct Pool {
  zkc_ctx* ctx;
  std::vector<void*> ptrs;
  explicit Pool(zkc_ctx* c) : ctx(c) {}
  ~Pool() {
    // error paths may leave side-stream work in flight: drain it before the buffers go back to the pool
    if (ctx->side_pending) { cudaStreamSynchronize(ctx->side_stream); ctx->side_pending = false; }
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);   // a staged upload may still be writing into a pool buffer
    for (void* p : ptrs) cudaFreeAsync(p, ctx->stream);
  }
  template <class T> int get(T** out, size_t count) {
    void* p = nullptr;
    cudaError_t e = cudaMallocAsync(&p, std::max<size_t>(count, 1) * sizeof(T), ctx->stream);
    if (e != cudaSuccess) return set_err(ctx, ZKC_ERR_OOM, std::string("cudaMallocAsync: ") + cudaGetErrorString(e));
    ptrs.push_back(p);
    *out = (T*)p;
    return ZKC_OK;
  }
};

// The vanishing argument's random polynomial, described so that the device can produce it (zkc_random_poly, validated).
struct RandomSpec {
  int kind = -1;                 // -1 = not given yet
  const Fr* scalars = nullptr;   // kind 0: n host scalars
  ChaChaKey key; int double_rounds = 10; uint64_t first_word = 0;   // kind 1
  std::vector<ChaChaKey> chunk_keys; uint64_t chunk_len = 0;        // kind 2
};

int random_spec_from_abi(zkc_ctx* ctx, const zkc_random_poly* r, uint64_t n, RandomSpec* out) {
  if (r->kind == 0) {
    if (!r->scalars) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_random_poly: kind 0 without scalars");
    out->scalars = (const Fr*)r->scalars;
  } else if (r->kind == 1) {
    if (r->rng_kind < 0 || r->rng_kind > 1) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_random_poly: unknown rng_kind");
    memcpy(out->key.k, r->seed, 32); out->double_rounds = r->rng_kind == 1 ? 6 : 10; out->first_word = r->first_word;
  } else if (r->kind == 2) {
    if (!r->seeds || !r->nseeds || !r->chunk_len || (n + r->chunk_len - 1) / r->chunk_len != r->nseeds)
      return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_random_poly: the chunks do not cover the polynomial exactly");
    out->chunk_keys.resize(r->nseeds);
    for (uint32_t j = 0; j < r->nseeds; ++j) memcpy(out->chunk_keys[j].k, r->seeds + 32 * (size_t)j, 32);
    out->chunk_len = r->chunk_len;
  } else {
    return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_random_poly: unknown kind");
  }
  out->kind = r->kind;
  return ZKC_OK;
}

struct Query { const Fr* poly; Fr point; Fr eval; };

struct RotationSet {
  std::vector<Fr> points;                       // ascending canonical order (BTreeSet<Fr>)
  std::vector<const Fr*> polys;                 // first-appearance order
  std::vector<std::vector<Fr>> evals;           // per poly, in point order
};

void build_rotation_sets(const std::vector<Query>& queries, std::vector<RotationSet>& sets, std::vector<Fr>& super_points) {
  auto less = [](const Fr& a, const Fr& b) { return host::fr_cmp_canonical(a, b) < 0; };
  auto same = [](const Fr& a, const Fr& b) { return fe_eq(a, b); };
  auto insert_sorted = [&](std::vector<Fr>& v, const Fr& p) {
    for (auto& e : v) if (same(e, p)) return;
    v.insert(std::upper_bound(v.begin(), v.end(), p, less), p);
  };
  std::vector<std::pair<const Fr*, std::vector<Fr>>> poly_points;
  for (auto& q : queries) {
    insert_sorted(super_points, q.point);
    bool found = false;
    for (auto& e : poly_points) if (e.first == q.poly) { insert_sorted(e.second, q.point); found = true; break; }
    if (!found) poly_points.push_back({q.poly, std::vector<Fr>{q.point}});
  }
  for (auto& pp : poly_points) {
    RotationSet* tgt = nullptr;
    for (auto& s : sets) {
      if (s.points.size() != pp.second.size()) continue;
      bool eq = true;
      for (size_t i = 0; i < s.points.size() && eq; ++i) eq = same(s.points[i], pp.second[i]);
      if (eq) { tgt = &s; break; }
    }
    if (!tgt) { sets.push_back(RotationSet()); tgt = &sets.back(); tgt->points = pp.second; }
    tgt->polys.push_back(pp.first);
    std::vector<Fr> ev;
    for (auto& pt : tgt->points)
      for (auto& q : queries) if (q.poly == pp.first && same(q.point, pt)) { ev.push_back(q.eval); break; }
    tgt->evals.push_back(ev);
  }
}

unsigned grid_for(uint64_t cnt, unsigned b) { return (unsigned)((cnt + b - 1) / b); }

enum { ST_LOOKUPS = 1, ST_PRODUCTS, ST_VANISHING, ST_QUOTIENT, ST_EVALS, ST_OPEN, ST_SHPLONK_W, ST_DONE, ST_FAILED };

}  // namespace

// One create_proof in flight: the device-resident state between prover rounds.  Every round is a method; the C entry
// points below (step API) and zkc_prove (driver with its own transcript + RNG) call the same methods.
struct zkc_prover {
  zkc_ctx* ctx;
  const zkc_pk* pk;
  Pool pool;
  int stage = 0;
  // shape
  uint64_t n = 0, en = 0, hn = 0, U = 0;   // hn: rows of the class-major extended coset that are evaluated (degree - 1 classes)
  uint32_t bf = 0, A = 0, I = 0, L = 0, Pn = 0, q = 0;
  bool team = false;
  std::vector<Segment> my_rows;        // row blocks of the (class-major) extended coset this process evaluates
  std::vector<std::pair<uint32_t, uint32_t>> my_classes;   // residue classes [first, second) those blocks touch
  uint64_t halo_lo = 0, halo_hi = 0;   // rotation reach inside a class: rows read before / after a row
  Fr omega, omega_inv, zeta, ONE, ZERO, DELTA;
  // columns
  Fr *inst_values = nullptr, *inst_polys = nullptr, *adv_values = nullptr, *adv_polys = nullptr;
  Fr *adv_cosets = nullptr, *inst_cosets = nullptr, *pz_cosets = nullptr, *lk_cosets = nullptr;
  Fr *lk_comp = nullptr, *lk_perm = nullptr, *lk_perm_polys = nullptr, *z_all = nullptr, *z_all_polys = nullptr;
  Fr *random_poly = nullptr, *hval = nullptr, *lk_comp_cosets = nullptr, *h_poly = nullptr;
  Fr *pz = nullptr, *pz_polys = nullptr, *lk_z = nullptr, *lk_z_polys = nullptr;
  const Fr** ptrs_dev = nullptr;
  DevQueries qlag, qext;
  MsmPending random_commit;
  bool random_enqueued = false;
  Fr theta, beta, gamma, y, x;
  // evaluations and opening queries
  std::vector<const Fr*> ev_polys; std::vector<Fr> ev_points, ev;
  std::vector<size_t> i_adv, i_fix, i_sig, i_pz, i_lk;
  size_t i_rand = 0, i_h = 0;
  std::vector<Query> queries;
  // multiopen
  Fr *acc = nullptr, *tmp1 = nullptr, *tmp2 = nullptr, *tmp3 = nullptr, *kd_exchange = nullptr;
  bool sliced = false;
  std::vector<std::pair<int, Segment>> my_coeffs;   // (team rank, coefficient slice)
  std::vector<RotationSet> sets; std::vector<Fr> super_points;
  std::vector<std::vector<std::vector<Fr>>> rcoef;   // [set][poly] -> r(X) coefficients
  Fr sh_y, sh_v;

  zkc_prover(zkc_ctx* c, const zkc_pk* p) : ctx(c), pk(p), pool(c) {}
  ~zkc_prover() {
    if (random_commit.active) { cudaEventSynchronize(random_commit.done); random_commit.active = false; }   // never leave a batch pending
    if (ctx->active_prover == this) ctx->active_prover = nullptr;
  }

  int commit(int basis, const Fr* polys, uint32_t ncols, std::vector<G1Affine>& out) { return commit_points(ctx, pk->srs, basis, polys, n, ncols, out); }

  // values -> coefficient form, `ncols` columns in place; team: the owner of a column transforms it and broadcasts the result
  int to_coeff(Fr* polys, uint32_t ncols) {
    if (!team) return dom_lagrange_to_coeff(ctx, pk->dom, polys, ncols);
    for (int r : team_ranks(ctx)) {
      uint32_t c0, c1;
      team_cols(ctx, ncols, r, &c0, &c1);
      if (c1 > c0) ZKC_TRY(dom_lagrange_to_coeff(ctx, pk->dom, polys + (size_t)c0 * n, c1 - c0));
    }
    ZKC_TRY(team_bcast_cols(ctx, polys, n, n, ncols));
    team_advance(ctx, ncols);
    return ZKC_OK;
  }
  // coefficient form -> extended coset, CLASS-MAJOR (ntt.cu: residue classes of the extended coset).  Team: every rank holds the
  // coefficient forms, so it transforms the classes its row block touches itself; coset rows cross a link only where two row
  // blocks share a class.
  int to_extended(const Fr* polys, Fr* cosets, uint32_t ncols) {
    if (!team || ctx->team_emulate) {
      for (const auto& cr : my_classes) ZKC_TRY(dom_coeff_to_classes(ctx, pk->dom, polys, n, cosets, ncols, cr.first, cr.second));
      return ZKC_OK;
    }
    // A class that lies inside one rank's row block is transformed there, all columns.  Where g > 1 row blocks share a class
    // (more ranks than classes, or a world size that does not divide them) those g ranks deal its columns among themselves,
    // each transforms its share and hands the rows the other g - 1 evaluate (+ rotation reach) to them: large point-to-point
    // messages inside a small group instead of g-fold redundant transforms.
    const int W = ctx->team_world, me = ctx->team_rank;
    // what travels: rows [p, p + len) of `cols` consecutive columns (column stride en), per peer and direction, in an order
    // both sides derive alike (class, then segment)
    struct Piece { Fr* p; uint64_t len; uint32_t cols; };
    std::vector<std::vector<Piece>> snd(W), rcv(W);
    for (const auto& cr : my_classes)
      for (uint32_t c = cr.first; c < cr.second; ++c) {
        std::vector<int> mem;
        std::vector<uint64_t> cum(1, 0);   // rows of class c owned by the members before member t
        for (int r = 0; r < W; ++r) {
          uint64_t lo, hi;
          shard_range(hn, W, r, &lo, &hi);
          if (hi > lo && lo < (uint64_t)(c + 1) * n && hi > (uint64_t)c * n) {
            mem.push_back(r);
            cum.push_back(cum.back() + std::min<uint64_t>(hi, (uint64_t)(c + 1) * n) - std::max<uint64_t>(lo, (uint64_t)c * n));
          }
        }
        const int g = (int)mem.size();
        // rows of class c that member o evaluates, widened by the rotation reach (cyclic inside the class): <= 2 segments
        auto rows_of = [&](int o, Segment seg[2]) -> int {
          const uint64_t ra = cum[o], len = (cum[o + 1] - cum[o]) + halo_lo + halo_hi;
          if (len >= n) { seg[0] = {0, n}; return 1; }
          const uint64_t start = (ra + n - halo_lo % n) % n;
          if (start + len <= n) { seg[0] = {start, len}; return 1; }
          seg[0] = {0, start + len - n}; seg[1] = {start, n - start};
          return 2;
        };
        int self = 0;
        while (mem[self] != me) ++self;
        // columns in proportion to the rows each member owns (so every rank transforms its fair share of class blocks whatever
        // way the row blocks cut the classes); the offset moves single-column batches from member to member
        const uint64_t off = ((uint64_t)ctx->team_rot * n / (uint64_t)W) % n;
        for (int t = 0; t < g; ++t) {
          const uint64_t a = ((uint64_t)ncols * cum[t] + (t ? off : 0)) / n, b = t + 1 == g ? ncols : ((uint64_t)ncols * cum[t + 1] + off) / n;
          if (b <= a) continue;
          Fr* blk = cosets + a * en + (uint64_t)c * n;
          Segment seg[2];
          if (mem[t] == me) {
            ZKC_TRY(dom_coeff_to_classes(ctx, pk->dom, polys + a * n, n, cosets + a * en, (uint32_t)(b - a), c, c + 1));
            for (int o = 0; o < g; ++o) {
              if (o == self) continue;
              const int ns = rows_of(o, seg);
              for (int q2 = 0; q2 < ns; ++q2) snd[mem[o]].push_back({blk + seg[q2].lo, seg[q2].len, (uint32_t)(b - a)});
            }
          } else {
            const int ns = rows_of(self, seg);
            for (int q2 = 0; q2 < ns; ++q2) rcv[mem[t]].push_back({blk + seg[q2].lo, seg[q2].len, (uint32_t)(b - a)});
          }
        }
      }
    // One message per peer and direction: the pieces are packed into a staging buffer with strided device copies, because one
    // large message moves at about twice the rate of the same bytes in column-sized ones (profiles/r02_p2p_2gpu.json).
    auto volume = [](const std::vector<std::vector<Piece>>& v) { uint64_t e = 0; for (auto& l : v) for (auto& pc : l) e += pc.len * pc.cols; return e; };
    const uint64_t nsend = volume(snd), nrecv = volume(rcv);
    if (nsend + nrecv) {
      ProfScope _p(ctx, "team.class_exchange");
      Fr *ssend = nullptr, *srecv = nullptr;
      ZKC_TRY(scratch_reserve(ctx, SCR_HOSTIO, std::max<uint64_t>(nsend, 1) * sizeof(Fr), (void**)&ssend));
      ZKC_TRY(scratch_reserve(ctx, SCR_HOSTIO2, std::max<uint64_t>(nrecv, 1) * sizeof(Fr), (void**)&srecv));
      cudaStream_t st = ctx->stream;
      auto strided = [&](Fr* packed, Fr* rows, const Piece& pc, bool pack) -> int {
        if (en * sizeof(Fr) < ((size_t)1 << 31)) {   // cudaMemcpy2D pitch limit (cudaDeviceProp::memPitch)
          ZKC_CUDA_TRY(ctx, pack ? cudaMemcpy2DAsync(packed, pc.len * sizeof(Fr), rows, en * sizeof(Fr), pc.len * sizeof(Fr), pc.cols, cudaMemcpyDeviceToDevice, st)
                                 : cudaMemcpy2DAsync(rows, en * sizeof(Fr), packed, pc.len * sizeof(Fr), pc.len * sizeof(Fr), pc.cols, cudaMemcpyDeviceToDevice, st));
        } else {
          for (uint32_t cc = 0; cc < pc.cols; ++cc)
            ZKC_CUDA_TRY(ctx, pack ? cudaMemcpyAsync(packed + cc * pc.len, rows + (uint64_t)cc * en, pc.len * sizeof(Fr), cudaMemcpyDeviceToDevice, st)
                                   : cudaMemcpyAsync(rows + (uint64_t)cc * en, packed + cc * pc.len, pc.len * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
        }
        return ZKC_OK;
      };
      std::vector<TeamXfer> xf;
      uint64_t so = 0, ro = 0;
      for (int r = 0; r < W; ++r) {
        const uint64_t s0 = so, r0 = ro;
        for (const Piece& pc : snd[r]) { ZKC_TRY(strided(ssend + so, pc.p, pc, true)); so += pc.len * pc.cols; }
        for (const Piece& pc : rcv[r]) ro += pc.len * pc.cols;
        if (so > s0) xf.push_back({r, true, ssend + s0, (so - s0) * sizeof(Fr)});
        if (ro > r0) xf.push_back({r, false, srecv + r0, (ro - r0) * sizeof(Fr)});
      }
      ZKC_TRY(team_exchange(ctx, xf, "team.class_exchange.nccl"));
      ro = 0;
      for (int r = 0; r < W; ++r)
        for (const Piece& pc : rcv[r]) { ZKC_TRY(strided(srecv + ro, pc.p, pc, false)); ro += pc.len * pc.cols; }
    }
    team_advance(ctx, ncols);
    return ZKC_OK;
  }
  const Fr* column_ptr(uint32_t kind, uint32_t idx, bool coset) const {
    if (kind == 0) return coset ? adv_cosets + (size_t)idx * en : adv_values + (size_t)idx * n;
    if (kind == 1) return coset ? pk->fixed_cosets + (size_t)idx * en : pk->fixed_values + (size_t)idx * n;
    return coset ? inst_cosets + (size_t)idx * en : inst_values + (size_t)idx * n;
  }
  PermSetArgs perm_args(uint32_t set, bool coset) const {
    const Cs& cs = pk->cs;
    PermSetArgs a;
    a.m = 0;
    Fr db = coset ? fe_mul(beta, zeta) : beta;
    for (uint32_t g = 0; g < set * cs.chunk_len; ++g) db = fe_mul(db, DELTA);
    for (uint32_t t = 0; t < cs.chunk_len && set * cs.chunk_len + t < cs.perm.size(); ++t) {
      const uint32_t g = set * cs.chunk_len + t;
      a.cols[t] = column_ptr(cs.perm[g].first, cs.perm[g].second, coset);
      a.sigmas[t] = coset ? pk->sigma_cosets + (size_t)g * en : pk->sigma_values + (size_t)g * n;
      a.delta_beta[t] = db;
      db = fe_mul(db, DELTA);
      a.m++;
    }
    return a;
  }
  int enqueue_random(const RandomSpec& rs);

  int begin(const zkc_fr* advice, int advice_on_device, const zkc_fr* const* instances, const size_t* instance_lens, const Fr* advice_tails,
            const RandomSpec* early_random, std::vector<G1Affine>& out);
  int lookups(const Fr& theta_, const Fr* tails, int lookup_fill, std::vector<G1Affine>& out);
  int products(const Fr& beta_, const Fr& gamma_, const Fr* tails, std::vector<G1Affine>& out);
  int vanishing(const RandomSpec* rs, G1Affine* out);
  int quotient(const Fr& y_, std::vector<G1Affine>& out);
  int evals(const Fr& x_, std::vector<Fr>& out);
  int open_prepare();
  int shplonk_h(const Fr& yy, const Fr& v, G1Affine* out);
  int shplonk_w(const Fr& u, G1Affine* out);
  int gwc(const Fr& v, std::vector<G1Affine>& out);
  // multiopen helpers (team: coefficient slices)
  int lincomb_s(Fr* out, const std::vector<const Fr*>& ps, const std::vector<Fr>& cf);
  int sub_low_s(Fr* a, const std::vector<Fr>& low);
  int scale_s(Fr* a, const Fr& f);
  int kate_s(const std::vector<Fr*>& ps, const std::vector<Fr>& roots, Fr* tmp);
};

// random polynomial on the side stream: generated (or uploaded) and committed; the point is collected in vanishing()
int zkc_prover::enqueue_random(const RandomSpec& rs) {
  SideScope side(ctx);
  cudaStream_t s = ctx->stream;
  if (rs.kind == 1 && (rs.first_word & 7) == 0) {
    k_chacha_fr<<<grid_for(n, 128), 128, 0, s>>>(random_poly, rs.key, rs.first_word, n, rs.double_rounds);
    ZKC_LAUNCH_CHECK(ctx);
  } else if (rs.kind == 1) {
    // a run that starts inside a 64-bit word pair never arises from Fr::random / fill_bytes(32) draws; serve it from the host generator
    uint8_t seed[32]; memcpy(seed, rs.key.k, 32);
    host::ChaCha20Rng g(seed, rs.double_rounds);
    g.seek(rs.first_word);
    std::vector<Fr> v(n);
    for (auto& e : v) e = g.fr_random();
    ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(random_poly, v.data(), n * sizeof(Fr), cudaMemcpyHostToDevice, s));
    ZKC_CUDA_TRY(ctx, cudaStreamSynchronize(s));
  } else if (rs.kind == 2) {
    ChaChaKey* keys;
    ZKC_TRY(scratch_reserve(ctx, SCR_MISC3, rs.chunk_keys.size() * sizeof(ChaChaKey), (void**)&keys));
    ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(keys, rs.chunk_keys.data(), rs.chunk_keys.size() * sizeof(ChaChaKey), cudaMemcpyHostToDevice, s));
    ZKC_CUDA_TRY(ctx, cudaStreamSynchronize(s));   // rs may not outlive this call
    k_chacha_fr_chunked<<<grid_for(n, 128), 128, 0, s>>>(random_poly, keys, rs.chunk_len, n);
    ZKC_LAUNCH_CHECK(ctx);
  } else {
    ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(random_poly, rs.scalars, n * sizeof(Fr), cudaMemcpyHostToDevice, s));
    ZKC_CUDA_TRY(ctx, cudaStreamSynchronize(s));   // the caller's buffer is only borrowed for this call
  }
  ZKC_TRY(srs_commit_enqueue(ctx, pk->srs, 0, random_poly, n, &random_commit));
  random_enqueued = true;
  return ZKC_OK;
}

// 1 + 2. instances and advice: upload, blinding policy, commit (Lagrange basis); coefficient forms and cosets on the side stream
int zkc_prover::begin(const zkc_fr* advice, int advice_on_device, const zkc_fr* const* instances, const size_t* instance_lens,
                      const Fr* advice_tails, const RandomSpec* early_random, std::vector<G1Affine>& out) {
  const Cs& cs = pk->cs;
  n = cs.n(); en = 1ull << pk->ext_k; U = cs.usable();
  bf = cs.blinding_factors; A = cs.num_advice; I = cs.num_instance; L = (uint32_t)cs.lookups.size(); Pn = cs.nsets();
  q = cs.degree - 1;
  cudaStream_t st = ctx->stream;
  // Team proving (dist.cuh): the same driver runs on every rank; MSMs split by point range, column transforms by column,
  // h(X) by extended-row block.  Collectives are issued on the stream of the kernels they depend on (one communicator per
  // stream), in the same host order on every rank, so the column exchanges of the side stream hide under the MSM phases.
  team = team_active(ctx);
  if (team) ctx->team_rot = 0;
  {
    int64_t rmin = -(int64_t)(bf + 1), rmax = 1;   // z(omega X), z(omega^-(bf+1) X), a'(omega^-1 X)
    for (auto* v : {&cs.aq, &cs.fq, &cs.iq}) for (auto& qq : *v) { rmin = std::min<int64_t>(rmin, qq.second); rmax = std::max<int64_t>(rmax, qq.second); }
    halo_lo = (uint64_t)(-rmin); halo_hi = (uint64_t)rmax;
  }
  // h(X) is evaluated on the class-major extended coset: a rank's row block [lo, hi) of the flat class-major index touches the
  // classes lo / n .. (hi - 1) / n, and every rotation of a row stays inside its class
  // Only q = degree - 1 of the 2^e classes are evaluated: h has fewer than q n coefficients, so q classes determine it
  // (ntt.cu, dom_classes_to_pieces) — a quarter less extended-domain work at degree 4, half at degree 5.
  hn = (uint64_t)q * n;
  if (!team) my_rows.push_back({0, hn});
  else for (int r : team_ranks(ctx)) { uint64_t lo, hi; shard_range(hn, ctx->team_world, r, &lo, &hi); if (hi > lo) my_rows.push_back({lo, hi - lo}); }
  for (const Segment& rb : my_rows) {
    const uint32_t c0 = (uint32_t)(rb.lo / n), c1 = (uint32_t)((rb.lo + rb.len - 1) / n) + 1;
    if (!my_classes.empty() && my_classes.back().second >= c0) my_classes.back().second = std::max(my_classes.back().second, c1);
    else my_classes.push_back({c0, c1});
  }
  zkc_domain_info di;
  zkc_domain_get_info(pk->dom, &di);
  memcpy(omega.v, &di.omega, 32); memcpy(omega_inv.v, &di.omega_inv, 32); memcpy(zeta.v, &di.g_coset, 32);
  ONE = fe_one<FrP>(); ZERO = fe_zero<FrP>(); DELTA = fr_from_raw_words(FR_DELTA_RAW);

  // instances: Lagrange columns zero-padded (KZG: the values themselves are absorbed by the caller's transcript)
  ZKC_TRY(pool.get(&inst_values, (size_t)I * n)); ZKC_TRY(pool.get(&inst_polys, (size_t)I * n));
  if (I) ZKC_CUDA_TRY(ctx, cudaMemsetAsync(inst_values, 0, (size_t)I * n * sizeof(Fr), st));
  for (uint32_t c = 0; c < I; ++c) {
    if (instance_lens[c] > U) return set_err(ctx, ZKC_ERR_INVALID_INSTANCES, "instance column longer than the usable rows");
    if (instance_lens[c])
      ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(inst_values + (size_t)c * n, instances[c], instance_lens[c] * sizeof(Fr), cudaMemcpyHostToDevice, st));
  }
  if (I) {
    ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(inst_polys, inst_values, (size_t)I * n * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
    ZKC_TRY(dom_lagrange_to_coeff(ctx, pk->dom, inst_polys, I));
  }

  // advice upload
  ZKC_TRY(pool.get(&adv_values, (size_t)A * n)); ZKC_TRY(pool.get(&adv_polys, (size_t)A * n));
  std::vector<std::pair<uint32_t, uint32_t>> staged;   // column groups of a staged upload, in flight on the copy stream
  if (A) {
    ProfScope _p(ctx, "prove.advice_h2d");
    const size_t cells = (size_t)A * n;
    // The staged path copies `keep` rows of every column with a 2-D copy (pitch = n * 32 bytes <= cudaDeviceProp::memPitch);
    // it overlaps the first commitments only for PINNED host memory (cudaMemcpy2DAsync from pageable memory blocks the host).
    const bool pitch_ok = n * sizeof(Fr) <= ((size_t)1 << 30);
    if (team && !advice_on_device && cells % (size_t)ctx->team_world == 0) {
      // every rank holds the same host witness: each uploads 1/world of it over its own PCIe link and the shares are
      // all-gathered over NVLink (world x less host-to-device traffic per rank)
      const size_t per = cells / (size_t)ctx->team_world;
      for (int r : team_ranks(ctx))
        ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(adv_values + (size_t)r * per, (const Fr*)advice + (size_t)r * per, per * sizeof(Fr), cudaMemcpyHostToDevice, st));
      ZKC_TRY(team_allgather(ctx, adv_values, per * sizeof(Fr)));
    } else if (!team && !advice_on_device && ctx->overlap && pitch_ok && cells * sizeof(Fr) >= ctx->tune.stage_min_bytes) {
      // Staged upload of a large host witness: groups of columns travel on a copy stream while the random polynomial is
      // generated and committed and the first groups are already being committed.  The copies skip the rows the blinding
      // policy overwrites, so they need no ordering against those writes.
      if (!ctx->copy_stream) ZKC_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
      const uint32_t ngroups = (uint32_t)std::min<size_t>(std::min<size_t>(A, 6), std::max<size_t>(1, cells * sizeof(Fr) >> 27));   // >= 128 MB each
      while (ctx->ev_copy.size() < ngroups) { cudaEvent_t e; ZKC_CUDA_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); ctx->ev_copy.push_back(e); }
      ZKC_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_fork, st));                       // adv_values is allocated in stream order on `st`
      ZKC_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_fork, 0));
      const uint64_t keep = advice_tails ? U : n - 1;                             // rows taken from the host
      for (uint32_t g = 0; g < ngroups; ++g) {
        uint64_t c0, c1;
        shard_range(A, (int)ngroups, (int)g, &c0, &c1);
        ZKC_CUDA_TRY(ctx, cudaMemcpy2DAsync(adv_values + c0 * n, n * sizeof(Fr), (const Fr*)advice + c0 * n, n * sizeof(Fr), keep * sizeof(Fr), c1 - c0,
                                            cudaMemcpyHostToDevice, ctx->copy_stream));
        ZKC_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_copy[g], ctx->copy_stream));
        staged.push_back({(uint32_t)c0, (uint32_t)c1});
      }
    } else {
      ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(adv_values, advice, cells * sizeof(Fr), advice_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    }
  }
  // blinding policy (SURVEY OPEN-1): axiom sets the last row to 1 and draws nothing; PSE fills the unusable rows with the caller's draws
  if (A && !advice_tails) {
    k_fill_rows<<<grid_for(A, 128), 128, 0, st>>>(adv_values, n, n - 1, 1, A, nullptr, ONE); ZKC_LAUNCH_CHECK(ctx);
  } else if (A) {
    Fr* tails_dev;
    ZKC_TRY(pool.get(&tails_dev, (size_t)A * (bf + 1)));
    ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(tails_dev, advice_tails, (size_t)A * (bf + 1) * sizeof(Fr), cudaMemcpyHostToDevice, st));
    k_fill_rows<<<grid_for((uint64_t)A * (bf + 1), 128), 128, 0, st>>>(adv_values, n, U, bf + 1, A, tails_dev, ONE); ZKC_LAUNCH_CHECK(ctx);
    ZKC_CUDA_TRY(ctx, cudaStreamSynchronize(st));   // the caller's tails are only borrowed for this call
  }
  // every buffer of the first rounds is allocated up front on the main stream, so side-stream work never sees an
  // allocation made after its fork point
  ZKC_TRY(pool.get(&ptrs_dev, (size_t)2 * (A + I) + 1));
  ZKC_TRY(pool.get(&adv_cosets, (size_t)A * en)); ZKC_TRY(pool.get(&inst_cosets, (size_t)I * en));
  ZKC_TRY(pool.get(&pz_cosets, (size_t)Pn * en)); ZKC_TRY(pool.get(&lk_cosets, (size_t)3 * L * en));
  if (team && ctx->tune.team_poison) {   // testing: rows a rank never receives must never be read
    ZKC_CUDA_TRY(ctx, cudaMemsetAsync(adv_cosets, 0xff, (size_t)A * en * sizeof(Fr), st));
    ZKC_CUDA_TRY(ctx, cudaMemsetAsync(pz_cosets, 0xff, (size_t)Pn * en * sizeof(Fr), st));
    ZKC_CUDA_TRY(ctx, cudaMemsetAsync(lk_cosets, 0xff, (size_t)3 * L * en * sizeof(Fr), st));
  }
  ZKC_TRY(pool.get(&lk_comp, (size_t)2 * L * n)); ZKC_TRY(pool.get(&lk_perm, (size_t)2 * L * n));
  ZKC_TRY(pool.get(&lk_perm_polys, (size_t)2 * L * n));
  ZKC_TRY(pool.get(&z_all, (size_t)(Pn + L) * n)); ZKC_TRY(pool.get(&z_all_polys, (size_t)(Pn + L) * n));
  pz = z_all; pz_polys = z_all_polys; lk_z = z_all + (size_t)Pn * n; lk_z_polys = z_all_polys + (size_t)Pn * n;
  ZKC_TRY(pool.get(&random_poly, n));
  // The vanishing argument's random polynomial depends on nothing but the RNG: when the caller can describe it now, the
  // polynomial and its commitment are produced first, on the side stream; the point is collected in round 4.
  if (early_random) ZKC_TRY(enqueue_random(*early_random));
  if (A + I) {
    // off the Fiat-Shamir critical path: coefficient forms and extended cosets of advice / instance columns are
    // not needed before h(X), so they run on the side stream underneath the latency-bound MSM phases
    SideScope side(ctx);
    if (!staged.empty()) ZKC_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[staged.size() - 1], 0));   // whole witness resident
    if (A) {
      ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(adv_polys, adv_values, (size_t)A * n * sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
      ZKC_TRY(to_coeff(adv_polys, A));
      ZKC_TRY(to_extended(adv_polys, adv_cosets, A));
    }
    if (I) ZKC_TRY(to_extended(inst_polys, inst_cosets, I));
  }
  out.clear();
  if (A && staged.empty()) {
    ZKC_TRY(commit(1, adv_values, A, out));
  } else if (A) {
    std::vector<G1Affine> part;
    for (size_t g = 0; g < staged.size(); ++g) {   // commit each group as soon as it has landed
      ZKC_CUDA_TRY(ctx, cudaStreamWaitEvent(st, ctx->ev_copy[g], 0));
      ZKC_TRY(commit_points(ctx, pk->srs, 1, adv_values + (size_t)staged[g].first * n, n, staged[g].second - staged[g].first, part));
      out.insert(out.end(), part.begin(), part.end());
    }
  }
  {
    std::vector<const Fr*> h(2 * (A + I) + 1, nullptr);
    for (uint32_t c = 0; c < A; ++c) { h[c] = adv_values + (size_t)c * n; h[A + I + c] = adv_cosets + (size_t)c * en; }
    for (uint32_t c = 0; c < I; ++c) { h[A + c] = inst_values + (size_t)c * n; h[2 * A + I + c] = inst_cosets + (size_t)c * en; }
    ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(ptrs_dev, h.data(), h.size() * sizeof(void*), cudaMemcpyHostToDevice, st));
    ZKC_CUDA_TRY(ctx, cudaStreamSynchronize(st));
  }
  {
    const uint32_t* qt = pk->qtab;
    const size_t na = cs.aq.size(), nf = cs.fq.size(), ni = cs.iq.size();
    qlag.aq_col = qt; qlag.aq_rot = (const int32_t*)(qt + na);
    qlag.fq_col = qt + 2 * na; qlag.fq_rot = (const int32_t*)(qt + 2 * na + nf);
    qlag.iq_col = qt + 2 * na + 2 * nf; qlag.iq_rot = (const int32_t*)(qt + 2 * na + 2 * nf + ni);
    qext = qlag;
    qlag.advice = ptrs_dev; qlag.instance = ptrs_dev + A; qlag.fixed = pk->fixed_val_ptrs;
    qext.advice = ptrs_dev + A + I; qext.instance = ptrs_dev + 2 * A + I; qext.fixed = pk->fixed_coset_ptrs;
  }
  stage = ST_LOOKUPS;
  return ZKC_OK;
}

// 4. lookups: theta-compression, permute_expression_pair, commit A', S'
//    per lookup: comp (2 x n: input, table), perm values (2 x n: A', S'), perm polys (2 x n), z values / poly
int zkc_prover::lookups(const Fr& theta_, const Fr* tails_host, int lookup_fill, std::vector<G1Affine>& out) {
  theta = theta_;
  out.clear();
  cudaStream_t st = ctx->stream;
  if (L) {
    Fr *ca, *ct, *tails;
    uint32_t *flags, *ranks, *replist, *counts;
    ZKC_TRY(pool.get(&ca, n)); ZKC_TRY(pool.get(&ct, n)); ZKC_TRY(pool.get(&tails, (size_t)L * 2 * (bf + 1)));
    ZKC_TRY(pool.get(&flags, 2 * n)); ZKC_TRY(pool.get(&ranks, 2 * n)); ZKC_TRY(pool.get(&replist, n)); ZKC_TRY(pool.get(&counts, 4));
    ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(tails, tails_host, (size_t)L * 2 * (bf + 1) * sizeof(Fr), cudaMemcpyHostToDevice, st));
    std::vector<G1Affine> pts;
    for (uint32_t l = 0; l < L; ++l) {
      ProfScope _p(ctx, "prove.lookup_permute");
      Fr* comp_in = lk_comp + (size_t)2 * l * n; Fr* comp_tab = comp_in + n;
      Fr* ap = lk_perm + (size_t)2 * l * n; Fr* sp = ap + n;
      const Fr* tl = tails + (size_t)l * 2 * (bf + 1);
      ZKC_TRY(eval_program(ctx, pk->lookups[l].first, qlag, comp_in, n, 1, theta, 0));
      ZKC_TRY(eval_program(ctx, pk->lookups[l].second, qlag, comp_tab, n, 1, theta, 0));
      k_lookup_prepare<<<grid_for(n, 256), 256, 0, st>>>(comp_in, ca, U, n); ZKC_LAUNCH_CHECK(ctx);
      k_lookup_prepare<<<grid_for(n, 256), 256, 0, st>>>(comp_tab, ct, U, n); ZKC_LAUNCH_CHECK(ctx);
      ZKC_TRY(sort_u256_padded(ctx, ca, n, U));
      ZKC_TRY(sort_u256_padded(ctx, ct, n, U));
      ZKC_CUDA_TRY(ctx, cudaMemsetAsync(counts, 0, 16, st));
      k_lookup_flags<<<grid_for(U, 128), 128, 0, st>>>(ca, ct, flags, flags + n, counts, U); ZKC_LAUNCH_CHECK(ctx);
      ZKC_TRY(u32_scan(ctx, flags, ranks, U, counts + 2));
      ZKC_TRY(u32_scan(ctx, flags + n, ranks + n, U, counts + 3));
      k_lookup_replist<<<grid_for(U, 256), 256, 0, st>>>(flags, ranks, replist, U); ZKC_LAUNCH_CHECK(ctx);
      // S' is assembled in canonical form inside `sp`
      k_lookup_assign<<<grid_for(U, 256), 256, 0, st>>>(ca, ct, flags, flags + n, ranks + n, replist, counts + 2, sp, U, lookup_fill); ZKC_LAUNCH_CHECK(ctx);
      uint32_t hc[4];
      ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(hc, counts, 16, cudaMemcpyDeviceToHost, st));
      ZKC_CUDA_TRY(ctx, cudaStreamSynchronize(st));
      if (hc[0] != hc[1] || hc[2] != hc[3])
        return set_err(ctx, ZKC_ERR_CONSTRAINT_SYSTEM_FAILURE, "lookup: an input value is not in the table (permute_expression_pair)");
      // blinding tails: A' first, then S' (A.6 step 5)
      k_lookup_finish<<<grid_for(n, 256), 256, 0, st>>>(ca, tl, ap, U, n); ZKC_LAUNCH_CHECK(ctx);
      k_lookup_finish<<<grid_for(n, 256), 256, 0, st>>>(sp, tl + (bf + 1), sp, U, n); ZKC_LAUNCH_CHECK(ctx);
      {
        SideScope side(ctx);   // A', S' coefficient forms and cosets (needed for h(X) and the openings)
        Fr* pp = lk_perm_polys + (size_t)2 * l * n;
        ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(pp, ap, (size_t)2 * n * sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
        ZKC_TRY(to_coeff(pp, 2));
        ZKC_TRY(to_extended(pp, lk_cosets + (size_t)3 * l * en + en, 2));
      }
      ZKC_TRY(commit(1, ap, 2, pts));
      out.insert(out.end(), pts.begin(), pts.end());
    }
  }
  stage = ST_PRODUCTS;
  return ZKC_OK;
}

// 6 + 7. permutation and lookup grand products.  Both depend only on (beta, gamma), so their denominators share
//        one batch inversion and their z columns one commitment launch.  One scan across all permutation sets chains
//        z_j[0] = z_{j-1}[U].
int zkc_prover::products(const Fr& beta_, const Fr& gamma_, const Fr* tails_host, std::vector<G1Affine>& out) {
  beta = beta_; gamma = gamma_;
  out.clear();
  cudaStream_t st = ctx->stream;
  if (Pn + L) {
    ProfScope _p(ctx, "prove.grand_products");
    // layout of num / den: [perm: Pn*U + 1] [lookup 0: U + 1] ... (the +1 pads make each scan emit z[U])
    const size_t perm_len = Pn ? (size_t)Pn * U + 1 : 0, lk_len = U + 1, total = perm_len + (size_t)L * lk_len;
    Fr *num, *den, *tails;
    ZKC_TRY(pool.get(&num, total)); ZKC_TRY(pool.get(&den, total)); ZKC_TRY(pool.get(&tails, (size_t)(Pn + L) * bf));
    ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(tails, tails_host, (size_t)(Pn + L) * bf * sizeof(Fr), cudaMemcpyHostToDevice, st));
    // rows are independent up to the scan: a team splits the flat [0, total) range, each rank builds and inverts its part
    std::vector<Segment> flat;
    if (team) for (int r : team_ranks(ctx)) { uint64_t lo, hi; shard_range(total, ctx->team_world, r, &lo, &hi); if (hi > lo) flat.push_back({lo, hi - lo}); }
    else flat.push_back({0, total});
    auto clip = [](const Segment& f, uint64_t base, uint64_t len, uint64_t* r0, uint64_t* rc) {   // rows of [base, base + len) inside f
      const uint64_t a = std::max(f.lo, base), b = std::min(f.lo + f.len, base + len);
      *r0 = a > base ? a - base : 0; *rc = b > a ? b - a : 0;
    };
    if (Pn) {
      ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(num + perm_len - 1, &ONE, sizeof(Fr), cudaMemcpyHostToDevice, st));
      ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(den + perm_len - 1, &ONE, sizeof(Fr), cudaMemcpyHostToDevice, st));
    }
    for (uint32_t l = 0; l < L; ++l) {
      ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(num + perm_len + (size_t)l * lk_len + U, &ONE, sizeof(Fr), cudaMemcpyHostToDevice, st));
      ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(den + perm_len + (size_t)l * lk_len + U, &ONE, sizeof(Fr), cudaMemcpyHostToDevice, st));
    }
    for (const Segment& f : flat) {
      uint64_t r0, rc;
      for (uint32_t s = 0; s < Pn; ++s) {
        clip(f, (uint64_t)s * U, U, &r0, &rc);
        if (!rc) continue;
        k_perm_num_den<<<grid_for(rc, 128), 128, 0, st>>>(perm_args(s, false), pk->omega_pows, beta, gamma, num + (size_t)s * U, den + (size_t)s * U, r0, rc);
        ZKC_LAUNCH_CHECK(ctx);
      }
      for (uint32_t l = 0; l < L; ++l) {
        const Fr* comp_in = lk_comp + (size_t)2 * l * n; const Fr* comp_tab = comp_in + n;
        const Fr* ap = lk_perm + (size_t)2 * l * n; const Fr* sp = ap + n;
        Fr* nl = num + perm_len + (size_t)l * lk_len; Fr* dl = den + perm_len + (size_t)l * lk_len;
        clip(f, perm_len + (uint64_t)l * lk_len, U, &r0, &rc);
        if (!rc) continue;
        k_lookup_num_den<<<grid_for(rc, 128), 128, 0, st>>>(comp_in, comp_tab, ap, sp, beta, gamma, nl, dl, r0, rc); ZKC_LAUNCH_CHECK(ctx);
      }
      ZKC_TRY(fr_batch_invert(ctx, den + f.lo, den + f.lo, f.len));
      k_pk_mul_vec<<<grid_for(f.len, 256), 256, 0, st>>>(num + f.lo, den + f.lo, num + f.lo, f.len); ZKC_LAUNCH_CHECK(ctx);
    }
    if (team) ZKC_TRY(team_allgather_flat(ctx, num, total));
    if (Pn) ZKC_TRY(fr_scan(ctx, num, den, perm_len, SCAN_MUL, 0, ONE));
    for (uint32_t l = 0; l < L; ++l)
      ZKC_TRY(fr_scan(ctx, num + perm_len + (size_t)l * lk_len, den + perm_len + (size_t)l * lk_len, lk_len, SCAN_MUL, 0, ONE));
    if (Pn) { k_assemble_z<<<grid_for((size_t)Pn * n, 256), 256, 0, st>>>(den, tails, pz, n, U, bf, Pn); ZKC_LAUNCH_CHECK(ctx); }
    for (uint32_t l = 0; l < L; ++l) {
      k_assemble_z<<<grid_for(n, 256), 256, 0, st>>>(den + perm_len + (size_t)l * lk_len, tails + (size_t)(Pn + l) * bf, lk_z + (size_t)l * n, n, U, bf, 1);
      ZKC_LAUNCH_CHECK(ctx);
    }
    {
      SideScope side(ctx);   // z coefficient forms and cosets
      ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(z_all_polys, z_all, (size_t)(Pn + L) * n * sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
      ZKC_TRY(to_coeff(z_all_polys, Pn + L));
      if (Pn) ZKC_TRY(to_extended(pz_polys, pz_cosets, Pn));
      for (uint32_t l = 0; l < L; ++l) ZKC_TRY(to_extended(lk_z_polys + (size_t)l * n, lk_cosets + (size_t)3 * l * en, 1));
    }
    ZKC_TRY(commit(1, z_all, Pn + L, out));   // synchronises: the borrowed tails have been consumed
  }
  stage = ST_VANISHING;
  return ZKC_OK;
}

// 8. vanishing argument: commitment to the random polynomial
int zkc_prover::vanishing(const RandomSpec* rs, G1Affine* out) {
  if (!random_enqueued) {
    if (!rs) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_prove_vanishing: the random polynomial was given neither here nor to zkc_prove_begin");
    ZKC_TRY(enqueue_random(*rs));
  } else if (rs) {
    return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_prove_vanishing: the random polynomial was already given to zkc_prove_begin");
  }
  zkc_g1 rp;
  ZKC_TRY(msm_finish(ctx, &random_commit, &rp));
  g1_from_abi(rp, *out);
  stage = ST_QUOTIENT;
  return ZKC_OK;
}

// 10 + 11. h(X) on the extended coset, division by X^n - 1, back to coefficients, commitments of the pieces
int zkc_prover::quotient(const Fr& y_, std::vector<G1Affine>& out) {
  y = y_;
  cudaStream_t st = ctx->stream;
  Fr* hcm;   // h on the class-major extended coset
  ZKC_TRY(pool.get(&hval, en)); ZKC_TRY(pool.get(&hcm, en)); ZKC_TRY(pool.get(&lk_comp_cosets, (size_t)2 * en));
  side_join(ctx);   // every coset produced on the side stream is complete from here on
  {
    ProfScope _p(ctx, "prove.quotient");
    const Fr* tw_ext = nullptr;
    if (Pn) ZKC_TRY(ntt_twiddles(ctx, pk->ext_k, &tw_ext));
    // class-major rows: the cyclic length every kernel sees is n (one residue class), rotations are not scaled
    for (const Segment& rb : my_rows) {
      const uint64_t r0 = rb.lo, rc = rb.len;
      ZKC_TRY(eval_program(ctx, pk->gates, qext, hcm, n, 1, y, 0, r0, rc));
      if (Pn) {
        PermFixedArgs fa; fa.nsets = Pn;
        for (uint32_t s = 0; s < Pn; ++s) fa.z[s] = pz_cosets + (size_t)s * en;
        const int64_t last_off = -(int64_t)(bf + 1);
        k_quot_perm_fixed<<<grid_for(rc, 128), 128, 0, st>>>(hcm, fa, pk->l0, pk->l_last, y, n, last_off, r0, rc); ZKC_LAUNCH_CHECK(ctx);
        for (uint32_t s = 0; s < Pn; ++s) {
          k_quot_perm_set<<<grid_for(rc, 128), 128, 0, st>>>(hcm, perm_args(s, true), pz_cosets + (size_t)s * en, pk->l_active, tw_ext, pk->ext_k,
                                                              pk->cs.k, beta, gamma, y, r0, rc);
          ZKC_LAUNCH_CHECK(ctx);
        }
      }
      for (uint32_t l = 0; l < L; ++l) {
        Fr* zc = lk_cosets + (size_t)3 * l * en; Fr* ac = zc + en; Fr* sc = ac + en;
        ZKC_TRY(eval_program(ctx, pk->lookups[l].first, qext, lk_comp_cosets, n, 1, theta, 0, r0, rc));
        ZKC_TRY(eval_program(ctx, pk->lookups[l].second, qext, lk_comp_cosets + en, n, 1, theta, 0, r0, rc));
        k_quot_lookup<<<grid_for(rc, 128), 128, 0, st>>>(hcm, zc, ac, sc, lk_comp_cosets, lk_comp_cosets + en, pk->l0, pk->l_last, pk->l_active, beta,
                                                          gamma, y, n, r0, rc);
        ZKC_LAUNCH_CHECK(ctx);
      }
      ZKC_TRY(dom_divide_by_vanishing_classes(ctx, pk->dom, hcm, r0, rc));
    }
    //     ... (team: every rank needs the whole quotient) and back to the coefficient forms of the q pieces
    if (!team) {
      ZKC_TRY(dom_classes_to_pieces(ctx, pk->dom, hcm, hval));
    } else {
      // the way back is the four-step transform with its exchange: every rank gathers the class values, the per-class inverse
      // transforms are dealt to the ranks (class c to rank c mod world) and broadcast, the q x q mix is replicated
      ZKC_TRY(team_allgather_rows(ctx, hcm, hn));
      for (int r : team_ranks(ctx))
        for (uint32_t c = (uint32_t)r; c < q; c += (uint32_t)ctx->team_world) ZKC_TRY(dom_classes_inverse(ctx, pk->dom, hcm, c, c + 1));
      ZKC_TRY(team_bcast_blocks(ctx, hcm, n, q));
      ZKC_TRY(dom_classes_mix(ctx, pk->dom, hcm, hval));
    }
  }
  ZKC_TRY(commit_points(ctx, pk->srs, 0, hval, n, q, out));
  stage = ST_EVALS;
  return ZKC_OK;
}

// 13. evaluations — one batched launch for every (polynomial, point) pair of the proof; `out` in transcript order (A.9)
int zkc_prover::evals(const Fr& x_, std::vector<Fr>& out) {
  x = x_;
  const Cs& cs = pk->cs;
  cudaStream_t st = ctx->stream;
  const Fr xn = fe_pow_u64(x, n);
  ZKC_TRY(pool.get(&h_poly, n));   // sum_i xn^i * piece_i
  {
    std::vector<const Fr*> ps; std::vector<Fr> cf;
    Fr pw = ONE;
    for (uint32_t i = 0; i < q; ++i) { ps.push_back(hval + (size_t)i * n); cf.push_back(pw); pw = fe_mul(pw, xn); }
    ZKC_TRY(fr_lincomb(ctx, h_poly, n, ps, cf));
  }
  auto want = [&](const Fr* poly, const Fr& pt) { ev_polys.push_back(poly); ev_points.push_back(pt); return ev_polys.size() - 1; };
  const Fr x_next = rotate_omega(x, omega, omega_inv, 1), x_inv = rotate_omega(x, omega, omega_inv, -1);
  const Fr x_last = rotate_omega(x, omega, omega_inv, -(int32_t)(bf + 1));
  for (auto& qq : cs.aq) i_adv.push_back(want(adv_polys + (size_t)qq.first * n, rotate_omega(x, omega, omega_inv, qq.second)));
  for (auto& qq : cs.fq) i_fix.push_back(want(pk->fixed_polys + (size_t)qq.first * n, rotate_omega(x, omega, omega_inv, qq.second)));
  i_rand = want(random_poly, x);
  for (size_t g = 0; g < cs.perm.size(); ++g) i_sig.push_back(want(pk->sigma_polys + g * n, x));
  for (uint32_t s = 0; s < Pn; ++s) {
    i_pz.push_back(want(pz_polys + (size_t)s * n, x));
    i_pz.push_back(want(pz_polys + (size_t)s * n, x_next));
    if (s + 1 != Pn) i_pz.push_back(want(pz_polys + (size_t)s * n, x_last));
  }
  for (uint32_t l = 0; l < L; ++l) {
    const Fr* zp = lk_z_polys + (size_t)l * n; const Fr* ap = lk_perm_polys + (size_t)2 * l * n; const Fr* sp = ap + n;
    i_lk.push_back(want(zp, x)); i_lk.push_back(want(zp, x_next)); i_lk.push_back(want(ap, x)); i_lk.push_back(want(ap, x_inv));
    i_lk.push_back(want(sp, x));
  }
  i_h = want(h_poly, x);
  side_join(ctx);   // coefficient forms come from the side stream (already joined by quotient(); harmless)
  if (!team) {
    ZKC_TRY(fr_eval_batch(ctx, ev_polys, n, ev_points, ev));
  } else {
    // the (polynomial, point) pairs are independent: each rank evaluates a contiguous share, 32-byte results all-gathered
    const size_t m = ev_polys.size(), W = (size_t)ctx->team_world, per = (m + W - 1) / W;
    Fr* ev_dev;
    ZKC_TRY(pool.get(&ev_dev, per * W));
    std::vector<Fr> all(per * W, ZERO);
    for (int r : team_ranks(ctx)) {
      const size_t j0 = std::min(m, (size_t)r * per), j1 = std::min(m, j0 + per);
      std::vector<Fr> part;
      ZKC_TRY(fr_eval_batch(ctx, std::vector<const Fr*>(ev_polys.begin() + j0, ev_polys.begin() + j1), n,
                            std::vector<Fr>(ev_points.begin() + j0, ev_points.begin() + j1), part));
      std::copy(part.begin(), part.end(), all.begin() + j0);
    }
    ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(ev_dev, all.data(), all.size() * sizeof(Fr), cudaMemcpyHostToDevice, st));
    ZKC_TRY(team_allgather(ctx, ev_dev, per * sizeof(Fr)));
    ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(all.data(), ev_dev, all.size() * sizeof(Fr), cudaMemcpyDeviceToHost, st));
    ZKC_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    ev.assign(all.begin(), all.begin() + m);
  }
  out.clear();
  for (size_t i : i_adv) out.push_back(ev[i]);
  for (size_t i : i_fix) out.push_back(ev[i]);
  out.push_back(ev[i_rand]);
  for (size_t i : i_sig) out.push_back(ev[i]);
  for (size_t i : i_pz) out.push_back(ev[i]);
  for (size_t i : i_lk) out.push_back(ev[i]);
  stage = ST_OPEN;
  return ZKC_OK;
}

// 14. opening queries in upstream order (A.10) + scratch columns of the multiopen argument
int zkc_prover::open_prepare() {
  if (!queries.empty()) return ZKC_OK;
  auto add_q = [&](size_t i) { queries.push_back(Query{ev_polys[i], ev_points[i], ev[i]}); };
  for (size_t i : i_adv) add_q(i);
  {
    size_t pos = 0;
    std::vector<size_t> lasts;
    for (uint32_t s = 0; s < Pn; ++s) {
      add_q(i_pz[pos]); add_q(i_pz[pos + 1]);
      if (s + 1 != Pn) { lasts.push_back(i_pz[pos + 2]); pos += 3; } else pos += 2;
    }
    for (size_t r = lasts.size(); r-- > 0;) add_q(lasts[r]);
  }
  for (uint32_t l = 0; l < L; ++l) {
    const size_t b = (size_t)5 * l;   // i_lk: z@x, z@x_next, a@x, a@x_inv, s@x
    add_q(i_lk[b]); add_q(i_lk[b + 2]); add_q(i_lk[b + 4]); add_q(i_lk[b + 3]); add_q(i_lk[b + 1]);
  }
  for (size_t i : i_fix) add_q(i);
  for (size_t i : i_sig) add_q(i);
  add_q(i_h);
  add_q(i_rand);
  ZKC_TRY(pool.get(&acc, n)); ZKC_TRY(pool.get(&tmp1, n)); ZKC_TRY(pool.get(&tmp2, n + n / 8 + 64)); ZKC_TRY(pool.get(&tmp3, n));   // tmp2: chunk partials of every level of the synthetic divisions
  return ZKC_OK;
}

int zkc_prover::lincomb_s(Fr* out, const std::vector<const Fr*>& ps, const std::vector<Fr>& cf) {
  for (auto& sl : my_coeffs) {
    std::vector<const Fr*> shifted(ps);
    for (auto& p : shifted) p += sl.second.lo;
    ZKC_TRY(fr_lincomb(ctx, out + sl.second.lo, sl.second.len, shifted, cf));
  }
  return ZKC_OK;
}
int zkc_prover::sub_low_s(Fr* a, const std::vector<Fr>& low) {
  for (auto& sl : my_coeffs) if (sl.second.lo == 0) ZKC_TRY(fr_sub_low(ctx, a, low));
  return ZKC_OK;
}
int zkc_prover::scale_s(Fr* a, const Fr& f) {
  for (auto& sl : my_coeffs) ZKC_TRY(fr_scale(ctx, a + sl.second.lo, sl.second.len, f));
  return ZKC_OK;
}
int zkc_prover::kate_s(const std::vector<Fr*>& ps, const std::vector<Fr>& roots, Fr* tmp) {
  if (!sliced) return fr_kate_division_batch(ctx, ps, roots, n, tmp);
  cudaStream_t st = ctx->stream;
  const size_t W = (size_t)ctx->team_world;
  for (size_t off = 0; off < ps.size(); off += KD_MAX_JOBS) {
    const size_t J = std::min<size_t>(KD_MAX_JOBS, ps.size() - off);
    const std::vector<Fr> rt(roots.begin() + off, roots.begin() + off + J);
    // value of every slice at the root, taken before the in-place divisions
    std::vector<Fr> E(W * KD_MAX_JOBS, ZERO);
    for (auto& sl : my_coeffs) {
      std::vector<const Fr*> cp(J);
      for (size_t j = 0; j < J; ++j) cp[j] = ps[off + j] + sl.second.lo;
      std::vector<Fr> e1;
      ZKC_TRY(fr_eval_batch(ctx, cp, sl.second.len, rt, e1));
      std::copy(e1.begin(), e1.end(), E.begin() + (size_t)sl.first * KD_MAX_JOBS);
    }
    ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(kd_exchange, E.data(), E.size() * sizeof(Fr), cudaMemcpyHostToDevice, st));
    ZKC_TRY(team_allgather(ctx, kd_exchange, KD_MAX_JOBS * sizeof(Fr)));
    ZKC_CUDA_TRY(ctx, cudaMemcpyAsync(E.data(), kd_exchange, E.size() * sizeof(Fr), cudaMemcpyDeviceToHost, st));
    ZKC_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    for (auto& sl : my_coeffs) {
      std::vector<Fr*> cp(J);
      for (size_t j = 0; j < J; ++j) cp[j] = ps[off + j] + sl.second.lo;
      ZKC_TRY(fr_kate_division_batch(ctx, cp, rt, sl.second.len, tmp));
      for (size_t j = 0; j < J; ++j) {
        Fr C = ZERO;   // value at the root of everything above this slice: C_{q-1} = E_q + z^(len_q) C_q
        for (int qq = (int)W - 1; qq > sl.first; --qq) {
          uint64_t lo, hi;
          shard_range(n, (int)W, qq, &lo, &hi);
          C = fe_add(E[(size_t)qq * KD_MAX_JOBS + j], fe_mul(fe_pow_u64(rt[j], hi - lo), C));
        }
        ZKC_TRY(fr_add_geometric(ctx, cp[j], sl.second.len, C, rt[j]));
      }
    }
  }
  return ZKC_OK;
}

// ---- SHPLONK (A.11), first message: h(X) = sum_i v^i * ( sum_j y^j (p_ij - r_ij) ) / Z_i ----
int zkc_prover::shplonk_h(const Fr& yy, const Fr& v, G1Affine* out) {
  ProfScope _p(ctx, "prove.shplonk");
  ZKC_TRY(open_prepare());
  sh_y = yy; sh_v = v;
  // Team proving: a point-range MSM reads only coefficients [lo, hi) of the polynomial it commits, so the
  // quotient polynomials are built slice by slice on the rank that will commit the slice: linear combinations are
  // element-wise, and a synthetic division of a slice needs one field element from the slices above it (the value of
  // their tail at the root), exchanged with a 32-byte-per-job all-gather.
  sliced = team && n >= 16ull * (uint64_t)ctx->team_world;
  if (sliced) for (int r : team_ranks(ctx)) { uint64_t lo, hi; shard_range(n, ctx->team_world, r, &lo, &hi); my_coeffs.push_back({r, {lo, hi - lo}}); }
  else my_coeffs.push_back({0, {0, n}});
  if (sliced) ZKC_TRY(pool.get(&kd_exchange, (size_t)ctx->team_world * KD_MAX_JOBS));
  build_rotation_sets(queries, sets, super_points);
  rcoef.assign(sets.size(), {});
  for (size_t s = 0; s < sets.size(); ++s) {
    const std::vector<std::vector<Fr>> basis = zkc::host::lagrange_basis(sets[s].points);   // one inversion per point, shared by the set
    for (size_t p = 0; p < sets[s].polys.size(); ++p) rcoef[s].push_back(zkc::host::interpolate_with_basis(basis, sets[s].evals[p]));
  }
  // numerators per set, then one batched division launch per "round" (the r-th root of every set that still has one),
  // then one linear combination
  Fr* setbuf;
  ZKC_TRY(pool.get(&setbuf, sets.size() * n));
  size_t max_roots = 0;
  for (size_t s = 0; s < sets.size(); ++s) {
    std::vector<Fr> cf; std::vector<Fr> low(sets[s].points.size(), ZERO);
    Fr py = ONE;
    for (size_t p = 0; p < sets[s].polys.size(); ++p) {
      cf.push_back(py);
      for (size_t i = 0; i < low.size(); ++i) low[i] = fe_add(low[i], fe_mul(py, rcoef[s][p][i]));
      py = fe_mul(py, yy);
    }
    ZKC_TRY(lincomb_s(setbuf + s * n, sets[s].polys, cf));
    ZKC_TRY(sub_low_s(setbuf + s * n, low));
    max_roots = std::max(max_roots, sets[s].points.size());
  }
  for (size_t r = 0; r < max_roots; ++r) {
    std::vector<Fr*> jp; std::vector<Fr> jr;
    for (size_t s = 0; s < sets.size(); ++s) if (sets[s].points.size() > r) { jp.push_back(setbuf + s * n); jr.push_back(sets[s].points[r]); }
    ZKC_TRY(kate_s(jp, jr, tmp2));
  }
  {
    std::vector<const Fr*> ps; std::vector<Fr> cf;
    Fr pv = ONE;
    for (size_t s = 0; s < sets.size(); ++s) { ps.push_back(setbuf + s * n); cf.push_back(pv); pv = fe_mul(pv, v); }
    ZKC_TRY(lincomb_s(acc, ps, cf));
  }
  std::vector<G1Affine> pts;
  ZKC_TRY(commit(0, acc, 1, pts));
  *out = pts[0];
  stage = ST_SHPLONK_W;
  return ZKC_OK;
}

// second message: L(X) = sum_i v^i z_i sum_j y^j (p_ij - r_ij(u)) - Z_T(u) h(X), divided by (X - u) and by z_0
int zkc_prover::shplonk_w(const Fr& u, G1Affine* out) {
  ProfScope _p(ctx, "prove.shplonk");
  std::vector<const Fr*> ps; std::vector<Fr> cf;
  Fr cst = ZERO, z0 = ZERO;
  Fr pv = ONE;
  for (size_t s = 0; s < sets.size(); ++s) {
    std::vector<Fr> diffs;
    for (auto& sp : super_points) {
      bool in = false;
      for (auto& p : sets[s].points) if (fe_eq(p, sp)) { in = true; break; }
      if (!in) diffs.push_back(sp);
    }
    const Fr zi = vanishing_eval(diffs, u);
    if (s == 0) z0 = zi;
    const Fr w = fe_mul(zi, pv);
    Fr py = ONE;
    for (size_t p = 0; p < sets[s].polys.size(); ++p) {
      const Fr c = fe_mul(w, py);
      // the same polynomial never sits in two sets, but may repeat across lincomb slots safely
      ps.push_back(sets[s].polys[p]); cf.push_back(c);
      cst = fe_add(cst, fe_mul(c, eval_small(rcoef[s][p], u)));
      py = fe_mul(py, sh_y);
    }
    pv = fe_mul(pv, sh_v);
  }
  const Fr zt = vanishing_eval(super_points, u);
  ps.push_back(acc); cf.push_back(fe_neg(zt));
  ZKC_TRY(lincomb_s(tmp1, ps, cf));
  ZKC_TRY(sub_low_s(tmp1, {cst}));
  if (sliced) ZKC_TRY(kate_s({tmp1}, {u}, tmp2));
  else ZKC_TRY(fr_kate_division(ctx, tmp1, tmp1, n, u, tmp2, tmp3));
  if (fe_is_zero(z0)) return set_err(ctx, ZKC_ERR_OPENING, "shplonk: z_diff of the first rotation set is zero");
  ZKC_TRY(scale_s(tmp1, fe_inv(z0)));
  std::vector<G1Affine> pts;
  ZKC_TRY(commit(0, tmp1, 1, pts));
  *out = pts[0];
  side_join(ctx);
  ZKC_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  stage = ST_DONE;
  return ZKC_OK;
}

// ---- GWC (A.11): one witness per distinct point, first-appearance order, powers of v ----
int zkc_prover::gwc(const Fr& v, std::vector<G1Affine>& out) {
  ProfScope _p(ctx, "prove.gwc");
  ZKC_TRY(open_prepare());
  std::vector<Fr> points;
  for (auto& qq : queries) {
    bool seen = false;
    for (auto& p : points) if (fe_eq(p, qq.point)) { seen = true; break; }
    if (!seen) points.push_back(qq.point);
  }
  Fr* wbuf;
  ZKC_TRY(pool.get(&wbuf, points.size() * n));
  std::vector<Fr*> jp;
  for (size_t pi = 0; pi < points.size(); ++pi) {
    const Fr& z = points[pi];
    std::vector<const Fr*> ps; std::vector<Fr> cf;
    Fr pvv = ONE, eacc = ZERO;
    for (auto& qq : queries) {
      if (!fe_eq(qq.point, z)) continue;
      ps.push_back(qq.poly); cf.push_back(pvv);
      eacc = fe_add(eacc, fe_mul(qq.eval, pvv));
      pvv = fe_mul(pvv, v);
    }
    ZKC_TRY(fr_lincomb(ctx, wbuf + pi * n, n, ps, cf));
    ZKC_TRY(fr_sub_low(ctx, wbuf + pi * n, {eacc}));
    jp.push_back(wbuf + pi * n);
  }
  ZKC_TRY(fr_kate_division_batch(ctx, jp, points, n, tmp2));
  ZKC_TRY(commit_points(ctx, pk->srs, 0, wbuf, n, (uint32_t)points.size(), out));   // one launch for every witness commitment
  side_join(ctx);
  ZKC_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  stage = ST_DONE;
  return ZKC_OK;
}

// ---- step API (include/zkcert_cuda.h) ------------------------------------------------------------------------------------------
namespace {
Fr fr_abi(const zkc_fr* p) { Fr r; memcpy(r.v, p, 32); return r; }
void points_out(const std::vector<G1Affine>& pts, zkc_g1_affine* out) { if (!pts.empty()) memcpy(out, pts.data(), pts.size() * sizeof(G1Affine)); }
// run one step: stage check, ctx lock, failure poisons the session
template <class F> int prover_step(zkc_prover* p, int expect, const char* name, F&& f) {
  if (!p) return ZKC_ERR_BAD_ARG;
  CtxLock lock(p->ctx);
  if (p->stage != expect) return set_err(p->ctx, ZKC_ERR_BAD_ARG, std::string(name) + ": called out of order (or after a failed step)");
  const int s = f();
  if (s != ZKC_OK) p->stage = ST_FAILED;
  return s;
}
int prover_create(zkc_ctx* ctx, const zkc_pk* pk, const zkc_fr* advice, const zkc_fr* const* instances, const size_t* instance_lens,
                  size_t num_instance_columns, const char* who, std::unique_ptr<zkc_prover>* out) {
  if (!ctx || !pk || (pk->cs.num_advice && !advice) || (pk->cs.num_instance && (!instances || !instance_lens)))
    return set_err(ctx, ZKC_ERR_BAD_ARG, std::string(who) + ": null argument");
  if (num_instance_columns != pk->cs.num_instance)
    return set_err(ctx, ZKC_ERR_INVALID_INSTANCES, std::string(who) + ": instances.len() != cs.num_instance_columns (plonk::Error::InvalidInstances)");
  if (ctx->active_prover) return set_err(ctx, ZKC_ERR_BAD_ARG, std::string(who) + ": another create_proof session is active on this ctx");
  out->reset(new zkc_prover(ctx, pk));
  ctx->active_prover = out->get();
  return ZKC_OK;
}
}  // namespace

extern "C" int zkc_prove_begin(zkc_ctx* ctx, const zkc_pk* pk, const zkc_fr* advice, int advice_on_device, const zkc_fr* const* instances,
                               const size_t* instance_lens, size_t num_instance_columns, const zkc_fr* advice_tails,
                               const zkc_random_poly* early_random, zkc_prover** out, zkc_g1_affine* advice_commitments) {
  if (!ctx || !out || (pk && pk->cs.num_advice && !advice_commitments)) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_prove_begin: null argument");
  *out = nullptr;
  CtxLock lock(ctx);
  std::unique_ptr<zkc_prover> p;
  ZKC_TRY(prover_create(ctx, pk, advice, instances, instance_lens, num_instance_columns, "zkc_prove_begin", &p));
  RandomSpec rs;
  if (early_random) ZKC_TRY(random_spec_from_abi(ctx, early_random, pk->cs.n(), &rs));
  std::vector<G1Affine> pts;
  ZKC_TRY(p->begin(advice, advice_on_device, instances, instance_lens, (const Fr*)advice_tails, early_random ? &rs : nullptr, pts));
  points_out(pts, advice_commitments);
  *out = p.release();
  return ZKC_OK;
}
extern "C" int zkc_prove_lookups(zkc_prover* p, const zkc_fr* theta, const zkc_fr* tails, int lookup_fill, zkc_g1_affine* out) {
  return prover_step(p, ST_LOOKUPS, "zkc_prove_lookups", [&]() -> int {
    if (!theta || (p->L && (!tails || !out)) || lookup_fill < 0 || lookup_fill > 1) return set_err(p->ctx, ZKC_ERR_BAD_ARG, "zkc_prove_lookups: bad argument");
    std::vector<G1Affine> pts;
    ZKC_TRY(p->lookups(fr_abi(theta), (const Fr*)tails, lookup_fill, pts));
    points_out(pts, out);
    return ZKC_OK;
  });
}
extern "C" int zkc_prove_products(zkc_prover* p, const zkc_fr* beta, const zkc_fr* gamma, const zkc_fr* tails, zkc_g1_affine* out) {
  return prover_step(p, ST_PRODUCTS, "zkc_prove_products", [&]() -> int {
    if (!beta || !gamma || ((p->Pn + p->L) && (!tails || !out))) return set_err(p->ctx, ZKC_ERR_BAD_ARG, "zkc_prove_products: null argument");
    std::vector<G1Affine> pts;
    ZKC_TRY(p->products(fr_abi(beta), fr_abi(gamma), (const Fr*)tails, pts));
    points_out(pts, out);
    return ZKC_OK;
  });
}
extern "C" int zkc_prove_vanishing(zkc_prover* p, const zkc_random_poly* random, zkc_g1_affine* out) {
  return prover_step(p, ST_VANISHING, "zkc_prove_vanishing", [&]() -> int {
    if (!out) return set_err(p->ctx, ZKC_ERR_BAD_ARG, "zkc_prove_vanishing: null argument");
    RandomSpec rs;
    if (random) ZKC_TRY(random_spec_from_abi(p->ctx, random, p->n, &rs));
    G1Affine pt;
    ZKC_TRY(p->vanishing(random ? &rs : nullptr, &pt));
    memcpy(out, &pt, sizeof pt);
    return ZKC_OK;
  });
}
extern "C" int zkc_prove_quotient(zkc_prover* p, const zkc_fr* y, zkc_g1_affine* out) {
  return prover_step(p, ST_QUOTIENT, "zkc_prove_quotient", [&]() -> int {
    if (!y || !out) return set_err(p->ctx, ZKC_ERR_BAD_ARG, "zkc_prove_quotient: null argument");
    std::vector<G1Affine> pts;
    ZKC_TRY(p->quotient(fr_abi(y), pts));
    points_out(pts, out);
    return ZKC_OK;
  });
}
extern "C" int zkc_prove_evals(zkc_prover* p, const zkc_fr* x, zkc_fr* out, size_t cap, size_t* count) {
  return prover_step(p, ST_EVALS, "zkc_prove_evals", [&]() -> int {
    if (!x || !count) return set_err(p->ctx, ZKC_ERR_BAD_ARG, "zkc_prove_evals: null argument");
    const Cs& cs = p->pk->cs;
    const size_t m = cs.aq.size() + cs.fq.size() + 1 + cs.perm.size() + (p->Pn ? 3 * (size_t)p->Pn - 1 : 0) + 5 * (size_t)p->L;
    *count = m;
    if (!out || cap < m) return set_err(p->ctx, ZKC_ERR_BAD_ARG, "zkc_prove_evals: output buffer too small");
    std::vector<Fr> ev;
    ZKC_TRY(p->evals(fr_abi(x), ev));
    memcpy(out, ev.data(), ev.size() * sizeof(Fr));
    return ZKC_OK;
  });
}
extern "C" int zkc_prove_open_shplonk_h(zkc_prover* p, const zkc_fr* y, const zkc_fr* v, zkc_g1_affine* out) {
  return prover_step(p, ST_OPEN, "zkc_prove_open_shplonk_h", [&]() -> int {
    if (!y || !v || !out) return set_err(p->ctx, ZKC_ERR_BAD_ARG, "zkc_prove_open_shplonk_h: null argument");
    G1Affine pt;
    ZKC_TRY(p->shplonk_h(fr_abi(y), fr_abi(v), &pt));
    memcpy(out, &pt, sizeof pt);
    return ZKC_OK;
  });
}
extern "C" int zkc_prove_open_shplonk_w(zkc_prover* p, const zkc_fr* u, zkc_g1_affine* out) {
  return prover_step(p, ST_SHPLONK_W, "zkc_prove_open_shplonk_w", [&]() -> int {
    if (!u || !out) return set_err(p->ctx, ZKC_ERR_BAD_ARG, "zkc_prove_open_shplonk_w: null argument");
    G1Affine pt;
    ZKC_TRY(p->shplonk_w(fr_abi(u), &pt));
    memcpy(out, &pt, sizeof pt);
    return ZKC_OK;
  });
}
extern "C" int zkc_prove_open_gwc(zkc_prover* p, const zkc_fr* v, zkc_g1_affine* out, size_t cap, size_t* count) {
  return prover_step(p, ST_OPEN, "zkc_prove_open_gwc", [&]() -> int {
    if (!v || !count) return set_err(p->ctx, ZKC_ERR_BAD_ARG, "zkc_prove_open_gwc: null argument");
    // distinct evaluation points = distinct rotations among all queries; count them before doing any work
    ZKC_TRY(p->open_prepare());
    std::vector<Fr> points;
    for (auto& qq : p->queries) {
      bool seen = false;
      for (auto& pt : points) if (fe_eq(pt, qq.point)) { seen = true; break; }
      if (!seen) points.push_back(qq.point);
    }
    *count = points.size();
    if (!out || cap < points.size()) return set_err(p->ctx, ZKC_ERR_BAD_ARG, "zkc_prove_open_gwc: output buffer too small");
    std::vector<G1Affine> pts;
    ZKC_TRY(p->gwc(fr_abi(v), pts));
    points_out(pts, out);
    return ZKC_OK;
  });
}
extern "C" void zkc_prove_end(zkc_prover* p) {
  if (!p) return;
  CtxLock lock(p->ctx);
  delete p;
}

// ---- create_proof with the library's own transcript and RNG: a driver over the same rounds ----------------------------------------
namespace {
// RNG cursor over the proof's ChaCha stream (host side): every value the prover needs from `rng`, in upstream order
struct Rng {
  host::ChaCha20Rng cpu;
  Rng(const uint8_t seed[32], int dr) : cpu(seed, dr) {}
  Fr draw() { return cpu.fr_random(); }
};
}  // namespace

extern "C" int zkc_prove(zkc_ctx* ctx, const zkc_pk* pk, const zkc_fr* advice, int advice_on_device, const zkc_fr* const* instances,
                         const size_t* instance_lens, size_t num_instance_columns, const zkc_prove_opts* opts, uint8_t* proof_out,
                         size_t proof_cap, size_t* proof_len) {
  if (!ctx || !pk || !opts || !proof_len) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_prove: null argument");
  CtxLock lock(ctx);
  // rank-independent checks come before any device work (team proving: every rank takes the same early exit)
  if (opts->rng_kind < 0 || opts->rng_kind > 1) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_prove: unknown rng_kind");
  if (opts->transcript < 0 || opts->transcript > 3 || opts->multiopen < 0 || opts->multiopen > 1)
    return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_prove: unknown transcript / multiopen");
  if (opts->lookup_fill < 0 || opts->lookup_fill > 1 || opts->random_poly < 0 || opts->random_poly > 1)
    return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_prove: unknown lookup_fill / random_poly");
  std::unique_ptr<zkc_prover> p;
  ZKC_TRY(prover_create(ctx, pk, advice, instances, instance_lens, num_instance_columns, "zkc_prove", &p));
  const Cs& cs = pk->cs;
  const uint64_t n = cs.n();
  const uint32_t bf = cs.blinding_factors, A = cs.num_advice, I = cs.num_instance, L = (uint32_t)cs.lookups.size(), Pn = cs.nsets();
  const uint32_t q = cs.degree - 1;
  uint64_t chunk_len = 0; uint32_t n_chunks = 0;
  if (opts->random_poly == 1) {
    // vanishing::Argument::commit, rayon variant: chunks of n / threads coefficients, threads (+1 if n % threads) generators
    const uint64_t T = opts->random_poly_threads;
    if (!T || T > n) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_prove: random_poly_threads must be in [1, n]");
    chunk_len = n / T; n_chunks = (uint32_t)(T + (n % T ? 1 : 0));
    if ((n + chunk_len - 1) / chunk_len != n_chunks) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_prove: upstream's zip_eq panics for this thread count");
  }
  Rng rng(opts->rng_seed, opts->rng_kind == 1 ? 6 : 10);
  Transcript tr(opts->transcript, opts->point_format);
  auto write_points = [&](const std::vector<G1Affine>& pts) -> int {
    for (auto& pt : pts) if (tr.write_point(pt)) return set_err(ctx, ZKC_ERR_TRANSCRIPT, "cannot write points at infinity to the transcript");
    return ZKC_OK;
  };
  // 0. vk, 1. instances: values are absorbed as scalars (KZG: QUERY_INSTANCE = false)
  tr.common_scalar(pk->transcript_repr);
  for (uint32_t c = 0; c < I; ++c)
    for (size_t i = 0; i < instance_lens[c]; ++i) { Fr v; memcpy(v.v, &instances[c][i], 32); tr.common_scalar(v); }
  // 2. advice.  PSE policy: bf + 1 draws per column; then (OPEN-2) one discarded blind per commitment
  std::vector<Fr> adv_tails;
  if (opts->advice_blinding) { adv_tails.resize((size_t)A * (bf + 1)); for (auto& v : adv_tails) v = rng.draw(); }
  if (opts->blind_draws) for (uint32_t c = 0; c < A; ++c) rng.draw();
  // The random polynomial depends on nothing but the RNG, and the keystream position of its draws is fixed by the constraint
  // system and the blinding policy: describe it now so that it is produced underneath the first rounds.
  const uint64_t words_before_random = rng.cpu.word_pos() + 16 * ((uint64_t)L * 2 * (bf + 1) + (opts->blind_draws ? 2 * L : 0) +
                                                                  (uint64_t)(Pn + L) * bf + (opts->blind_draws ? (Pn + L) : 0));
  RandomSpec rs;
  uint64_t words_after_random;
  if (opts->random_poly == 0) {
    rs.kind = 1; memcpy(rs.key.k, opts->rng_seed, 32); rs.double_rounds = rng.cpu.double_rounds; rs.first_word = words_before_random;
    words_after_random = words_before_random + 16 * n;
  } else {
    host::ChaCha20Rng peek = rng.cpu;
    peek.seek(words_before_random);
    rs.kind = 2; rs.chunk_len = chunk_len; rs.chunk_keys.resize(n_chunks);
    for (auto& k : rs.chunk_keys) peek.fill_bytes((uint8_t*)k.k, 32);
    words_after_random = peek.word_pos();
  }
  std::vector<G1Affine> pts;
  ZKC_TRY(p->begin(advice, advice_on_device, instances, instance_lens, opts->advice_blinding ? adv_tails.data() : nullptr, &rs, pts));
  ZKC_TRY(write_points(pts));
  // 3. theta; 4. lookups: per lookup bf + 1 draws for A', bf + 1 for S', then the two discarded blinds
  const Fr theta = tr.squeeze_challenge();
  {
    std::vector<Fr> t((size_t)L * 2 * (bf + 1));
    for (uint32_t l = 0; l < L; ++l) {
      for (uint32_t i = 0; i < 2 * (bf + 1); ++i) t[(size_t)l * 2 * (bf + 1) + i] = rng.draw();
      if (opts->blind_draws) { rng.draw(); rng.draw(); }
    }
    ZKC_TRY(p->lookups(theta, t.data(), opts->lookup_fill, pts));
    ZKC_TRY(write_points(pts));
  }
  // 5. beta, gamma; 6 + 7. grand products: tails in upstream draw order (every permutation set, then every lookup product)
  const Fr beta = tr.squeeze_challenge();
  const Fr gamma = tr.squeeze_challenge();
  {
    std::vector<Fr> t((size_t)(Pn + L) * bf);
    for (uint32_t s = 0; s < Pn + L; ++s) {
      for (uint32_t i = 0; i < bf; ++i) t[(size_t)s * bf + i] = rng.draw();
      if (opts->blind_draws) rng.draw();
    }
    ZKC_TRY(p->products(beta, gamma, t.data(), pts));
    ZKC_TRY(write_points(pts));
  }
  // 8. vanishing argument: the draws were produced on the device from the same stream
  if (rng.cpu.word_pos() != words_before_random) return set_err(ctx, ZKC_ERR_SYNTHESIS, "internal: RNG draw schedule mismatch before the vanishing argument");
  rng.cpu.seek(words_after_random);
  if (opts->blind_draws) rng.draw();
  {
    G1Affine rp;
    ZKC_TRY(p->vanishing(nullptr, &rp));
    ZKC_TRY(write_points({rp}));
  }
  // 9. y; 10 + 11. h(X)
  const Fr y = tr.squeeze_challenge();
  if (opts->blind_draws) for (uint32_t i = 0; i < q; ++i) rng.draw();
  ZKC_TRY(p->quotient(y, pts));
  ZKC_TRY(write_points(pts));
  // 12. x; 13. evaluations
  const Fr x = tr.squeeze_challenge();
  {
    std::vector<Fr> ev;
    ZKC_TRY(p->evals(x, ev));
    for (auto& e : ev) tr.write_scalar(e);
  }
  // 14. multiopen
  if (opts->multiopen == 0) {
    const Fr yy = tr.squeeze_challenge();
    const Fr v = tr.squeeze_challenge();
    G1Affine pt;
    ZKC_TRY(p->shplonk_h(yy, v, &pt));
    ZKC_TRY(write_points({pt}));
    const Fr u = tr.squeeze_challenge();
    ZKC_TRY(p->shplonk_w(u, &pt));
    ZKC_TRY(write_points({pt}));
  } else {
    const Fr v = tr.squeeze_challenge();
    ZKC_TRY(p->gwc(v, pts));
    ZKC_TRY(write_points(pts));
  }
  *proof_len = tr.proof.size();
  if (!proof_out || proof_cap < tr.proof.size()) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_prove: proof buffer too small");
  memcpy(proof_out, tr.proof.data(), tr.proof.size());
  return ZKC_OK;
}

// Fr::random stream exposed for hosts that want to cross-check their RNG restatement
extern "C" int zkc_rng_fr_random(const uint8_t seed[32], int rng_kind, uint64_t skip, zkc_fr* out, size_t count) {
  if (!seed || !out || rng_kind < 0 || rng_kind > 1) return ZKC_ERR_BAD_ARG;
  host::ChaCha20Rng r(seed, rng_kind == 1 ? 6 : 10);
  r.counter = skip; r.pos = 16;
  for (size_t i = 0; i < count; ++i) { Fr f = r.fr_random(); memcpy(&out[i], f.v, 32); }
  return ZKC_OK;
}
extern "C" void zkc_seed_from_u64(uint64_t state, uint8_t seed[32]) { host::seed_from_u64(state, seed); }

// Compact witness upload: SHA256-bit style circuits assign almost only {0,1} cells, so the
// host may hand columns over as bit / byte / u16 / u64 arrays; they are expanded to Montgomery form on the
// device and the proof is identical to the one zkc_prove emits for the expanded columns.
extern "C" int zkc_prove_compact(zkc_ctx* ctx, const zkc_pk* pk, const zkc_advice_column* cols, const zkc_fr* const* instances,
                                 const size_t* instance_lens, size_t num_instance_columns, const zkc_prove_opts* opts, uint8_t* proof_out,
                                 size_t proof_cap, size_t* proof_len) {
  if (!ctx || !pk || !cols) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_prove_compact: null argument");
  CtxLock lock(ctx);
  const uint64_t n = pk->cs.n();
  const uint32_t A = pk->cs.num_advice;
  Fr* adv = nullptr;
  uint8_t* stage = nullptr;
  cudaStream_t st = ctx->stream;
  ZKC_CUDA_TRY(ctx, cudaMallocAsync((void**)&adv, std::max<size_t>((size_t)A * n, 1) * sizeof(Fr), st));
  size_t stage_bytes = 0, off = 0;
  auto col_bytes = [&](int kind) -> size_t { return kind == 0 ? n * 32 : kind == 1 ? (n + 7) / 8 : kind == 2 ? n : kind == 3 ? n * 2 : n * 8; };
  for (uint32_t c = 0; c < A; ++c) {
    if (cols[c].kind < 0 || cols[c].kind > 4 || !cols[c].data) { cudaFreeAsync(adv, st); return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_prove_compact: bad column"); }
    if (cols[c].kind) stage_bytes += (col_bytes(cols[c].kind) + 15) & ~(size_t)15;
  }
  int status = ZKC_OK;
  if (stage_bytes && cudaMallocAsync((void**)&stage, stage_bytes, st) != cudaSuccess) status = set_err(ctx, ZKC_ERR_OOM, "zkc_prove_compact: out of memory");
  {
    ProfScope _p(ctx, "prove.advice_h2d");
    for (uint32_t c = 0; c < A && status == ZKC_OK; ++c) {
      const size_t b = col_bytes(cols[c].kind);
      cudaError_t e;
      if (cols[c].kind == 0) {
        e = cudaMemcpyAsync(adv + (size_t)c * n, cols[c].data, b, cudaMemcpyHostToDevice, st);
      } else {
        e = cudaMemcpyAsync(stage + off, cols[c].data, b, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) {
          k_expand_compact<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(stage + off, adv + (size_t)c * n, n, cols[c].kind);
          ctx->launches++;
          e = cudaGetLastError();
        }
        off += (b + 15) & ~(size_t)15;
      }
      if (e != cudaSuccess) status = set_err(ctx, ZKC_ERR_CUDA, cudaGetErrorString(e));
    }
  }
  if (status == ZKC_OK) status = zkc_prove(ctx, pk, (const zkc_fr*)adv, 1, instances, instance_lens, num_instance_columns, opts, proof_out, proof_cap, proof_len);
  if (stage) cudaFreeAsync(stage, st);
  cudaFreeAsync(adv, st);
  return status;
}

// Poseidon spec (Grain-generated round constants and MDS) of the SDK transcript, canonical little-endian; for cross-checks
extern "C" int zkc_poseidon_spec(zkc_fr* constants /* 65*3 */, zkc_fr* mds /* 3*3 */) {
  if (!constants || !mds) return ZKC_ERR_BAD_ARG;
  const Fr *c, *m;
  host::poseidon_spec(&c, &m);
  for (int i = 0; i < 195; ++i) { Fr v = fe_to_canonical(c[i]); memcpy(&constants[i], v.v, 32); }
  for (int i = 0; i < 9; ++i) { Fr v = fe_to_canonical(m[i]); memcpy(&mds[i], v.v, 32); }
  return ZKC_OK;
}
