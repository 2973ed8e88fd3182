// Context, device-memory helpers and element-wise Fr/Fq vector kernels of libzkcert_cuda.so.
#include "common.cuh"

#include <cstdlib>

using namespace zkc;

namespace {
int env_int(const char* name, int lo, int hi) { const char* e = getenv(name); if (!e) return 0; const int x = atoi(e); return x >= lo && x <= hi ? x : 0; }
void tunables_from_env(Tunables& v) {
  v.msm_c = env_int("ZKC_MSM_C", 3, 20); v.msm_c_pre = env_int("ZKC_MSM_C_PRE", 3, 20); v.msm_T = env_int("ZKC_MSM_T", 4, 128);
  v.ntt_two_pass_max = env_int("ZKC_NTT_TWO_PASS_MAX", 12, 22); v.msm_accum_occ = env_int("ZKC_MSM_ACCUM_OCC", 3, 4);
  if (const char* e = getenv("ZKC_STAGE_MIN_BYTES")) v.stage_min_bytes = (size_t)strtoull(e, nullptr, 10);
  v.team_poison = getenv("ZKC_TEAM_POISON") != nullptr;
  v.team_commit_by_column = getenv("ZKC_TEAM_COMMIT_BY_COLUMN") != nullptr;
  v.no_program_factoring = getenv("ZKC_NO_PROGRAM_FACTORING") != nullptr;
}
}  // namespace

extern "C" int zkc_ctx_set_tunable(zkc_ctx* c, const char* name, int64_t value) {
  if (!c || !name) return ZKC_ERR_BAD_ARG;
  CtxLock lock(c);
  const std::string n(name);
  Tunables& t = c->tune;
  if (n == "msm_c") t.msm_c = (int)value;
  else if (n == "msm_c_pre") t.msm_c_pre = (int)value;
  else if (n == "msm_T") t.msm_T = (int)value;
  else if (n == "msm_accum_occ") t.msm_accum_occ = (int)value;
  else if (n == "ntt_two_pass_max") t.ntt_two_pass_max = (int)value;
  else if (n == "stage_min_bytes") t.stage_min_bytes = value < 0 ? ((size_t)4 << 20) : (size_t)value;
  else if (n == "team_poison") t.team_poison = value != 0;
  else if (n == "no_program_factoring") t.no_program_factoring = value != 0;
  else if (n == "team_commit_by_column") t.team_commit_by_column = value != 0;
  else return set_err(c, ZKC_ERR_BAD_ARG, "zkc_ctx_set_tunable: unknown name " + n);
  return ZKC_OK;
}

extern "C" const char* zkc_version(void) { return "halo2-zkcert_b200 0.2 (sm_100a)"; }

extern "C" int zkc_ctx_create(int device, zkc_ctx** out) {
  if (!out) return ZKC_ERR_BAD_ARG;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  // No CPU fallback: without a device every entry point fails loudly.
  if (e != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) return ZKC_ERR_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return ZKC_ERR_CUDA;
  zkc_ctx* c = new zkc_ctx();
  c->dev = device;
  tunables_from_env(c->tune);
  if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return ZKC_ERR_CUDA; }
  c->stream = c->own_stream;
  if (cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_msm_main, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_msm_side, cudaEventDisableTiming) != cudaSuccess) { delete c; return ZKC_ERR_CUDA; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
  // keep the stream-ordered pool's memory across synchronisations: per-proof temporaries are
  // re-allocated from it without going back to the driver
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t thr = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  *out = c;
  return ZKC_OK;
}

extern "C" void zkc_ctx_destroy(zkc_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->dev);
  cudaStreamSynchronize(c->stream);
  zkc_team_leave(c);
  for (auto& b : c->scratch) if (b.p) cudaFree(b.p);
  for (auto& kv : c->twiddles) cudaFree(kv.second);
  for (int i = 0; i < 2; ++i) if (c->pinned[i]) cudaFreeHost(c->pinned[i]);
  for (auto e : c->ev_copy) cudaEventDestroy(e);
  if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
  if (c->ev_msm_main) cudaEventDestroy(c->ev_msm_main);
  if (c->ev_msm_side) cudaEventDestroy(c->ev_msm_side);
  for (auto& r : c->prof_pending) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  for (auto e : c->prof_pool) cudaEventDestroy(e);
  if (c->side_stream) { cudaStreamSynchronize(c->side_stream); cudaStreamDestroy(c->side_stream); }
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete c;
}

extern "C" const char* zkc_last_error(const zkc_ctx* c) { return c ? c->err.c_str() : "no context (no CUDA device?)"; }
extern "C" uint64_t zkc_ctx_launch_count(const zkc_ctx* c) { return c ? c->launches : 0; }

extern "C" int zkc_ctx_set_stream(zkc_ctx* c, void* s, int external) {
  if (!c) return ZKC_ERR_BAD_ARG;
  CtxLock lock(c);
  ZKC_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  c->stream = external ? (cudaStream_t)s : c->own_stream;
  return ZKC_OK;
}
extern "C" int zkc_ctx_set_overlap(zkc_ctx* c, int on) {
  if (!c) return ZKC_ERR_BAD_ARG;
  CtxLock lock(c);
  ZKC_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  ZKC_CUDA_TRY(c, cudaStreamSynchronize(c->side_stream));
  c->overlap = on != 0;
  return ZKC_OK;
}
extern "C" int zkc_ctx_sync(zkc_ctx* c) {
  if (!c) return ZKC_ERR_BAD_ARG;
  CtxLock lock(c);
  ZKC_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return ZKC_OK;
}

extern "C" int zkc_profile_enable(zkc_ctx* c, int on) {
  if (!c) return ZKC_ERR_BAD_ARG;
  CtxLock lock(c);
  c->profiling = on != 0;
  return ZKC_OK;
}
extern "C" int zkc_profile_report(zkc_ctx* c, char* buf, size_t cap) {
  if (!c || !buf || cap == 0) return ZKC_ERR_BAD_ARG;
  CtxLock lock(c);
  ZKC_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  for (auto& r : c->prof_pending) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) { auto& a = c->prof_acc[r.name]; a.first += ms; a.second++; }
    c->prof_pool.push_back(r.e0); c->prof_pool.push_back(r.e1);
  }
  c->prof_pending.clear();
  std::string js = "{";
  bool first = true;
  for (auto& kv : c->prof_acc) {
    char tmp[256];
    snprintf(tmp, sizeof tmp, "%s\"%s\": {\"ms\": %.6f, \"n\": %llu}", first ? "" : ", ", kv.first.c_str(), kv.second.first,
             (unsigned long long)kv.second.second);
    js += tmp; first = false;
  }
  for (auto& kv : c->stats) {
    char tmp[256];
    snprintf(tmp, sizeof tmp, "%s\"count:%s\": {\"ms\": 0, \"n\": %llu}", first ? "" : ", ", kv.first.c_str(), (unsigned long long)kv.second);
    js += tmp; first = false;
  }
  js += "}";
  c->prof_acc.clear();
  c->stats.clear();
  snprintf(buf, cap, "%s", js.c_str());
  return ZKC_OK;
}

// Timeline of the phases recorded since profiling was enabled / last reported: [["name", start_ms, duration_ms], ...]
// relative to the first phase.  Consumes the pending records like zkc_profile_report.
extern "C" int zkc_profile_timeline(zkc_ctx* c, char* buf, size_t cap) {
  if (!c || !buf || cap == 0) return ZKC_ERR_BAD_ARG;
  CtxLock lock(c);
  ZKC_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  if (c->side_stream) ZKC_CUDA_TRY(c, cudaStreamSynchronize(c->side_stream));
  std::string js = "[";
  bool first = true;
  for (auto& r : c->prof_pending) {
    float start = 0, dur = 0;
    if (cudaEventElapsedTime(&start, c->prof_pending[0].e0, r.e0) == cudaSuccess && cudaEventElapsedTime(&dur, r.e0, r.e1) == cudaSuccess) {
      char tmp[256];
      snprintf(tmp, sizeof tmp, "%s[\"%s\", %.4f, %.4f]", first ? "" : ", ", r.name.c_str(), start, dur);
      js += tmp; first = false;
    }
    c->prof_pool.push_back(r.e0); c->prof_pool.push_back(r.e1);
  }
  c->prof_pending.clear();
  js += "]";
  snprintf(buf, cap, "%s", js.c_str());
  return ZKC_OK;
}

extern "C" int zkc_dev_alloc(zkc_ctx* c, size_t bytes, void** dptr) {
  if (!c || !dptr) return ZKC_ERR_BAD_ARG;
  CtxLock lock(c);
  ZKC_CUDA_TRY(c, cudaMalloc(dptr, bytes ? bytes : 1));
  return ZKC_OK;
}
extern "C" int zkc_dev_free(zkc_ctx* c, void* dptr) {
  if (!c) return ZKC_ERR_BAD_ARG;
  CtxLock lock(c);
  ZKC_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  ZKC_CUDA_TRY(c, cudaFree(dptr));
  return ZKC_OK;
}
extern "C" int zkc_h2d(zkc_ctx* c, void* dptr, const void* hptr, size_t bytes) {
  if (!c) return ZKC_ERR_BAD_ARG;
  CtxLock lock(c);
  ZKC_CUDA_TRY(c, cudaMemcpyAsync(dptr, hptr, bytes, cudaMemcpyHostToDevice, c->stream));
  ZKC_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return ZKC_OK;
}
extern "C" int zkc_d2h(zkc_ctx* c, void* hptr, const void* dptr, size_t bytes) {
  if (!c) return ZKC_ERR_BAD_ARG;
  CtxLock lock(c);
  ZKC_CUDA_TRY(c, cudaMemcpyAsync(hptr, dptr, bytes, cudaMemcpyDeviceToHost, c->stream));
  ZKC_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return ZKC_OK;
}

// ---- element-wise vector ops -------------------------------------------------------------------
namespace zkc {

template <class P>
__global__ void k_vec_op(int op, const Fe<P>* a, const Fe<P>* b, Fe<P>* out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fe<P> x = fe_load(a + i), r;
  switch (op) {
    case ZKC_OP_ADD: r = fe_add(x, fe_load(b + i)); break;
    case ZKC_OP_SUB: r = fe_sub(x, fe_load(b + i)); break;
    case ZKC_OP_MUL: r = fe_mul(x, fe_load(b + i)); break;
    case ZKC_OP_FROM_CANONICAL: r = fe_from_canonical(x); break;
    case ZKC_OP_TO_CANONICAL: r = fe_to_canonical(x); break;
    case ZKC_OP_NEG: r = fe_neg(x); break;
    case ZKC_OP_SQUARE: r = fe_sqr(x); break;
    default: r = x;
  }
  fe_store(out + i, r);
}

// Batch inversion (Montgomery's trick) with zeros passed through, as halo2's `batch_invert`.
// Each thread owns BI_CHUNK elements strided by the grid so that global accesses stay coalesced.
#define BI_CHUNK 8
template <class P>
__global__ void k_batch_inv(const Fe<P>* a, Fe<P>* out, size_t n, size_t stride) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= stride) return;
  Fe<P> v[BI_CHUNK], pre[BI_CHUNK];
  Fe<P> acc = fe_one<P>();
#pragma unroll
  for (int j = 0; j < BI_CHUNK; ++j) {
    const size_t i = t + (size_t)j * stride;
    v[j] = (i < n) ? fe_load(a + i) : fe_zero<P>();
    pre[j] = acc;
    if (!fe_is_zero(v[j])) acc = fe_mul(acc, v[j]);
  }
  acc = fe_inv(acc);
#pragma unroll
  for (int j = BI_CHUNK - 1; j >= 0; --j) {
    const size_t i = t + (size_t)j * stride;
    if (i < n) {
      if (fe_is_zero(v[j])) { fe_store(out + i, v[j]); }
      else { fe_store(out + i, fe_mul(acc, pre[j])); acc = fe_mul(acc, v[j]); }
    }
  }
}

template <class P>
int vec_op_impl(zkc_ctx* ctx, int op, const Fe<P>* a, const Fe<P>* b, Fe<P>* out, size_t n) {
  if (n == 0) return ZKC_OK;
  if (op == ZKC_OP_INV) {
    const size_t stride = (n + BI_CHUNK - 1) / BI_CHUNK;
    k_batch_inv<P><<<(unsigned)((stride + 127) / 128), 128, 0, ctx->stream>>>(a, out, n, stride);
  } else {
    k_vec_op<P><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(op, a, b, out, n);
  }
  ZKC_LAUNCH_CHECK(ctx);
  return ZKC_OK;
}

int fr_batch_invert_scan(zkc_ctx* ctx, const Fr* a, Fr* out, uint64_t n);   // poly.cu
// small batches: per-thread Montgomery trick (one Fermat inversion per 8 elements); large ones: two product scans
// and a single inversion (6 products per element instead of ~50)
int fr_batch_invert(zkc_ctx* ctx, const Fr* a, Fr* out, size_t n) {
  if (n >= (1u << 15)) return fr_batch_invert_scan(ctx, a, out, n);
  return vec_op_impl<FrP>(ctx, ZKC_OP_INV, a, nullptr, out, n);
}

__global__ void k_mul_inplace(Fr* a, const Fr* b, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) fe_store(a + i, fe_mul(fe_load(a + i), fe_load(b + i)));
}

}  // namespace zkc

namespace zkc { int fr_powers(zkc_ctx* ctx, Fr* out, uint64_t n, const Fr& base, const Fr& first); }
extern "C" int zkc_fr_powers_dev(zkc_ctx* ctx, zkc_fr* out, size_t n, const zkc_fr* base, const zkc_fr* first) {
  if (!ctx || !out || !base || !first) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_fr_powers_dev: null argument");
  CtxLock lock(ctx);
  Fr b, f; memcpy(b.v, base, 32); memcpy(f.v, first, 32);
  return fr_powers(ctx, (Fr*)out, n, b, f);
}

// poly::batch_invert_assigned: Assigned::Rational(num, den) cells -> num * den^-1 (den = 0 -> 0, as upstream's
// `Assigned::evaluate`); Trivial cells pass den = 1.
extern "C" int zkc_batch_invert_assigned_dev(zkc_ctx* ctx, const zkc_fr* num, const zkc_fr* den, zkc_fr* out, size_t n) {
  if (!ctx || !num || !den || !out) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_batch_invert_assigned_dev: null argument");
  if (n == 0) return ZKC_OK;
  CtxLock lock(ctx);
  ZKC_TRY(fr_batch_invert(ctx, (const Fr*)den, (Fr*)out, n));
  k_mul_inplace<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((Fr*)out, (const Fr*)num, n);
  ZKC_LAUNCH_CHECK(ctx);
  return ZKC_OK;
}

extern "C" int zkc_field_vec_op_dev(zkc_ctx* ctx, int field, int op, const void* a, const void* b, void* out, size_t n) {
  if (!ctx || !a || !out) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_field_vec_op_dev: null argument");
  if ((op == ZKC_OP_ADD || op == ZKC_OP_SUB || op == ZKC_OP_MUL) && !b) return set_err(ctx, ZKC_ERR_BAD_ARG, "binary op needs b");
  if (op < 0 || op > ZKC_OP_SQUARE) return set_err(ctx, ZKC_ERR_BAD_ARG, "unknown op");
  CtxLock lock(ctx);
  if (field == 0 && op == ZKC_OP_INV) return fr_batch_invert(ctx, (const Fr*)a, (Fr*)out, n);
  if (field == 0) return vec_op_impl<FrP>(ctx, op, (const Fr*)a, (const Fr*)b, (Fr*)out, n);
  if (field == 1) return vec_op_impl<FqP>(ctx, op, (const Fq*)a, (const Fq*)b, (Fq*)out, n);
  return set_err(ctx, ZKC_ERR_BAD_ARG, "field must be 0 (Fr) or 1 (Fq)");
}
