// Shared host-side plumbing for libzkcert_cuda.so: context, error handling, scratch memory.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include "../../include/zkcert_cuda.h"
#include "ff.cuh"

namespace zkc {

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
};

// Debug / sweep overrides, per ctx.  The environment (ZKC_MSM_C, ZKC_MSM_C_PRE, ZKC_MSM_T, ZKC_NTT_TWO_PASS_MAX,
// ZKC_STAGE_MIN_BYTES, ZKC_TEAM_POISON) is read ONCE, when the ctx is created — never on the proving path; tests and the
// sweeps under tools/ change a value on a live ctx with zkc_ctx_set_tunable.
struct Tunables {
  int msm_c = 0, msm_c_pre = 0, msm_T = 0, ntt_two_pass_max = 0;   // 0 = library default
  int msm_accum_occ = 0;                      // k_msm_accum CTAs per SM: 3 = no register cap (134), else 4 (128 registers)
  size_t stage_min_bytes = (size_t)4 << 20;   // host witnesses at least this large are uploaded in stages on a copy stream
  bool no_program_factoring = false;         // keys loaded while set keep the gate / lookup programs as parsed (host/cs.h optimize_program off)
  bool team_commit_by_column = false;         // team proving: batches of >= world commitments are dealt by column instead of by point range
  bool team_poison = false;                   // team proving: rows a rank never receives are filled with 0xff
};

}  // namespace zkc

// The opaque context: one per GPU.  All public calls lock `mu` (thread-safe per ctx).
struct zkc_ctx {
  int dev = 0;
  cudaStream_t stream = nullptr;      // stream all kernels are launched on
  cudaStream_t own_stream = nullptr;  // created with the ctx
  cudaStream_t side_stream = nullptr; // second stream: work that is not on the Fiat-Shamir critical path (SideScope)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_msm_main = nullptr, ev_msm_side = nullptr;
  cudaStream_t copy_stream = nullptr; // staged witness upload (zkc_prove): host-to-device copies that run under the first MSMs
  std::vector<cudaEvent_t> ev_copy;
  bool side_pending = false;
  bool overlap = true;                // zkc_ctx_set_overlap: 0 serialises side work on the main stream (clean per-kernel timing)
  std::string err;
  zkc::Tunables tune;
  struct zkc_prover* active_prover = nullptr;   // step API: one create_proof session per ctx at a time
  uint64_t launches = 0;
  std::recursive_mutex mu;
  int sm_count = 148;
  // grow-only scratch arenas (index = purpose), freed with the ctx
  zkc::DevBuf scratch[16];   // [0..8) main stream, [8..16) side stream
  // twiddle tables: log_n -> device table of omega_n^i, i < n/2 (canonical root of unity)
  std::map<uint32_t, zkc::Fr*> twiddles;
  void* pinned[2] = {nullptr, nullptr};  // small pinned staging buffers for results ([1]: side stream)
  size_t pinned_bytes[2] = {0, 0};
  // optional per-phase CUDA-event timers (zkc_profile_*): name -> accumulated ms / count
  bool profiling = false;
  struct ProfRec { std::string name; cudaEvent_t e0, e1; };
  std::vector<ProfRec> prof_pending;
  std::vector<cudaEvent_t> prof_pool;
  std::map<std::string, std::pair<double, uint64_t>> prof_acc;
  std::map<std::string, uint64_t> stats;   // work counters (e.g. msm.madds), reported with the profile
  // team proving (dist.cuh): this GPU's place among the GPUs that share one create_proof
  int team_rank = 0, team_world = 1;
  int team_rot = 0;                        // rank that takes the first column block of the next partitioned batch (load balance)
  bool team_emulate = false;               // zkc_team_emulate: all shards run here, one after the other, no collectives
  void* team_comm[2] = {nullptr, nullptr}; // ncclComm_t: [0] collectives issued on the main stream, [1] on the side stream
};

namespace zkc {

inline int set_err(zkc_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}

#define ZKC_CUDA_TRY(ctx, expr)                                                                     \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      return zkc::set_err(ctx, _e == cudaErrorMemoryAllocation ? ZKC_ERR_OOM : ZKC_ERR_CUDA,       \
                          std::string(#expr) + ": " + cudaGetErrorString(_e));                     \
    }                                                                                               \
  } while (0)

#define ZKC_TRY(expr)                 \
  do {                                \
    int _s = (expr);                  \
    if (_s != ZKC_OK) return _s;      \
  } while (0)

#define ZKC_LAUNCH_CHECK(ctx)                                  \
  do {                                                         \
    (ctx)->launches++;                                         \
    ZKC_CUDA_TRY(ctx, cudaGetLastError());                     \
  } while (0)

// Ensure scratch arena `slot` holds at least `bytes`; contents are NOT preserved on growth.
inline int scratch_reserve(zkc_ctx* ctx, int slot, size_t bytes, void** out) {
  DevBuf& b = ctx->scratch[slot + (ctx->stream == ctx->side_stream && ctx->side_stream ? 8 : 0)];
  if (b.bytes < bytes) {
    if (b.p) { ZKC_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)); ZKC_CUDA_TRY(ctx, cudaFree(b.p)); b.p = nullptr; b.bytes = 0; }
    size_t want = bytes + (bytes >> 3);
    ZKC_CUDA_TRY(ctx, cudaMalloc(&b.p, want));
    b.bytes = want;
  }
  *out = b.p;
  return ZKC_OK;
}

inline int pinned_reserve(zkc_ctx* ctx, size_t bytes, void** out, int i = 0) {
  if (ctx->pinned_bytes[i] < bytes) {
    if (ctx->pinned[i]) { ZKC_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)); ZKC_CUDA_TRY(ctx, cudaFreeHost(ctx->pinned[i])); ctx->pinned[i] = nullptr; ctx->pinned_bytes[i] = 0; }
    ZKC_CUDA_TRY(ctx, cudaMallocHost(&ctx->pinned[i], bytes + 4096));
    ctx->pinned_bytes[i] = bytes + 4096;
  }
  *out = ctx->pinned[i];
  return ZKC_OK;
}

// RAII CUDA-event timer around a phase; active only while profiling is enabled on the ctx.
struct ProfScope {
  zkc_ctx* c; int idx = -1;
  ProfScope(zkc_ctx* ctx, const char* name) : c(ctx) {
    if (!c->profiling) return;
    cudaEvent_t ev[2];
    for (int i = 0; i < 2; ++i) {
      if (!c->prof_pool.empty()) { ev[i] = c->prof_pool.back(); c->prof_pool.pop_back(); }
      else cudaEventCreate(&ev[i]);
    }
    cudaEventRecord(ev[0], c->stream);
    c->prof_pending.push_back({name, ev[0], ev[1]});
    idx = (int)c->prof_pending.size() - 1;
  }
  ~ProfScope() { if (idx >= 0) cudaEventRecord(c->prof_pending[idx].e1, c->stream); }
};

// Enqueue the enclosed work on the side stream, ordered after everything issued so far on the main stream.
// side_join() makes the main stream wait for all side work issued so far.
struct SideScope {
  zkc_ctx* c; cudaStream_t saved;
  explicit SideScope(zkc_ctx* ctx) : c(ctx), saved(ctx->stream) {
    if (!c->overlap) return;
    cudaEventRecord(c->ev_fork, saved);
    cudaStreamWaitEvent(c->side_stream, c->ev_fork, 0);
    c->stream = c->side_stream;
  }
  ~SideScope() {
    if (!c->overlap) return;
    cudaEventRecord(c->ev_join, c->side_stream); c->side_pending = true; c->stream = saved;
  }
};
inline void side_join(zkc_ctx* c) {
  if (c->side_pending) { cudaStreamWaitEvent(c->stream, c->ev_join, 0); c->side_pending = false; }
}

struct CtxLock {
  zkc_ctx* c;
  explicit CtxLock(zkc_ctx* ctx) : c(ctx) { c->mu.lock(); cudaSetDevice(c->dev); }
  ~CtxLock() { c->mu.unlock(); }
};

// scratch slots
enum { SCR_NTT = 0, SCR_MSM = 1, SCR_MSM2 = 2, SCR_HOSTIO = 3, SCR_HOSTIO2 = 4, SCR_MISC = 5, SCR_MISC2 = 6, SCR_MISC3 = 7 };

// Fr constants shared by host code (canonical integers, little-endian 32-bit words)
static const uint32_t FR_ROOT_OF_UNITY_RAW[8] = {0x60c37c9cu, 0xd34f1ed9u, 0xd39329c8u, 0x3215cf6du, 0x3dd31f74u, 0x98865ea9u, 0x166d18b7u, 0x03ddb9f5u};
static const uint32_t FR_ZETA_RAW[8] = {0x36636f23u, 0xb8ca0b2du, 0xec2bc5e9u, 0xcc37a73fu, 0x3fd84104u, 0x048b6e19u, 0xe131a029u, 0x30644e72u};
static const uint32_t FR_ZETA_ALT_RAW[8] = {0xb99c90ddu, 0x8b17ea66u, 0x8d8daaa7u, 0x5bfc4108u, 0x41a91758u, 0xb3c4d79du, 0u, 0u};
static const uint32_t FR_DELTA_RAW[8] = {0xe533e9a2u, 0x870e56bbu, 0x5e963f25u, 0x5b5f898eu, 0xd4c86e71u, 0x64ec26aau, 0x22c6f0cau, 0x09226b6eu};
static const uint32_t FR_S = 28;

inline Fr fr_from_raw_words(const uint32_t w[8]) { Fr t; for (int i = 0; i < 8; ++i) t.v[i] = w[i]; return fe_from_canonical(t); }
inline Fr fr_from_u64(uint64_t x) { Fr t = fe_zero<FrP>(); t.v[0] = (uint32_t)x; t.v[1] = (uint32_t)(x >> 32); return fe_from_canonical(t); }
inline Fr fr_root_of_unity(uint32_t log_n) {  // canonical 2^log_n-th root: ROOT^(2^(S - log_n))
  Fr w = fr_from_raw_words(FR_ROOT_OF_UNITY_RAW);
  for (uint32_t i = log_n; i < FR_S; ++i) w = fe_sqr(w);
  return w;
}
inline Fr fr_pow(Fr x, uint64_t e) { return fe_pow_u64(x, e); }

}  // namespace zkc
