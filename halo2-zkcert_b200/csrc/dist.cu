// Team plumbing of libzkcert_cuda.so: NCCL communicator per ctx (bound with dlopen), the partition helpers and the
// four collectives the sharded create_proof needs (dist.cuh).  SURVEY.md §8e; the reference has no multi-GPU path —
// its CPU `best_multiexp` splits by point range across rayon threads, which is the split kept here across GPUs.
#include "dist.cuh"
#include <dlfcn.h>
#include <algorithm>
#include <cstdlib>

namespace zkc {
namespace {

// the slice of the NCCL ABI used here (nccl.h 2.x: stable since 2.7 for send/recv)
struct NcclId { char internal[128]; };
typedef void* NcclComm;
enum { kNcclUint8 = 1 };
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // a process that already holds torch's bundled libnccl gets that copy (same soname); otherwise the system one
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) return;
    bool all = true;
    auto sym = [&](const char* n) { void* p = dlsym(api.handle, n); if (!p) all = false; return p; };
    api.GetUniqueId = (int (*)(NcclId*))sym("ncclGetUniqueId");
    api.CommInitRank = (int (*)(NcclComm*, int, NcclId, int))sym("ncclCommInitRank");
    api.CommDestroy = (int (*)(NcclComm))sym("ncclCommDestroy");
    api.AllGather = (int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t))sym("ncclAllGather");
    api.Broadcast = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))sym("ncclBroadcast");
    api.Send = (int (*)(const void*, size_t, int, int, NcclComm, cudaStream_t))sym("ncclSend");
    api.Recv = (int (*)(void*, size_t, int, int, NcclComm, cudaStream_t))sym("ncclRecv");
    api.GroupStart = (int (*)())sym("ncclGroupStart");
    api.GroupEnd = (int (*)())sym("ncclGroupEnd");
    api.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    api.ok = all;
  });
  return api;
}

int nccl_fail(zkc_ctx* ctx, const char* what, int rc) {
  return set_err(ctx, ZKC_ERR_CUDA, std::string(what) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(rc) : "nccl error"));
}
#define ZKC_NCCL_TRY(ctx, expr)                                 \
  do {                                                          \
    int _rc = (expr);                                           \
    if (_rc != 0) return nccl_fail(ctx, #expr, _rc);            \
  } while (0)

bool real_comm(const zkc_ctx* ctx) { return ctx->team_world > 1 && !ctx->team_emulate; }
// one communicator per stream: collectives of the two streams may be in flight at the same time
NcclComm comm_of(const zkc_ctx* ctx) { return (NcclComm)ctx->team_comm[ctx->side_stream && ctx->stream == ctx->side_stream ? 1 : 0]; }

}  // namespace

std::vector<int> team_ranks(const zkc_ctx* ctx) {
  std::vector<int> r;
  if (ctx->team_emulate) for (int i = 0; i < ctx->team_world; ++i) r.push_back(i);
  else r.push_back(ctx->team_rank);
  return r;
}

int team_allgather(zkc_ctx* ctx, void* buf, size_t bytes_per_rank) {
  if (!real_comm(ctx) || !bytes_per_rank) return ZKC_OK;
  ZKC_NCCL_TRY(ctx, nccl().AllGather((char*)buf + (size_t)ctx->team_rank * bytes_per_rank, buf, bytes_per_rank, kNcclUint8, comm_of(ctx),
                                     ctx->stream));
  return ZKC_OK;
}

int team_bcast_cols(zkc_ctx* ctx, Fr* base, uint64_t stride, uint64_t len, uint32_t ncols) {
  if (!real_comm(ctx) || !ncols || !len) return ZKC_OK;
  ProfScope _p(ctx, "team.bcast_cols");
  ZKC_NCCL_TRY(ctx, nccl().GroupStart());
  for (int r = 0; r < ctx->team_world; ++r) {
    uint32_t c0, c1;
    team_cols(ctx, ncols, r, &c0, &c1);
    if (c1 <= c0) continue;
    // an owner's columns are adjacent: when they are also dense (stride == len) they travel as ONE message — a 512 MB
    // message moves at twice the rate of eight 64 MB ones (profiles/r02_p2p_2gpu.json)
    const uint32_t per = stride == len ? c1 - c0 : 1;
    for (uint32_t c = c0; c < c1; c += per) {
      Fr* p = base + (uint64_t)c * stride;
      int rc = nccl().Broadcast(p, p, (size_t)per * len * sizeof(Fr), kNcclUint8, r, comm_of(ctx), ctx->stream);
      if (rc != 0) { nccl().GroupEnd(); return nccl_fail(ctx, "ncclBroadcast", rc); }
    }
  }
  ZKC_NCCL_TRY(ctx, nccl().GroupEnd());
  return ZKC_OK;
}

int team_bcast_blocks(zkc_ctx* ctx, Fr* base, uint64_t len, uint32_t nblocks) {
  if (!real_comm(ctx) || !nblocks || !len) return ZKC_OK;
  ProfScope _p(ctx, "team.bcast_cols");
  ZKC_NCCL_TRY(ctx, nccl().GroupStart());
  for (uint32_t b = 0; b < nblocks; ++b) {
    Fr* p = base + (uint64_t)b * len;
    int rc = nccl().Broadcast(p, p, len * sizeof(Fr), kNcclUint8, (int)(b % (uint32_t)ctx->team_world), comm_of(ctx), ctx->stream);
    if (rc != 0) { nccl().GroupEnd(); return nccl_fail(ctx, "ncclBroadcast", rc); }
  }
  ZKC_NCCL_TRY(ctx, nccl().GroupEnd());
  return ZKC_OK;
}

int team_exchange(zkc_ctx* ctx, const std::vector<TeamXfer>& ops, const char* what) {
  if (!real_comm(ctx) || ops.empty()) return ZKC_OK;
  ProfScope _p(ctx, what);
  ZKC_NCCL_TRY(ctx, nccl().GroupStart());
  for (const TeamXfer& x : ops) {
    const int rc = x.send ? nccl().Send(x.p, x.bytes, kNcclUint8, x.peer, comm_of(ctx), ctx->stream)
                          : nccl().Recv(x.p, x.bytes, kNcclUint8, x.peer, comm_of(ctx), ctx->stream);
    if (rc != 0) { nccl().GroupEnd(); return nccl_fail(ctx, "ncclSend/ncclRecv", rc); }
  }
  ZKC_NCCL_TRY(ctx, nccl().GroupEnd());
  return ZKC_OK;
}

int team_allgather_rows(zkc_ctx* ctx, Fr* col, uint64_t en) {
  if (!real_comm(ctx)) return ZKC_OK;
  ProfScope _p(ctx, "team.allgather_rows");
  return team_allgather_flat(ctx, col, en);
}

int team_allgather_flat(zkc_ctx* ctx, Fr* col, uint64_t en) {
  if (!real_comm(ctx)) return ZKC_OK;
  const int W = ctx->team_world;
  if (en % (uint64_t)W == 0) return team_allgather(ctx, col, (size_t)(en / W) * sizeof(Fr));
  ZKC_NCCL_TRY(ctx, nccl().GroupStart());
  for (int r = 0; r < W; ++r) {
    uint64_t lo, hi;
    shard_range(en, W, r, &lo, &hi);
    if (hi == lo) continue;
    int rc = nccl().Broadcast(col + lo, col + lo, (hi - lo) * sizeof(Fr), kNcclUint8, r, comm_of(ctx), ctx->stream);
    if (rc != 0) { nccl().GroupEnd(); return nccl_fail(ctx, "ncclBroadcast", rc); }
  }
  ZKC_NCCL_TRY(ctx, nccl().GroupEnd());
  return ZKC_OK;
}

}  // namespace zkc

using namespace zkc;

extern "C" int zkc_team_unique_id(uint8_t id[ZKC_TEAM_ID_BYTES]) {
  if (!id) return ZKC_ERR_BAD_ARG;
  if (!nccl().ok) return ZKC_ERR_CUDA;
  NcclId nid;
  if (nccl().GetUniqueId(&nid) != 0) return ZKC_ERR_CUDA;
  memcpy(id, nid.internal, ZKC_TEAM_ID_BYTES);
  return ZKC_OK;
}

extern "C" int zkc_team_init(zkc_ctx* ctx, int rank, int world, const uint8_t id[ZKC_TEAM_ID_BYTES]) {
  if (!ctx || world < 1 || rank < 0 || rank >= world) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_team_init: bad rank / world");
  CtxLock lock(ctx);
  if (ctx->team_comm[0] || ctx->team_world > 1) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_team_init: the context already belongs to a team");
  if (world == 1) return ZKC_OK;
  if (!id) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_team_init: null id");
  if (!nccl().ok) return set_err(ctx, ZKC_ERR_CUDA, "zkc_team_init: libnccl.so.2 not found (team proving needs NCCL)");
  NcclId nid;
  memcpy(nid.internal, id, ZKC_TEAM_ID_BYTES);
  NcclComm comm = nullptr, comm2 = nullptr;
  ZKC_NCCL_TRY(ctx, nccl().CommInitRank(&comm, world, nid, rank));
  // second communicator (side stream): its id travels over the first one
  NcclId nid2;
  memset(&nid2, 0, sizeof nid2);
  if (rank == 0 && nccl().GetUniqueId(&nid2) != 0) { nccl().CommDestroy(comm); return set_err(ctx, ZKC_ERR_CUDA, "zkc_team_init: ncclGetUniqueId failed"); }
  void* d = nullptr;
  cudaError_t e = cudaMalloc(&d, sizeof nid2);
  if (e == cudaSuccess) e = cudaMemcpy(d, &nid2, sizeof nid2, cudaMemcpyHostToDevice);
  int rc = e == cudaSuccess ? nccl().Broadcast(d, d, sizeof nid2, kNcclUint8, 0, comm, ctx->stream) : 0;
  if (e == cudaSuccess && rc == 0) e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess && rc == 0) e = cudaMemcpy(&nid2, d, sizeof nid2, cudaMemcpyDeviceToHost);
  if (d) cudaFree(d);
  if (e == cudaSuccess && rc == 0) rc = nccl().CommInitRank(&comm2, world, nid2, rank);
  if (e != cudaSuccess || rc != 0) {
    nccl().CommDestroy(comm);
    return e != cudaSuccess ? set_err(ctx, ZKC_ERR_CUDA, std::string("zkc_team_init: ") + cudaGetErrorString(e)) : nccl_fail(ctx, "zkc_team_init (second communicator)", rc);
  }
  ctx->team_comm[0] = comm; ctx->team_comm[1] = comm2; ctx->team_rank = rank; ctx->team_world = world; ctx->team_emulate = false;
  // NCCL connects peers lazily at the first collective of each kind: do that here, on both communicators, not inside a proof
  void* w = nullptr;
  if (cudaMalloc(&w, 256 * (size_t)world * 2) == cudaSuccess) {
    cudaMemset(w, 0, 256 * (size_t)world * 2);
    for (NcclComm cm : {comm, comm2}) {
      nccl().AllGather((char*)w + 256 * (size_t)rank, w, 256, kNcclUint8, cm, ctx->stream);
      nccl().Broadcast(w, w, 256, kNcclUint8, 0, cm, ctx->stream);
      nccl().GroupStart();
      for (int p = 0; p < world; ++p) {
        if (p == rank) continue;
        nccl().Send((char*)w + 256 * (size_t)rank, 256, kNcclUint8, p, cm, ctx->stream);
        nccl().Recv((char*)w + 256 * ((size_t)world + p), 256, kNcclUint8, p, cm, ctx->stream);
      }
      nccl().GroupEnd();
    }
    cudaStreamSynchronize(ctx->stream);
    cudaFree(w);
  }
  return ZKC_OK;
}

extern "C" int zkc_team_emulate(zkc_ctx* ctx, int world) {
  if (!ctx || world < 1 || world > 64) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_team_emulate: bad world");
  CtxLock lock(ctx);
  if (ctx->team_comm[0]) return set_err(ctx, ZKC_ERR_BAD_ARG, "zkc_team_emulate: the context belongs to a real team");
  ZKC_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->team_world = world; ctx->team_rank = 0; ctx->team_emulate = world > 1;
  return ZKC_OK;
}

extern "C" int zkc_team_leave(zkc_ctx* ctx) {
  if (!ctx) return ZKC_ERR_BAD_ARG;
  CtxLock lock(ctx);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->side_stream) cudaStreamSynchronize(ctx->side_stream);
  for (auto& c : ctx->team_comm) if (c) { nccl().CommDestroy((NcclComm)c); c = nullptr; }
  ctx->team_world = 1; ctx->team_rank = 0; ctx->team_emulate = false;
  return ZKC_OK;
}

extern "C" int zkc_team_info(const zkc_ctx* ctx, int* rank, int* world, int* emulated) {
  if (!ctx) return ZKC_ERR_BAD_ARG;
  if (rank) *rank = ctx->team_rank;
  if (world) *world = ctx->team_world;
  if (emulated) *emulated = ctx->team_emulate ? 1 : 0;
  return ZKC_OK;
}

// The partition arithmetic, exposed for host-side planning and tests (no device work).
extern "C" int zkc_team_shard_range(uint64_t total, int world, int rank, uint64_t* lo, uint64_t* hi) {
  if (world < 1 || rank < 0 || rank >= world || !lo || !hi) return ZKC_ERR_BAD_ARG;
  shard_range(total, world, rank, lo, hi);
  return ZKC_OK;
}
// residue classes [c0, c1) of the extended coset that `rank`'s row block of the class-major extended domain touches; only the
// first `num_classes` = degree - 1 classes are evaluated (prover.cu: the classes the rank transforms from the replicated
// coefficient forms)
extern "C" int zkc_team_classes(uint32_t k, uint32_t num_classes, int world, int rank, uint32_t* c0, uint32_t* c1) {
  if (world < 1 || rank < 0 || rank >= world || !c0 || !c1 || k > 27 || num_classes < 1 || num_classes > 256) return ZKC_ERR_BAD_ARG;
  uint64_t lo, hi;
  shard_range((uint64_t)num_classes << k, world, rank, &lo, &hi);
  if (hi == lo) { *c0 = *c1 = 0; return ZKC_OK; }
  *c0 = (uint32_t)(lo >> k); *c1 = (uint32_t)((hi - 1) >> k) + 1;
  return ZKC_OK;
}
