// Team proving: one create_proof spread over the GPUs of a node.  One process per GPU;
// every rank runs the same deterministic host driver (same transcript, same RNG cursor) and the O(n)
// device work is partitioned:
//   * MSM / commitments  by POINT RANGE   — the split best_multiexp makes across CPU threads; the only
//                                           exchange is an all-gather of the c bit-plane sums per column
//   * lagrange_to_coeff  by COLUMN        — the owner broadcasts the coefficient form
//   * coeff_to_extended  by RESIDUE CLASS — the extended coset splits into 2^(extended_k - k) cosets of the size-n
//                                           subgroup (ntt.cu, dom_coeff_to_classes); a rank transforms the classes its
//                                           row block touches from the replicated coefficient forms: no exchange
//   * h(X) evaluation    by EXTENDED ROW  — contiguous blocks of the class-major extended coset (rotations stay inside a
//                                           class), then an all-gather of the quotient values
// The collectives are NCCL over NVLink/NVSwitch, issued on the stream the kernels run on.  libnccl is
// bound at run time (dlopen), so the library loads on hosts without it.
// ZKC_TEAM_EMULATE=W runs the W shards of every partitioned step one after the other on ONE GPU with the
// collectives elided (same arithmetic, same slicing) — the single-GPU test of the sharding logic.
#pragma once
#include "common.cuh"

namespace zkc {

struct Segment { uint64_t lo, len; };

// contiguous [start, end) of `total` items for `rank` (the point-range split of best_multiexp)
inline void shard_range(uint64_t total, int world, int rank, uint64_t* lo, uint64_t* hi) {
  const uint64_t base = total / (uint64_t)world, rem = total % (uint64_t)world;
  *lo = (uint64_t)rank * base + std::min<uint64_t>((uint64_t)rank, rem);
  *hi = *lo + base + ((uint64_t)rank < rem ? 1 : 0);
}

// ranks whose shards this process computes: {rank} normally, {0..world) under ZKC_TEAM_EMULATE
std::vector<int> team_ranks(const zkc_ctx* ctx);
inline bool team_active(const zkc_ctx* ctx) { return ctx->team_world > 1; }

// in-place all-gather of equally sized blocks: block r of `buf` (bytes_per_rank each) comes from rank r
int team_allgather(zkc_ctx* ctx, void* buf, size_t bytes_per_rank);
// column c of base[ncols][stride] (first `len` elements) is broadcast from the rank that owns it
int team_bcast_cols(zkc_ctx* ctx, Fr* base, uint64_t stride, uint64_t len, uint32_t ncols);
// block b of base[nblocks][len] is broadcast from rank b mod world
int team_bcast_blocks(zkc_ctx* ctx, Fr* base, uint64_t len, uint32_t nblocks);
// point-to-point transfers of one step, issued as ONE NCCL group (every send has its matching receive in the peer's group)
struct TeamXfer { int peer; bool send; void* p; size_t bytes; };
int team_exchange(zkc_ctx* ctx, const std::vector<TeamXfer>& ops, const char* what);
// in-place all-gather of the row blocks of one length-en column
int team_allgather_rows(zkc_ctx* ctx, Fr* col, uint64_t en);
// in-place all-gather of a flat array split with shard_range(total, world, r)
int team_allgather_flat(zkc_ctx* ctx, Fr* base, uint64_t total);
// columns [c0, c1) of a batch of `ncols` that rank `r` transforms
// (blocks are dealt starting at rank ctx->team_rot, which team_advance moves past the ranks that just received the
// larger blocks, so that batches of few columns do not all land on rank 0)
inline void team_cols(const zkc_ctx* ctx, uint32_t ncols, int r, uint32_t* c0, uint32_t* c1) {
  uint64_t lo, hi;
  const int W = ctx->team_world;
  shard_range(ncols, W, ((r - ctx->team_rot) % W + W) % W, &lo, &hi);
  *c0 = (uint32_t)lo; *c1 = (uint32_t)hi;
}
inline void team_advance(zkc_ctx* ctx, uint32_t ncols) { ctx->team_rot = (int)((ctx->team_rot + ncols % (uint32_t)ctx->team_world) % (uint32_t)ctx->team_world); }

}  // namespace zkc
