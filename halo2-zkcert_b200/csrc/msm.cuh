// MSM geometry and the asynchronous enqueue / finish split shared by msm.cu and prover.cu.
#pragma once
#include "common.cuh"
#include "ec.cuh"

namespace zkc {

struct MsmGeom {
  uint32_t c;          // window bits
  uint32_t W;          // number of windows: W*c >= 255
  uint32_t NB;         // buckets per window = 2^(c-1)
  uint32_t sets;       // bucket sets per column: W (generic bases) or 1 (precomputed tables)
  uint64_t n;          // points per column
  uint32_t ncols;
  uint32_t T;          // entries per accumulate thread
  uint64_t sstride;    // distance between scalar columns (n, or the full column length when a point range is processed)
  uint64_t bstride;    // distance between window tables (precomputed layout [w][i])
  ZKC_HD uint64_t nbtot() const { return (uint64_t)ncols * sets * NB; }
  ZKC_HD uint64_t emax() const { return (uint64_t)ncols * W * n; }
};


// one enqueued batch: results land in pinned host memory; `done` fires after the D2H copy
struct MsmPending {
  MsmGeom g;
  uint32_t nc = 0;
  uint64_t n = 0;
  void* hU = nullptr;
  size_t ubytes = 0;   // bytes of bit-plane sums per shard
  int shards = 1;      // team proving: one block of `ubytes` per rank, summed on the host
  bool by_column = false;   // team proving, column split: rank r's block holds the sums of ITS columns (shard_range(nc, shards, r))
  cudaEvent_t done = nullptr;
  bool active = false;
};

// `team`: the batch is a commitment against resident bases shared by a team (dist.cuh): each rank walks its point range
// and the bit-plane sums are all-gathered before the D2H copy.
int msm_enqueue(zkc_ctx* ctx, const Fr* scalars, const G1Affine* bases, uint64_t n, uint32_t nc, uint32_t c, bool precomputed, MsmPending* pend,
                int result_slot, bool team = false);
int msm_finish(zkc_ctx* ctx, MsmPending* pend, zkc_g1* out);
int srs_commit_enqueue(zkc_ctx* ctx, const zkc_srs* s, int basis, const Fr* poly, uint64_t len, MsmPending* pend);

}  // namespace zkc
