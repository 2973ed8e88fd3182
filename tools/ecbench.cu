// Latency / throughput microbenchmark of XYZZ point addition on the B200 (development tool).
#include <cstdio>
#include <cuda_runtime.h>
#include "../halo2-zkcert_b200/csrc/ec.cuh"
using namespace zkc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

// the variant round 1 measured: shuffles with a run-time quad mask and ?: selection (a divergent branch region per word)
__device__ __forceinline__ Fq fq_sel4_br(uint32_t r, const Fq& a0, const Fq& a1, const Fq& a2, const Fq& a3) {
  Fq o;
#pragma unroll
  for (int i = 0; i < 8; ++i) o.v[i] = r == 0 ? a0.v[i] : (r == 1 ? a1.v[i] : (r == 2 ? a2.v[i] : a3.v[i]));
  return o;
}
__device__ __forceinline__ Fq fq_bcast_m(uint32_t qmask, const Fq& v, int src) {
  Fq o;
#pragma unroll
  for (int i = 0; i < 8; ++i) o.v[i] = __shfl_sync(qmask, v.v[i], src);
  return o;
}
__device__ __forceinline__ void xyzz_add_quad_masked(G1Xyzz& p, const G1Xyzz& q) {
  const uint32_t lane = threadIdx.x & 31, r = lane & 3, base = lane & ~3u;
  const uint32_t qmask = 0xFu << base;
  Fq m = fe_mul(fq_sel4_br(r, p.x, q.x, p.y, q.y), fq_sel4_br(r, q.zz, p.zz, q.zzz, p.zzz));
  const Fq u1 = fq_bcast_m(qmask, m, base), u2 = fq_bcast_m(qmask, m, base + 1), s1 = fq_bcast_m(qmask, m, base + 2), s2 = fq_bcast_m(qmask, m, base + 3);
  const Fq pp_ = fe_sub(u2, u1), rr = fe_sub(s2, s1);
  m = fe_mul(fq_sel4_br(r, pp_, rr, p.zz, p.zzz), fq_sel4_br(r, pp_, rr, q.zz, q.zzz));
  const Fq pp = fq_bcast_m(qmask, m, base), r2 = fq_bcast_m(qmask, m, base + 1), zz12 = fq_bcast_m(qmask, m, base + 2), zzz12 = fq_bcast_m(qmask, m, base + 3);
  m = fe_mul(fq_sel4_br(r, pp_, u1, zz12, zz12), pp);
  const Fq ppp = fq_bcast_m(qmask, m, base), qq = fq_bcast_m(qmask, m, base + 1), zz3 = fq_bcast_m(qmask, m, base + 2);
  const Fq x3 = fe_sub(fe_sub(r2, ppp), fe_dbl(qq));
  m = fe_mul(fq_sel4_br(r, rr, s1, zzz12, zzz12), fq_sel4_br(r, fe_sub(qq, x3), ppp, ppp, ppp));
  const Fq t1 = fq_bcast_m(qmask, m, base), t2 = fq_bcast_m(qmask, m, base + 1), zzz3 = fq_bcast_m(qmask, m, base + 2);
  p.x = x3; p.y = fe_sub(t1, t2); p.zz = zz3; p.zzz = zzz3;
}

__global__ void k_chain(G1Xyzz* io, int iters, int mode) {
  G1Xyzz p = xyzz_load(io + 0), q = xyzz_load(io + 1);
  for (int i = 0; i < iters; ++i) { if (mode == 0) xyzz_add(p, q); else if (mode == 1) xyzz_add_quad_masked(p, q); else xyzz_add_quad(p, q); }
  if (threadIdx.x == 0 && blockIdx.x == 0) xyzz_store(io + 2 + mode, p);
}
__global__ void k_mulchain(Fq* io, int iters) {
  Fq a = fe_load(io), b = fe_load(io + 1);
  for (int i = 0; i < iters; ++i) a = fe_mul(a, b);
  if (threadIdx.x == 0 && blockIdx.x == 0) fe_store(io + 2, a);
}

int main() {
  // two arbitrary (not on-curve) XYZZ values suffice: the formulas are polynomial; both variants must agree
  G1Xyzz h[5];
  uint64_t s = 1234567;
  uint32_t* w = (uint32_t*)h;
  for (int i = 0; i < 64; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; w[i] = (uint32_t)s; if (i % 8 == 7) w[i] &= 0x1fffffff; }
  G1Xyzz* d; CK(cudaMalloc(&d, sizeof(h))); CK(cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice));
  Fq* df; CK(cudaMalloc(&df, 3 * sizeof(Fq))); CK(cudaMemcpy(df, h, 2 * sizeof(Fq), cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 2000;
  float ms;
  for (int blocks : {1, 148 * 4}) for (int threads : {32, 256}) {
    for (int mode = 0; mode < 3; ++mode) {
      k_chain<<<blocks, threads>>>(d, 10, mode); CK(cudaDeviceSynchronize());
      cudaEventRecord(e0); k_chain<<<blocks, threads>>>(d, iters, mode); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      cudaEventElapsedTime(&ms, e0, e1);
      printf("xyzz_add %-5s blocks=%4d threads=%3d : %.3f us per dependent add\n", mode == 0 ? "plain" : mode == 1 ? "qmask" : "quad", blocks, threads, ms * 1e3 / iters);
    }
    k_mulchain<<<blocks, threads>>>(df, 10); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); k_mulchain<<<blocks, threads>>>(df, iters * 10); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms, e0, e1);
    printf("fe_mul          blocks=%4d threads=%3d : %.3f us per dependent mul\n", blocks, threads, ms * 1e3 / (iters * 10));
  }
  CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
  printf("variants agree: %d %d\n", memcmp(&h[2], &h[3], sizeof(G1Xyzz)) == 0, memcmp(&h[2], &h[4], sizeof(G1Xyzz)) == 0);
  return 0;
}
