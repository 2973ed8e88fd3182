// Integer-pipe microbenchmark + on-device self-check of ff.cuh (run on the B200 via gpurun).
//   * raw IMAD / IMAD.WIDE.U32.X issue rates (register-resident dependent chains, many warps)
//   * fe_mul throughput (the IMAD.WIDE even/odd Montgomery product) for Fr and Fq
//   * a 32-bit mad.lo.cc/madc.hi.cc CIOS product for comparison (the "classic" formulation)
// Prints one JSON object; the numbers feed DESIGN.md's integer roofline.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../halo2-zkcert_b200/csrc/ff.cuh"
using namespace zkc;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void k_imad(uint32_t* out, int iters) {
  uint32_t a = threadIdx.x * 2654435761u + 1, b = blockIdx.x * 40503u + 7, c0 = 1, c1 = 2, c2 = 3, c3 = 4;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) { c0 = c0 * a + b; c1 = c1 * a + b; c2 = c2 * a + b; c3 = c3 * a + b; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0 ^ c1 ^ c2 ^ c3;
}
__global__ void k_imadwide(u64* out, int iters) {
  uint32_t a = threadIdx.x * 2654435761u + 1, b = blockIdx.x * 40503u + 7;
  u64 c0 = 1, c1 = 2, c2 = 3, c3 = 4;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      asm("{\n\t.reg .u64 t;\n\t"
          "mul.wide.u32 t,%4,%5;\n\tadd.cc.u64 %0,%0,t;\n\t"
          "mul.wide.u32 t,%4,%5;\n\taddc.cc.u64 %1,%1,t;\n\t"
          "mul.wide.u32 t,%4,%5;\n\taddc.cc.u64 %2,%2,t;\n\t"
          "mul.wide.u32 t,%4,%5;\n\taddc.u64 %3,%3,t;\n\t}"
          : "+l"(c0), "+l"(c1), "+l"(c2), "+l"(c3) : "r"(a), "r"(b));
      a += (uint32_t)c3;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0 ^ c1 ^ c2 ^ c3;
}

template <class P> __global__ void k_mul(Fe<P>* out, const Fe<P>* in, int iters) {
  int tid = blockIdx.x * blockDim.x + threadIdx.x;
  Fe<P> a = fe_load(in + 2 * tid), b = fe_load(in + 2 * tid + 1);
  for (int it = 0; it < iters; ++it) { a = fe_mul(a, b); b = fe_mul(b, a); }
  fe_store(out + tid, fe_add(a, b));
}
template <class P> __global__ void k_ops(Fe<P>* out, const Fe<P>* in, int n) {
  int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n) return;
  Fe<P> a = fe_load(in + 2 * tid), b = fe_load(in + 2 * tid + 1);
  fe_store(out + 4 * tid, fe_mul(a, b));
  fe_store(out + 4 * tid + 1, fe_add(a, b));
  fe_store(out + 4 * tid + 2, fe_sub(a, b));
  fe_store(out + 4 * tid + 3, fe_mul(fe_inv(a), a));
}

// classic 32-bit CIOS with mad.lo.cc / madc.hi.cc (for comparison only)
__device__ __forceinline__ void classic_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#ifdef __CUDA_ARCH__
  uint32_t t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0, t6 = 0, t7 = 0, t8 = 0;
#define ROWLO(x0,x1,x2,x3,x4,x5,x6,x7,mm) asm("mad.lo.cc.u32 %0,%9,%17,%0;\n\tmadc.lo.cc.u32 %1,%10,%17,%1;\n\tmadc.lo.cc.u32 %2,%11,%17,%2;\n\tmadc.lo.cc.u32 %3,%12,%17,%3;\n\tmadc.lo.cc.u32 %4,%13,%17,%4;\n\tmadc.lo.cc.u32 %5,%14,%17,%5;\n\tmadc.lo.cc.u32 %6,%15,%17,%6;\n\tmadc.lo.cc.u32 %7,%16,%17,%7;\n\taddc.u32 %8,%8,0;" : "+r"(t0),"+r"(t1),"+r"(t2),"+r"(t3),"+r"(t4),"+r"(t5),"+r"(t6),"+r"(t7),"+r"(t8) : "r"(x0),"r"(x1),"r"(x2),"r"(x3),"r"(x4),"r"(x5),"r"(x6),"r"(x7),"r"(mm))
#define ROWHI(x0,x1,x2,x3,x4,x5,x6,x7,mm) asm("mad.hi.cc.u32 %0,%8,%16,%0;\n\tmadc.hi.cc.u32 %1,%9,%16,%1;\n\tmadc.hi.cc.u32 %2,%10,%16,%2;\n\tmadc.hi.cc.u32 %3,%11,%16,%3;\n\tmadc.hi.cc.u32 %4,%12,%16,%4;\n\tmadc.hi.cc.u32 %5,%13,%16,%5;\n\tmadc.hi.cc.u32 %6,%14,%16,%6;\n\tmadc.hi.u32 %7,%15,%16,%7;" : "+r"(t1),"+r"(t2),"+r"(t3),"+r"(t4),"+r"(t5),"+r"(t6),"+r"(t7),"+r"(t8) : "r"(x0),"r"(x1),"r"(x2),"r"(x3),"r"(x4),"r"(x5),"r"(x6),"r"(x7),"r"(mm))
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint32_t bi = b[i];
    ROWLO(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], bi);
    ROWHI(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], bi);
    uint32_t m = t0 * FrP::INV;
    ROWLO(FrP::M(0), FrP::M(1), FrP::M(2), FrP::M(3), FrP::M(4), FrP::M(5), FrP::M(6), FrP::M(7), m);
    ROWHI(FrP::M(0), FrP::M(1), FrP::M(2), FrP::M(3), FrP::M(4), FrP::M(5), FrP::M(6), FrP::M(7), m);
    t0 = t1; t1 = t2; t2 = t3; t3 = t4; t4 = t5; t5 = t6; t6 = t7; t7 = t8; t8 = 0;
  }
  uint32_t t[8] = {t0, t1, t2, t3, t4, t5, t6, t7};
  reduce_once<FrP>(t);
  for (int i = 0; i < 8; ++i) r[i] = t[i];
#endif
}
__global__ void k_classic(Fr* out, const Fr* in, int iters) {
  int tid = blockIdx.x * blockDim.x + threadIdx.x;
  Fr a = fe_load(in + 2 * tid), b = fe_load(in + 2 * tid + 1);
  for (int it = 0; it < iters; ++it) { classic_mul(a.v, a.v, b.v); classic_mul(b.v, b.v, a.v); }
  fe_store(out + tid, fe_add(a, b));
}

template <class F> float time_ms(F f, int reps = 5) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

template <class P> int check_ops(const char* name) {
  const int n = 4096;
  std::vector<Fe<P>> in(2 * n), out(4 * n);
  uint64_t s = 88172645463325252ULL;
  for (auto& e : in) { for (int i = 0; i < 8; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; e.v[i] = (uint32_t)s; } e.v[7] &= 0x1fffffff; }
  for (int i = 0; i < 8; ++i) { in[0].v[i] = 0; in[1].v[i] = P::M(i); } in[1].v[0] -= 1;   // 0 * (p-1)
  for (int i = 0; i < 8; ++i) { in[2].v[i] = P::M(i); in[3].v[i] = P::M(i); } in[2].v[0] -= 1; in[3].v[0] -= 1;
  Fe<P>*din, *dout; CK(cudaMalloc(&din, sizeof(Fe<P>) * 2 * n)); CK(cudaMalloc(&dout, sizeof(Fe<P>) * 4 * n));
  CK(cudaMemcpy(din, in.data(), sizeof(Fe<P>) * 2 * n, cudaMemcpyHostToDevice));
  k_ops<P><<<n / 128, 128>>>(dout, din, n); CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out.data(), dout, sizeof(Fe<P>) * 4 * n, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int i = 0; i < n; ++i) {
    Fe<P> a = in[2 * i], b = in[2 * i + 1];
    Fe<P> m = fe_mul(a, b), ad = fe_add(a, b), sb = fe_sub(a, b);
    Fe<P> one = fe_is_zero(a) ? fe_zero<P>() : fe_one<P>();
    if (!fe_eq(m, out[4 * i]) || !fe_eq(ad, out[4 * i + 1]) || !fe_eq(sb, out[4 * i + 2]) || !fe_eq(one, out[4 * i + 3])) {
      if (bad < 3) printf("# %s mismatch at %d\n", name, i);
      ++bad;
    }
  }
  cudaFree(din); cudaFree(dout);
  return bad;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  int bad_fr = check_ops<FrP>("fr"), bad_fq = check_ops<FqP>("fq");
  const int blocks = sms * 8, threads = 256, iters = 2000;
  uint32_t* o32; u64* o64; CK(cudaMalloc(&o32, blocks * threads * 4)); CK(cudaMalloc(&o64, blocks * threads * 8));
  float ms_imad = time_ms([&] { k_imad<<<blocks, threads>>>(o32, iters); });
  float ms_wide = time_ms([&] { k_imadwide<<<blocks, threads>>>(o64, iters); });
  double n_imad = (double)blocks * threads * iters * 64.0;
  std::vector<Fr> in(2 * blocks * threads);
  uint64_t s = 1234567;
  for (auto& e : in) { for (int i = 0; i < 8; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; e.v[i] = (uint32_t)s; } e.v[7] &= 0x1fffffff; }
  Fr *din, *dout; CK(cudaMalloc(&din, sizeof(Fr) * in.size())); CK(cudaMalloc(&dout, sizeof(Fr) * blocks * threads));
  CK(cudaMemcpy(din, in.data(), sizeof(Fr) * in.size(), cudaMemcpyHostToDevice));
  const int mit = 200;
  float ms_fr = time_ms([&] { k_mul<FrP><<<blocks, threads>>>(dout, din, mit); });
  float ms_fq = time_ms([&] { k_mul<FqP><<<blocks, threads>>>((Fq*)dout, (const Fq*)din, mit); });
  float ms_cl = time_ms([&] { k_classic<<<blocks, threads>>>(dout, din, mit); });
  double n_mul = (double)blocks * threads * mit * 2.0;
  printf("{\"gpu\":\"%s\",\"sms\":%d,\"clock_khz\":%d,\"check_fr_bad\":%d,\"check_fq_bad\":%d,"
         "\"imad_per_s\":%.4g,\"imad_wide_per_s\":%.4g,\"fr_mul_per_s\":%.4g,\"fq_mul_per_s\":%.4g,\"classic_mul_per_s\":%.4g}\n",
         prop.name, sms, prop.clockRate, bad_fr, bad_fq, n_imad / (ms_imad * 1e-3), n_imad / (ms_wide * 1e-3),
         n_mul / (ms_fr * 1e-3), n_mul / (ms_fq * 1e-3), n_mul / (ms_cl * 1e-3));
  return (bad_fr || bad_fq) ? 2 : 0;
}
