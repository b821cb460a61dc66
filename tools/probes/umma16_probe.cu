// Probe for tcgen05.mma kind::f16 (fp16 / bf16 operands, fp32 accumulate) next to kind::tf32 on sm_100a:
//  (1) check: the un-swizzled K-major canonical layout [k/8][row][8 halfs] (16-byte units, SBO = 128 B, LBO = chunk pitch)
//      with a row-shifted start address, against an exact CPU product of the same 16-bit values;
//  (2) rate:  cycles per MMA of one CTA and chip-wide TFLOP/s (CUDA events, 148 CTAs) for a given kind / N --
//      the measured tensor peak of SS-mode, cta_group::1, M = 128 MMAs that the conv kernels' roofline is quoted against.
// usage: umma16_probe check <fmt 0=f16 1=bf16> N shift | umma16_probe rate <kind 0=tf32 1=f16 2=bf16> N reps [nbuf nacc albo blbo nbbuf]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// kind 0: tf32 (a/b format 2), kind 1: f16 (format 0), kind 2: bf16 (format 1)
__device__ __forceinline__ uint32_t make_idesc(int kind, int M, int N) {
  const uint32_t fmt = kind == 0 ? 2u : (kind == 1 ? 0u : 1u);
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <int KIND>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (KIND == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
                 "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
                 "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity, long long max_spin) {
  for (long long i = 0; i < max_spin; ++i) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}

// A [144 rows][K], B [N][K] as raw 16-bit patterns from global (row-major), K = 32: two K = 16 MMAs.
template <int K>
__global__ void __launch_bounds__(128) check_kernel(const uint16_t* A, const uint16_t* B, float* D, int N, int shift, int fmt, int* status) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  constexpr int AR = 128 + 16;
  uint16_t* sA = reinterpret_cast<uint16_t*>(smem);
  uint16_t* sB = sA + (K / 8) * AR * 8;
  for (int i = threadIdx.x; i < AR * K; i += blockDim.x) { const int r = i / K, k = i % K; sA[((k / 8) * AR + r) * 8 + (k % 8)] = A[i]; }
  for (int i = threadIdx.x; i < N * K; i += blockDim.x) { const int r = i / K, k = i % K; sB[((k / 8) * N + r) * 8 + (k % 8)] = B[i]; }
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;\n" ::"r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tm = tmem_base;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(fmt == 0 ? 1 : 2, 128, N);
    for (int j = 0; j < K / 16; ++j) {
      const uint64_t ad = make_desc(smem_u32(sA) + shift * 16 + j * 2 * AR * 16, AR * 16, 128);
      const uint64_t bd = make_desc(smem_u32(sB) + j * 2 * N * 16, N * 16, 128);
      umma<1>(tm, ad, bd, idesc, j > 0);
    }
    umma_commit(&bar);
  }
  const bool ok = mbar_wait_bounded(&bar, 0, 20000000LL);
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  if (!ok) { if (threadIdx.x == 0) *status = 1; }
  else {
    for (int c0 = 0; c0 < N; c0 += 8) {
      uint32_t v[8];
      const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16) + c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      for (int i = 0; i < 8; ++i) D[threadIdx.x * N + c0 + i] = __uint_as_float(v[i]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;\n" ::"r"(tm) : "memory");
}

template <int KIND>
__global__ void __launch_bounds__(128) rate_kernel(int N, int reps, int nbuf, int nacc, int albo, int blbo, int nbbuf, long long* cycles, int* status) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  uint32_t* s = reinterpret_cast<uint32_t*>(smem);
  const int total = (nbuf * 2 * albo + nbbuf * 2 * blbo) / 4 + 64;
  // small integers in every format: 0x3c00 = 1.0 (f16), 0x3f80 = 1.0 (bf16), tf32 1.0f
  const uint32_t one = KIND == 0 ? 0x3f800000u : (KIND == 1 ? 0x3c003c00u : 0x3f803f80u);
  for (int i = threadIdx.x; i < total; i += blockDim.x) s[i] = ((i * 37) % 17) < 8 ? one : 0u;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tm = tmem_base;
  if (warp == 0) {
    const uint32_t idesc = make_idesc(KIND, 128, N);
    const uint32_t a0 = smem_u32(s), b0 = a0 + nbuf * (2 * albo);
    uint32_t elected;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(elected));
    const long long t0 = clock64();
    int ib = 0, ish = 0, iacc = 0, ibb = 0;
    for (int r = 0; r < reps; ++r) {
      const uint64_t ad = make_desc(a0 + ib * (2 * albo) + ish * 16, albo, 128);
      const uint64_t bd = make_desc(b0 + ibb * (2 * blbo), blbo, 128);
      if (++ibb == nbbuf) ibb = 0;
      if (elected) umma<KIND>(tm + iacc * N, ad, bd, idesc, r >= nacc);
      if (++ib == nbuf) ib = 0;
      if (++ish == 3) ish = 0;
      if (++iacc == nacc) iacc = 0;
    }
    if (elected) umma_commit(&bar);
    __syncwarp();
    const bool ok = mbar_wait_bounded(&bar, 0, 2000000000LL);
    const long long t1 = clock64();
    if (threadIdx.x == 0) { if (!ok) *status = 1; cycles[blockIdx.x] = t1 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tm) : "memory");
}

static float h2f(uint16_t h, int fmt) {
  if (fmt == 1) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }
  const int s = h >> 15, e = (h >> 10) & 31, m = h & 1023;
  float v = e == 0 ? ldexpf((float)m, -24) : ldexpf((float)(m | 1024), e - 25);
  return s ? -v : v;
}

int main(int argc, char** argv) {
  if (argc < 5) return 2;
  int* status; cudaMalloc(&status, 4); cudaMemset(status, 0, 4);
  if (!strcmp(argv[1], "check")) {
    const int fmt = atoi(argv[2]), N = atoi(argv[3]), shift = atoi(argv[4]);
    constexpr int K = 32, AR = 144;
    std::vector<uint16_t> A(AR * K), B(N * K);
    std::vector<float> D(128 * N);
    srand(1);
    auto rnd = [&]() -> uint16_t {  // random sign, exponent around 1, random mantissa
      const int s = rand() & 1, m = rand();
      if (fmt == 1) return (uint16_t)((s << 15) | ((125 + rand() % 4) << 7) | (m & 127));
      return (uint16_t)((s << 15) | ((13 + rand() % 4) << 10) | (m & 1023));
    };
    for (auto& v : A) v = rnd();
    for (auto& v : B) v = rnd();
    uint16_t *dA, *dB; float* dD;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, D.size() * 4);
    const int smem = (K / 8) * AR * 16 + (K / 8) * N * 16;
    check_kernel<K><<<1, 128, smem>>>(dA, dB, dD, N, shift, fmt, status);
    cudaError_t e = cudaDeviceSynchronize();
    int st = 0; cudaMemcpy(&st, status, 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess || st) { printf("check16 fmt=%d N=%d shift=%d: FAILED (%s, status %d)\n", fmt, N, shift, cudaGetErrorString(e), st); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double err = 0, mx = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
      double se = 0;
      for (int k = 0; k < K; ++k) se += (double)h2f(A[(m + shift) * K + k], fmt) * h2f(B[n * K + k], fmt);
      err = fmax(err, fabs(D[m * N + n] - se)); mx = fmax(mx, fabs(se));
    }
    printf("check16 fmt=%s N=%d shift=%d: max|D|=%.3f, max error vs exact product of the 16-bit inputs %.3e (fp32 accumulate: ~1e-7 relative expected)\n",
           fmt ? "bf16" : "f16", N, shift, mx, err);
    return 0;
  }
  const int kind = atoi(argv[2]), N = atoi(argv[3]), reps = atoi(argv[4]);
  const int nbuf = argc > 5 ? atoi(argv[5]) : 8, nacc = argc > 6 ? atoi(argv[6]) : 1, albo = argc > 7 ? atoi(argv[7]) : 2304,
            blbo = argc > 8 ? atoi(argv[8]) : 4096, nbbuf = argc > 9 ? atoi(argv[9]) : 1;
  long long* cyc; cudaMalloc(&cyc, 148 * 8);
  const int smem = nbuf * 2 * albo + nbbuf * 2 * blbo + 256;
  auto launch = [&]() {
    if (kind == 0) { cudaFuncSetAttribute(rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); rate_kernel<0><<<148, 128, smem>>>(N, reps, nbuf, nacc, albo, blbo, nbbuf, cyc, status); }
    else if (kind == 1) { cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); rate_kernel<1><<<148, 128, smem>>>(N, reps, nbuf, nacc, albo, blbo, nbbuf, cyc, status); }
    else { cudaFuncSetAttribute(rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); rate_kernel<2><<<148, 128, smem>>>(N, reps, nbuf, nacc, albo, blbo, nbbuf, cyc, status); }
  };
  launch();
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  launch();
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  int st = 0; cudaMemcpy(&st, status, 4, cudaMemcpyDeviceToHost);
  std::vector<long long> h(148); cudaMemcpy(h.data(), cyc, 148 * 8, cudaMemcpyDeviceToHost);
  long long mn = h[0], mxc = h[0]; for (auto v : h) { mn = v < mn ? v : mn; mxc = v > mxc ? v : mxc; }
  const int Kmma = kind == 0 ? 8 : 16;
  const double flop = 148.0 * reps * 2.0 * 128.0 * N * Kmma;
  printf("rate16 kind=%s N=%d reps=%d nbuf=%d nacc=%d albo=%d blbo=%d nbbuf=%d: %s status %d; cycles/MMA min %.2f max %.2f; %.3f ms -> %.1f TFLOP/s chip-wide (148 CTAs, M=128, K=%d, SS mode, cta_group::1)\n",
         kind == 0 ? "tf32" : (kind == 1 ? "f16" : "bf16"), N, reps, nbuf, nacc, albo, blbo, nbbuf, cudaGetErrorString(e), st, (double)mn / reps, (double)mxc / reps,
         ms, flop / (ms * 1e-3) / 1e12, Kmma);
  return 0;
}
