// Probe for tcgen05.mma kind::tf32 on sm_100a: (1) shared-memory descriptor format for the un-swizzled K-major
// canonical layout incl. a row-shifted start address, (2) tf32 input handling (truncate vs round),
// (3) issue rate versus N (is a small-N MMA bound by the A-operand shared-memory read?).
// usage: umma_probe check N shift | umma_probe rate N reps
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // layout type 0 = no swizzle, base offset 0
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;                    // D format f32
  d |= 2u << 7;                    // A format tf32
  d |= 2u << 10;                   // B format tf32
  d |= (uint32_t)(N >> 3) << 17;   // N
  d |= (uint32_t)(M >> 4) << 24;   // M
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
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

// A [128 + 16 rows][K], B [N][K] from global (row-major), K = 32.  Canonical K-major no-swizzle: [k/4][row][4].
template <int K>
__global__ void __launch_bounds__(128) check_kernel(const float* A, const float* B, float* D, int N, int shift, int* status) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  constexpr int AR = 128 + 16;  // rows staged for A (so that a shifted start stays in range)
  float* sA = reinterpret_cast<float*>(smem);
  float* sB = sA + (K / 4) * AR * 4;
  for (int i = threadIdx.x; i < AR * K; i += blockDim.x) {
    const int r = i / K, k = i % K;
    sA[((k / 4) * AR + r) * 4 + (k % 4)] = A[i];
  }
  for (int i = threadIdx.x; i < N * K; i += blockDim.x) {
    const int r = i / K, k = i % K;
    sB[((k / 4) * N + r) * 4 + (k % 4)] = B[i];
  }
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;\n" ::"r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic-proxy smem writes -> visible to the MMA (async proxy)
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tm = tmem_base;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(128, N);
    for (int j = 0; j < K / 8; ++j) {
      const uint64_t ad = make_desc(smem_u32(sA) + shift * 16 + j * 2 * AR * 16, AR * 16, 128);
      const uint64_t bd = make_desc(smem_u32(sB) + j * 2 * N * 16, N * 16, 128);
      umma_tf32(tm, ad, bd, idesc, j > 0);
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
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;\n" ::"r"(tm) : "memory");
}

__global__ void __launch_bounds__(128) rate_kernel(int N, int reps, int nbuf, int nacc, int albo, int blbo, int nbbuf, long long* cycles, int* status) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  float* s = reinterpret_cast<float*>(smem);
  const int total = (nbuf * 2 * albo + nbbuf * 2 * blbo) / 4 + 64;
  for (int i = threadIdx.x; i < total; i += blockDim.x) s[i] = (float)((i * 37) % 17) * 0.125f;
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
  if (warp == 0) {   // warp-uniform branch: descriptors are uniform values, one elected lane issues (CUTLASS pattern)
    const uint32_t idesc = make_idesc(128, N);
    const uint32_t a0 = smem_u32(s), b0 = a0 + nbuf * (2 * albo);
    uint32_t elected;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(elected));
    const long long t0 = clock64();
    int ib = 0, ish = 0, iacc = 0, ibb = 0;
    for (int r = 0; r < reps; ++r) {
      const uint64_t ad = make_desc(a0 + ib * (2 * albo) + ish * 16, albo, 128);  // rotating A buffers, shifted starts
      const uint64_t bd = make_desc(b0 + ibb * (2 * blbo), blbo, 128);  // rotating B buffers: defeats any operand reuse
      if (++ibb == nbbuf) ibb = 0;
      if (elected) umma_tf32(tm + iacc * N, ad, bd, idesc, r >= nacc);
      if (++ib == nbuf) ib = 0;
      if (++ish == 3) ish = 0;
      if (++iacc == nacc) iacc = 0;
    }
    if (elected) umma_commit(&bar);
    __syncwarp();
    const bool ok = mbar_wait_bounded(&bar, 0, 200000000LL);
    const long long t1 = clock64();
    if (threadIdx.x == 0) { if (!ok) *status = 1; cycles[blockIdx.x] = t1 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;\n" ::"r"(tm) : "memory");
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; memcpy(&x, &u, 4); return x; }
static float tf32_round(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x00000fffu + ((u >> 13) & 1u); u &= 0xffffe000u; memcpy(&x, &u, 4); return x; }

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  int* status; cudaMalloc(&status, 4); cudaMemset(status, 0, 4);
  if (!strcmp(argv[1], "check")) {
    const int N = atoi(argv[2]), shift = atoi(argv[3]);
    constexpr int K = 32, AR = 144;
    std::vector<float> A(AR * K), B(N * K), D(128 * N);
    srand(1);
    for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto& v : B) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, D.size() * 4);
    const int smem = (K / 4) * AR * 16 + (K / 4) * N * 16;
    check_kernel<K><<<1, 128, smem>>>(dA, dB, dD, N, shift, status);
    cudaError_t e = cudaDeviceSynchronize();
    int st = 0; cudaMemcpy(&st, status, 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess || st) { printf("check N=%d shift=%d: FAILED (%s, status %d)\n", N, shift, cudaGetErrorString(e), st); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double e_exact = 0, e_trunc = 0, e_round = 0, mx = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
      double se = 0, st_ = 0, sr = 0;
      for (int k = 0; k < K; ++k) {
        const float a = A[(m + shift) * K + k], b = B[n * K + k];
        se += (double)a * b; st_ += (double)tf32_trunc(a) * tf32_trunc(b); sr += (double)tf32_round(a) * tf32_round(b);
      }
      const double d = D[m * N + n];
      e_exact = fmax(e_exact, fabs(d - se)); e_trunc = fmax(e_trunc, fabs(d - st_)); e_round = fmax(e_round, fabs(d - sr)); mx = fmax(mx, fabs(se));
    }
    printf("check N=%d shift=%d: max|D|=%.3f err vs exact %.3e, vs tf32-truncated inputs %.3e, vs tf32-rounded inputs %.3e\n", N, shift, mx, e_exact, e_trunc, e_round);
    return 0;
  }
  const int N = atoi(argv[2]), reps = atoi(argv[3]), nbuf = argc > 4 ? atoi(argv[4]) : 8, nacc = argc > 5 ? atoi(argv[5]) : 1, albo = argc > 6 ? atoi(argv[6]) : 2304, blbo = argc > 7 ? atoi(argv[7]) : 4096, nbbuf = argc > 8 ? atoi(argv[8]) : 1;
  long long* cyc; cudaMalloc(&cyc, 148 * 8);
  const int smem = nbuf * 2 * albo + nbbuf * 2 * blbo + 256;
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int it = 0; it < 2; ++it) rate_kernel<<<148, 128, smem>>>(N, reps, nbuf, nacc, albo, blbo, nbbuf, cyc, status);
  cudaError_t e = cudaDeviceSynchronize();
  int st = 0; cudaMemcpy(&st, status, 4, cudaMemcpyDeviceToHost);
  std::vector<long long> h(148); cudaMemcpy(h.data(), cyc, 148 * 8, cudaMemcpyDeviceToHost);
  long long mn = h[0], mxc = h[0]; for (auto v : h) { mn = v < mn ? v : mn; mxc = v > mxc ? v : mxc; }
  printf("rate N=%d reps=%d nbuf=%d nacc=%d albo=%d blbo=%d nbbuf=%d: %s status %d; cycles/MMA min %.2f max %.2f (floor 128*N/256 = %.1f; A bytes/MMA 4096 -> %.1f B/cycle)\n", N, reps, nbuf, nacc, albo, blbo, nbbuf,
         cudaGetErrorString(e), st, (double)mn / reps, (double)mxc / reps, 128.0 * N / 256.0, 4096.0 / ((double)mn / reps));
  return 0;
}
