// Shared helpers for the deepatlas_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define DA_API extern "C" __attribute__((visibility("default")))

// error codes (negative = library, positive = cudaError_t)
#define DA_OK 0
#define DA_ERR_BAD_ARG (-1)
#define DA_ERR_UNSUPPORTED (-2)
#define DA_ERR_WORKSPACE (-3)

void da_set_error(const char* fmt, ...);
// checks cudaGetLastError() and adds `nkernels` to the library's launch counter (da_launch_count)
int da_check_launch(const char* what, int nkernels = 1);

#define DA_REQUIRE(cond, ...)            \
  do {                                   \
    if (!(cond)) {                       \
      da_set_error(__VA_ARGS__);         \
      return DA_ERR_BAD_ARG;             \
    }                                    \
  } while (0)

static __host__ __device__ inline int64_t da_cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t da_align(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-DEVICE attribute: one process may drive several GPUs
// (tests, a user who moves a model), so "configured once" must be tracked per device.  first() returns true until it
// has been called once for the current device (a benign race between threads just sets the attribute twice).
struct DaPerDeviceOnce {
  unsigned long long done[2] = {0ull, 0ull};
  bool first() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) d = 0;
    d &= 127;
    const unsigned long long bit = 1ull << (d & 63);
    const unsigned long long old = __atomic_fetch_or(&done[d >> 6], bit, __ATOMIC_RELAXED);
    return (old & bit) == 0;
  }
};

// number of SMs on a B200; grids of persistent / grid-stride kernels are sized in multiples of it
#define DA_NUM_SMS 148

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum via shuffle tree + one smem hop; result valid in thread 0.  Fixed order => deterministic.
template <typename T, int NWARPS>
__device__ __forceinline__ T block_sum(T v, T* smem /* NWARPS */) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem[w] = v;
  __syncthreads();
  T r = 0;
  if (w == 0) {
    r = (lane < NWARPS) ? smem[lane] : T(0);
    r = warp_sum(r);
  }
  return r;
}
