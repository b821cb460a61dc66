// Library-level entry points: version, thread-local error string, launch checking.
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void da_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static unsigned long long g_launches = 0;

int da_check_launch(const char* what, int nkernels) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    da_set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  __atomic_fetch_add(&g_launches, (unsigned long long)nkernels, __ATOMIC_RELAXED);
  return DA_OK;
}

// number of kernels this library has launched in this process (bench.py reports the delta over the timed region)
DA_API int64_t da_launch_count(void) { return (int64_t)__atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

DA_API int da_version(void) { return 100; }  // 0.1.0
DA_API const char* da_last_error(void) { return g_err; }

// Fills dst with zeros (used by wrappers that cannot rely on the caller's allocator zeroing).
DA_API int da_memset_zero(void* dst, int64_t bytes, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(dst, 0, (size_t)bytes, stream);
  if (e != cudaSuccess) {
    da_set_error("da_memset_zero: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return DA_OK;
}
