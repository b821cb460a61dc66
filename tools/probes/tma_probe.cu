// Probe: which TMA (cp.async.bulk.tensor.5d) coordinate / box configurations the hardware accepts.
// usage: tma_probe x0 y0 z0 bx [W H D C]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../deepatlas_b200/csrc/tma.cuh"

void da_set_error(const char* fmt, ...) { fprintf(stderr, "error: %s\n", fmt); }
int da_check_launch(const char*, int) { return 0; }
da_encode_tiled_fn da_get_encode_tiled() {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) return (da_encode_tiled_fn)p;
  return nullptr;
}

__global__ void probe(const __grid_constant__ CUtensorMap m, float* out, int x0, int y0, int z0, int nfloat) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  __syncthreads();
  if (threadIdx.x == 0) { mbar_expect_tx(&bar, nfloat * 4); tma_load_5d(smem, &m, &bar, x0, y0, z0, 0, 0); }
  mbar_wait(&bar, 0);
  const float* s = (const float*)smem;
  for (int i = threadIdx.x; i < nfloat; i += blockDim.x) out[i] = s[i];
}

int main(int argc, char** argv) {
  int x0 = atoi(argv[1]), y0 = atoi(argv[2]), z0 = atoi(argv[3]), bx = atoi(argv[4]);
  int W = argc > 5 ? atoi(argv[5]) : 32, H = argc > 6 ? atoi(argv[6]) : 32, D = argc > 7 ? atoi(argv[7]) : 32, C = argc > 8 ? atoi(argv[8]) : 8;
  const int by = 10, bz = 6, bc = 4;
  size_t n = (size_t)W * H * D * C;
  std::vector<float> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = (float)(i % 1000003) + 1.0f;
  float *d, *o;
  cudaMalloc(&d, n * 4); cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
  int nf = bx * by * bz * bc;
  cudaMalloc(&o, nf * 4);
  CUtensorMap m;
  int rc = da_make_volume_map(&m, d, 1, C, D, H, W, bx, by, bz, bc);
  if (rc) { printf("encode failed\n"); return 2; }
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, nf * 4 + 128);
  probe<<<1, 128, nf * 4, 0>>>(m, o, x0, y0, z0, nf);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("cfg x0=%d y0=%d z0=%d bx=%d W=%d: CUDA ERROR %s\n", x0, y0, z0, bx, W, cudaGetErrorString(e)); return 1; }
  std::vector<float> r(nf);
  cudaMemcpy(r.data(), o, nf * 4, cudaMemcpyDeviceToHost);
  size_t bad = 0;
  for (int c = 0; c < bc; ++c) for (int z = 0; z < bz; ++z) for (int y = 0; y < by; ++y) for (int x = 0; x < bx; ++x) {
    int gx = x0 + x, gy = y0 + y, gz = z0 + z;
    float want = 0.f;
    if (gx >= 0 && gx < W && gy >= 0 && gy < H && gz >= 0 && gz < D && c < C) want = h[(((size_t)c * D + gz) * H + gy) * W + gx];
    if (r[((c * bz + z) * by + y) * bx + x] != want) ++bad;
  }
  printf("cfg x0=%d y0=%d z0=%d bx=%d W=%d: ok, mismatches=%zu of %d\n", x0, y0, z0, bx, W, bad, nf);
  return 0;
}
