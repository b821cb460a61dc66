// eval: the validation step that follows the hot path every epoch.
//
// The reference moves the logits to the host, takes torch.max(pred, 1)[1] and then, for every class c >= 1, calls
// scipy.spatial.distance.dice on two full-volume boolean arrays (models/segmentation.py:188-194,
// lib/evalMetrics.py:58-68) -- 31 passes over the volume on one CPU core.  Here one pass over the logits produces the
// label map (first maximum wins, as torch.max) and, per class, the exact integer counts
//   P_c = #[argmax = c],  T_c = #[truth = c],  I_c = #[argmax = c and truth = c]
// from which the host evaluates 1 - dice_dissimilarity = 2 I / (P + T) in float64 exactly as scipy does.
// Integer atomics only: results are bit-exact and order independent.
#include "common.cuh"

namespace {

constexpr int EV_THREADS = 256;
constexpr int EV_MAXC = 64;

__device__ __forceinline__ int ev_label(const void* t, int kind, int64_t i) {
  if (kind == 0) return (int)((const uint8_t*)t)[i];
  if (kind == 1) return (int)((const int64_t*)t)[i];
  return ((const int32_t*)t)[i];
}

// counts [N][3][C] (unsigned 64-bit, zeroed by the entry point); pred (nullable) [N][V] uint8
__global__ void __launch_bounds__(EV_THREADS) argmax_counts_kernel(const float* __restrict__ logits, const void* __restrict__ truth,
                                                                   int kind, int C, int64_t V, unsigned long long* __restrict__ counts,
                                                                   uint8_t* __restrict__ pred) {
  __shared__ unsigned int h[3][EV_MAXC];
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < 3 * EV_MAXC; i += EV_THREADS) (&h[0][0])[i] = 0u;
  __syncthreads();
  const float* s = logits + (int64_t)n * C * V;
  for (int64_t v = (int64_t)blockIdx.x * EV_THREADS + threadIdx.x; v < V; v += (int64_t)gridDim.x * EV_THREADS) {
    float best = s[v];
    int arg = 0;
    for (int c = 1; c < C; ++c) {
      const float x = s[(int64_t)c * V + v];
      if (x > best || (x != x && best == best)) { best = x; arg = c; }  // first maximum; a NaN wins like in torch.max
    }
    if (pred) pred[(int64_t)n * V + v] = (uint8_t)arg;
    atomicAdd(&h[0][arg], 1u);
    if (truth) {
      const int lab = ev_label(truth, kind, (int64_t)n * V + v);
      if (lab >= 0 && lab < C) {
        atomicAdd(&h[1][lab], 1u);
        if (lab == arg) atomicAdd(&h[2][arg], 1u);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * C; i += EV_THREADS) {
    const int q = i / C, c = i - q * C;
    const unsigned int val = h[q][c];
    if (val) atomicAdd(counts + ((int64_t)n * 3 + q) * C + c, (unsigned long long)val);
  }
}

}  // namespace

// logits [N,C,V] fp32; truth (nullable) [N,V] labels (kind 0 uint8, 1 int64, 3 int32); counts [N,3,C] int64 = (P, T, I);
// pred (nullable) [N,V] uint8 receives the argmax label map.  C <= 64.
DA_API int da_argmax_counts(const float* logits, const void* truth, int truth_kind, int N, int C, int64_t V, int64_t* counts,
                            uint8_t* pred, cudaStream_t stream) {
  DA_REQUIRE(logits && counts, "da_argmax_counts: null pointer");
  DA_REQUIRE(C >= 1 && C <= EV_MAXC, "da_argmax_counts: unsupported class count %d (1..64)", C);
  DA_REQUIRE(truth_kind == 0 || truth_kind == 1 || truth_kind == 3, "da_argmax_counts: bad label kind");
  cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(int64_t) * (size_t)N * 3 * C, stream);
  if (e != cudaSuccess) { da_set_error("da_argmax_counts memset: %s", cudaGetErrorString(e)); return (int)e; }
  int64_t nb = da_cdiv(V, (int64_t)EV_THREADS * 4);
  const int64_t cap = (int64_t)DA_NUM_SMS * 8;
  if (nb > cap) nb = cap;
  if (nb < 1) nb = 1;
  argmax_counts_kernel<<<dim3((unsigned)nb, N), EV_THREADS, 0, stream>>>(logits, truth, truth_kind, C, V, (unsigned long long*)counts, pred);
  return da_check_launch("da_argmax_counts");
}
