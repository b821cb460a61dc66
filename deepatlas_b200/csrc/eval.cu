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

constexpr int LC_BINS = 256;

__device__ __forceinline__ int lc_label(const void* t, int kind, int64_t i) {
  if (kind == 4) return (int)((const float*)t)[i];   // float label maps: truncation, as mask.long() (lib/transforms.py:687)
  return ev_label(t, kind, i);
}

// counts [N][3][bins]: (#[a == c], #[b == c], #[a == c and b == c]) for two label maps; labels outside [0, bins) are skipped
__global__ void __launch_bounds__(EV_THREADS) label_overlap_kernel(const void* __restrict__ a, int kind_a, const void* __restrict__ b,
                                                                   int kind_b, int bins, int64_t V,
                                                                   unsigned long long* __restrict__ counts) {
  __shared__ unsigned int h[3][LC_BINS];
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < 3 * LC_BINS; i += EV_THREADS) (&h[0][0])[i] = 0u;
  __syncthreads();
  for (int64_t v = (int64_t)blockIdx.x * EV_THREADS + threadIdx.x; v < V; v += (int64_t)gridDim.x * EV_THREADS) {
    const int la = lc_label(a, kind_a, (int64_t)n * V + v), lb = lc_label(b, kind_b, (int64_t)n * V + v);
    if (la >= 0 && la < bins) atomicAdd(&h[0][la], 1u);
    if (lb >= 0 && lb < bins) {
      atomicAdd(&h[1][lb], 1u);
      if (la == lb) atomicAdd(&h[2][lb], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * bins; i += EV_THREADS) {
    const int q = i / bins, c = i - q * bins;
    const unsigned int val = h[q][c];
    if (val) atomicAdd(counts + ((int64_t)n * 3 + q) * bins + c, (unsigned long long)val);
  }
}

}  // namespace

// Two label maps a, b [N,V] (kind 0 uint8, 1 int64, 3 int32, 4 fp32 truncated) -> counts [N,3,bins] int64 =
// (#[a == c], #[b == c], #[a == b == c]); bins <= 256.  The integer half of DiceLossOnLabel (lib/loss.py:348-391).
DA_API int da_label_overlap_counts(const void* a, int kind_a, const void* b, int kind_b, int N, int bins, int64_t V,
                                   int64_t* counts, cudaStream_t stream) {
  DA_REQUIRE(a && b && counts, "da_label_overlap_counts: null pointer");
  DA_REQUIRE(bins >= 1 && bins <= LC_BINS, "da_label_overlap_counts: unsupported bin count %d (1..256)", bins);
  auto okk = [](int k) { return k == 0 || k == 1 || k == 3 || k == 4; };
  DA_REQUIRE(okk(kind_a) && okk(kind_b), "da_label_overlap_counts: bad label kind");
  cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(int64_t) * (size_t)N * 3 * bins, stream);
  if (e != cudaSuccess) { da_set_error("da_label_overlap_counts memset: %s", cudaGetErrorString(e)); return (int)e; }
  int64_t nb = da_cdiv(V, (int64_t)EV_THREADS * 4);
  const int64_t cap = (int64_t)DA_NUM_SMS * 8;
  if (nb > cap) nb = cap;
  if (nb < 1) nb = 1;
  label_overlap_kernel<<<dim3((unsigned)nb, N), EV_THREADS, 0, stream>>>(a, kind_a, b, kind_b, bins, V, (unsigned long long*)counts);
  return da_check_launch("da_label_overlap_counts");
}

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
