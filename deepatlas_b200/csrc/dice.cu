// softmax_dice: the reduction half of DiceLossMultiClass (lib/loss.py:410-476).
//
// The reference runs softmax (loss.py:427), materialises a float32 one-hot of the labels
// (lib/transforms.py:675-689, 629 MB at 32 classes / 160x192x160) and three full-tensor sums
// (loss.py:449-450,472).  Here one pass reads the logits once (planar NCDHW, coalesced along W per
// class plane) and the raw uint8/int64 labels, keeps the class vector in registers, and produces
// per class  S = sum p,  T = sum t,  I = sum p*t  through a warp-shuffle tree, one smem hop per
// block and a fixed-order second stage (deterministic).  The scalar weighting / score formula on
// the (N,C) sums stays in the host-side mirror of the reference class.
// Algorithmic bytes fwd+bwd: 3*C*V*4 + 2*V.
#include "common.cuh"

namespace {

enum TargetKind { TK_U8 = 0, TK_I64 = 1, TK_SOFT = 2, TK_I32 = 3 };

__device__ __forceinline__ int load_label(const void* t, int kind, int64_t i) {
  if (kind == TK_U8) return (int)((const uint8_t*)t)[i];
  if (kind == TK_I64) return (int)((const int64_t*)t)[i];
  return ((const int32_t*)t)[i];
}

template <int CP>
__device__ __forceinline__ void load_probs(const float* __restrict__ s, int64_t V, int64_t v, int C,
                                           bool softmax, float (&p)[CP]) {
#pragma unroll
  for (int c = 0; c < CP; ++c) p[c] = (c < C) ? s[(int64_t)c * V + v] : -INFINITY;
  if (softmax) {
    float m = p[0];
#pragma unroll
    for (int c = 1; c < CP; ++c) m = fmaxf(m, p[c]);
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < CP; ++c) {
      p[c] = (c < C) ? expf(p[c] - m) : 0.f;
      sum += p[c];
    }
    const float inv = 1.0f / sum;
#pragma unroll
    for (int c = 0; c < CP; ++c) p[c] *= inv;
  } else {
#pragma unroll
    for (int c = 0; c < CP; ++c)
      if (c >= C) p[c] = 0.f;
  }
}

constexpr int DICE_THREADS = 128;

// partials layout: [N][gridDim.x][3][C]
template <int CP>
__global__ void __launch_bounds__(DICE_THREADS) dice_sums_kernel(const float* __restrict__ source,
                                                                 const void* __restrict__ target, int kind,
                                                                 int softmax, int C, int64_t V,
                                                                 float* __restrict__ partials, float* __restrict__ probs_out) {
  const int n = blockIdx.y;
  const float* s = source + (int64_t)n * C * V;
  float* po = probs_out ? probs_out + (int64_t)n * C * V : nullptr;
  float aS[CP], aT[CP], aI[CP], p[CP];
#pragma unroll
  for (int c = 0; c < CP; ++c) aS[c] = aT[c] = aI[c] = 0.f;
  for (int64_t v = (int64_t)blockIdx.x * DICE_THREADS + threadIdx.x; v < V;
       v += (int64_t)gridDim.x * DICE_THREADS) {
    load_probs<CP>(s, V, v, C, softmax != 0, p);
    if (po) {  // the probabilities are needed again downstream (anatomy term): written here instead of by a second pass
#pragma unroll
      for (int c = 0; c < CP; ++c)
        if (c < C) po[(int64_t)c * V + v] = p[c];
    }
    if (kind == TK_SOFT) {
      const float* t = (const float*)target + (int64_t)n * C * V;
#pragma unroll
      for (int c = 0; c < CP; ++c) {
        const float tv = (c < C) ? t[(int64_t)c * V + v] : 0.f;
        aS[c] += p[c]; aT[c] += tv; aI[c] += p[c] * tv;
      }
    } else {
      const int lab = load_label(target, kind, (int64_t)n * V + v);
#pragma unroll
      for (int c = 0; c < CP; ++c) {
        const bool hit = (lab == c);
        aS[c] += p[c];
        aT[c] += hit ? 1.f : 0.f;
        aI[c] += hit ? p[c] : 0.f;
      }
    }
  }
  __shared__ float red[DICE_THREADS / 32][3 * CP];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < CP; ++c) {
    const float a = warp_sum(aS[c]), b = warp_sum(aT[c]), d = warp_sum(aI[c]);
    if (lane == 0) { red[w][c] = a; red[w][CP + c] = b; red[w][2 * CP + c] = d; }
  }
  __syncthreads();
  float* out = partials + ((int64_t)n * gridDim.x + blockIdx.x) * 3 * C;
  for (int i = threadIdx.x; i < 3 * C; i += DICE_THREADS) {
    const int q = i / C, c = i - q * C;
    float acc = 0.f;
#pragma unroll
    for (int ww = 0; ww < DICE_THREADS / 32; ++ww) acc += red[ww][q * CP + c];
    out[i] = acc;
  }
}

// one warp per output element: lanes stride over the blocks' partial rows, fixed-order fp64 fold
__global__ void __launch_bounds__(256) dice_finalize_kernel(const float* __restrict__ partials, int nblocks, int C3,
                                                            float* __restrict__ sums) {
  const int n = blockIdx.y, lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= C3) return;
  const float* p = partials + (int64_t)n * nblocks * C3 + i;
  double acc = 0.0;
  for (int b = lane; b < nblocks; b += 32) acc += (double)p[(int64_t)b * C3];
  acc = warp_sum(acc);
  if (lane == 0) sums[(int64_t)n * C3 + i] = (float)acc;
}

template <int CP>
__global__ void __launch_bounds__(DICE_THREADS) dice_bwd_kernel(
    const float* __restrict__ source, const void* __restrict__ target, int kind, int softmax, int C, int64_t V,
    const float* __restrict__ gS, const float* __restrict__ gT, const float* __restrict__ gI,
    float* __restrict__ grad_source, float* __restrict__ grad_target, const float* __restrict__ gprob) {
  const int n = blockIdx.y;
  __shared__ float sg[3][CP];
  for (int i = threadIdx.x; i < 3 * CP; i += DICE_THREADS) {
    const int q = i / CP, c = i - q * CP;
    const float* gp = q == 0 ? gS : (q == 1 ? gT : gI);
    sg[q][c] = (c < C && gp) ? gp[n * C + c] : 0.f;
  }
  __syncthreads();
  const float* s = source + (int64_t)n * C * V;
  float* gs = grad_source ? grad_source + (int64_t)n * C * V : nullptr;
  float p[CP], g[CP];
  for (int64_t v = (int64_t)blockIdx.x * DICE_THREADS + threadIdx.x; v < V;
       v += (int64_t)gridDim.x * DICE_THREADS) {
    load_probs<CP>(s, V, v, C, softmax != 0, p);
    if (kind == TK_SOFT) {
      const float* t = (const float*)target + (int64_t)n * C * V;
      float* gt = grad_target ? grad_target + (int64_t)n * C * V : nullptr;
#pragma unroll
      for (int c = 0; c < CP; ++c) {
        const float tv = (c < C) ? t[(int64_t)c * V + v] : 0.f;
        g[c] = sg[0][c] + sg[2][c] * tv;
        if (gt && c < C) gt[(int64_t)c * V + v] = sg[1][c] + sg[2][c] * p[c];
      }
    } else {
      const int lab = load_label(target, kind, (int64_t)n * V + v);
#pragma unroll
      for (int c = 0; c < CP; ++c) g[c] = sg[0][c] + ((lab == c) ? sg[2][c] : 0.f);
    }
    if (!gs) continue;
    if (gprob) {  // a second consumer of the probabilities: its gradient joins the Dice part before the softmax Jacobian
      const float* gq = gprob + (int64_t)n * C * V;
#pragma unroll
      for (int c = 0; c < CP; ++c)
        if (c < C) g[c] += gq[(int64_t)c * V + v];
    }
    if (softmax) {
      float dot = 0.f;
#pragma unroll
      for (int c = 0; c < CP; ++c) dot += p[c] * g[c];
#pragma unroll
      for (int c = 0; c < CP; ++c)
        if (c < C) gs[(int64_t)c * V + v] = p[c] * (g[c] - dot);
    } else {
#pragma unroll
      for (int c = 0; c < CP; ++c)
        if (c < C) gs[(int64_t)c * V + v] = g[c];
    }
  }
}


// The hot instantiation of the backward -- label target, softmax inside, optionally a second gradient arriving at the
// probabilities (GP) -- as its own kernel: the general one carries four uniform branches (target kind, softmax,
// grad_target, gprob) and at CP = 32 ptxas keeps both sides alive, 255 registers plus 1.4 KB of spills per thread.
// Here the upstream gradient of channel c is formed on the fly (never stored as an array).
template <int CP, bool GP>
__global__ void __launch_bounds__(DICE_THREADS) softmax_dice_bwd_label_kernel(
    const float* __restrict__ source, const void* __restrict__ target, int kind, int C, int64_t V, const float* __restrict__ gS,
    const float* __restrict__ gI, const float* __restrict__ gprob, float* __restrict__ grad_source) {
  const int n = blockIdx.y;
  __shared__ float sg[2][CP];
  for (int i = threadIdx.x; i < 2 * CP; i += DICE_THREADS) {
    const int q = i / CP, c = i - q * CP;
    sg[q][c] = (c < C) ? (q == 0 ? gS : gI)[n * C + c] : 0.f;
  }
  __syncthreads();
  const float* s = source + (int64_t)n * C * V;
  const float* gq = GP ? gprob + (int64_t)n * C * V : nullptr;
  float* gs = grad_source + (int64_t)n * C * V;
  float p[CP], u[CP];
  for (int64_t v = (int64_t)blockIdx.x * DICE_THREADS + threadIdx.x; v < V; v += (int64_t)gridDim.x * DICE_THREADS) {
    if (GP) {
#pragma unroll
      for (int c = 0; c < CP; ++c) u[c] = (c < C) ? gq[(int64_t)c * V + v] : 0.f;
    }
    load_probs<CP>(s, V, v, C, true, p);
    const int lab = load_label(target, kind, (int64_t)n * V + v);
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < CP; ++c) {
      const float g = sg[0][c] + ((lab == c) ? sg[1][c] : 0.f) + (GP ? u[c] : 0.f);
      if (GP) u[c] = g;
      dot = fmaf(p[c], g, dot);
    }
#pragma unroll
    for (int c = 0; c < CP; ++c) {
      const float g = GP ? u[c] : sg[0][c] + ((lab == c) ? sg[1][c] : 0.f);
      if (c < C) gs[(int64_t)c * V + v] = p[c] * (g - dot);
    }
  }
}


// channel softmax (F.softmax(dim=1), lib/loss.py:427 semantics) materialised once for the anatomy branch,
// where the probabilities are themselves warped.
template <int CP>
__global__ void __launch_bounds__(DICE_THREADS) softmax_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                   int C, int64_t V) {
  const int n = blockIdx.y;
  const float* s = x + (int64_t)n * C * V;
  float* o = y + (int64_t)n * C * V;
  float p[CP];
  for (int64_t v = (int64_t)blockIdx.x * DICE_THREADS + threadIdx.x; v < V; v += (int64_t)gridDim.x * DICE_THREADS) {
    load_probs<CP>(s, V, v, C, true, p);
#pragma unroll
    for (int c = 0; c < CP; ++c)
      if (c < C) o[(int64_t)c * V + v] = p[c];
  }
}

// dx = y * (dy - sum_c y*dy)
template <int CP>
__global__ void __launch_bounds__(DICE_THREADS) softmax_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy,
                                                                   float* __restrict__ dx, int C, int64_t V) {
  const int n = blockIdx.y;
  const int64_t base = (int64_t)n * C * V;
  float p[CP], g[CP];
  for (int64_t v = (int64_t)blockIdx.x * DICE_THREADS + threadIdx.x; v < V; v += (int64_t)gridDim.x * DICE_THREADS) {
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < CP; ++c) {
      p[c] = (c < C) ? y[base + (int64_t)c * V + v] : 0.f;
      g[c] = (c < C) ? dy[base + (int64_t)c * V + v] : 0.f;
      dot = fmaf(p[c], g[c], dot);
    }
#pragma unroll
    for (int c = 0; c < CP; ++c)
      if (c < C) dx[base + (int64_t)c * V + v] = p[c] * (g[c] - dot);
  }
}

inline int dice_blocks(int64_t V) {
  int64_t b = da_cdiv(V, (int64_t)DICE_THREADS * 4);
  const int64_t cap = (int64_t)DA_NUM_SMS * 8;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

DA_API int64_t da_dice_workspace_bytes(int N, int C, int64_t V) {
  return (int64_t)sizeof(float) * N * dice_blocks(V) * 3 * C;
}

#define DICE_DISPATCH(CALL)                                                        \
  if (C <= 4) { CALL(4); } else if (C <= 8) { CALL(8); } else if (C <= 16) { CALL(16); } \
  else if (C <= 32) { CALL(32); } else { CALL(64); }

// source [N,C,V] fp32 (logits if apply_softmax else probabilities); target: labels [N,V]
// (kind 0 = uint8, 1 = int64, 3 = int32) or soft [N,C,V] fp32 (kind 2).  sums [N,3,C] = S,T,I.
DA_API int da_dice_sums_fwd(const float* source, const void* target, int target_kind, int apply_softmax,
                            int N, int C, int64_t V, float* sums, void* workspace, int64_t workspace_bytes,
                            cudaStream_t stream) {
  DA_REQUIRE(source && target && sums && workspace, "da_dice_sums_fwd: null pointer");
  DA_REQUIRE(C >= 1 && C <= 64, "da_dice_sums_fwd: unsupported class count %d (1..64)", C);
  DA_REQUIRE(target_kind >= 0 && target_kind <= 3, "da_dice_sums_fwd: bad target kind");
  if (workspace_bytes < da_dice_workspace_bytes(N, C, V)) {
    da_set_error("da_dice_sums_fwd: workspace too small");
    return DA_ERR_WORKSPACE;
  }
  const int nb = dice_blocks(V);
  dim3 grid(nb, N);
#define CALL(CP) dice_sums_kernel<CP><<<grid, DICE_THREADS, 0, stream>>>(source, target, target_kind, apply_softmax, C, V, (float*)workspace, nullptr)
  DICE_DISPATCH(CALL)
#undef CALL
  int rc = da_check_launch("da_dice_sums_fwd");
  if (rc) return rc;
  dim3 g2((3 * C + 7) / 8, N);
  dice_finalize_kernel<<<g2, 256, 0, stream>>>((const float*)workspace, nb, 3 * C, sums);
  return da_check_launch("da_dice_sums_fwd/finalize");
}

// softmax + Dice sums + the probabilities themselves in ONE pass over the logits: for a prediction that feeds both the
// supervised Dice term and, as probabilities, a second consumer (the anatomy term of the joint step).  Replaces
// da_dice_sums_fwd(apply_softmax) + da_softmax_fwd; probs [N,C,V].
DA_API int da_softmax_dice_fwd(const float* logits, const void* target, int target_kind, int N, int C, int64_t V, float* sums,
                               float* probs, void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  DA_REQUIRE(logits && target && sums && probs && workspace, "da_softmax_dice_fwd: null pointer");
  DA_REQUIRE(C >= 1 && C <= 64, "da_softmax_dice_fwd: unsupported class count %d (1..64)", C);
  DA_REQUIRE(target_kind >= 0 && target_kind <= 3, "da_softmax_dice_fwd: bad target kind");
  if (workspace_bytes < da_dice_workspace_bytes(N, C, V)) { da_set_error("da_softmax_dice_fwd: workspace too small"); return DA_ERR_WORKSPACE; }
  const int nb = dice_blocks(V);
  dim3 grid(nb, N);
#define CALL(CP) dice_sums_kernel<CP><<<grid, DICE_THREADS, 0, stream>>>(logits, target, target_kind, 1, C, V, (float*)workspace, probs)
  DICE_DISPATCH(CALL)
#undef CALL
  int rc = da_check_launch("da_softmax_dice_fwd");
  if (rc) return rc;
  dim3 g2((3 * C + 7) / 8, N);
  dice_finalize_kernel<<<g2, 256, 0, stream>>>((const float*)workspace, nb, 3 * C, sums);
  return da_check_launch("da_softmax_dice_fwd/finalize");
}

// Backward of da_softmax_dice_fwd: grad_logits = softmax Jacobian applied to (Dice part from gS, gI [+ gT for a soft
// target]) + grad_probs, in one pass (replaces da_dice_sums_bwd + da_softmax_bwd + the sum of their two results).
// grad_probs nullable.
DA_API int da_softmax_dice_bwd(const float* logits, const void* target, int target_kind, int N, int C, int64_t V, const float* gS,
                               const float* gT, const float* gI, const float* grad_probs, float* grad_logits, cudaStream_t stream) {
  DA_REQUIRE(logits && target && gS && gI && grad_logits, "da_softmax_dice_bwd: null pointer");
  DA_REQUIRE(C >= 1 && C <= 64, "da_softmax_dice_bwd: unsupported class count %d (1..64)", C);
  dim3 grid(dice_blocks(V), N);
  if (target_kind != TK_SOFT && C <= 32) {
    if (grad_probs) {
#define CALL(CP) softmax_dice_bwd_label_kernel<CP, true><<<grid, DICE_THREADS, 0, stream>>>(logits, target, target_kind, C, V, gS, gI, grad_probs, grad_logits)
      DICE_DISPATCH(CALL)
#undef CALL
    } else {
#define CALL(CP) softmax_dice_bwd_label_kernel<CP, false><<<grid, DICE_THREADS, 0, stream>>>(logits, target, target_kind, C, V, gS, gI, nullptr, grad_logits)
      DICE_DISPATCH(CALL)
#undef CALL
    }
    return da_check_launch("da_softmax_dice_bwd/label");
  }
#define CALL(CP) dice_bwd_kernel<CP><<<grid, DICE_THREADS, 0, stream>>>(logits, target, target_kind, 1, C, V, gS, gT, gI, grad_logits, nullptr, grad_probs)
  DICE_DISPATCH(CALL)
#undef CALL
  return da_check_launch("da_softmax_dice_bwd");
}

// gS,gT,gI [N,C]: upstream gradients w.r.t. the three sums (gT may be null).
DA_API int da_dice_sums_bwd(const float* source, const void* target, int target_kind, int apply_softmax,
                            int N, int C, int64_t V, const float* gS, const float* gT, const float* gI,
                            float* grad_source, float* grad_target, cudaStream_t stream) {
  DA_REQUIRE(source && target && gS && gI, "da_dice_sums_bwd: null pointer");
  DA_REQUIRE(C >= 1 && C <= 64, "da_dice_sums_bwd: unsupported class count %d (1..64)", C);
  DA_REQUIRE(grad_target == nullptr || target_kind == TK_SOFT, "da_dice_sums_bwd: grad_target needs a soft target");
  dim3 grid(dice_blocks(V), N);
  if (apply_softmax && target_kind != TK_SOFT && grad_source && C <= 32) {   // (at 64 classes the general kernel is the spill-free one)
#define CALL(CP) softmax_dice_bwd_label_kernel<CP, false><<<grid, DICE_THREADS, 0, stream>>>(source, target, target_kind, C, V, gS, gI, nullptr, grad_source)
    DICE_DISPATCH(CALL)
#undef CALL
    return da_check_launch("da_dice_sums_bwd/label");
  }
#define CALL(CP) dice_bwd_kernel<CP><<<grid, DICE_THREADS, 0, stream>>>(source, target, target_kind, apply_softmax, C, V, gS, gT, gI, grad_source, grad_target, nullptr)
  DICE_DISPATCH(CALL)
#undef CALL
  return da_check_launch("da_dice_sums_bwd");
}

// y = softmax over the channel axis of x [N,C,V]
DA_API int da_softmax_fwd(const float* x, float* y, int N, int C, int64_t V, cudaStream_t stream) {
  DA_REQUIRE(x && y, "da_softmax_fwd: null pointer");
  DA_REQUIRE(C >= 1 && C <= 64, "da_softmax_fwd: unsupported class count %d (1..64)", C);
  dim3 grid(dice_blocks(V), N);
#define CALL(CP) softmax_fwd_kernel<CP><<<grid, DICE_THREADS, 0, stream>>>(x, y, C, V)
  DICE_DISPATCH(CALL)
#undef CALL
  return da_check_launch("da_softmax_fwd");
}

DA_API int da_softmax_bwd(const float* y, const float* dy, float* dx, int N, int C, int64_t V, cudaStream_t stream) {
  DA_REQUIRE(y && dy && dx, "da_softmax_bwd: null pointer");
  DA_REQUIRE(C >= 1 && C <= 64, "da_softmax_bwd: unsupported class count %d (1..64)", C);
  dim3 grid(dice_blocks(V), N);
#define CALL(CP) softmax_bwd_kernel<CP><<<grid, DICE_THREADS, 0, stream>>>(y, dy, dx, C, V)
  DICE_DISPATCH(CALL)
#undef CALL
  return da_check_launch("da_softmax_bwd");
}
