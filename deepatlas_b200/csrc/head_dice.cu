// head_dice: the 1x1x1 class head of the U-Net (lib/network_factory/unets.py:250, Conv3d(16, n_classes, 1)) fused with
// softmax + the Dice sums of DiceLossMultiClass (lib/loss.py:427-476), forward and backward.
//
// Unfused, the head writes the logits (C*V*4 B = 629 MB at 32 classes / 160x192x160), the Dice pass reads them, the
// backward writes their gradient, and the head's data and weight gradients read that again twice.  Here the logits and
// their gradient live in registers only: the forward reads the 16 feature planes once (+ writes the probabilities when a
// second consumer needs them), the backward reads the features (+ the probabilities' gradient) and writes the 16
// feature-gradient planes; the head's weight gradient dW[c][k] = sum_v dlogit[c][v] * f[k][v] is formed per block from a
// shared-memory copy of the block's 256 voxels (register tiles of 4 c x 4 k, accumulators resident for the block's whole
// walk, one fixed-order reduction at the end: deterministic).
// Thread = two adjacent voxels (8-byte accesses, a warp touches 256 contiguous bytes of every plane); the 512 head
// weights are broadcast from shared memory as float4 (one LDS.128 per 8 FMA).
#include "common.cuh"

namespace {

constexpr int HD_K = 16;       // head input channels
constexpr int HD_CP = 32;      // classes, padded
constexpr int HD_THREADS = 128;
constexpr int HD_VB = 2 * HD_THREADS;   // voxels per block iteration
constexpr int HD_PITCH = HD_VB + 4;     // row pitch of the staged tiles (floats): rows r, r+1, .. start 4 banks apart

__device__ __forceinline__ int hd_label(const void* t, int kind, int64_t i) {
  if (kind == 0) return (int)((const uint8_t*)t)[i];
  if (kind == 1) return (int)((const int64_t*)t)[i];
  return ((const int32_t*)t)[i];
}

// logits of two voxels from their 16 features: acc = bias + sum_k f[k] * W[c][k]; swT is [k][c] (c >= C: weight 0,
// bias -inf, so that the padded classes drop out of the softmax)
__device__ __forceinline__ void hd_logits(const float2 (&fv)[HD_K], const float* __restrict__ swT, const float* __restrict__ sb,
                                          float (&a0)[HD_CP], float (&a1)[HD_CP]) {
#pragma unroll
  for (int c = 0; c < HD_CP; ++c) a0[c] = a1[c] = sb[c];
#pragma unroll
  for (int k = 0; k < HD_K; ++k) {
    const float4* w4 = reinterpret_cast<const float4*>(swT + k * HD_CP);
#pragma unroll
    for (int q = 0; q < HD_CP / 4; ++q) {
      const float4 w = w4[q];
      a0[4 * q + 0] = fmaf(fv[k].x, w.x, a0[4 * q + 0]); a1[4 * q + 0] = fmaf(fv[k].y, w.x, a1[4 * q + 0]);
      a0[4 * q + 1] = fmaf(fv[k].x, w.y, a0[4 * q + 1]); a1[4 * q + 1] = fmaf(fv[k].y, w.y, a1[4 * q + 1]);
      a0[4 * q + 2] = fmaf(fv[k].x, w.z, a0[4 * q + 2]); a1[4 * q + 2] = fmaf(fv[k].y, w.z, a1[4 * q + 2]);
      a0[4 * q + 3] = fmaf(fv[k].x, w.w, a0[4 * q + 3]); a1[4 * q + 3] = fmaf(fv[k].y, w.w, a1[4 * q + 3]);
    }
  }
}

// in-place channel softmax (F.softmax(dim=1) semantics: max-shifted exponentials).  The kernels are instruction-issue
// bound (ncu: 56-68 % of the issue slots) and the 64 exponentials per thread and step are a fifth of their instructions
// as expf; ex2.approx on the max-shifted argument (<= 0) is 2 ulp accurate: probabilities move by ~2e-7 relative.
__device__ __forceinline__ void hd_softmax(float (&p)[HD_CP]) {
  float m = p[0];
#pragma unroll
  for (int c = 1; c < HD_CP; ++c) m = fmaxf(m, p[c]);
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < HD_CP; ++c) {
    p[c] = __expf(p[c] - m);   // padded classes: exp(-inf) = 0
    sum += p[c];
  }
  const float inv = 1.0f / sum;
#pragma unroll
  for (int c = 0; c < HD_CP; ++c) p[c] *= inv;
}

__device__ __forceinline__ void hd_stage_weights(const float* __restrict__ w, const float* __restrict__ bias, int C, float* swT,
                                                 float* sb) {
  for (int i = threadIdx.x; i < HD_K * HD_CP; i += HD_THREADS) {
    const int k = i / HD_CP, c = i % HD_CP;
    swT[i] = c < C ? w[c * HD_K + k] : 0.f;
  }
  for (int c = threadIdx.x; c < HD_CP; c += HD_THREADS) sb[c] = c < C ? (bias ? bias[c] : 0.f) : -INFINITY;
}

// partials: [N][gridDim.x][3][C] = S, T, I of this block's voxels (the layout of the unfused Dice pass)
template <bool PROBS>
__global__ void __launch_bounds__(HD_THREADS) head_dice_fwd_kernel(const float* __restrict__ feat, const float* __restrict__ w,
                                                                   const float* __restrict__ bias, const void* __restrict__ target,
                                                                   int kind, int C, int64_t V, float* __restrict__ partials,
                                                                   float* __restrict__ probs) {
  __shared__ __align__(16) float swT[HD_K * HD_CP];
  __shared__ float sb[HD_CP];
  __shared__ int sT[HD_CP];
  __shared__ float red[HD_THREADS / 32][2 * HD_CP];
  const int n = blockIdx.y;
  hd_stage_weights(w, bias, C, swT, sb);
  if (threadIdx.x < HD_CP) sT[threadIdx.x] = 0;
  __syncthreads();
  const float* f = feat + (int64_t)n * HD_K * V;
  float* po = PROBS ? probs + (int64_t)n * C * V : nullptr;
  float aS[HD_CP], aI[HD_CP];
#pragma unroll
  for (int c = 0; c < HD_CP; ++c) aS[c] = aI[c] = 0.f;
  const int64_t npairs = V / 2;
  for (int64_t q = (int64_t)blockIdx.x * HD_THREADS + threadIdx.x; q < npairs; q += (int64_t)gridDim.x * HD_THREADS) {
    const int64_t v = 2 * q;
    float2 fv[HD_K];
#pragma unroll
    for (int k = 0; k < HD_K; ++k) fv[k] = __ldg(reinterpret_cast<const float2*>(f + (int64_t)k * V + v));
    const int l0 = hd_label(target, kind, (int64_t)n * V + v), l1 = hd_label(target, kind, (int64_t)n * V + v + 1);
    float p0[HD_CP], p1[HD_CP];
    hd_logits(fv, swT, sb, p0, p1);
    hd_softmax(p0);
    hd_softmax(p1);
    if (PROBS) {
#pragma unroll
      for (int c = 0; c < HD_CP; ++c)
        if (c < C) *reinterpret_cast<float2*>(po + (int64_t)c * V + v) = make_float2(p0[c], p1[c]);
    }
#pragma unroll
    for (int c = 0; c < HD_CP; ++c) {
      aS[c] += p0[c] + p1[c];
      aI[c] += ((l0 == c) ? p0[c] : 0.f) + ((l1 == c) ? p1[c] : 0.f);
    }
    if ((unsigned)l0 < (unsigned)C) atomicAdd(&sT[l0], 1);   // label counts are integers: exact whatever the order
    if ((unsigned)l1 < (unsigned)C) atomicAdd(&sT[l1], 1);
  }
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < HD_CP; ++c) {
    const float a = warp_sum(aS[c]), d = warp_sum(aI[c]);
    if (lane == 0) { red[wp][c] = a; red[wp][HD_CP + c] = d; }
  }
  __syncthreads();
  float* out = partials + ((int64_t)n * gridDim.x + blockIdx.x) * 3 * C;
  for (int i = threadIdx.x; i < 3 * C; i += HD_THREADS) {
    const int qq = i / C, c = i - qq * C;
    float acc;
    if (qq == 1) {
      acc = (float)sT[c];   // < 2^24 voxels per block
    } else {
      acc = 0.f;
#pragma unroll
      for (int ww = 0; ww < HD_THREADS / 32; ++ww) acc += red[ww][(qq ? HD_CP : 0) + c];
    }
    out[i] = acc;
  }
}

// one warp per output element: lanes stride over the blocks' partial rows, fixed-order fp64 fold
__global__ void __launch_bounds__(256) hd_finalize_sums_kernel(const float* __restrict__ partials, int nblocks, int C3,
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

// Backward.  Dynamic shared memory: sdl [CP][PITCH] (logit gradients of the block's 256 voxels), sf [K][PITCH] (their
// features), then swT [K][CP], sw [CP][K], sb, sg [2][CP].  wpart: [N * gridDim.x][CP * K + CP] (weight and bias
// gradient partial sums of this block, padded classes included as zeros).
constexpr int HD_WPART = HD_CP * HD_K + HD_CP;
constexpr int HD_BWD_SMEM = (HD_CP * HD_PITCH + HD_K * HD_PITCH + 2 * HD_K * HD_CP + HD_CP + 2 * HD_CP) * 4;

template <bool GP>
__global__ void __launch_bounds__(HD_THREADS) head_dice_bwd_kernel(const float* __restrict__ feat, const float* __restrict__ w,
                                                                   const float* __restrict__ bias, const void* __restrict__ target,
                                                                   int kind, int C, int64_t V, const float* __restrict__ gS,
                                                                   const float* __restrict__ gI, const float* __restrict__ gprob,
                                                                   float* __restrict__ gfeat, float* __restrict__ wpart) {
  extern __shared__ __align__(16) float hd_smem[];
  float* sdl = hd_smem;
  float* sf = sdl + HD_CP * HD_PITCH;
  float* swT = sf + HD_K * HD_PITCH;
  float* sw = swT + HD_K * HD_CP;
  float* sb = sw + HD_CP * HD_K;
  float* sg = sb + HD_CP;
  const int n = blockIdx.y;
  hd_stage_weights(w, bias, C, swT, sb);
  for (int i = threadIdx.x; i < HD_CP * HD_K; i += HD_THREADS) sw[i] = (i / HD_K) < C ? w[i] : 0.f;
  for (int i = threadIdx.x; i < 2 * HD_CP; i += HD_THREADS) {
    const int qq = i / HD_CP, c = i - qq * HD_CP;
    sg[i] = c < C ? (qq == 0 ? gS : gI)[n * C + c] : 0.f;
  }
  __syncthreads();
  const float* f = feat + (int64_t)n * HD_K * V;
  const float* gq = GP ? gprob + (int64_t)n * C * V : nullptr;
  float* gf = gfeat + (int64_t)n * HD_K * V;
  // weight-gradient tile of this thread: classes ct + 8 j, features kt + 4 i, voxel quarter = warp
  const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
  const int ct = lane >> 2, kt = lane & 3;
  float wacc[4][4], bacc[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    bacc[j] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) wacc[j][i] = 0.f;
  }
  const int64_t npairs = V / 2;
  const int64_t nit = da_cdiv(npairs, (int64_t)gridDim.x * HD_THREADS);
  for (int64_t it = 0; it < nit; ++it) {
    const int64_t q = (it * gridDim.x + blockIdx.x) * HD_THREADS + threadIdx.x;
    const bool live = q < npairs;
    const int64_t v = 2 * q;
    float2 fv[HD_K];
#pragma unroll
    for (int k = 0; k < HD_K; ++k) fv[k] = live ? __ldg(reinterpret_cast<const float2*>(f + (int64_t)k * V + v)) : make_float2(0.f, 0.f);
    int l0 = -1, l1 = -1;
    if (live) { l0 = hd_label(target, kind, (int64_t)n * V + v); l1 = hd_label(target, kind, (int64_t)n * V + v + 1); }
#pragma unroll
    for (int k = 0; k < HD_K; ++k) *reinterpret_cast<float2*>(sf + k * HD_PITCH + 2 * threadIdx.x) = fv[k];
    float p0[HD_CP], p1[HD_CP];
    hd_logits(fv, swT, sb, p0, p1);
    hd_softmax(p0);
    hd_softmax(p1);
    // upstream gradient at the probabilities: Dice part (+ the second consumer's), then the softmax Jacobian
    float g0[HD_CP], g1[HD_CP];
    float dot0 = 0.f, dot1 = 0.f;
#pragma unroll
    for (int c = 0; c < HD_CP; ++c) {
      float2 u = make_float2(0.f, 0.f);
      if (GP) { if (c < C && live) u = __ldg(reinterpret_cast<const float2*>(gq + (int64_t)c * V + v)); }
      g0[c] = sg[c] + ((l0 == c) ? sg[HD_CP + c] : 0.f) + u.x;
      g1[c] = sg[c] + ((l1 == c) ? sg[HD_CP + c] : 0.f) + u.y;
      dot0 = fmaf(p0[c], g0[c], dot0);
      dot1 = fmaf(p1[c], g1[c], dot1);
    }
    float d0[HD_K], d1[HD_K];
#pragma unroll
    for (int k = 0; k < HD_K; ++k) d0[k] = d1[k] = 0.f;
#pragma unroll
    for (int c = 0; c < HD_CP; ++c) {
      const float e0 = live ? p0[c] * (g0[c] - dot0) : 0.f, e1 = live ? p1[c] * (g1[c] - dot1) : 0.f;
      *reinterpret_cast<float2*>(sdl + c * HD_PITCH + 2 * threadIdx.x) = make_float2(e0, e1);
      const float4* w4 = reinterpret_cast<const float4*>(sw + c * HD_K);
#pragma unroll
      for (int qk = 0; qk < HD_K / 4; ++qk) {
        const float4 wv = w4[qk];
        d0[4 * qk + 0] = fmaf(e0, wv.x, d0[4 * qk + 0]); d1[4 * qk + 0] = fmaf(e1, wv.x, d1[4 * qk + 0]);
        d0[4 * qk + 1] = fmaf(e0, wv.y, d0[4 * qk + 1]); d1[4 * qk + 1] = fmaf(e1, wv.y, d1[4 * qk + 1]);
        d0[4 * qk + 2] = fmaf(e0, wv.z, d0[4 * qk + 2]); d1[4 * qk + 2] = fmaf(e1, wv.z, d1[4 * qk + 2]);
        d0[4 * qk + 3] = fmaf(e0, wv.w, d0[4 * qk + 3]); d1[4 * qk + 3] = fmaf(e1, wv.w, d1[4 * qk + 3]);
      }
    }
    if (live) {
#pragma unroll
      for (int k = 0; k < HD_K; ++k) *reinterpret_cast<float2*>(gf + (int64_t)k * V + v) = make_float2(d0[k], d1[k]);
    }
    __syncthreads();
    // weight gradient of the block's 256 voxels: this warp's quarter, four voxels per step
    {
      const float* pa = sdl + ct * HD_PITCH + wq * (HD_VB / 4);
      const float* pb = sf + kt * HD_PITCH + wq * (HD_VB / 4);
#pragma unroll 4
      for (int s = 0; s < HD_VB / 16; ++s) {
        float4 a[4], b[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) a[j] = *reinterpret_cast<const float4*>(pa + 8 * j * HD_PITCH + 4 * s);
#pragma unroll
        for (int i = 0; i < 4; ++i) b[i] = *reinterpret_cast<const float4*>(pb + 4 * i * HD_PITCH + 4 * s);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            wacc[j][i] = fmaf(a[j].x, b[i].x, fmaf(a[j].y, b[i].y, fmaf(a[j].z, b[i].z, fmaf(a[j].w, b[i].w, wacc[j][i]))));
          if (kt == 0) bacc[j] += (a[j].x + a[j].y) + (a[j].z + a[j].w);
        }
      }
    }
    __syncthreads();
  }
  // fold the four voxel quarters (fixed order) through the staging area, one partial row per block
  float* red = sdl;   // [4 quarters][HD_WPART]
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = ct + 8 * j;
#pragma unroll
    for (int i = 0; i < 4; ++i) red[wq * HD_WPART + c * HD_K + kt + 4 * i] = wacc[j][i];
    if (kt == 0) red[wq * HD_WPART + HD_CP * HD_K + c] = bacc[j];
  }
  __syncthreads();
  float* out = wpart + ((int64_t)n * gridDim.x + blockIdx.x) * HD_WPART;
  for (int i = threadIdx.x; i < HD_WPART; i += HD_THREADS)
    out[i] = (red[i] + red[HD_WPART + i]) + (red[2 * HD_WPART + i] + red[3 * HD_WPART + i]);
}

// grad_weight [C][K], grad_bias [C] (nullable) = fixed-order fp64 sums over the blocks' partial rows (warp per element)
__global__ void __launch_bounds__(256) hd_finalize_wgrad_kernel(const float* __restrict__ wpart, int nrows, int C,
                                                                float* __restrict__ gw, float* __restrict__ gb, int accumulate) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= HD_WPART) return;
  const bool is_b = i >= HD_CP * HD_K;
  const int c = is_b ? i - HD_CP * HD_K : i / HD_K;
  if (c >= C || (is_b && !gb)) return;
  double acc = 0.0;
  for (int r = lane; r < nrows; r += 32) acc += (double)wpart[(int64_t)r * HD_WPART + i];
  acc = warp_sum(acc);
  if (lane == 0) {
    if (is_b) gb[c] = (float)(accumulate ? acc + (double)gb[c] : acc);
    else gw[i] = (float)(accumulate ? acc + (double)gw[i] : acc);
  }
}

inline int hd_blocks(int64_t V) {
  int64_t b = da_cdiv(V / 2, (int64_t)HD_THREADS);
  const int64_t cap = (int64_t)DA_NUM_SMS * 3;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

// 1 when the fused kernels apply: 16 head inputs, <= 32 classes, an even number of voxels per volume
DA_API int64_t da_head_dice_supported(int K, int C, int64_t V) { return (K == HD_K && C >= 1 && C <= HD_CP && V >= 2 && (V & 1) == 0) ? 1 : 0; }

DA_API int64_t da_head_dice_workspace_bytes(int N, int C, int64_t V) {
  const int64_t nb = hd_blocks(V);
  const int64_t fwd = (int64_t)sizeof(float) * N * nb * 3 * C, bwd = (int64_t)sizeof(float) * N * nb * HD_WPART;
  return fwd > bwd ? fwd : bwd;
}

// feat [N,16,V] fp32; weight [C,16] (nn.Conv3d(16, C, 1).weight), bias [C] nullable; target: labels [N,V] (kind 0 = uint8,
// 1 = int64, 3 = int32).  sums [N,3,C] = S, T, I of softmax(head(feat)) against one-hot(target); probs [N,C,V] nullable.
DA_API int da_head_dice_fwd(const float* feat, const float* weight, const float* bias, const void* target, int target_kind, int N,
                            int K, int C, int64_t V, float* sums, float* probs, void* workspace, int64_t workspace_bytes,
                            cudaStream_t stream) {
  DA_REQUIRE(feat && weight && target && sums && workspace, "da_head_dice_fwd: null pointer");
  DA_REQUIRE(da_head_dice_supported(K, C, V), "da_head_dice_fwd: unsupported shape (K = %d, C = %d, V = %lld)", K, C, (long long)V);
  DA_REQUIRE(target_kind == 0 || target_kind == 1 || target_kind == 3, "da_head_dice_fwd: label target only");
  DA_REQUIRE(((uintptr_t)feat & 7) == 0 && ((uintptr_t)probs & 7) == 0, "da_head_dice_fwd: tensors must be 8-byte aligned");
  if (workspace_bytes < da_head_dice_workspace_bytes(N, C, V)) { da_set_error("da_head_dice_fwd: workspace too small"); return DA_ERR_WORKSPACE; }
  const int nb = hd_blocks(V);
  dim3 grid(nb, N);
  if (probs) head_dice_fwd_kernel<true><<<grid, HD_THREADS, 0, stream>>>(feat, weight, bias, target, target_kind, C, V, (float*)workspace, probs);
  else head_dice_fwd_kernel<false><<<grid, HD_THREADS, 0, stream>>>(feat, weight, bias, target, target_kind, C, V, (float*)workspace, nullptr);
  int rc = da_check_launch("da_head_dice_fwd");
  if (rc) return rc;
  dim3 g2((3 * C + 7) / 8, N);
  hd_finalize_sums_kernel<<<g2, 256, 0, stream>>>((const float*)workspace, nb, 3 * C, sums);
  return da_check_launch("da_head_dice_fwd/finalize");
}

// gS, gI [N,C]: gradients of the loss w.r.t. the S and I sums (T does not depend on the network); grad_probs [N,C,V]
// nullable (the second consumer's gradient at the probabilities).  grad_feat [N,16,V], grad_weight [C,16], grad_bias [C]
// nullable; accumulate = 1: grad_weight / grad_bias += the result.
DA_API int da_head_dice_bwd(const float* feat, const float* weight, const float* bias, const void* target, int target_kind, int N,
                            int K, int C, int64_t V, const float* gS, const float* gI, const float* grad_probs, float* grad_feat,
                            float* grad_weight, float* grad_bias, int accumulate, void* workspace, int64_t workspace_bytes,
                            cudaStream_t stream) {
  DA_REQUIRE(feat && weight && target && gS && gI && grad_feat && grad_weight && workspace, "da_head_dice_bwd: null pointer");
  DA_REQUIRE(da_head_dice_supported(K, C, V), "da_head_dice_bwd: unsupported shape (K = %d, C = %d, V = %lld)", K, C, (long long)V);
  DA_REQUIRE(target_kind == 0 || target_kind == 1 || target_kind == 3, "da_head_dice_bwd: label target only");
  DA_REQUIRE(((uintptr_t)feat & 7) == 0 && ((uintptr_t)grad_probs & 7) == 0 && ((uintptr_t)grad_feat & 7) == 0,
             "da_head_dice_bwd: tensors must be 8-byte aligned");
  if (workspace_bytes < da_head_dice_workspace_bytes(N, C, V)) { da_set_error("da_head_dice_bwd: workspace too small"); return DA_ERR_WORKSPACE; }
  static DaPerDeviceOnce configured;
  if (configured.first()) {
    cudaFuncSetAttribute(head_dice_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, HD_BWD_SMEM);
    cudaFuncSetAttribute(head_dice_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, HD_BWD_SMEM);
  }
  const int nb = hd_blocks(V);
  dim3 grid(nb, N);
  if (grad_probs)
    head_dice_bwd_kernel<true><<<grid, HD_THREADS, HD_BWD_SMEM, stream>>>(feat, weight, bias, target, target_kind, C, V, gS, gI, grad_probs,
                                                                          grad_feat, (float*)workspace);
  else
    head_dice_bwd_kernel<false><<<grid, HD_THREADS, HD_BWD_SMEM, stream>>>(feat, weight, bias, target, target_kind, C, V, gS, gI, nullptr,
                                                                           grad_feat, (float*)workspace);
  int rc = da_check_launch("da_head_dice_bwd");
  if (rc) return rc;
  hd_finalize_wgrad_kernel<<<(HD_WPART + 7) / 8, 256, 0, stream>>>((const float*)workspace, N * nb, C, grad_weight, grad_bias, accumulate);
  return da_check_launch("da_head_dice_bwd/finalize");
}
