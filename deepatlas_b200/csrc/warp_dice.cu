// warp_dice: the anatomy-similarity term of the joint step,  dice(grid_sample(P, phi), onehot(S_t))  with P the
// moving segmentation's class probabilities (reference ingredients: F.grid_sample as called at
// lib/network_factory/voxel_morph.py:90-91, DiceLossMultiClass lib/loss.py:410-476 with a label target,
// mask_to_one_hot lib/transforms.py:675-689), fused so that
//   * forward never materialises the warped C-channel map (629 MB at C = 32, 160x192x160): every voxel's C sampled
//     values go straight into the three Dice sums  S_c = sum S_w,  T_c = #[S_t = c],  I_c = sum S_w [S_t = c];
//   * backward uses the structure of the Dice gradient, dL/dS_w[c][v] = gS_c + gI_c [S_t(v) = c]:
//       dL/dP[c][q] = gS_c * Wsum[q] + gI_c * L_c[q],   Wsum[q] = sum_v w(v->q),   L_c[q] = sum_{v: S_t(v)=c} w(v->q)
//     so the trilinear scatter needs 16 scalar reductions per voxel (8 into Wsum, 8 into the label's plane) instead
//     of 8*C, followed by one dense pass.  dL/dphi keeps the 8*C gathers of the forward.
// Sums are deterministic (shuffle tree + fixed-order second stage); the scatter uses fp32 atomics like ATen's
// grid_sampler_3d_backward.
#include "common.cuh"
#include "warp_common.cuh"

namespace {

constexpr int WD_THREADS = 128;

__device__ __forceinline__ int wd_label(const void* t, int kind, int64_t i) {
  if (kind == 0) return (int)((const uint8_t*)t)[i];
  if (kind == 1) return (int)((const int64_t*)t)[i];
  return ((const int32_t*)t)[i];
}

// partials [N][gridDim.x][3][C]
template <int CP, bool ADD_ID>
__global__ void __launch_bounds__(WD_THREADS) warp_dice_fwd_kernel(const float* __restrict__ src, const float* __restrict__ field,
                                                                   const void* __restrict__ labels, int kind, WarpGeom g,
                                                                   float* __restrict__ partials) {
  const int n = blockIdx.y, C = g.C;
  const int64_t Vo = (int64_t)g.Do * g.Ho * g.Wo, Vs = (int64_t)g.D * g.H * g.W;
  const float* s = src + (int64_t)n * C * Vs;
  float aS[CP], aT[CP], aI[CP];
#pragma unroll
  for (int c = 0; c < CP; ++c) aS[c] = aT[c] = aI[c] = 0.f;
  for (int64_t v = (int64_t)blockIdx.x * WD_THREADS + threadIdx.x; v < Vo; v += (int64_t)gridDim.x * WD_THREADS) {
    const int x = (int)(v % g.Wo), y = (int)((v / g.Wo) % g.Ho), z = (int)(v / ((int64_t)g.Wo * g.Ho));
    float px, py, pz;
    load_phi<ADD_ID>(field + (int64_t)n * 3 * Vo, Vo, v, x, y, z, g, px, py, pz);
    Corners c;
    make_corners(unnormalize(px, g.W), unnormalize(py, g.H), unnormalize(pz, g.D), g, c);
    const int lab = wd_label(labels, kind, (int64_t)n * Vo + v);
#pragma unroll
    for (int ch = 0; ch < CP; ++ch) {
      if (ch < C) {
        const float* sp = s + (int64_t)ch * Vs;
        float p = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (c.ok[k]) p += __ldg(sp + c.off[k]) * c.w[k];
        const bool hit = (lab == ch);
        aS[ch] += p;
        aT[ch] += hit ? 1.f : 0.f;
        aI[ch] += hit ? p : 0.f;
      }
    }
  }
  __shared__ float red[WD_THREADS / 32][3 * CP];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int ch = 0; ch < CP; ++ch) {
    const float a = warp_sum(aS[ch]), b = warp_sum(aT[ch]), d = warp_sum(aI[ch]);
    if (lane == 0) { red[w][ch] = a; red[w][CP + ch] = b; red[w][2 * CP + ch] = d; }
  }
  __syncthreads();
  float* out = partials + ((int64_t)n * gridDim.x + blockIdx.x) * 3 * C;
  for (int i = threadIdx.x; i < 3 * C; i += WD_THREADS) {
    const int q = i / C, ch = i - q * C;
    float acc = 0.f;
#pragma unroll
    for (int ww = 0; ww < WD_THREADS / 32; ++ww) acc += red[ww][q * CP + ch];
    out[i] = acc;
  }
}

__global__ void warp_dice_finalize_kernel(const float* __restrict__ partials, int nblocks, int C3, float* __restrict__ sums) {
  const int n = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C3) return;
  const float* p = partials + (int64_t)n * nblocks * C3 + i;
  double acc = 0.0;
  for (int b = 0; b < nblocks; ++b) acc += (double)p[(int64_t)b * C3];
  sums[(int64_t)n * C3 + i] = (float)acc;
}

// scatter (Wsum, L) + grad_field.  lacc = grad_prob buffer used as the L accumulator (zeroed by the entry point).
template <int CP, bool ADD_ID>
__global__ void __launch_bounds__(WD_THREADS) warp_dice_bwd_kernel(const float* __restrict__ src, const float* __restrict__ field,
                                                                   const void* __restrict__ labels, int kind,
                                                                   const float* __restrict__ gS, const float* __restrict__ gI,
                                                                   WarpGeom g, float* __restrict__ wsum, float* __restrict__ lacc,
                                                                   float* __restrict__ gfield) {
  const int n = blockIdx.y, C = g.C;
  __shared__ float sg[2][CP];
  for (int i = threadIdx.x; i < 2 * CP; i += WD_THREADS) {
    const int q = i / CP, ch = i - q * CP;
    sg[q][ch] = (ch < C) ? (q == 0 ? gS : gI)[n * C + ch] : 0.f;
  }
  __syncthreads();
  const int64_t Vo = (int64_t)g.Do * g.Ho * g.Wo, Vs = (int64_t)g.D * g.H * g.W;
  const float* s = src + (int64_t)n * C * Vs;
  float* ws = wsum ? wsum + (int64_t)n * Vs : nullptr;
  float* la = lacc ? lacc + (int64_t)n * C * Vs : nullptr;
  for (int64_t v = (int64_t)blockIdx.x * WD_THREADS + threadIdx.x; v < Vo; v += (int64_t)gridDim.x * WD_THREADS) {
    const int x = (int)(v % g.Wo), y = (int)((v / g.Wo) % g.Ho), z = (int)(v / ((int64_t)g.Wo * g.Ho));
    float px, py, pz;
    load_phi<ADD_ID>(field + (int64_t)n * 3 * Vo, Vo, v, x, y, z, g, px, py, pz);
    Corners c;
    make_corners(unnormalize(px, g.W), unnormalize(py, g.H), unnormalize(pz, g.D), g, c);
    const int lab = wd_label(labels, kind, (int64_t)n * Vo + v);
    if (ws) {
      float* lp = (la && lab >= 0 && lab < C) ? la + (int64_t)lab * Vs : nullptr;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (!c.ok[k]) continue;
        atomicAdd(ws + c.off[k], c.w[k]);
        if (lp) atomicAdd(lp + c.off[k], c.w[k]);
      }
    }
    if (gfield) {
      // sum over channels of the upstream gradient times the corner values, per corner
      float cv[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) cv[k] = 0.f;
#pragma unroll 4
      for (int ch = 0; ch < C; ++ch) {
        const float gc = sg[0][ch] + ((lab == ch) ? sg[1][ch] : 0.f);
        const float* sp = s + (int64_t)ch * Vs;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (c.ok[k]) cv[k] = fmaf(__ldg(sp + c.off[k]), gc, cv[k]);
      }
      float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int dx = k & 1, dy = (k >> 1) & 1, dz = k >> 2;
        const float val = cv[k];
        gx += (dx ? val : -val) * c.fy[dy] * c.fz[dz];
        gy += (dy ? val : -val) * c.fx[dx] * c.fz[dz];
        gz += (dz ? val : -val) * c.fx[dx] * c.fy[dy];
      }
      float* gf = gfield + (int64_t)n * 3 * Vo;
      gf[v] = gx * (0.5f * (float)(g.W - 1));
      gf[Vo + v] = gy * (0.5f * (float)(g.H - 1));
      gf[2 * Vo + v] = gz * (0.5f * (float)(g.D - 1));
    }
  }
}

// grad_prob[c][q] = gS_c * Wsum[q] + gI_c * L_c[q]   (in place on the L accumulator)
__global__ void __launch_bounds__(256) warp_dice_bwd_dense_kernel(const float* __restrict__ wsum, const float* __restrict__ gS,
                                                                  const float* __restrict__ gI, int C, int64_t Vs,
                                                                  float* __restrict__ gp) {
  const int ch = blockIdx.y, n = blockIdx.z;
  const float a = gS[n * C + ch], b = gI[n * C + ch];
  const float* ws = wsum + (int64_t)n * Vs;
  float* p = gp + ((int64_t)n * C + ch) * Vs;
  if ((Vs & 3) == 0) {
    const float4* w4 = reinterpret_cast<const float4*>(ws);
    float4* p4 = reinterpret_cast<float4*>(p);
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < Vs / 4; i += (int64_t)gridDim.x * 256) {
      const float4 w = w4[i];
      float4 l = p4[i];
      l.x = fmaf(b, l.x, a * w.x); l.y = fmaf(b, l.y, a * w.y); l.z = fmaf(b, l.z, a * w.z); l.w = fmaf(b, l.w, a * w.w);
      p4[i] = l;
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < Vs; i += (int64_t)gridDim.x * 256) p[i] = fmaf(b, p[i], a * ws[i]);
  }
}

inline int wd_blocks(int64_t V) {
  int64_t b = da_cdiv(V, (int64_t)WD_THREADS * 2);
  const int64_t cap = (int64_t)DA_NUM_SMS * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

#define WD_DISPATCH(CALL)                                                                   \
  if (C <= 4) { CALL(4); } else if (C <= 8) { CALL(8); } else if (C <= 16) { CALL(16); } else { CALL(32); }

// workspace of the forward (block partials) / backward (Wsum [N, D*H*W])
DA_API int64_t da_warp_dice_fwd_workspace_bytes(int N, int C, int64_t Vo) { return (int64_t)sizeof(float) * N * wd_blocks(Vo) * 3 * C + 256; }
DA_API int64_t da_warp_dice_bwd_workspace_bytes(int N, int64_t Vs) { return (int64_t)sizeof(float) * N * Vs + 256; }

// prob [N,C,D,H,W]; field [N,3,Do,Ho,Wo] (+ identity if add_identity); labels [N,Do,Ho,Wo] (kind 0 uint8, 1 int64, 3 int32).
// sums [N,3,C] = (S, T, I) of the warped map against the labels.  C <= 32.
DA_API int da_warp_dice_sums_fwd(const float* prob, const float* field, int add_identity, const void* labels, int label_kind,
                                 int N, int C, int D, int H, int W, int Do, int Ho, int Wo, float* sums, void* workspace,
                                 int64_t workspace_bytes, cudaStream_t stream) {
  DA_REQUIRE(prob && field && labels && sums && workspace, "da_warp_dice_sums_fwd: null pointer");
  DA_REQUIRE(C >= 1 && C <= 32, "da_warp_dice_sums_fwd: unsupported class count %d (1..32)", C);
  DA_REQUIRE(label_kind == 0 || label_kind == 1 || label_kind == 3, "da_warp_dice_sums_fwd: bad label kind");
  const int64_t Vo = (int64_t)Do * Ho * Wo;
  if (workspace_bytes < da_warp_dice_fwd_workspace_bytes(N, C, Vo)) { da_set_error("da_warp_dice_sums_fwd: workspace too small"); return DA_ERR_WORKSPACE; }
  WarpGeom g{N, C, D, H, W, Do, Ho, Wo};
  const int nb = wd_blocks(Vo);
  dim3 grid(nb, N);
#define CALL(CP)                                                                                                          \
  do {                                                                                                                    \
    if (add_identity) warp_dice_fwd_kernel<CP, true><<<grid, WD_THREADS, 0, stream>>>(prob, field, labels, label_kind, g, (float*)workspace); \
    else warp_dice_fwd_kernel<CP, false><<<grid, WD_THREADS, 0, stream>>>(prob, field, labels, label_kind, g, (float*)workspace);            \
  } while (0)
  WD_DISPATCH(CALL)
#undef CALL
  int rc = da_check_launch("da_warp_dice_sums_fwd");
  if (rc) return rc;
  dim3 g2((3 * C + 127) / 128, N);
  warp_dice_finalize_kernel<<<g2, 128, 0, stream>>>((const float*)workspace, nb, 3 * C, sums);
  return da_check_launch("da_warp_dice_sums_fwd/finalize");
}

// gS, gI [N,C]: upstream gradients w.r.t. S and I.  grad_prob (nullable) [N,C,D,H,W]; grad_field (nullable) [N,3,Do,Ho,Wo].
DA_API int da_warp_dice_sums_bwd(const float* prob, const float* field, int add_identity, const void* labels, int label_kind,
                                 const float* gS, const float* gI, int N, int C, int D, int H, int W, int Do, int Ho, int Wo,
                                 float* grad_prob, float* grad_field, void* workspace, int64_t workspace_bytes,
                                 cudaStream_t stream) {
  DA_REQUIRE(prob && field && labels && gS && gI, "da_warp_dice_sums_bwd: null pointer");
  DA_REQUIRE(grad_prob || grad_field, "da_warp_dice_sums_bwd: nothing to compute");
  DA_REQUIRE(C >= 1 && C <= 32, "da_warp_dice_sums_bwd: unsupported class count %d (1..32)", C);
  const int64_t Vo = (int64_t)Do * Ho * Wo, Vs = (int64_t)D * H * W;
  float* wsum = nullptr;
  if (grad_prob) {
    DA_REQUIRE(workspace, "da_warp_dice_sums_bwd: workspace needed for grad_prob");
    if (workspace_bytes < da_warp_dice_bwd_workspace_bytes(N, Vs)) { da_set_error("da_warp_dice_sums_bwd: workspace too small"); return DA_ERR_WORKSPACE; }
    wsum = (float*)workspace;
    cudaError_t e = cudaMemsetAsync(wsum, 0, sizeof(float) * (size_t)N * Vs, stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(grad_prob, 0, sizeof(float) * (size_t)N * C * Vs, stream);
    if (e != cudaSuccess) { da_set_error("da_warp_dice_sums_bwd memset: %s", cudaGetErrorString(e)); return (int)e; }
  }
  WarpGeom g{N, C, D, H, W, Do, Ho, Wo};
  dim3 grid(wd_blocks(Vo), N);
#define CALL(CP)                                                                                                                   \
  do {                                                                                                                             \
    if (add_identity) warp_dice_bwd_kernel<CP, true><<<grid, WD_THREADS, 0, stream>>>(prob, field, labels, label_kind, gS, gI, g, wsum, grad_prob, grad_field); \
    else warp_dice_bwd_kernel<CP, false><<<grid, WD_THREADS, 0, stream>>>(prob, field, labels, label_kind, gS, gI, g, wsum, grad_prob, grad_field);            \
  } while (0)
  WD_DISPATCH(CALL)
#undef CALL
  int rc = da_check_launch("da_warp_dice_sums_bwd");
  if (rc || !grad_prob) return rc;
  int64_t nb = da_cdiv(Vs / 4 > 0 ? Vs / 4 : Vs, 256);
  if (nb > 4096) nb = 4096;
  warp_dice_bwd_dense_kernel<<<dim3((unsigned)nb, C, N), 256, 0, stream>>>(wsum, gS, gI, C, Vs, grad_prob);
  return da_check_launch("da_warp_dice_sums_bwd/dense");
}
