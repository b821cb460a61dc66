// warp_dice: the anatomy-similarity term of the joint step,  dice(grid_sample(P, phi), onehot(S_t))  with P the
// moving segmentation's class probabilities (reference ingredients: F.grid_sample as called at
// lib/network_factory/voxel_morph.py:90-91, DiceLossMultiClass lib/loss.py:410-476 with a label target,
// mask_to_one_hot lib/transforms.py:675-689), fused so that
//   * forward never materialises the warped C-channel map (629 MB at C = 32, 160x192x160): every voxel's C sampled
//     values go straight into the three Dice sums  S_c = sum S_w,  T_c = #[S_t = c],  I_c = sum S_w [S_t = c];
//   * backward uses the structure of the Dice gradient, dL/dS_w[c][v] = gS_c + gI_c [S_t(v) = c]:
//       dL/dP[c][q] = gS_c * Wsum[q] + gI_c * L_c[q],   Wsum[q] = sum_v w(v->q),   L_c[q] = sum_{v: S_t(v)=c} w(v->q)
//     so the trilinear scatter needs scalar reductions per voxel instead of 8*C, followed by one dense pass;
//   * the same Wsum turns the forward's S_c into a dense dot product and dL/dphi into 16 gathers (see "Scatter
//     formulation" below): neither direction reads 8*C values per voxel.
// T and I sums are deterministic (private shared-memory columns, shuffle tree, fixed-order second stage); Wsum and L
// use fp32 atomics like ATen's grid_sampler_3d_backward, so S_c carries that ordering noise (~1e-7 relative) unless
// the caller asks for the gather kernel (wsum = null).
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

// one warp per output element: lanes stride over the blocks' partial rows, fixed-order fp64 fold
__global__ void __launch_bounds__(256) warp_dice_finalize_kernel(const float* __restrict__ partials, int nblocks, int C3,
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

// ---------------------------------------------------------------------------------------------------------------------
// Scatter formulation (default; the gather kernels above stay as the bitwise-deterministic variant):
//   S_c = sum_v sum_k w_k(v) P[c][q_k(v)] = sum_q P[c][q] Wsum[q]         one 8-scalar scatter + one dense dot product
//   I_c = sum_{v: S_t(v)=c} sum_k w_k P[c][q_k]                          8 gathers per voxel (only the label's plane)
//   dL/dphi(v) needs sum_c (gS_c + gI_c [S_t(v)=c]) P[c][q_k] = Q[q_k] + gI_lab P[lab][q_k],   Q = sum_c gS_c P[c]
// so neither direction touches 8*C values per voxel: forward 8 atomics + 8 gathers, backward 8 atomics + 16 gathers,
// plus dense streaming passes over P.  Wsum is produced by the forward and handed to the backward.
constexpr int WD2_THREADS = 128;

// scatter Wsum; partial T, I per block: partials [N][gridDim.x][2][C]
template <bool ADD_ID>
__global__ void __launch_bounds__(WD2_THREADS) wd2_fwd_kernel(const float* __restrict__ src, const float* __restrict__ field,
                                                              const void* __restrict__ labels, int kind, WarpGeom g,
                                                              float* __restrict__ wsum, float* __restrict__ partials) {
  __shared__ float sT[32][WD2_THREADS], sI[32][WD2_THREADS];  // one private column per thread: no atomics, fixed order
  const int n = blockIdx.y, C = g.C, tid = threadIdx.x;
  for (int c = 0; c < C; ++c) { sT[c][tid] = 0.f; sI[c][tid] = 0.f; }
  const int64_t Vo = (int64_t)g.Do * g.Ho * g.Wo, Vs = (int64_t)g.D * g.H * g.W;
  const float* s = src + (int64_t)n * C * Vs;
  float* ws = wsum + (int64_t)n * Vs;
  for (int64_t v = (int64_t)blockIdx.x * WD2_THREADS + tid; v < Vo; v += (int64_t)gridDim.x * WD2_THREADS) {
    const int x = (int)(v % g.Wo), y = (int)((v / g.Wo) % g.Ho), z = (int)(v / ((int64_t)g.Wo * g.Ho));
    float px, py, pz;
    load_phi<ADD_ID>(field + (int64_t)n * 3 * Vo, Vo, v, x, y, z, g, px, py, pz);
    Corners c;
    make_corners(unnormalize(px, g.W), unnormalize(py, g.H), unnormalize(pz, g.D), g, c);
    const int lab = wd_label(labels, kind, (int64_t)n * Vo + v);
    const bool lv = lab >= 0 && lab < C;
    const float* sp = s + (int64_t)(lv ? lab : 0) * Vs;
    float p = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (c.ok[k]) {
        atomicAdd(ws + c.off[k], c.w[k]);
        if (lv) p += __ldg(sp + c.off[k]) * c.w[k];
      }
    if (lv) { sT[lab][tid] += 1.f; sI[lab][tid] += p; }
  }
  __shared__ float red[WD2_THREADS / 32][2 * 32];
  const int lane = tid & 31, w = tid >> 5;
  for (int ch = 0; ch < C; ++ch) {
    const float a = warp_sum(sT[ch][tid]), b = warp_sum(sI[ch][tid]);
    if (lane == 0) { red[w][ch] = a; red[w][32 + ch] = b; }
  }
  __syncthreads();
  float* out = partials + ((int64_t)n * gridDim.x + blockIdx.x) * 2 * C;
  for (int i = tid; i < 2 * C; i += WD2_THREADS) {
    const int q = i / C, ch = i - q * C;
    float acc = 0.f;
#pragma unroll
    for (int ww = 0; ww < WD2_THREADS / 32; ++ww) acc += red[ww][q * 32 + ch];
    out[i] = acc;
  }
}

// partial S per block: partials [N][gridDim.x][C];  S_c = <P[c], Wsum>
template <int CP>
__global__ void __launch_bounds__(256) wd2_dot_kernel(const float* __restrict__ src, const float* __restrict__ wsum, int C, int64_t Vs,
                                                      float* __restrict__ partials) {
  const int n = blockIdx.y;
  const float* s = src + (int64_t)n * C * Vs;
  const float* ws = wsum + (int64_t)n * Vs;
  float acc[CP];
#pragma unroll
  for (int c = 0; c < CP; ++c) acc[c] = 0.f;
  if ((Vs & 3) == 0) {
    const int64_t V4 = Vs >> 2;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < V4; i += (int64_t)gridDim.x * 256) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(ws) + i);
#pragma unroll
      for (int c = 0; c < CP; ++c)
        if (c < C) {
          const float4 p = __ldcs(reinterpret_cast<const float4*>(s + (int64_t)c * Vs) + i);
          acc[c] += (p.x * w.x + p.y * w.y) + (p.z * w.z + p.w * w.w);
        }
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < Vs; i += (int64_t)gridDim.x * 256) {
      const float w = ws[i];
#pragma unroll
      for (int c = 0; c < CP; ++c)
        if (c < C) acc[c] += s[(int64_t)c * Vs + i] * w;
    }
  }
  __shared__ float red[8][CP];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < CP; ++c) {
    const float a = warp_sum(acc[c]);
    if (lane == 0) red[wp][c] = a;
  }
  __syncthreads();
  float* out = partials + ((int64_t)n * gridDim.x + blockIdx.x) * C;
  for (int c = threadIdx.x; c < C; c += 256) {
    float a = 0.f;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) a += red[ww][c];
    out[c] = a;
  }
}

// sums [N][3][C] from the two partial arrays
__global__ void __launch_bounds__(256) wd2_finalize_kernel(const float* __restrict__ pS, int nbS, const float* __restrict__ pTI, int nbTI,
                                                           int C, float* __restrict__ sums) {
  // one warp per output element (a single thread per element walked up to 1184 partial rows serially: 158 us)
  const int n = blockIdx.y, lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= 3 * C) return;
  const int q = i / C, ch = i - q * C;
  double acc = 0.0;
  if (q == 0) {
    const float* p = pS + (int64_t)n * nbS * C + ch;
    for (int b = lane; b < nbS; b += 32) acc += (double)p[(int64_t)b * C];
  } else {
    const float* p = pTI + (int64_t)n * nbTI * 2 * C + (q - 1) * C + ch;
    for (int b = lane; b < nbTI; b += 32) acc += (double)p[(int64_t)b * 2 * C];
  }
  acc = warp_sum(acc);
  if (lane == 0) sums[(int64_t)n * 3 * C + i] = (float)acc;
}

// Q[q] = sum_c gS_c P[c][q]
template <int CP>
__global__ void __launch_bounds__(256) wd2_q_kernel(const float* __restrict__ src, const float* __restrict__ gS, int C, int64_t Vs,
                                                    float* __restrict__ Q) {
  const int n = blockIdx.y;
  const float* s = src + (int64_t)n * C * Vs;
  float* q = Q + (int64_t)n * Vs;
  float g[CP];
#pragma unroll
  for (int c = 0; c < CP; ++c) g[c] = c < C ? gS[n * C + c] : 0.f;
  if ((Vs & 3) == 0) {
    const int64_t V4 = Vs >> 2;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < V4; i += (int64_t)gridDim.x * 256) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int c = 0; c < CP; ++c)
        if (c < C) {
          const float4 p = __ldg(reinterpret_cast<const float4*>(s + (int64_t)c * Vs) + i);
          a.x = fmaf(g[c], p.x, a.x); a.y = fmaf(g[c], p.y, a.y); a.z = fmaf(g[c], p.z, a.z); a.w = fmaf(g[c], p.w, a.w);
        }
      reinterpret_cast<float4*>(q)[i] = a;
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < Vs; i += (int64_t)gridDim.x * 256) {
      float a = 0.f;
#pragma unroll
      for (int c = 0; c < CP; ++c)
        if (c < C) a = fmaf(g[c], s[(int64_t)c * Vs + i], a);
      q[i] = a;
    }
  }
}

// scatter L (and Wsum when the forward's is not at hand) + grad_field from Q and the label's plane
template <bool ADD_ID>
__global__ void __launch_bounds__(WD2_THREADS) wd2_bwd_kernel(const float* __restrict__ src, const float* __restrict__ field,
                                                              const void* __restrict__ labels, int kind,
                                                              const float* __restrict__ gI, const float* __restrict__ Q, WarpGeom g,
                                                              float* __restrict__ wsum, float* __restrict__ lacc,
                                                              float* __restrict__ gfield) {
  const int n = blockIdx.y, C = g.C;
  __shared__ float sg[32];
  if (threadIdx.x < 32) sg[threadIdx.x] = threadIdx.x < C ? gI[n * C + threadIdx.x] : 0.f;
  __syncthreads();
  const int64_t Vo = (int64_t)g.Do * g.Ho * g.Wo, Vs = (int64_t)g.D * g.H * g.W;
  const float* s = src + (int64_t)n * C * Vs;
  const float* qn = Q ? Q + (int64_t)n * Vs : nullptr;
  float* ws = wsum ? wsum + (int64_t)n * Vs : nullptr;
  float* la = lacc ? lacc + (int64_t)n * C * Vs : nullptr;
  for (int64_t v = (int64_t)blockIdx.x * WD2_THREADS + threadIdx.x; v < Vo; v += (int64_t)gridDim.x * WD2_THREADS) {
    const int x = (int)(v % g.Wo), y = (int)((v / g.Wo) % g.Ho), z = (int)(v / ((int64_t)g.Wo * g.Ho));
    float px, py, pz;
    load_phi<ADD_ID>(field + (int64_t)n * 3 * Vo, Vo, v, x, y, z, g, px, py, pz);
    Corners c;
    make_corners(unnormalize(px, g.W), unnormalize(py, g.H), unnormalize(pz, g.D), g, c);
    const int lab = wd_label(labels, kind, (int64_t)n * Vo + v);
    const bool lv = lab >= 0 && lab < C;
    const int64_t loff = (int64_t)(lv ? lab : 0) * Vs;
    float cv[8];
    const float gl = lv ? sg[lab] : 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      cv[k] = 0.f;
      if (!c.ok[k]) continue;
      if (ws) atomicAdd(ws + c.off[k], c.w[k]);
      if (la && lv) atomicAdd(la + loff + c.off[k], c.w[k]);
      if (gfield) cv[k] = fmaf(gl, __ldg(s + loff + c.off[k]), __ldg(qn + c.off[k]));
    }
    if (gfield) {
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

inline int wd2_dense_blocks(int64_t Vs) {
  int64_t nb = da_cdiv((Vs & 3) == 0 ? Vs / 4 : Vs, (int64_t)256 * 4);
  const int64_t cap = (int64_t)DA_NUM_SMS * 8;
  return (int)(nb < 1 ? 1 : (nb > cap ? cap : nb));
}

inline int wd_blocks(int64_t V) {
  int64_t b = da_cdiv(V, (int64_t)WD_THREADS * 2);
  const int64_t cap = (int64_t)DA_NUM_SMS * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

#define WD_DISPATCH(CALL)                                                                   \
  if (C <= 4) { CALL(4); } else if (C <= 8) { CALL(8); } else if (C <= 16) { CALL(16); } else { CALL(32); }

// workspace of the forward (block partials) / backward (Q and, when the forward's Wsum is not passed, Wsum: 2 x [N, D*H*W])
DA_API int64_t da_warp_dice_fwd_workspace_bytes(int N, int C, int64_t Vo, int64_t Vs) {
  return (int64_t)sizeof(float) * N * ((int64_t)wd_blocks(Vo) * 3 * C + (int64_t)wd2_dense_blocks(Vs) * C) + 256;
}
DA_API int64_t da_warp_dice_bwd_workspace_bytes(int N, int64_t Vs) { return (int64_t)sizeof(float) * 2 * N * Vs + 256; }

// prob [N,C,D,H,W]; field [N,3,Do,Ho,Wo] (+ identity if add_identity); labels [N,Do,Ho,Wo] (kind 0 uint8, 1 int64, 3 int32).
// sums [N,3,C] = (S, T, I) of the warped map against the labels.  C <= 32.
// wsum (nullable) [N,D,H,W]: when given, the scatter formulation runs and leaves Wsum there for the backward; when
// null, the gather kernel computes the same sums with a fixed summation order (bitwise reproducible, ~5x slower).
DA_API int da_warp_dice_sums_fwd(const float* prob, const float* field, int add_identity, const void* labels, int label_kind,
                                 int N, int C, int D, int H, int W, int Do, int Ho, int Wo, float* sums, float* wsum,
                                 void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  DA_REQUIRE(prob && field && labels && sums && workspace, "da_warp_dice_sums_fwd: null pointer");
  DA_REQUIRE(C >= 1 && C <= 32, "da_warp_dice_sums_fwd: unsupported class count %d (1..32)", C);
  DA_REQUIRE(label_kind == 0 || label_kind == 1 || label_kind == 3, "da_warp_dice_sums_fwd: bad label kind");
  const int64_t Vo = (int64_t)Do * Ho * Wo, Vs = (int64_t)D * H * W;
  if (workspace_bytes < da_warp_dice_fwd_workspace_bytes(N, C, Vo, Vs)) { da_set_error("da_warp_dice_sums_fwd: workspace too small"); return DA_ERR_WORKSPACE; }
  WarpGeom g{N, C, D, H, W, Do, Ho, Wo};
  const int nb = wd_blocks(Vo);
  dim3 grid(nb, N);
  if (wsum) {
    cudaError_t e = cudaMemsetAsync(wsum, 0, sizeof(float) * (size_t)N * Vs, stream);
    if (e != cudaSuccess) { da_set_error("da_warp_dice_sums_fwd memset: %s", cudaGetErrorString(e)); return (int)e; }
    float* pTI = (float*)workspace;
    float* pS = pTI + (int64_t)N * nb * 2 * C;
    if (add_identity) wd2_fwd_kernel<true><<<grid, WD2_THREADS, 0, stream>>>(prob, field, labels, label_kind, g, wsum, pTI);
    else wd2_fwd_kernel<false><<<grid, WD2_THREADS, 0, stream>>>(prob, field, labels, label_kind, g, wsum, pTI);
    int rc = da_check_launch("da_warp_dice_sums_fwd/scatter");
    if (rc) return rc;
    const int nbd = wd2_dense_blocks(Vs);
#define CALL(CP) wd2_dot_kernel<CP><<<dim3(nbd, N), 256, 0, stream>>>(prob, wsum, C, Vs, pS)
    WD_DISPATCH(CALL)
#undef CALL
    rc = da_check_launch("da_warp_dice_sums_fwd/dot");
    if (rc) return rc;
    wd2_finalize_kernel<<<dim3((3 * C + 7) / 8, N), 256, 0, stream>>>(pS, nbd, pTI, nb, C, sums);
    return da_check_launch("da_warp_dice_sums_fwd/finalize");
  }
#define CALL(CP)                                                                                                          \
  do {                                                                                                                    \
    if (add_identity) warp_dice_fwd_kernel<CP, true><<<grid, WD_THREADS, 0, stream>>>(prob, field, labels, label_kind, g, (float*)workspace); \
    else warp_dice_fwd_kernel<CP, false><<<grid, WD_THREADS, 0, stream>>>(prob, field, labels, label_kind, g, (float*)workspace);            \
  } while (0)
  WD_DISPATCH(CALL)
#undef CALL
  int rc = da_check_launch("da_warp_dice_sums_fwd");
  if (rc) return rc;
  dim3 g2((3 * C + 7) / 8, N);
  warp_dice_finalize_kernel<<<g2, 256, 0, stream>>>((const float*)workspace, nb, 3 * C, sums);
  return da_check_launch("da_warp_dice_sums_fwd/finalize");
}

// gS, gI [N,C]: upstream gradients w.r.t. S and I.  grad_prob (nullable) [N,C,D,H,W]; grad_field (nullable) [N,3,Do,Ho,Wo].
// wsum (nullable): the forward's Wsum; when null it is recomputed into the workspace.
DA_API int da_warp_dice_sums_bwd(const float* prob, const float* field, int add_identity, const void* labels, int label_kind,
                                 const float* gS, const float* gI, const float* wsum_fwd, int N, int C, int D, int H, int W,
                                 int Do, int Ho, int Wo, float* grad_prob, float* grad_field, void* workspace,
                                 int64_t workspace_bytes, cudaStream_t stream) {
  DA_REQUIRE(prob && field && labels && gS && gI && workspace, "da_warp_dice_sums_bwd: null pointer");
  DA_REQUIRE(grad_prob || grad_field, "da_warp_dice_sums_bwd: nothing to compute");
  DA_REQUIRE(C >= 1 && C <= 32, "da_warp_dice_sums_bwd: unsupported class count %d (1..32)", C);
  const int64_t Vo = (int64_t)Do * Ho * Wo, Vs = (int64_t)D * H * W;
  if (workspace_bytes < da_warp_dice_bwd_workspace_bytes(N, Vs)) { da_set_error("da_warp_dice_sums_bwd: workspace too small"); return DA_ERR_WORKSPACE; }
  float* Q = (float*)workspace;
  float* wsum_new = nullptr;
  const float* wsum = wsum_fwd;
  cudaError_t e = cudaSuccess;
  if (grad_prob) {
    if (!wsum) {
      wsum_new = Q + (int64_t)N * Vs;
      wsum = wsum_new;
      e = cudaMemsetAsync(wsum_new, 0, sizeof(float) * (size_t)N * Vs, stream);
    }
    if (e == cudaSuccess) e = cudaMemsetAsync(grad_prob, 0, sizeof(float) * (size_t)N * C * Vs, stream);
    if (e != cudaSuccess) { da_set_error("da_warp_dice_sums_bwd memset: %s", cudaGetErrorString(e)); return (int)e; }
  }
  int rc;
  if (grad_field) {
    const int nbd = wd2_dense_blocks(Vs);
#define CALL(CP) wd2_q_kernel<CP><<<dim3(nbd, N), 256, 0, stream>>>(prob, gS, C, Vs, Q)
    WD_DISPATCH(CALL)
#undef CALL
    rc = da_check_launch("da_warp_dice_sums_bwd/q");
    if (rc) return rc;
  }
  WarpGeom g{N, C, D, H, W, Do, Ho, Wo};
  dim3 grid(wd_blocks(Vo), N);
  if (add_identity) wd2_bwd_kernel<true><<<grid, WD2_THREADS, 0, stream>>>(prob, field, labels, label_kind, gI, grad_field ? Q : nullptr, g, wsum_new, grad_prob, grad_field);
  else wd2_bwd_kernel<false><<<grid, WD2_THREADS, 0, stream>>>(prob, field, labels, label_kind, gI, grad_field ? Q : nullptr, g, wsum_new, grad_prob, grad_field);
  rc = da_check_launch("da_warp_dice_sums_bwd");
  if (rc || !grad_prob) return rc;
  int64_t nb = da_cdiv(Vs / 4 > 0 ? Vs / 4 : Vs, 256);
  if (nb > 4096) nb = 4096;
  warp_dice_bwd_dense_kernel<<<dim3((unsigned)nb, C, N), 256, 0, stream>>>(wsum, gS, gI, C, Vs, grad_prob);
  return da_check_launch("da_warp_dice_sums_bwd/dense");
}
