// The remaining entries of the reference's loss registry (lib/loss.py:739-750):
//   'ncc'  NormalizedCrossCorrelationLoss (lib/loss.py:485-501)      -> pair moments
//   'mse'  nn.MSELoss, 'L2' L2Loss (lib/loss.py:733-736)              -> pair moments
//   'gradient' gradientLoss incl. its '+' quirk (lib/loss.py:625-671) -> gradient sums
//   'cross_entropy' nn.CrossEntropyLoss, 'focal' FocalLoss (lib/loss.py:120-186),
//   'soft_cross_entropy' SoftCrossEntropy (lib/loss.py:96-117)        -> channel log-softmax terms
// All of them are one streaming pass over planar NCDHW data followed by a shuffle-tree reduction (fp64 block
// partials, fixed-order second stage => deterministic); the closing formulas on the handful of sums are evaluated by
// the host-side mirror (deepatlas_b200/losses.py).  Every kernel is HBM-bound: algorithmic bytes are one read of each
// input in the forward and one read of each input + one write of each gradient in the backward.
#include "common.cuh"

namespace {

constexpr int LX_THREADS = 256;
constexpr int LX_WARPS = LX_THREADS / 32;
constexpr int LX_BLOCKS = DA_NUM_SMS * 4;  // blocks per sample (grid.x); grid.y = N

// ------------------------------------------------------------------------------------------------------------------
// pair moments: per sample n the nine numbers
//   0 Sa = sum a        1 Sb = sum b        2 Saa = sum a^2      3 Sbb = sum b^2      4 Sab = sum a*b
//   5 Sdd = sum (a-b)^2 6 cAA = sum (a-ma)^2 7 cBB = sum (b-mb)^2 8 cAB = sum (a-ma)(b-mb)
// products of two fp32 values are exact in fp64, so the centred moments (closed in fp64 by the finalize kernel) carry
// no cancellation error.  b may be null (L2Loss): the b terms are 0 and Sdd = Saa.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LX_THREADS) pair_moments_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                                  int64_t V, double* __restrict__ partials) {
  __shared__ double red[LX_WARPS];
  const int n = blockIdx.y;
  const float* an = a + (int64_t)n * V;
  const float* bn = b ? b + (int64_t)n * V : nullptr;
  double s[6] = {0, 0, 0, 0, 0, 0};
  for (int64_t i = (int64_t)blockIdx.x * LX_THREADS + threadIdx.x; i < V; i += (int64_t)gridDim.x * LX_THREADS) {
    const double x = (double)__ldg(an + i);
    const double y = bn ? (double)__ldg(bn + i) : 0.0;
    const double d = x - y;
    s[0] += x; s[1] += y; s[2] += x * x; s[3] += y * y; s[4] += x * y; s[5] += d * d;
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double t = block_sum<double, LX_WARPS>(s[k], red);
    if (threadIdx.x == 0) partials[((int64_t)n * gridDim.x + blockIdx.x) * 6 + k] = t;
  }
}

__global__ void pair_moments_finalize_kernel(const double* __restrict__ partials, int nb, int N, int64_t V,
                                             float* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double s[6] = {0, 0, 0, 0, 0, 0};
  for (int b = 0; b < nb; ++b)
    for (int k = 0; k < 6; ++k) s[k] += partials[((int64_t)n * nb + b) * 6 + k];
  float* o = out + (int64_t)n * 9;
  for (int k = 0; k < 6; ++k) o[k] = (float)s[k];
  const double v = (double)V;
  o[6] = (float)(s[2] - s[0] * s[0] / v);
  o[7] = (float)(s[3] - s[1] * s[1] / v);
  o[8] = (float)(s[4] - s[0] * s[1] / v);
}

// out[n][i] = coef[n][0]*a[n][i] + coef[n][1]*b[n][i] + coef[n][2]: the gradient of any function of the nine moments
__global__ void __launch_bounds__(LX_THREADS) affine2_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                             const float* __restrict__ coef, int64_t V,
                                                             float* __restrict__ out) {
  const int n = blockIdx.y;
  const float ca = coef[n * 3 + 0], cb = coef[n * 3 + 1], cc = coef[n * 3 + 2];
  const float* an = a + (int64_t)n * V;
  const float* bn = b ? b + (int64_t)n * V : nullptr;
  float* on = out + (int64_t)n * V;
  for (int64_t i = (int64_t)blockIdx.x * LX_THREADS + threadIdx.x; i < V; i += (int64_t)gridDim.x * LX_THREADS)
    on[i] = ca * __ldg(an + i) + (bn ? cb * __ldg(bn + i) : 0.f) + cc;
}

// ------------------------------------------------------------------------------------------------------------------
// gradientLoss (lib/loss.py:655-670): per (n, channel) the three sums of f(r) with
//   r0 = u[d+2] - u[d]   over d in [0, D-2), all h, w           (lib/loss.py:655)
//   r1 = u[h+2] + u[h]   over h in [0, H-2)   -- the reference ADDS here (lib/loss.py:657), kept
//   r2 = u[w+2] + u[w]   over w in [0, W-2)   -- likewise (lib/loss.py:659)
// f = square (norm 'L2') or abs (any other norm: the reference only squares under 'L2').
// ------------------------------------------------------------------------------------------------------------------
struct GGeo { int D, H, W; int64_t sH, sD, V; };

__global__ void __launch_bounds__(LX_THREADS) gradient_fwd_kernel(const float* __restrict__ u, GGeo g, int l1,
                                                                  double* __restrict__ partials) {
  __shared__ double red[LX_WARPS];
  const int nc = blockIdx.y;
  const float* uc = u + (int64_t)nc * g.V;
  float acc[3] = {0.f, 0.f, 0.f};
  for (int64_t i = (int64_t)blockIdx.x * LX_THREADS + threadIdx.x; i < g.V; i += (int64_t)gridDim.x * LX_THREADS) {
    const int x = (int)(i % g.W), y = (int)((i / g.W) % g.H), z = (int)(i / g.sD);
    const float c = __ldg(uc + i);
    if (z + 2 < g.D) { const float r = __ldg(uc + i + 2 * g.sD) - c; acc[0] += l1 ? fabsf(r) : r * r; }
    if (y + 2 < g.H) { const float r = __ldg(uc + i + 2 * g.sH) + c; acc[1] += l1 ? fabsf(r) : r * r; }
    if (x + 2 < g.W) { const float r = __ldg(uc + i + 2) + c; acc[2] += l1 ? fabsf(r) : r * r; }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double t = block_sum<double, LX_WARPS>((double)acc[k], red);
    if (threadIdx.x == 0) partials[((int64_t)nc * gridDim.x + blockIdx.x) * 3 + k] = t;
  }
}

__global__ void sums_finalize_kernel(const double* __restrict__ partials, int nb, int rows, int K, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * K) return;
  const int r = i / K, k = i - r * K;
  double acc = 0;
  for (int b = 0; b < nb; ++b) acc += partials[((int64_t)r * nb + b) * K + k];
  out[i] = (float)acc;
}

__device__ __forceinline__ float dfun(float r, int l1) { return l1 ? (r > 0.f ? 1.f : (r < 0.f ? -1.f : 0.f)) : 2.f * r; }

// gather form of the transposed stencils: deterministic, one read of u's neighbourhood, one write
__global__ void __launch_bounds__(LX_THREADS) gradient_bwd_kernel(const float* __restrict__ u, const float* __restrict__ gsums,
                                                                  GGeo g, int l1, float* __restrict__ gu) {
  const int nc = blockIdx.y;
  const float* uc = u + (int64_t)nc * g.V;
  const float g0 = gsums[nc * 3 + 0], g1 = gsums[nc * 3 + 1], g2 = gsums[nc * 3 + 2];
  for (int64_t i = (int64_t)blockIdx.x * LX_THREADS + threadIdx.x; i < g.V; i += (int64_t)gridDim.x * LX_THREADS) {
    const int x = (int)(i % g.W), y = (int)((i / g.W) % g.H), z = (int)(i / g.sD);
    const float c = __ldg(uc + i);
    float acc = 0.f;
    if (z >= 2) acc += g0 * dfun(c - __ldg(uc + i - 2 * g.sD), l1);
    if (z + 2 < g.D) acc -= g0 * dfun(__ldg(uc + i + 2 * g.sD) - c, l1);
    if (y >= 2) acc += g1 * dfun(c + __ldg(uc + i - 2 * g.sH), l1);
    if (y + 2 < g.H) acc += g1 * dfun(__ldg(uc + i + 2 * g.sH) + c, l1);
    if (x >= 2) acc += g2 * dfun(c + __ldg(uc + i - 2), l1);
    if (x + 2 < g.W) acc += g2 * dfun(__ldg(uc + i + 2) + c, l1);
    gu[(int64_t)nc * g.V + i] = acc;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// channel log-softmax terms.  One thread = one voxel; the C planes are read with the warp along W (coalesced).
//   mode 0  cross entropy, class-index target:  term = -w[t] * logp[t], weight = w[t]   (t == ignore_index: skipped)
//   mode 1  FocalLoss (lib/loss.py:149-186): term = -alpha[t] * (1 - probs)^gamma * log_p with
//           log_p = log_softmax(x)[t] and probs = F.nll_loss(P, t) = -P[t]  (so the factor is (1 + P[t])^gamma: the
//           reference's sign slip, kept), P = softmax(x) if focal_softmax else x
//   mode 2  SoftCrossEntropy, softmax=True  (lib/loss.py:114): term = -sum_c t_c * log_softmax(x)_c, soft target
//   mode 3  SoftCrossEntropy, softmax=False (lib/loss.py:116): term = -sum_c t_c * log(max(x_c, 1e-8))
// out2 = (sum of terms, sum of weights [mode 0] or number of voxels N*V [modes 1-3]).
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t load_label(const void* t, int kind, int64_t i) {
  if (kind == 0) return (int64_t) reinterpret_cast<const uint8_t*>(t)[i];
  if (kind == 1) return reinterpret_cast<const int64_t*>(t)[i];
  return (int64_t) reinterpret_cast<const int32_t*>(t)[i];
}

struct XentArgs {
  const float* x; const void* target; int kind, mode, C; int64_t V;
  const float* cw; float gamma; int focal_softmax; int64_t ignore_index;
};

__device__ __forceinline__ void softmax_stats(const float* __restrict__ xv, int C, int64_t V, float& m, float& s) {
  m = -INFINITY;
  for (int c = 0; c < C; ++c) m = fmaxf(m, __ldg(xv + (int64_t)c * V));
  s = 0.f;
  for (int c = 0; c < C; ++c) s += expf(__ldg(xv + (int64_t)c * V) - m);
}

__global__ void __launch_bounds__(LX_THREADS) xent_fwd_kernel(XentArgs a, double* __restrict__ partials) {
  __shared__ double red[LX_WARPS];
  const int n = blockIdx.y;
  double term = 0.0, wsum = 0.0;
  for (int64_t v = (int64_t)blockIdx.x * LX_THREADS + threadIdx.x; v < a.V; v += (int64_t)gridDim.x * LX_THREADS) {
    const float* xv = a.x + (int64_t)n * a.C * a.V + v;
    if (a.mode == 3) {
      const float* tv = reinterpret_cast<const float*>(a.target) + (int64_t)n * a.C * a.V + v;
      float acc = 0.f;
      for (int c = 0; c < a.C; ++c) acc -= __ldg(tv + (int64_t)c * a.V) * logf(fmaxf(__ldg(xv + (int64_t)c * a.V), 1e-8f));
      term += (double)acc; wsum += 1.0;
      continue;
    }
    float m, s;
    softmax_stats(xv, a.C, a.V, m, s);
    const float lse = m + logf(s);
    if (a.mode == 2) {
      const float* tv = reinterpret_cast<const float*>(a.target) + (int64_t)n * a.C * a.V + v;
      float acc = 0.f;
      for (int c = 0; c < a.C; ++c) acc -= __ldg(tv + (int64_t)c * a.V) * (__ldg(xv + (int64_t)c * a.V) - lse);
      term += (double)acc; wsum += 1.0;
      continue;
    }
    const int64_t t = load_label(a.target, a.kind, (int64_t)n * a.V + v);
    if (a.mode == 0 && t == a.ignore_index) continue;
    if (t < 0 || t >= a.C) { term += (double)NAN; continue; }  // the reference raises on an out-of-range label
    const float xt = __ldg(xv + t * a.V);
    const float lp = xt - lse;
    const float w = a.cw ? __ldg(a.cw + t) : 1.f;
    if (a.mode == 0) {
      term += (double)(-w * lp); wsum += (double)w;
    } else {
      const float P = a.focal_softmax ? expf(lp) : xt;
      term += (double)(-w * powf(1.f + P, a.gamma) * lp); wsum += 1.0;
    }
  }
  const double t0 = block_sum<double, LX_WARPS>(term, red);
  const double t1 = block_sum<double, LX_WARPS>(wsum, red);
  if (threadIdx.x == 0) {
    partials[((int64_t)n * gridDim.x + blockIdx.x) * 2 + 0] = t0;
    partials[((int64_t)n * gridDim.x + blockIdx.x) * 2 + 1] = t1;
  }
}

__global__ void xent_finalize_kernel(const double* __restrict__ partials, int count, float* __restrict__ out2) {
  // one warp, fixed order
  double a = 0, b = 0;
  for (int i = threadIdx.x; i < count; i += 32) { a += partials[2 * (int64_t)i]; b += partials[2 * (int64_t)i + 1]; }
  a = warp_sum(a); b = warp_sum(b);
  if (threadIdx.x == 0) { out2[0] = (float)a; out2[1] = (float)b; }
}

// gscale: device scalar = d loss / d out2[0]
__global__ void __launch_bounds__(LX_THREADS) xent_bwd_kernel(XentArgs a, const float* __restrict__ gscale,
                                                              float* __restrict__ gx, float* __restrict__ gt) {
  const int n = blockIdx.y;
  const float gs = __ldg(gscale);
  for (int64_t v = (int64_t)blockIdx.x * LX_THREADS + threadIdx.x; v < a.V; v += (int64_t)gridDim.x * LX_THREADS) {
    const int64_t base = (int64_t)n * a.C * a.V + v;
    const float* xv = a.x + base;
    if (a.mode == 3) {
      const float* tv = reinterpret_cast<const float*>(a.target) + base;
      for (int c = 0; c < a.C; ++c) {
        const float xc = __ldg(xv + (int64_t)c * a.V), tc = __ldg(tv + (int64_t)c * a.V);
        if (gx) gx[base + (int64_t)c * a.V] = xc >= 1e-8f ? -gs * tc / xc : 0.f;  // clamp(min) passes the gradient where x >= min
        if (gt) gt[base + (int64_t)c * a.V] = -gs * logf(fmaxf(xc, 1e-8f));
      }
      continue;
    }
    float m, s;
    softmax_stats(xv, a.C, a.V, m, s);
    const float lse = m + logf(s);
    if (a.mode == 2) {
      const float* tv = reinterpret_cast<const float*>(a.target) + base;
      float tsum = 0.f;
      for (int c = 0; c < a.C; ++c) tsum += __ldg(tv + (int64_t)c * a.V);
      for (int c = 0; c < a.C; ++c) {
        const float lpc = __ldg(xv + (int64_t)c * a.V) - lse;
        if (gx) gx[base + (int64_t)c * a.V] = gs * (expf(lpc) * tsum - __ldg(tv + (int64_t)c * a.V));
        if (gt) gt[base + (int64_t)c * a.V] = -gs * lpc;
      }
      continue;
    }
    const int64_t t = load_label(a.target, a.kind, (int64_t)n * a.V + v);
    const bool skip = (a.mode == 0 && t == a.ignore_index) || t < 0 || t >= a.C;
    if (skip) {
      for (int c = 0; c < a.C; ++c) gx[base + (int64_t)c * a.V] = 0.f;
      continue;
    }
    const float xt = __ldg(xv + t * a.V);
    const float lp = xt - lse;
    const float w = (a.cw ? __ldg(a.cw + t) : 1.f) * gs;
    if (a.mode == 0) {
      for (int c = 0; c < a.C; ++c) {
        const float pc = expf(__ldg(xv + (int64_t)c * a.V) - lse);
        gx[base + (int64_t)c * a.V] = w * (pc - (c == t ? 1.f : 0.f));
      }
    } else {
      // L = -alpha q^gamma lp, q = 1 + P_t:  dL/dx_c = -alpha [gamma q^(gamma-1) lp dP_t/dx_c + q^gamma ([c==t] - p_c)]
      const float pt = expf(lp);
      const float P = a.focal_softmax ? pt : xt;
      const float q = 1.f + P;
      const float qg = powf(q, a.gamma);
      const float dq = a.gamma * powf(q, a.gamma - 1.f) * lp;  // multiplies dP_t/dx_c
      for (int c = 0; c < a.C; ++c) {
        const float pc = expf(__ldg(xv + (int64_t)c * a.V) - lse);
        const float onehot = c == t ? 1.f : 0.f;
        const float dP = a.focal_softmax ? pt * (onehot - pc) : onehot;
        gx[base + (int64_t)c * a.V] = -w * (dq * dP + qg * (onehot - pc));
      }
    }
  }
}

inline int blocks_for(int64_t V) {
  const int64_t b = da_cdiv(V, LX_THREADS);
  return (int)(b < LX_BLOCKS ? (b < 1 ? 1 : b) : LX_BLOCKS);
}

}  // namespace

DA_API int64_t da_pair_moments_workspace_bytes(int N) { return (int64_t)sizeof(double) * LX_BLOCKS * 6 * N; }

// a, b [N,V] (b nullable); out [N,9] (see pair_moments_kernel)
DA_API int da_pair_moments_fwd(const float* a, const float* b, int N, int64_t V, float* out, void* workspace,
                               int64_t workspace_bytes, cudaStream_t stream) {
  DA_REQUIRE(a && out && workspace, "da_pair_moments_fwd: null pointer");
  DA_REQUIRE(N >= 1 && N <= 65535 && V >= 1, "da_pair_moments_fwd: bad extents");
  if (workspace_bytes < da_pair_moments_workspace_bytes(N)) { da_set_error("da_pair_moments_fwd: workspace too small"); return DA_ERR_WORKSPACE; }
  const int nb = blocks_for(V);
  pair_moments_kernel<<<dim3(nb, N), LX_THREADS, 0, stream>>>(a, b, V, (double*)workspace);
  pair_moments_finalize_kernel<<<(N + 63) / 64, 64, 0, stream>>>((const double*)workspace, nb, N, V, out);
  return da_check_launch("da_pair_moments_fwd", 2);
}

// out [N,V] = coef[n,0]*a + coef[n,1]*b + coef[n,2]   (coef on the device, [N,3]; b nullable)
DA_API int da_affine2(const float* a, const float* b, const float* coef, int N, int64_t V, float* out, cudaStream_t stream) {
  DA_REQUIRE(a && coef && out, "da_affine2: null pointer");
  DA_REQUIRE(N >= 1 && N <= 65535 && V >= 1, "da_affine2: bad extents");
  affine2_kernel<<<dim3(blocks_for(V), N), LX_THREADS, 0, stream>>>(a, b, coef, V, out);
  return da_check_launch("da_affine2");
}

DA_API int64_t da_gradient_loss_workspace_bytes(int N, int C) { return (int64_t)sizeof(double) * LX_BLOCKS * 3 * N * C; }

// u [N,C,D,H,W]; sums [N,C,3]; norm_l1: 0 = squares ('L2'), 1 = absolute values
DA_API int da_gradient_loss_fwd(const float* u, int N, int C, int D, int H, int W, int norm_l1, float* sums, void* workspace,
                                int64_t workspace_bytes, cudaStream_t stream) {
  DA_REQUIRE(u && sums && workspace, "da_gradient_loss_fwd: null pointer");
  DA_REQUIRE(N >= 1 && C >= 1 && (int64_t)N * C <= 65535, "da_gradient_loss_fwd: bad extents");
  DA_REQUIRE(D >= 3 && H >= 3 && W >= 3, "da_gradient_loss_fwd: extent must be >= 3 per axis");
  if (workspace_bytes < da_gradient_loss_workspace_bytes(N, C)) { da_set_error("da_gradient_loss_fwd: workspace too small"); return DA_ERR_WORKSPACE; }
  GGeo g{D, H, W, (int64_t)W, (int64_t)H * W, (int64_t)D * H * W};
  const int nb = blocks_for(g.V);
  gradient_fwd_kernel<<<dim3(nb, N * C), LX_THREADS, 0, stream>>>(u, g, norm_l1, (double*)workspace);
  sums_finalize_kernel<<<(N * C * 3 + 63) / 64, 64, 0, stream>>>((const double*)workspace, nb, N * C, 3, sums);
  return da_check_launch("da_gradient_loss_fwd", 2);
}

// grad_sums [N,C,3] upstream; grad_u [N,C,D,H,W]
DA_API int da_gradient_loss_bwd(const float* u, const float* grad_sums, int N, int C, int D, int H, int W, int norm_l1,
                                float* grad_u, cudaStream_t stream) {
  DA_REQUIRE(u && grad_sums && grad_u, "da_gradient_loss_bwd: null pointer");
  DA_REQUIRE(N >= 1 && C >= 1 && (int64_t)N * C <= 65535, "da_gradient_loss_bwd: bad extents");
  GGeo g{D, H, W, (int64_t)W, (int64_t)H * W, (int64_t)D * H * W};
  gradient_bwd_kernel<<<dim3(blocks_for(g.V), N * C), LX_THREADS, 0, stream>>>(u, grad_sums, g, norm_l1, grad_u);
  return da_check_launch("da_gradient_loss_bwd");
}

DA_API int64_t da_xent_workspace_bytes(int N) { return (int64_t)sizeof(double) * LX_BLOCKS * 2 * N; }

static int xent_check(const char* who, const float* x, const void* target, int kind, int mode, int N, int C) {
  DA_REQUIRE(x && target, "%s: null pointer", who);
  DA_REQUIRE(mode >= 0 && mode <= 3, "%s: mode %d (0 cross entropy, 1 focal, 2/3 soft cross entropy)", who, mode);
  DA_REQUIRE(N >= 1 && N <= 65535 && C >= 1, "%s: bad extents", who);
  if (mode >= 2) DA_REQUIRE(kind == 2, "%s: soft cross entropy takes an fp32 target of the shape of x (kind 2)", who);
  else DA_REQUIRE(kind == 0 || kind == 1 || kind == 3, "%s: label kind %d (0 uint8, 1 int64, 3 int32)", who, kind);
  return DA_OK;
}

// x [N,C,V]; target: labels [N,V] (kind 0 uint8 / 1 int64 / 3 int32) for modes 0-1, fp32 [N,C,V] (kind 2) for modes 2-3;
// class_weight (nullable, [C]): CrossEntropyLoss weight / FocalLoss alpha; out2 [2] = (sum of terms, sum of weights)
DA_API int da_xent_fwd(const float* x, const void* target, int target_kind, int mode, int N, int C, int64_t V,
                       const float* class_weight, float gamma, int focal_softmax, int64_t ignore_index, float* out2,
                       void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  int rc = xent_check("da_xent_fwd", x, target, target_kind, mode, N, C);
  if (rc) return rc;
  DA_REQUIRE(out2 && workspace, "da_xent_fwd: null pointer");
  if (workspace_bytes < da_xent_workspace_bytes(N)) { da_set_error("da_xent_fwd: workspace too small"); return DA_ERR_WORKSPACE; }
  XentArgs a{x, target, target_kind, mode, C, V, class_weight, gamma, focal_softmax, ignore_index};
  const int nb = blocks_for(V);
  xent_fwd_kernel<<<dim3(nb, N), LX_THREADS, 0, stream>>>(a, (double*)workspace);
  xent_finalize_kernel<<<1, 32, 0, stream>>>((const double*)workspace, nb * N, out2);
  return da_check_launch("da_xent_fwd", 2);
}

// grad_scale: DEVICE scalar d loss / d out2[0]; grad_x [N,C,V] (nullable only in the soft modes); grad_target (nullable,
// soft modes only) [N,C,V]
DA_API int da_xent_bwd(const float* x, const void* target, int target_kind, int mode, int N, int C, int64_t V,
                       const float* class_weight, float gamma, int focal_softmax, int64_t ignore_index,
                       const float* grad_scale, float* grad_x, float* grad_target, cudaStream_t stream) {
  int rc = xent_check("da_xent_bwd", x, target, target_kind, mode, N, C);
  if (rc) return rc;
  DA_REQUIRE(grad_scale, "da_xent_bwd: null pointer");
  DA_REQUIRE(mode >= 2 ? (grad_x || grad_target) : (grad_x && !grad_target), "da_xent_bwd: gradient buffers do not fit the mode");
  XentArgs a{x, target, target_kind, mode, C, V, class_weight, gamma, focal_softmax, ignore_index};
  xent_bwd_kernel<<<dim3(blocks_for(V), N), LX_THREADS, 0, stream>>>(a, grad_scale, grad_x, grad_target);
  return da_check_launch("da_xent_bwd");
}

// ------------------------------------------------------------------------------------------------------------------
// Multi-scale LNCCLoss (lib/loss.py:512-586; in the file but not in the registry): per scale a k^3 box filter with
// dilation d and stride s (F.conv3d with a ones kernel, padding 0) of I, J, I^2, J^2, I*J, then
//   lncc = cross^2 / (Ivar * Jvar + 1e-5),  cross = IJs - Is*Js/n,  Ivar = I2s - Is^2/n,  Jvar = J2s - Js^2/n,  n = k^3
// and 1 - mean(lncc).  One thread = one window: the five sums are accumulated in fp64 (the reference's fp32 expressions
// cancel), the per-window derivatives with respect to the five sums are kept for the backward, which is a gather over
// the windows that contain a voxel (deterministic).  Windows are sparse in the volume (stride 2..10), so brute force over
// the taps is a few hundred M loads per scale, mostly L1 hits.
// ------------------------------------------------------------------------------------------------------------------
namespace {

struct MsGeo { int D, H, W, k, dil, stride, Do, Ho, Wo; };

// coef [N][5][Do*Ho*Wo]: d lncc / d (Is, I2s, IJs, Js, J2s); partials [gridDim.x] sums of lncc
__global__ void __launch_bounds__(LX_THREADS) lncc_ms_fwd_kernel(const float* __restrict__ I, const float* __restrict__ J, int N, MsGeo g,
                                                                 float* __restrict__ coef, double* __restrict__ partials) {
  __shared__ double red[LX_WARPS];
  const int64_t Vo = (int64_t)g.Do * g.Ho * g.Wo, V = (int64_t)g.D * g.H * g.W;
  const double nwin = (double)g.k * g.k * g.k;
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * LX_THREADS + threadIdx.x; i < (int64_t)N * Vo; i += (int64_t)gridDim.x * LX_THREADS) {
    const int64_t n = i / Vo, o = i - n * Vo;
    const int xo = (int)(o % g.Wo), yo = (int)((o / g.Wo) % g.Ho), zo = (int)(o / ((int64_t)g.Wo * g.Ho));
    const float* Ip = I + n * V + ((int64_t)zo * g.stride * g.H + (int64_t)yo * g.stride) * g.W + (int64_t)xo * g.stride;
    const float* Jp = J + n * V + (Ip - (I + n * V));
    double sI = 0, sJ = 0, sII = 0, sJJ = 0, sIJ = 0;
    for (int a = 0; a < g.k; ++a)
      for (int b = 0; b < g.k; ++b) {
        const int64_t row = ((int64_t)a * g.dil * g.H + (int64_t)b * g.dil) * g.W;
        float fI = 0.f, fJ = 0.f, fII = 0.f, fJJ = 0.f, fIJ = 0.f;
        for (int c = 0; c < g.k; ++c) {
          const float x = __ldg(Ip + row + c * g.dil), y = __ldg(Jp + row + c * g.dil);
          fI += x; fJ += y; fII = fmaf(x, x, fII); fJJ = fmaf(y, y, fJJ); fIJ = fmaf(x, y, fIJ);
        }
        sI += fI; sJ += fJ; sII += fII; sJJ += fJJ; sIJ += fIJ;
      }
    const double cross = sIJ - sI * sJ / nwin, iv = sII - sI * sI / nwin, jv = sJJ - sJ * sJ / nwin;
    const double den = iv * jv + 1e-5;
    acc += cross * cross / den;
    if (coef) {
      const double cIJ = 2.0 * cross / den, cII = -cross * cross * jv / (den * den), cJJ = -cross * cross * iv / (den * den);
      float* cp = coef + n * 5 * Vo + o;
      cp[0] = (float)(-cIJ * sJ / nwin - 2.0 * cII * sI / nwin);
      cp[Vo] = (float)cII;
      cp[2 * Vo] = (float)cIJ;
      cp[3 * Vo] = (float)(-cIJ * sI / nwin - 2.0 * cJJ * sJ / nwin);
      cp[4 * Vo] = (float)cJJ;
    }
  }
  const double t = block_sum<double, LX_WARPS>(acc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

__global__ void lncc_ms_finalize_kernel(const double* __restrict__ partials, int nb, float* __restrict__ out) {
  double a = 0;
  for (int i = threadIdx.x; i < nb; i += 32) a += partials[i];
  a = warp_sum(a);
  if (threadIdx.x == 0) out[0] = (float)a;
}

// grad[n][p] (+)= scale * sum over windows o containing p of (c1 + 2*self[p]*c2 + other[p]*c3), (c1, c2, c3) = coefficient
// fields (Is, I2s, IJs) for the gradient of I, (Js, J2s, IJs) for the gradient of J
__global__ void __launch_bounds__(LX_THREADS) lncc_ms_bwd_kernel(const float* __restrict__ self, const float* __restrict__ other,
                                                                 const float* __restrict__ coef, int which, const float* __restrict__ gscale,
                                                                 float scale, int N, MsGeo g, int accumulate, float* __restrict__ grad) {
  const int64_t Vo = (int64_t)g.Do * g.Ho * g.Wo, V = (int64_t)g.D * g.H * g.W;
  const float gs = __ldg(gscale) * scale;
  const int ext = g.dil * (g.k - 1);
  for (int64_t i = (int64_t)blockIdx.x * LX_THREADS + threadIdx.x; i < (int64_t)N * V; i += (int64_t)gridDim.x * LX_THREADS) {
    const int64_t n = i / V, p = i - n * V;
    const int x = (int)(p % g.W), y = (int)((p / g.W) % g.H), z = (int)(p / ((int64_t)g.W * g.H));
    const float* c1 = coef + (n * 5 + (which ? 3 : 0)) * Vo;
    const float* c2 = coef + (n * 5 + (which ? 4 : 1)) * Vo;
    const float* c3 = coef + (n * 5 + 2) * Vo;
    // windows along one axis: o*stride <= q <= o*stride + ext and (q - o*stride) % dil == 0
    auto lo = [&](int q) { const int t = q - ext; return t <= 0 ? 0 : (t + g.stride - 1) / g.stride; };
    float a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int zo = lo(z); zo < g.Do && zo * g.stride <= z; ++zo) {
      if ((z - zo * g.stride) % g.dil) continue;
      for (int yo = lo(y); yo < g.Ho && yo * g.stride <= y; ++yo) {
        if ((y - yo * g.stride) % g.dil) continue;
        for (int xo = lo(x); xo < g.Wo && xo * g.stride <= x; ++xo) {
          if ((x - xo * g.stride) % g.dil) continue;
          const int64_t o = ((int64_t)zo * g.Ho + yo) * g.Wo + xo;
          a1 += __ldg(c1 + o); a2 += __ldg(c2 + o); a3 += __ldg(c3 + o);
        }
      }
    }
    const float r = gs * (a1 + 2.f * __ldg(self + i) * a2 + __ldg(other + i) * a3);
    grad[i] = accumulate ? grad[i] + r : r;
  }
}

inline int ms_out(int n, int k, int dil, int stride) { return (n - dil * (k - 1) - 1) / stride + 1; }

}  // namespace

DA_API int64_t da_lncc_ms_workspace_bytes(void) { return (int64_t)sizeof(double) * LX_BLOCKS * 2; }
DA_API int64_t da_lncc_ms_coef_bytes(int N, int D, int H, int W, int k, int dil, int stride) {
  const int Do = ms_out(D, k, dil, stride), Ho = ms_out(H, k, dil, stride), Wo = ms_out(W, k, dil, stride);
  if (Do < 1 || Ho < 1 || Wo < 1) return 0;
  return (int64_t)sizeof(float) * N * 5 * Do * Ho * Wo;
}

// One scale of LNCCLoss: I, J [N,1,D,H,W]; out_sum [1] = sum over all windows of lncc (the caller divides by the window
// count); coef (nullable) receives the per-window derivatives for da_lncc_ms_bwd.
DA_API int da_lncc_ms_fwd(const float* I, const float* J, int N, int D, int H, int W, int k, int dil, int stride, float* out_sum,
                          float* coef, void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  DA_REQUIRE(I && J && out_sum && workspace, "da_lncc_ms_fwd: null pointer");
  DA_REQUIRE(k >= 1 && dil >= 1 && stride >= 1, "da_lncc_ms_fwd: bad window");
  MsGeo g{D, H, W, k, dil, stride, ms_out(D, k, dil, stride), ms_out(H, k, dil, stride), ms_out(W, k, dil, stride)};
  DA_REQUIRE(g.Do >= 1 && g.Ho >= 1 && g.Wo >= 1, "da_lncc_ms_fwd: window %d (dilation %d) does not fit the volume", k, dil);
  if (workspace_bytes < da_lncc_ms_workspace_bytes()) { da_set_error("da_lncc_ms_fwd: workspace too small"); return DA_ERR_WORKSPACE; }
  const int nb = blocks_for((int64_t)N * g.Do * g.Ho * g.Wo);
  lncc_ms_fwd_kernel<<<nb, LX_THREADS, 0, stream>>>(I, J, N, g, coef, (double*)workspace);
  lncc_ms_finalize_kernel<<<1, 32, 0, stream>>>((const double*)workspace, nb, out_sum);
  return da_check_launch("da_lncc_ms_fwd", 2);
}

// grad (+)= grad_scale[0] * scale * d(sum of lncc)/d(I or J); which: 0 = gradient of I, 1 = gradient of J; accumulate: add to grad
DA_API int da_lncc_ms_bwd(const float* I, const float* J, const float* coef, int which, const float* grad_scale, float scale, int N,
                          int D, int H, int W, int k, int dil, int stride, int accumulate, float* grad, cudaStream_t stream) {
  DA_REQUIRE(I && J && coef && grad_scale && grad, "da_lncc_ms_bwd: null pointer");
  MsGeo g{D, H, W, k, dil, stride, ms_out(D, k, dil, stride), ms_out(H, k, dil, stride), ms_out(W, k, dil, stride)};
  DA_REQUIRE(g.Do >= 1 && g.Ho >= 1 && g.Wo >= 1, "da_lncc_ms_bwd: window does not fit the volume");
  const int64_t total = (int64_t)N * D * H * W;
  const int64_t b = da_cdiv(total, LX_THREADS);
  const int grid = (int)(b > (int64_t)DA_NUM_SMS * 16 ? (int64_t)DA_NUM_SMS * 16 : b);
  lncc_ms_bwd_kernel<<<grid, LX_THREADS, 0, stream>>>(which ? J : I, which ? I : J, coef, which, grad_scale, scale, N, g, accumulate, grad);
  return da_check_launch("da_lncc_ms_bwd");
}
