// Batch-norm / activation / pooling / nearest-upsampling kernels of the encoder-decoders.
//
//   nn.BatchNorm3d, train mode with batch 1  (lib/network_factory/unets.py:31,51,116,130)
//   nn.LeakyReLU(0.01) / nn.ReLU             (unets.py:5-6,32 ; lib/network_factory/modules.py:58)
//   nn.MaxPool3d(2)                          (unets.py:84-86,230)  first-maximum tie rule kept
//   F.interpolate(size=) default 'nearest'   (lib/network_factory/voxel_morph.py:72-80)
// All are HBM-bound streaming kernels over planar NCDHW fp32; reductions are shuffle trees with a
// fixed-order fp64 second stage (deterministic).
#include "common.cuh"

namespace {

constexpr int BN_THREADS = 256;
#ifndef DA_BN_SPLITS
#define DA_BN_SPLITS 32
#endif
constexpr int BN_SPLITS = DA_BN_SPLITS;  // blocks per channel for the statistics passes
// -DDA_BN_REVERSE=1 makes the statistics passes walk their tensor from its END (the producing kernel's last ~100 MB
// should still be in the 126 MB L2, and blocks are dispatched in blockIdx order).  Measured on one box against the
// forward order (tools/ab_bench.py, three rounds): 44.87 vs 44.75 ms per step, i.e. nothing -- left off.
#ifndef DA_BN_REVERSE
#define DA_BN_REVERSE 0
#endif
constexpr bool BN_REVERSE = DA_BN_REVERSE != 0;

__device__ __forceinline__ float act_fwd(float z, int act, float slope) { return (act && z <= 0.f) ? z * slope : z; }
__device__ __forceinline__ float act_grad(float z, int act, float slope) { return (act && z <= 0.f) ? slope : 1.f; }

// block-wide min or max (valid in thread 0); red: BN_THREADS / 32 floats
__device__ __forceinline__ float block_minmax(float v, bool is_max, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float w = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, w) : fminf(v, w);
  }
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  if (lane == 0) red[wp] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < BN_THREADS / 32; ++i) v = is_max ? fmaxf(v, red[i]) : fminf(v, red[i]);
  }
  return v;
}

// partials [C][BN_SPLITS][4] (sum, sumsq, min, max) in fp64.  vec: V % 4 == 0 and 16-byte aligned base -> 16-byte loads, four
// independent ones in flight per thread (a scalar one-load-per-iteration loop left half of the HBM bandwidth unused)
__global__ void __launch_bounds__(BN_THREADS) bn_stats_kernel(const float* __restrict__ x, int N, int C, int64_t V, int vec,
                                                              double* __restrict__ partials, float* __restrict__ amax_slot) {
  __shared__ double red[BN_THREADS / 32];
  __shared__ float redm[2][BN_THREADS / 32];
  const int c = blockIdx.x, s = BN_REVERSE ? (int)gridDim.y - 1 - (int)blockIdx.y : (int)blockIdx.y;
  if (amax_slot && c == 0 && s == 0 && threadIdx.x == 0) *amax_slot = 0.f;   // the finalize kernel folds into it with atomicMax
  double a1 = 0, a2 = 0;
  float vmin = INFINITY, vmax = -INFINITY;   // value range of the channel: bounds the layer's output (bn_finalize_kernel)
  for (int n = 0; n < N; ++n) {
    const float* p = x + ((int64_t)n * C + c) * V;
    if (vec) {
      const int64_t V4 = V >> 2, chunk = da_cdiv(V4, BN_SPLITS);
      const int64_t lo = (int64_t)s * chunk, hi = min(V4, lo + chunk);
      const float4* p4 = reinterpret_cast<const float4*>(p);
      float f1 = 0.f, f2 = 0.f;
      int k = 0;
#pragma unroll 4
      for (int64_t i = lo + threadIdx.x; i < hi; i += BN_THREADS) {
        const float4 v = __ldg(p4 + i);
        f1 += (v.x + v.y) + (v.z + v.w);
        f2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, f2))));
        vmin = fminf(fminf(vmin, fminf(v.x, v.y)), fminf(v.z, v.w));
        vmax = fmaxf(fmaxf(vmax, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
        if (++k == 8) { a1 += (double)f1; a2 += (double)f2; f1 = f2 = 0.f; k = 0; }
      }
      a1 += (double)f1; a2 += (double)f2;
    } else {
      const int64_t chunk = da_cdiv(V, BN_SPLITS);
      const int64_t lo = (int64_t)s * chunk, hi = min(V, lo + chunk);
      float f1 = 0.f, f2 = 0.f;
      int k = 0;
      for (int64_t i = lo + threadIdx.x; i < hi; i += BN_THREADS) {
        const float v = p[i];
        f1 += v; f2 = fmaf(v, v, f2);
        vmin = fminf(vmin, v); vmax = fmaxf(vmax, v);
        if (++k == 32) { a1 += (double)f1; a2 += (double)f2; f1 = f2 = 0.f; k = 0; }
      }
      a1 += (double)f1; a2 += (double)f2;
    }
  }
  const double b1 = block_sum<double, BN_THREADS / 32>(a1, red);
  const double b2 = block_sum<double, BN_THREADS / 32>(a2, red);
  const float m0 = block_minmax(vmin, false, redm[0]), m1 = block_minmax(vmax, true, redm[1]);
  if (threadIdx.x == 0) {
    partials[((int64_t)c * BN_SPLITS + s) * 4 + 0] = b1;
    partials[((int64_t)c * BN_SPLITS + s) * 4 + 1] = b2;
    partials[((int64_t)c * BN_SPLITS + s) * 4 + 2] = (double)m0;
    partials[((int64_t)c * BN_SPLITS + s) * 4 + 3] = (double)m1;
  }
}

// warp = channel; amax_y (nullable, zeroed by the statistics kernel): upper bound of max|act(gamma * xhat + beta)| over the whole tensor, from the channels'
// value ranges -- the next convolution's tensor-core path scales its fp16 operand pairs by it and skips its own pass
__global__ void __launch_bounds__(128) bn_finalize_kernel(const double* __restrict__ partials, int C, double M, float eps, float momentum,
                                                          float* __restrict__ mean, float* __restrict__ invstd,
                                                          float* __restrict__ running_mean, float* __restrict__ running_var,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          int act, float slope, float* __restrict__ amax_y) {
  static_assert(BN_SPLITS <= 32, "one lane per split");
  __shared__ float redb[4];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float bound = 0.f;
  for (int c = blockIdx.x * nw + wp; c < C; c += gridDim.x * nw) {   // warp = channel, lane = split: fixed-order butterfly folds
    const bool live = lane < BN_SPLITS;
    const double* q = partials + ((int64_t)c * BN_SPLITS + (live ? lane : 0)) * 4;
    double s1 = live ? q[0] : 0.0, s2 = live ? q[1] : 0.0, lo = live ? q[2] : INFINITY, hi = live ? q[3] : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    const double mu = s1 / M;
    double var = s2 / M - mu * mu;
    if (var < 0) var = 0;
    const double is = 1.0 / sqrt(var + (double)eps);
    if (lane == 0) {
      mean[c] = (float)mu;
      invstd[c] = (float)is;
      if (running_mean) running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * mu);
      if (running_var) {
        const double unbiased = M > 1 ? var * M / (M - 1) : var;
        running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * unbiased);
      }
    }
    // act(gamma * xhat + beta) is monotone in xhat: its extreme magnitudes sit at the ends of the channel's value range
    const double ga = gamma ? (double)gamma[c] : 1.0, be = beta ? (double)beta[c] : 0.0;
    double y0 = ga * (lo - mu) * is + be, y1 = ga * (hi - mu) * is + be;
    if (act && y0 <= 0) y0 *= (double)slope;
    if (act && y1 <= 0) y1 *= (double)slope;
    bound = fmaxf(bound, (float)(fmax(fabs(y0), fabs(y1)) * 1.00001));   // (+ the apply pass's fp32 round-off)
  }
  if (amax_y) {   // (every lane of a warp holds the warp's bound)
    if (lane == 0) redb[wp] = bound;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int i = 1; i < nw; ++i) bound = fmaxf(bound, redb[i]);
      atomicMax(reinterpret_cast<unsigned int*>(amax_y), __float_as_uint(bound));   // non-negative floats order as unsigned ints
    }
  }
}

// y = act((x - mean) * invstd * gamma + beta); grid (blocks, C, N)
// Apply passes: DA_EW_THREADS threads per block.  With -DDA_EW_THREADS=128 (<= 32 registers) a block fits NEXT TO a
// resident tcgen05 convolution CTA (640 threads x 96 registers leave 4 K of the SM's 64 K registers), so that these
// HBM-bound passes of one branch of the step could run under the tensor kernels of another.  Measured same-box against
// 256 (three A/B rounds): 41.89 vs 42.02 ms per step -- within the noise; 256 stays (no spills in the backward pass).
#ifndef DA_EW_THREADS
#define DA_EW_THREADS 256
#endif
constexpr int EW_T = DA_EW_THREADS;
__global__ void __launch_bounds__(EW_T, 2048 / EW_T) bn_act_fwd_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                                         const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, int C, int64_t V, int act,
                                                         float slope, float* __restrict__ y) {
  const int c = blockIdx.y, n = blockIdx.z;
  const float mu = mean[c], is = invstd[c], ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
  const int64_t base = ((int64_t)n * C + c) * V;
  const float4* x4 = reinterpret_cast<const float4*>(x + base);
  float4* y4 = reinterpret_cast<float4*>(y + base);
  const bool vec = ((V & 3) == 0);
  if (vec) {
    for (int64_t i = (int64_t)blockIdx.x * EW_T + threadIdx.x; i < V / 4; i += (int64_t)gridDim.x * EW_T) {
      float4 v = x4[i];
      v.x = act_fwd((v.x - mu) * is * ga + be, act, slope);
      v.y = act_fwd((v.y - mu) * is * ga + be, act, slope);
      v.z = act_fwd((v.z - mu) * is * ga + be, act, slope);
      v.w = act_fwd((v.w - mu) * is * ga + be, act, slope);
      y4[i] = v;
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * EW_T + threadIdx.x; i < V; i += (int64_t)gridDim.x * EW_T)
      y[base + i] = act_fwd((x[base + i] - mu) * is * ga + be, act, slope);
  }
}

// backward statistics: s1 = sum g, s2 = sum g*xhat with g = dy * act'(z)
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_stats_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                  const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  int N, int C, int64_t V, int act, float slope, int vec,
                                                                  double* __restrict__ partials, float* __restrict__ amax_slot) {
  __shared__ double red[BN_THREADS / 32];
  const int c = blockIdx.x, s = BN_REVERSE ? (int)gridDim.y - 1 - (int)blockIdx.y : (int)blockIdx.y;
  if (amax_slot && c == 0 && s == 0 && threadIdx.x == 0) *amax_slot = 0.f;   // the finalize kernel folds into it with atomicMax
  const float mu = mean[c], is = invstd[c], ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
  double a1 = 0, a2 = 0;
  float mg = 0.f, mx = 0.f;   // max|dy|, max|xhat| of the channel: bound the layer's input gradient (bn_bwd_finalize_kernel)
  __shared__ float redm[2][BN_THREADS / 32];
  for (int n = 0; n < N; ++n) {
    const int64_t base = ((int64_t)n * C + c) * V;
    if (vec) {
      const int64_t V4 = V >> 2, chunk = da_cdiv(V4, BN_SPLITS);
      const int64_t lo = (int64_t)s * chunk, hi = min(V4, lo + chunk);
      const float4* x4 = reinterpret_cast<const float4*>(x + base);
      const float4* d4 = reinterpret_cast<const float4*>(dy + base);
      float f1 = 0.f, f2 = 0.f;
      int k = 0;
#pragma unroll 2
      for (int64_t i = lo + threadIdx.x; i < hi; i += BN_THREADS) {
        const float4 xv = __ldg(x4 + i), dv = __ldg(d4 + i);
        const float xh0 = (xv.x - mu) * is, xh1 = (xv.y - mu) * is, xh2 = (xv.z - mu) * is, xh3 = (xv.w - mu) * is;
        const float g0 = dv.x * act_grad(xh0 * ga + be, act, slope), g1 = dv.y * act_grad(xh1 * ga + be, act, slope);
        const float g2 = dv.z * act_grad(xh2 * ga + be, act, slope), g3 = dv.w * act_grad(xh3 * ga + be, act, slope);
        f1 += (g0 + g1) + (g2 + g3);
        f2 = fmaf(g0, xh0, fmaf(g1, xh1, fmaf(g2, xh2, fmaf(g3, xh3, f2))));
        mg = fmaxf(fmaxf(mg, fmaxf(fabsf(dv.x), fabsf(dv.y))), fmaxf(fabsf(dv.z), fabsf(dv.w)));
        mx = fmaxf(fmaxf(mx, fmaxf(fabsf(xh0), fabsf(xh1))), fmaxf(fabsf(xh2), fabsf(xh3)));
        if (++k == 8) { a1 += (double)f1; a2 += (double)f2; f1 = f2 = 0.f; k = 0; }
      }
      a1 += (double)f1; a2 += (double)f2;
    } else {
      const int64_t chunk = da_cdiv(V, BN_SPLITS);
      const int64_t lo = (int64_t)s * chunk, hi = min(V, lo + chunk);
      float f1 = 0.f, f2 = 0.f;
      int k = 0;
      for (int64_t i = lo + threadIdx.x; i < hi; i += BN_THREADS) {
        const float xh = (x[base + i] - mu) * is;
        const float g = dy[base + i] * act_grad(xh * ga + be, act, slope);
        f1 += g; f2 = fmaf(g, xh, f2);
        mg = fmaxf(mg, fabsf(dy[base + i])); mx = fmaxf(mx, fabsf(xh));
        if (++k == 32) { a1 += (double)f1; a2 += (double)f2; f1 = f2 = 0.f; k = 0; }
      }
      a1 += (double)f1; a2 += (double)f2;
    }
  }
  const double b1 = block_sum<double, BN_THREADS / 32>(a1, red);
  const double b2 = block_sum<double, BN_THREADS / 32>(a2, red);
  const float m0 = block_minmax(mg, true, redm[0]), m1 = block_minmax(mx, true, redm[1]);
  if (threadIdx.x == 0) {
    partials[((int64_t)c * BN_SPLITS + s) * 4 + 0] = b1;
    partials[((int64_t)c * BN_SPLITS + s) * 4 + 1] = b2;
    partials[((int64_t)c * BN_SPLITS + s) * 4 + 2] = (double)m0;
    partials[((int64_t)c * BN_SPLITS + s) * 4 + 3] = (double)m1;
  }
}

// warp = channel; amax_dx (nullable, zeroed by the statistics kernel): upper bound of max|dx| over the whole tensor,
//   |dx| <= |gamma invstd| (max|g| + |s1| / M + max|xhat| |s2| / M)   with |g| <= |dy| (slopes <= 1)
__global__ void __launch_bounds__(128) bn_bwd_finalize_kernel(const double* __restrict__ partials, int C, float* __restrict__ dgamma,
                                                              float* __restrict__ dbeta, float* __restrict__ s12,
                                                              const float* __restrict__ invstd, const float* __restrict__ gamma, double invM,
                                                              int training, float* __restrict__ amax_dx, int accumulate) {
  __shared__ float redb[4];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float bound = 0.f;
  for (int c = blockIdx.x * nw + wp; c < C; c += gridDim.x * nw) {   // warp = channel, lane = split
    const bool live = lane < BN_SPLITS;
    const double* q = partials + ((int64_t)c * BN_SPLITS + (live ? lane : 0)) * 4;
    double s1 = live ? q[0] : 0.0, s2 = live ? q[1] : 0.0, mg = live ? q[2] : 0.0, mx = live ? q[3] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      mg = fmax(mg, __shfl_xor_sync(0xffffffffu, mg, o));
      mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) {
      if (dbeta) dbeta[c] = (float)(accumulate ? s1 + (double)dbeta[c] : s1);
      if (dgamma) dgamma[c] = (float)(accumulate ? s2 + (double)dgamma[c] : s2);
      s12[2 * c] = (float)s1;
      s12[2 * c + 1] = (float)s2;
    }
    const double k = fabs((gamma ? (double)gamma[c] : 1.0) * (double)invstd[c]);
    const double b = training ? k * (mg + fabs(s1) * invM + mx * fabs(s2) * invM) : k * mg;
    bound = fmaxf(bound, (float)(b * 1.00001));
  }
  if (amax_dx) {
    if (lane == 0) redb[wp] = bound;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int i = 1; i < nw; ++i) bound = fmaxf(bound, redb[i]);
      atomicMax(reinterpret_cast<unsigned int*>(amax_dx), __float_as_uint(bound));
    }
  }
}

// dx = gamma*invstd*(g - s1/M - xhat*s2/M)   (training)  |  gamma*invstd*g   (eval)
__global__ void __launch_bounds__(EW_T, 2048 / EW_T) bn_act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                         const float* __restrict__ mean, const float* __restrict__ invstd,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         const float* __restrict__ s12, int C, int64_t V, float invM,
                                                         int training, int act, float slope, int vec, float* __restrict__ dx) {
  const int c = blockIdx.y, n = blockIdx.z;
  const float mu = mean[c], is = invstd[c], ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
  const float m1 = training ? s12[2 * c] * invM : 0.f, m2 = training ? s12[2 * c + 1] * invM : 0.f;
  const float k = ga * is;
  const int64_t base = ((int64_t)n * C + c) * V;
  if (vec) {
    const float4* x4 = reinterpret_cast<const float4*>(x + base);
    const float4* d4 = reinterpret_cast<const float4*>(dy + base);
    float4* o4 = reinterpret_cast<float4*>(dx + base);
#pragma unroll 2
    for (int64_t i = (int64_t)blockIdx.x * EW_T + threadIdx.x; i < (V >> 2); i += (int64_t)gridDim.x * EW_T) {
      const float4 xv = __ldg(x4 + i), dv = __ldg(d4 + i);
      float4 o;
      float xh, g;
      xh = (xv.x - mu) * is; g = dv.x * act_grad(xh * ga + be, act, slope); o.x = k * (g - m1 - xh * m2);
      xh = (xv.y - mu) * is; g = dv.y * act_grad(xh * ga + be, act, slope); o.y = k * (g - m1 - xh * m2);
      xh = (xv.z - mu) * is; g = dv.z * act_grad(xh * ga + be, act, slope); o.z = k * (g - m1 - xh * m2);
      xh = (xv.w - mu) * is; g = dv.w * act_grad(xh * ga + be, act, slope); o.w = k * (g - m1 - xh * m2);
      o4[i] = o;
    }
    return;
  }
  for (int64_t i = (int64_t)blockIdx.x * EW_T + threadIdx.x; i < V; i += (int64_t)gridDim.x * EW_T) {
    const float xh = (x[base + i] - mu) * is;
    const float g = dy[base + i] * act_grad(xh * ga + be, act, slope);
    dx[base + i] = k * (g - m1 - xh * m2);
  }
}

// dx = dy * (y > 0 ? 1 : slope)
__global__ void __launch_bounds__(256) act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float slope,
                                                      int64_t total, int vec, float* __restrict__ dx) {
  if (vec) {
    const float4* d4 = reinterpret_cast<const float4*>(dy);
    const float4* y4 = reinterpret_cast<const float4*>(y);
    float4* o4 = reinterpret_cast<float4*>(dx);
#pragma unroll 2
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < (total >> 2); i += (int64_t)gridDim.x * 256) {
      const float4 d = __ldg(d4 + i), v = __ldg(y4 + i);
      o4[i] = make_float4(v.x > 0.f ? d.x : d.x * slope, v.y > 0.f ? d.y : d.y * slope, v.z > 0.f ? d.z : d.z * slope,
                          v.w > 0.f ? d.w : d.w * slope);
    }
    return;
  }
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256)
    dx[i] = y[i] > 0.f ? dy[i] : dy[i] * slope;
}

// ---- max pool 2x2x2 stride 2 (floor) ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) maxpool2_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t NC,
                                                           int D, int H, int W, int Do, int Ho, int Wo) {
  const int64_t total = NC * Do * Ho * Wo;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int xo = (int)(i % Wo);
    int64_t t = i / Wo;
    const int yo = (int)(t % Ho); t /= Ho;
    const int zo = (int)(t % Do);
    const int64_t nc = t / Do;
    const float* p = x + ((nc * D + 2 * zo) * H + 2 * yo) * W + 2 * xo;
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float v = p[((int64_t)(k >> 2) * H + ((k >> 1) & 1)) * W + (k & 1)];
      if (v > m || v != v) m = v;  // ATen: (val > maxval) || isnan(val)
    }
    y[i] = m;
  }
}

// dx is fully written when D,H,W are even; odd trailing planes are zeroed by the entry point's memset.
__global__ void __launch_bounds__(256) maxpool2_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                           float* __restrict__ dx, int64_t NC, int D, int H, int W, int Do,
                                                           int Ho, int Wo) {
  const int64_t total = NC * Do * Ho * Wo;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int xo = (int)(i % Wo);
    int64_t t = i / Wo;
    const int yo = (int)(t % Ho); t /= Ho;
    const int zo = (int)(t % Do);
    const int64_t nc = t / Do;
    const int64_t off = ((nc * D + 2 * zo) * H + 2 * yo) * W + 2 * xo;
    const float* p = x + off;
    float m = -INFINITY;
    int arg = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float v = p[((int64_t)(k >> 2) * H + ((k >> 1) & 1)) * W + (k & 1)];
      if (v > m || v != v) { m = v; arg = k; }
    }
    const float g = dy[i];
#pragma unroll
    for (int k = 0; k < 8; ++k) dx[off + ((int64_t)(k >> 2) * H + ((k >> 1) & 1)) * W + (k & 1)] = (k == arg) ? g : 0.f;
  }
}

// ---- nearest upsampling to an explicit size ------------------------------------------------------------------
__device__ __forceinline__ int nearest_src(int dst, float scale, int in) {
  const int s = (int)floorf((float)dst * scale);
  return s < in - 1 ? s : in - 1;
}

__global__ void __launch_bounds__(256) upsample_nearest_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                   int64_t NC, int D, int H, int W, int Do, int Ho, int Wo) {
  const float sz = (float)D / (float)Do, sy = (float)H / (float)Ho, sx = (float)W / (float)Wo;
  const int64_t total = NC * Do * Ho * Wo;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int xo = (int)(i % Wo);
    int64_t t = i / Wo;
    const int yo = (int)(t % Ho); t /= Ho;
    const int zo = (int)(t % Do);
    const int64_t nc = t / Do;
    y[i] = x[((nc * D + nearest_src(zo, sz, D)) * H + nearest_src(yo, sy, H)) * W + nearest_src(xo, sx, W)];
  }
}

__device__ __forceinline__ void dst_range(int src, float scale, int in, int out, int& lo, int& hi) {
  // all dst with nearest_src(dst) == src form a contiguous range (the map is monotone)
  int d = (int)floorf((float)src / scale) - 2;
  if (d < 0) d = 0;
  while (d < out && nearest_src(d, scale, in) < src) ++d;
  lo = d;
  while (d < out && nearest_src(d, scale, in) == src) ++d;
  hi = d;
}

__global__ void __launch_bounds__(256) upsample_nearest_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx,
                                                                   int64_t NC, int D, int H, int W, int Do, int Ho, int Wo) {
  const float sz = (float)D / (float)Do, sy = (float)H / (float)Ho, sx = (float)W / (float)Wo;
  const int64_t total = NC * D * H * W;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int xs = (int)(i % W);
    int64_t t = i / W;
    const int ys = (int)(t % H); t /= H;
    const int zs = (int)(t % D);
    const int64_t nc = t / D;
    int z0, z1, y0, y1, x0, x1;
    dst_range(zs, sz, D, Do, z0, z1);
    dst_range(ys, sy, H, Ho, y0, y1);
    dst_range(xs, sx, W, Wo, x0, x1);
    float acc = 0.f;
    for (int z = z0; z < z1; ++z)
      for (int y = y0; y < y1; ++y)
        for (int xx = x0; xx < x1; ++xx) acc += dy[((nc * Do + z) * Ho + y) * Wo + xx];
    dx[i] = acc;
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

inline int ew_grid(int64_t total, int per_block = 256) {
  int64_t b = da_cdiv(total, per_block);
  const int64_t cap = (int64_t)DA_NUM_SMS * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

DA_API int64_t da_bn_workspace_bytes(int C) { return (int64_t)sizeof(double) * C * BN_SPLITS * 4 + (int64_t)sizeof(float) * 2 * C + 256; }

// Training-mode statistics: mean/invstd [C] out; running stats (nullable) updated in place with `momentum`
// (unbiased variance), as nn.BatchNorm3d does.
DA_API int da_bn_stats_ex(const float* x, int N, int C, int64_t V, float eps, float momentum, float* mean, float* invstd,
                          float* running_mean, float* running_var, const float* gamma, const float* beta, int act, float slope,
                          float* amax_y, void* workspace, int64_t workspace_bytes, cudaStream_t stream);
DA_API int da_bn_stats(const float* x, int N, int C, int64_t V, float eps, float momentum, float* mean, float* invstd,
                       float* running_mean, float* running_var, void* workspace, int64_t workspace_bytes,
                       cudaStream_t stream) {
  return da_bn_stats_ex(x, N, C, V, eps, momentum, mean, invstd, running_mean, running_var, nullptr, nullptr, 0, 0.f, nullptr,
                        workspace, workspace_bytes, stream);
}

// The same; with amax_y (one device float) it also leaves max|act(gamma * xhat + beta)| (up to round-off, from above),
// the magnitude of the layer's output that da_bn_act_fwd is about to write (gamma / beta / act / slope as passed there):
// the consumer's da_conv3d_fwd_ex takes it as a valid max-abs slot.
DA_API int da_bn_stats_ex(const float* x, int N, int C, int64_t V, float eps, float momentum, float* mean, float* invstd,
                          float* running_mean, float* running_var, const float* gamma, const float* beta, int act, float slope,
                          float* amax_y, void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  DA_REQUIRE(x && mean && invstd && workspace, "da_bn_stats: null pointer");
  if (workspace_bytes < da_bn_workspace_bytes(C)) { da_set_error("da_bn_stats: workspace too small"); return DA_ERR_WORKSPACE; }
  dim3 grid(C, BN_SPLITS);
  bn_stats_kernel<<<grid, BN_THREADS, 0, stream>>>(x, N, C, V, ((V & 3) == 0 && aligned16(x)) ? 1 : 0, (double*)workspace, amax_y);
  bn_finalize_kernel<<<(C + 3) / 4, 128, 0, stream>>>((const double*)workspace, C, (double)N * (double)V, eps, momentum, mean, invstd,
                                            running_mean, running_var, gamma, beta, act, slope, amax_y);
  return da_check_launch("da_bn_stats", 2);
}

// Eval-mode helper: mean = running_mean, invstd = rsqrt(running_var + eps) is formed by the caller.
DA_API int da_bn_act_fwd(const float* x, const float* mean, const float* invstd, const float* gamma, const float* beta,
                         int N, int C, int64_t V, int act, float slope, float* y, cudaStream_t stream) {
  DA_REQUIRE(x && mean && invstd && y, "da_bn_act_fwd: null pointer");
  const int nbx = ew_grid(V, 4 * EW_T) > 512 * (256 / EW_T) ? 512 * (256 / EW_T) : ew_grid(V, 4 * EW_T);
  dim3 grid(nbx, C, N);
  bn_act_fwd_kernel<<<grid, EW_T, 0, stream>>>(x, mean, invstd, gamma, beta, C, V, act, slope, y);
  return da_check_launch("da_bn_act_fwd");
}

DA_API int da_bn_act_bwd_ex(const float* dy, const float* x, const float* mean, const float* invstd, const float* gamma,
                            const float* beta, int N, int C, int64_t V, int training, int act, float slope, float* dx,
                            float* dgamma, float* dbeta, float* amax_dx, int accumulate, void* workspace, int64_t workspace_bytes,
                            cudaStream_t stream);
DA_API int da_bn_act_bwd(const float* dy, const float* x, const float* mean, const float* invstd, const float* gamma,
                         const float* beta, int N, int C, int64_t V, int training, int act, float slope, float* dx,
                         float* dgamma, float* dbeta, void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  return da_bn_act_bwd_ex(dy, x, mean, invstd, gamma, beta, N, C, V, training, act, slope, dx, dgamma, dbeta, nullptr, 0, workspace,
                          workspace_bytes, stream);
}

// The same; amax_dx (one device float, nullable) receives an upper bound of max|dx| (a valid max-abs slot for the
// da_conv3d_dgrad_ex / _wgrad_ex calls that consume dx); accumulate = 1: dgamma / dbeta += the result.
DA_API int da_bn_act_bwd_ex(const float* dy, const float* x, const float* mean, const float* invstd, const float* gamma,
                            const float* beta, int N, int C, int64_t V, int training, int act, float slope, float* dx,
                            float* dgamma, float* dbeta, float* amax_dx, int accumulate, void* workspace, int64_t workspace_bytes,
                            cudaStream_t stream) {
  DA_REQUIRE(dy && x && mean && invstd && dx && workspace, "da_bn_act_bwd: null pointer");
  if (workspace_bytes < da_bn_workspace_bytes(C)) { da_set_error("da_bn_act_bwd: workspace too small"); return DA_ERR_WORKSPACE; }
  double* partials = (double*)workspace;
  float* s12 = (float*)(partials + (int64_t)C * BN_SPLITS * 4);
  dim3 g1(C, BN_SPLITS);
  const int vec = ((V & 3) == 0 && aligned16(x) && aligned16(dy) && aligned16(dx)) ? 1 : 0;
  const double invM = 1.0 / ((double)N * (double)V);
  bn_bwd_stats_kernel<<<g1, BN_THREADS, 0, stream>>>(dy, x, mean, invstd, gamma, beta, N, C, V, act, slope, vec, partials, amax_dx);
  bn_bwd_finalize_kernel<<<(C + 3) / 4, 128, 0, stream>>>(partials, C, dgamma, dbeta, s12, invstd, gamma, invM, training, amax_dx, accumulate);
  const int nbx = ew_grid(V, 4 * EW_T) > 512 * (256 / EW_T) ? 512 * (256 / EW_T) : ew_grid(V, 4 * EW_T);
  dim3 g2(nbx, C, N);
  bn_act_bwd_kernel<<<g2, EW_T, 0, stream>>>(dy, x, mean, invstd, gamma, beta, s12, C, V, (float)invM, training, act, slope, vec, dx);
  return da_check_launch("da_bn_act_bwd", 3);
}

DA_API int da_act_bwd(const float* dy, const float* y, float slope, int64_t total, float* dx, cudaStream_t stream) {
  DA_REQUIRE(dy && y && dx, "da_act_bwd: null pointer");
  const int vec = ((total & 3) == 0 && aligned16(dy) && aligned16(y) && aligned16(dx)) ? 1 : 0;
  act_bwd_kernel<<<ew_grid(vec ? total / 4 : total), 256, 0, stream>>>(dy, y, slope, total, vec, dx);
  return da_check_launch("da_act_bwd");
}

DA_API int da_maxpool2_fwd(const float* x, float* y, int64_t NC, int D, int H, int W, cudaStream_t stream) {
  DA_REQUIRE(x && y, "da_maxpool2_fwd: null pointer");
  DA_REQUIRE(D >= 2 && H >= 2 && W >= 2, "da_maxpool2_fwd: extent too small");
  maxpool2_fwd_kernel<<<ew_grid(NC * (D / 2) * (H / 2) * (W / 2)), 256, 0, stream>>>(x, y, NC, D, H, W, D / 2, H / 2, W / 2);
  return da_check_launch("da_maxpool2_fwd");
}

DA_API int da_maxpool2_bwd(const float* dy, const float* x, float* dx, int64_t NC, int D, int H, int W, cudaStream_t stream) {
  DA_REQUIRE(dy && x && dx, "da_maxpool2_bwd: null pointer");
  if ((D | H | W) & 1) {
    cudaError_t e = cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)NC * D * H * W, stream);
    if (e != cudaSuccess) { da_set_error("da_maxpool2_bwd memset: %s", cudaGetErrorString(e)); return (int)e; }
  }
  maxpool2_bwd_kernel<<<ew_grid(NC * (D / 2) * (H / 2) * (W / 2)), 256, 0, stream>>>(dy, x, dx, NC, D, H, W, D / 2, H / 2, W / 2);
  return da_check_launch("da_maxpool2_bwd");
}

DA_API int da_upsample_nearest_fwd(const float* x, float* y, int64_t NC, int D, int H, int W, int Do, int Ho, int Wo,
                                   cudaStream_t stream) {
  DA_REQUIRE(x && y, "da_upsample_nearest_fwd: null pointer");
  upsample_nearest_fwd_kernel<<<ew_grid(NC * Do * Ho * Wo), 256, 0, stream>>>(x, y, NC, D, H, W, Do, Ho, Wo);
  return da_check_launch("da_upsample_nearest_fwd");
}

DA_API int da_upsample_nearest_bwd(const float* dy, float* dx, int64_t NC, int D, int H, int W, int Do, int Ho, int Wo,
                                   cudaStream_t stream) {
  DA_REQUIRE(dy && dx, "da_upsample_nearest_bwd: null pointer");
  upsample_nearest_bwd_kernel<<<ew_grid(NC * D * H * W), 256, 0, stream>>>(dy, dx, NC, D, H, W, Do, Ho, Wo);
  return da_check_launch("da_upsample_nearest_bwd");
}
