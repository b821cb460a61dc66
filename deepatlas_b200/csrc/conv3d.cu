// conv3d: the 3D convolutions of the segmentation / registration encoder-decoders.
//
// Covers nn.Conv3d k3 (stride 1/2, pad 1) and k1 as built by unets.convBlock
// (lib/network_factory/unets.py:24-39,98,250), modules.convBlock (lib/network_factory/modules.py:48)
// and the flow head (lib/network_factory/voxel_morph.py:57); nn.ConvTranspose3d k3 s1 p1
// (unets.py:89-96 via UNet.decoder :124-137) runs through the same kernels with a flipped /
// transposed weight repack.  The skip concatenations (unets.py:275,157-171; voxel_morph.py:65-82) are never
// materialised: every kernel takes two channel-planar sources.
//
// Layout: planar NCDHW fp32 (the reference's own), W contiguous -> every warp access is a
// coalesced row segment.  Weights are repacked per call to [cin][tap][cout_pad] so that one
// thread's output-channel vector is a broadcast LDS.128.
//
// Kernels in this file (exact fp32 FFMA; tensor-core variants live in conv3d_mma.cu):
//   repack_weights_kernel      (Cout,Cin,k^3) | (Cin,Cout,k^3)  ->  [a][tap][b_pad]
//   conv3d_direct_kernel<KS,CO>     any stride/pad, two sources, bias + optional leaky/ReLU epilogue
//   conv3d_tiled_kernel<CK,CO>      k3 s1 p1 hot path: smem halo tile, 4 voxels x CO channels per thread
//   conv3d_dgrad_s2_kernel<CO>      gather form of the stride-2 transposed conv
//   conv3d_wgrad_kernel<KS,CI,CO>   lane = voxel along W, register accumulators, butterfly reduce
//   reduce_partials_kernel          fixed-order second stage (deterministic)
#include <cuda_fp16.h>

#include "common.cuh"
#include "tma.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// weight repack:  dst[a][tap'][b_pad]  from src laid out (d0, d1, T).
//   a_is_dim0: a indexes d0 (else d1);  b indexes the other one, offset by b_off;  flip: tap' = T-1-tap
// ------------------------------------------------------------------------------------------------
__global__ void repack_weights_kernel(const float* __restrict__ src, float* __restrict__ dst, int d0, int d1, int T,
                                      int a_is_dim0, int flip, int A, int a_off, int B, int b_off, int Bpad) {
  const int64_t total = (int64_t)A * T * Bpad;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i % Bpad);
    const int t = (int)((i / Bpad) % T);
    const int a = (int)(i / ((int64_t)Bpad * T));
    float v = 0.f;
    if (b < B) {
      const int ts = flip ? (T - 1 - t) : t;
      const int i0 = a_is_dim0 ? (a + a_off) : (b + b_off);
      const int i1 = a_is_dim0 ? (b + b_off) : (a + a_off);
      v = src[((int64_t)i0 * d1 + i1) * T + ts];
    }
    dst[i] = v;
  }
}

struct ConvGeom {
  int N;
  int C1, C2;          // channels of the two planar sources (C2 may be 0)
  int Di, Hi, Wi;      // input extent
  int Do, Ho, Wo;      // output extent
  int Cout, Cop;       // real / padded output channels (packed weight pitch)
  int stride, pad;
  int act;             // 0 none, 1 leaky/relu with slope
  float slope;
};

// ------------------------------------------------------------------------------------------------
// generic direct convolution
// ------------------------------------------------------------------------------------------------
constexpr int DIRECT_THREADS = 128;
constexpr int DIRECT_CC = 8;  // input channels staged per weight chunk

template <int KS, int CO>
__global__ void __launch_bounds__(DIRECT_THREADS) conv3d_direct_kernel(const float* __restrict__ x1,
                                                                       const float* __restrict__ x2,
                                                                       const float* __restrict__ wp,
                                                                       const float* __restrict__ bias,
                                                                       float* __restrict__ out, ConvGeom g) {
  constexpr int T = KS * KS * KS;
  __shared__ __align__(16) float sw[DIRECT_CC * T * CO];
  const int n = blockIdx.z, cog = blockIdx.y;
  const int64_t Vo = (int64_t)g.Do * g.Ho * g.Wo, Vi = (int64_t)g.Di * g.Hi * g.Wi;
  const int64_t v = (int64_t)blockIdx.x * DIRECT_THREADS + threadIdx.x;
  const bool live = v < Vo;
  const int xo = (int)(v % g.Wo), yo = (int)((v / g.Wo) % g.Ho), zo = (int)(v / ((int64_t)g.Wo * g.Ho));
  const int zi0 = zo * g.stride - g.pad, yi0 = yo * g.stride - g.pad, xi0 = xo * g.stride - g.pad;
  uint32_t mask = 0;
  if (live) {
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int kz = t / (KS * KS), ky = (t / KS) % KS, kx = t % KS;
      const int zi = zi0 + kz, yi = yi0 + ky, xi = xi0 + kx;
      if (zi >= 0 && zi < g.Di && yi >= 0 && yi < g.Hi && xi >= 0 && xi < g.Wi) mask |= 1u << t;
    }
  }
  const int64_t base = ((int64_t)zi0 * g.Hi + yi0) * g.Wi + xi0;
  float acc[CO];
#pragma unroll
  for (int c = 0; c < CO; ++c) acc[c] = 0.f;
  const int Cin = g.C1 + g.C2;
  for (int c0 = 0; c0 < Cin; c0 += DIRECT_CC) {
    const int cc = min(DIRECT_CC, Cin - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < cc * T * CO; i += DIRECT_THREADS) {
      const int co = i % CO, rest = i / CO;  // rest = ci_local*T + tap
      sw[i] = wp[((int64_t)c0 * T + rest) * g.Cop + cog * CO + co];
    }
    __syncthreads();
    if (!live) continue;
    for (int cl = 0; cl < cc; ++cl) {
      const int ci = c0 + cl;
      const float* src = (ci < g.C1) ? x1 + ((int64_t)n * g.C1 + ci) * Vi : x2 + ((int64_t)n * g.C2 + (ci - g.C1)) * Vi;
      src += base;
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const int kz = t / (KS * KS), ky = (t / KS) % KS, kx = t % KS;
        const float xv = (mask >> t) & 1u ? __ldg(src + ((int64_t)kz * g.Hi + ky) * g.Wi + kx) : 0.f;
        const float4* w4 = reinterpret_cast<const float4*>(sw + (cl * T + t) * CO);
#pragma unroll
        for (int q = 0; q < CO / 4; ++q) {
          const float4 w = w4[q];
          acc[4 * q + 0] = fmaf(xv, w.x, acc[4 * q + 0]);
          acc[4 * q + 1] = fmaf(xv, w.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(xv, w.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(xv, w.w, acc[4 * q + 3]);
        }
      }
    }
  }
  if (!live) return;
#pragma unroll
  for (int c = 0; c < CO; ++c) {
    const int co = cog * CO + c;
    if (co >= g.Cout) break;
    float r = acc[c] + (bias ? bias[co] : 0.f);
    if (g.act) r = r > 0.f ? r : r * g.slope;
    out[((int64_t)n * g.Cout + co) * Vo + v] = r;
  }
}

// ------------------------------------------------------------------------------------------------
// tiled k3 s1 p1 convolution: output tile 4(z) x 8(y) x 32(x), 256 threads, each thread 4 voxels
// along x times CO output channels; input halo tile [CK][6][10][36] in shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int TZ = 4, TY = 8, TX = 32, VX = 4;
constexpr int HZ = TZ + 2, HY = TY + 2, HXP = 36;  // halo extents; x pitch padded to a multiple of 4 (>= TX+2)
constexpr int TILED_THREADS = TZ * TY * (TX / VX);  // 256

template <int CK, int CO>
__global__ void __launch_bounds__(TILED_THREADS, 2) conv3d_tiled_kernel(const float* __restrict__ x1,
                                                                        const float* __restrict__ x2,
                                                                        const float* __restrict__ wp,
                                                                        const float* __restrict__ bias,
                                                                        float* __restrict__ out, ConvGeom g,
                                                                        int tiles_x, int tiles_y) {
  extern __shared__ __align__(16) float smem[];
  float* sx = smem;                          // [CK][HZ][HY][HXP]
  float* sw = smem + CK * HZ * HY * HXP;     // [CK][27][CO]
  const int n = blockIdx.z, cog = blockIdx.y;
  int tb = blockIdx.x;
  const int bx = tb % tiles_x; tb /= tiles_x;
  const int by = tb % tiles_y;
  const int bz = tb / tiles_y;
  const int X0 = bx * TX, Y0 = by * TY, Z0 = bz * TZ;
  const int tx = threadIdx.x % (TX / VX), ty = (threadIdx.x / (TX / VX)) % TY, tz = threadIdx.x / ((TX / VX) * TY);
  const int64_t Vi = (int64_t)g.Di * g.Hi * g.Wi;

  float acc[VX][CO];
#pragma unroll
  for (int i = 0; i < VX; ++i)
#pragma unroll
    for (int c = 0; c < CO; ++c) acc[i][c] = 0.f;

  const int Cin = g.C1 + g.C2;
  for (int c0 = 0; c0 < Cin; c0 += CK) {
    const int cc = min(CK, Cin - c0);
    __syncthreads();
    // halo fill: rows of TX+2 (=34) floats; global x from X0-1
    for (int i = threadIdx.x; i < CK * HZ * HY * HXP; i += TILED_THREADS) {
      const int hx = i % HXP;
      int r = i / HXP;
      const int hy = r % HY; r /= HY;
      const int hz = r % HZ;
      const int cl = r / HZ;
      float v = 0.f;
      const int gx = X0 - 1 + hx, gy = Y0 - 1 + hy, gz = Z0 - 1 + hz;
      if (cl < cc && hx < TX + 2 && gx >= 0 && gx < g.Wi && gy >= 0 && gy < g.Hi && gz >= 0 && gz < g.Di) {
        const int ci = c0 + cl;
        const float* src = (ci < g.C1) ? x1 + ((int64_t)n * g.C1 + ci) * Vi : x2 + ((int64_t)n * g.C2 + (ci - g.C1)) * Vi;
        v = __ldg(src + ((int64_t)gz * g.Hi + gy) * g.Wi + gx);
      }
      sx[i] = v;
    }
    for (int i = threadIdx.x; i < CK * 27 * CO; i += TILED_THREADS) {
      const int co = i % CO, rest = i / CO;
      sw[i] = (rest < cc * 27) ? wp[((int64_t)c0 * 27 + rest) * g.Cop + cog * CO + co] : 0.f;
    }
    __syncthreads();
#pragma unroll 1
    for (int cl = 0; cl < CK; ++cl) {
#pragma unroll
      for (int kz = 0; kz < 3; ++kz) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const float* row = sx + ((cl * HZ + tz + kz) * HY + ty + ky) * HXP + tx * VX;
          const float4 a = *reinterpret_cast<const float4*>(row);
          const float2 b = *reinterpret_cast<const float2*>(row + 4);
          const float in[6] = {a.x, a.y, a.z, a.w, b.x, b.y};
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float4* w4 = reinterpret_cast<const float4*>(sw + (cl * 27 + (kz * 3 + ky) * 3 + kx) * CO);
#pragma unroll
            for (int q = 0; q < CO / 4; ++q) {
              const float4 w = w4[q];
#pragma unroll
              for (int i = 0; i < VX; ++i) {
                acc[i][4 * q + 0] = fmaf(in[i + kx], w.x, acc[i][4 * q + 0]);
                acc[i][4 * q + 1] = fmaf(in[i + kx], w.y, acc[i][4 * q + 1]);
                acc[i][4 * q + 2] = fmaf(in[i + kx], w.z, acc[i][4 * q + 2]);
                acc[i][4 * q + 3] = fmaf(in[i + kx], w.w, acc[i][4 * q + 3]);
              }
            }
          }
        }
      }
    }
  }
  const int z = Z0 + tz, y = Y0 + ty, x = X0 + tx * VX;
  if (z >= g.Do || y >= g.Ho || x >= g.Wo) return;
  const int64_t Vo = (int64_t)g.Do * g.Ho * g.Wo;
  const bool vec = (x + VX <= g.Wo) && ((g.Wo & 3) == 0);
#pragma unroll
  for (int c = 0; c < CO; ++c) {
    const int co = cog * CO + c;
    if (co >= g.Cout) break;
    const float bv = bias ? bias[co] : 0.f;
    float r[VX];
#pragma unroll
    for (int i = 0; i < VX; ++i) {
      r[i] = acc[i][c] + bv;
      if (g.act) r[i] = r[i] > 0.f ? r[i] : r[i] * g.slope;
    }
    float* o = out + ((int64_t)n * g.Cout + co) * Vo + ((int64_t)z * g.Ho + y) * g.Wo + x;
    if (vec) {
      *reinterpret_cast<float4*>(o) = make_float4(r[0], r[1], r[2], r[3]);
    } else {
#pragma unroll
      for (int i = 0; i < VX; ++i)
        if (x + i < g.Wo) o[i] = r[i];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// stride-2 (k3, pad 1) data gradient, gather form:
//   dX[ci][q] = sum_{co} sum_{k : (q+1-k) even, o=(q+1-k)/2 in range} dY[co][o] * W[co][ci][k]
// wp packed as [co][tap][ci_pad] (no flip).
// ------------------------------------------------------------------------------------------------
template <int CO>
__global__ void __launch_bounds__(DIRECT_THREADS) conv3d_dgrad_s2_kernel(const float* __restrict__ dy,
                                                                         const float* __restrict__ wp,
                                                                         float* __restrict__ dx, int N, int Cdy,
                                                                         int Cdx, int Cdxp, int Do, int Ho, int Wo,
                                                                         int Di, int Hi, int Wi) {
  __shared__ __align__(16) float sw[DIRECT_CC * 27 * CO];
  const int n = blockIdx.z, cig = blockIdx.y;
  const int64_t Vi = (int64_t)Di * Hi * Wi, Vo = (int64_t)Do * Ho * Wo;
  const int64_t v = (int64_t)blockIdx.x * DIRECT_THREADS + threadIdx.x;
  const bool live = v < Vi;
  const int xq = (int)(v % Wi), yq = (int)((v / Wi) % Hi), zq = (int)(v / ((int64_t)Wi * Hi));
  // per-axis: tap k valid iff (q+1-k) even and 0 <= (q+1-k)/2 < O
  uint32_t mask = 0;
  int oz[3], oy[3], ox[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int az = zq + 1 - k, ay = yq + 1 - k, ax = xq + 1 - k;
    oz[k] = (az >= 0 && !(az & 1) && (az >> 1) < Do) ? (az >> 1) : -1;
    oy[k] = (ay >= 0 && !(ay & 1) && (ay >> 1) < Ho) ? (ay >> 1) : -1;
    ox[k] = (ax >= 0 && !(ax & 1) && (ax >> 1) < Wo) ? (ax >> 1) : -1;
  }
  if (live) {
#pragma unroll
    for (int t = 0; t < 27; ++t)
      if (oz[t / 9] >= 0 && oy[(t / 3) % 3] >= 0 && ox[t % 3] >= 0) mask |= 1u << t;
  }
  float acc[CO];
#pragma unroll
  for (int c = 0; c < CO; ++c) acc[c] = 0.f;
  for (int c0 = 0; c0 < Cdy; c0 += DIRECT_CC) {
    const int cc = min(DIRECT_CC, Cdy - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < cc * 27 * CO; i += DIRECT_THREADS) {
      const int co = i % CO, rest = i / CO;
      sw[i] = wp[((int64_t)c0 * 27 + rest) * Cdxp + cig * CO + co];
    }
    __syncthreads();
    if (!live || mask == 0) continue;
    for (int cl = 0; cl < cc; ++cl) {
      const float* src = dy + ((int64_t)n * Cdy + c0 + cl) * Vo;
#pragma unroll
      for (int t = 0; t < 27; ++t) {
        if (!((mask >> t) & 1u)) continue;
        const float gv = __ldg(src + ((int64_t)oz[t / 9] * Ho + oy[(t / 3) % 3]) * Wo + ox[t % 3]);
        const float4* w4 = reinterpret_cast<const float4*>(sw + (cl * 27 + t) * CO);
#pragma unroll
        for (int q = 0; q < CO / 4; ++q) {
          const float4 w = w4[q];
          acc[4 * q + 0] = fmaf(gv, w.x, acc[4 * q + 0]);
          acc[4 * q + 1] = fmaf(gv, w.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(gv, w.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(gv, w.w, acc[4 * q + 3]);
        }
      }
    }
  }
  if (!live) return;
#pragma unroll
  for (int c = 0; c < CO; ++c) {
    const int ci = cig * CO + c;
    if (ci >= Cdx) break;
    dx[((int64_t)n * Cdx + ci) * Vi + v] = acc[c];
  }
}

// ------------------------------------------------------------------------------------------------
// weight gradient.  lane = output voxel along W; a warp owns one task = (kz,ky, 4 input channels,
// 8 output channels) with the three kx taps in registers (96 accumulators) and walks the rows of its
// region; the accumulators are folded across lanes with a halving butterfly (1 shuffle per value).
// partials: [region][Cout][Cin_total][T]  (PyTorch weight order, so the second stage writes grad_weight)
// ------------------------------------------------------------------------------------------------
constexpr int WG_THREADS = 256;
constexpr int WG_CI = 4, WG_CO = 8;

template <int NV>
__device__ __forceinline__ void butterfly_reduce(float (&v)[NV], int lane) {
  // NV = 32*G values per lane; on exit v[g] (g<G) of lane l holds the warp total of element g*32+l
  static_assert(NV % 32 == 0, "NV must be a multiple of 32");
  constexpr int G = NV / 32;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    // fold the 32 values v[g*32 .. g*32+31] in place down to v[g*32]
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
      const bool up = (lane & s) != 0;
#pragma unroll
      for (int i = 0; i < s; ++i) {
        const float keep = up ? v[g * 32 + i + s] : v[g * 32 + i];
        const float send = up ? v[g * 32 + i] : v[g * 32 + i + s];
        v[g * 32 + i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
      }
    }
  }
#pragma unroll
  for (int g = 1; g < G; ++g) v[g] = v[g * 32];
}

template <int KS>
__global__ void __launch_bounds__(WG_THREADS) conv3d_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                  float* __restrict__ partials, int N, int C,
                                                                  int ci_off, int Cin_total, int Cout, int co_off,
                                                                  int64_t region_stride, int Di, int Hi,
                                                                  int Wi, int Do, int Ho, int Wo, int stride, int pad,
                                                                  int64_t rows_per_region, int64_t total_rows) {
  constexpr int KX = KS;                 // taps along x held in registers
  constexpr int T = KS * KS * KS;
  constexpr int NACC = KX * WG_CI * WG_CO;  // 96 (k3) or 32 (k1)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nCiB = (C + WG_CI - 1) / WG_CI, nCoB = (Cout + WG_CO - 1) / WG_CO;
  const int ntasks = KS * KS * nCiB * nCoB;
  const int task = blockIdx.y * (WG_THREADS / 32) + warp;
  if (task >= ntasks) return;
  const int cob = task % nCoB;
  const int cib = (task / nCoB) % nCiB;
  const int kzy = task / (nCoB * nCiB);
  const int kz = kzy / KS, ky = kzy % KS;
  const int xb = (Wo + 31) / 32;
  const int64_t Vi = (int64_t)Di * Hi * Wi, Vo = (int64_t)Do * Ho * Wo;

  float acc[NACC < 32 ? 32 : NACC];
#pragma unroll
  for (int i = 0; i < (NACC < 32 ? 32 : NACC); ++i) acc[i] = 0.f;

  const int64_t r0 = (int64_t)blockIdx.x * rows_per_region;
  const int64_t r1 = min(total_rows, r0 + rows_per_region);
  for (int64_t rb = r0; rb < r1; ++rb) {
    const int bxi = (int)(rb % xb);
    int64_t t = rb / xb;
    const int yo = (int)(t % Ho); t /= Ho;
    const int zo = (int)(t % Do);
    const int n = (int)(t / Do);
    const int zi = zo * stride + kz - pad, yi = yo * stride + ky - pad;
    if (zi < 0 || zi >= Di || yi < 0 || yi >= Hi) continue;  // warp-uniform
    const int xo = bxi * 32 + lane;
    const bool live = xo < Wo;
    float dv[WG_CO];
#pragma unroll
    for (int c = 0; c < WG_CO; ++c) {
      const int co = cob * WG_CO + c;
      dv[c] = (live && co < Cout) ? __ldg(dy + ((int64_t)n * Cout + co) * Vo + ((int64_t)zo * Ho + yo) * Wo + xo) : 0.f;
    }
#pragma unroll
    for (int c = 0; c < WG_CI; ++c) {
      const int ci = cib * WG_CI + c;
      const float* xr = x + ((int64_t)n * C + (ci < C ? ci : 0)) * Vi + ((int64_t)zi * Hi + yi) * Wi;
#pragma unroll
      for (int kx = 0; kx < KX; ++kx) {
        const int xi = xo * stride + kx - pad;
        const float xv = (live && ci < C && xi >= 0 && xi < Wi) ? __ldg(xr + xi) : 0.f;
#pragma unroll
        for (int o = 0; o < WG_CO; ++o) acc[(kx * WG_CI + c) * WG_CO + o] = fmaf(xv, dv[o], acc[(kx * WG_CI + c) * WG_CO + o]);
      }
    }
  }
  butterfly_reduce<(NACC < 32 ? 32 : NACC)>(acc, lane);
  constexpr int G = (NACC < 32 ? 32 : NACC) / 32;
  float* pr = partials + (int64_t)blockIdx.x * region_stride;
#pragma unroll
  for (int gi = 0; gi < G; ++gi) {
    const int e = gi * 32 + lane;
    if (e >= NACC) continue;
    const int o = e % WG_CO, c = (e / WG_CO) % WG_CI, kx = e / (WG_CO * WG_CI);
    const int co = cob * WG_CO + o, ci = cib * WG_CI + c;
    if (co < Cout && ci < C) pr[((int64_t)(co_off + co) * Cin_total + ci_off + ci) * T + (kz * KS + ky) * KS + kx] = acc[gi];
  }
}


// ------------------------------------------------------------------------------------------------
// tiled k3 s1 p1 weight gradient (the hot one: 1/3 of the step's FLOPs).
//   dW[co][ci][kz][ky][kx] = sum_v dY[co][v] * X[ci][v + (kz,ky,kx) - 1]
// Block = 9 warps, warp w owns (kz,ky) = (w/3, w%3) for one group of 4 input x 8 output channels
// (96 register accumulators: 3 kx x 4 ci x 8 co).  The block walks a region of 4x8x32 output tiles;
// per tile the X halo [4][6][10][36] and the dY tile [8][4][8][32] are staged in shared memory with
// cp.async (double buffered, zero-filled outside the volume), each lane consumes 4 consecutive x
// (LDS.128) -> 384 FFMA per 12 LDS.  Lanes are folded with the halving butterfly once per block.
// The bias gradient (sum of dY) is accumulated by the centre warp of the ci-group-0 blocks.
// ------------------------------------------------------------------------------------------------
constexpr int WT_WARPS = 9, WT_THREADS = WT_WARPS * 32;
constexpr int WT_SX = WG_CI * HZ * HY * HXP;   // 8640 floats
constexpr int WT_SD = WG_CO * TZ * TY * TX;    // 8192 floats
constexpr int WT_STAGE = WT_SX + WT_SD;
constexpr size_t WT_SMEM = sizeof(float) * 2 * WT_STAGE;

__device__ __forceinline__ void cp_async4(float* dst, const float* src, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(src), "r"(sz));
}
__device__ __forceinline__ void cp_async16(float* dst, const float* src, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

struct WgTiledArgs {
  const float* x;        // [N][C][Di][Hi][Wi]   (the operand that is read with the 3x3x3 halo)
  const float* dy;       // [N][Cout][Di][Hi][Wi]
  float* partials;       // [region][Cout_total? see strides]
  float* bias_partials;  // [region][Cout] or null
  int N, C, ci_off, Cin_total, Cout, co_off;
  int64_t region_stride;
  int D, H, W;
  int tiles_x, tiles_y, tiles_z;
  int tiles_per_region, ntiles;
  int nCoB;
};

__global__ void __launch_bounds__(WT_THREADS, 1) conv3d_wgrad_tiled_kernel(WgTiledArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int kz = warp / 3, ky = warp % 3;
  const int cob = blockIdx.x % a.nCoB, cib = blockIdx.x / a.nCoB;
  const int region = blockIdx.y;
  const int64_t V = (int64_t)a.D * a.H * a.W;
  const bool vec_ok = (a.W & 3) == 0;

  float acc[96];
#pragma unroll
  for (int i = 0; i < 96; ++i) acc[i] = 0.f;
  float bacc[WG_CO];
#pragma unroll
  for (int o = 0; o < WG_CO; ++o) bacc[o] = 0.f;
  const bool do_bias = a.bias_partials != nullptr && cib == 0 && warp == 4;

  const int t0 = region * a.tiles_per_region;
  const int t1 = min(a.ntiles, t0 + a.tiles_per_region);

  auto stage_tile = [&](int t, int s) {
    float* sx = smem + s * WT_STAGE;
    float* sd = sx + WT_SX;
    int tb = t;
    const int bx = tb % a.tiles_x; tb /= a.tiles_x;
    const int by = tb % a.tiles_y; tb /= a.tiles_y;
    const int bz = tb % a.tiles_z;
    const int n = tb / a.tiles_z;
    const int X0 = bx * TX, Y0 = by * TY, Z0 = bz * TZ;
    // X halo: rows of 34 used floats (pitch 36)
    for (int i = threadIdx.x; i < WG_CI * HZ * HY * (TX + 2); i += WT_THREADS) {
      const int hx = i % (TX + 2);
      int r = i / (TX + 2);
      const int hy = r % HY; r /= HY;
      const int hz = r % HZ;
      const int cl = r / HZ;
      const int gx = X0 - 1 + hx, gy = Y0 - 1 + hy, gz = Z0 - 1 + hz;
      const int ci = cib * WG_CI + cl;
      const bool ok = ci < a.C && gx >= 0 && gx < a.W && gy >= 0 && gy < a.H && gz >= 0 && gz < a.D;
      const float* src = ok ? a.x + ((int64_t)n * a.C + ci) * V + ((int64_t)gz * a.H + gy) * a.W + gx : a.x;
      cp_async4(sx + ((cl * HZ + hz) * HY + hy) * HXP + hx, src, ok);
    }
    // dY tile
    if (vec_ok) {
      for (int i = threadIdx.x; i < WG_CO * TZ * TY * (TX / 4); i += WT_THREADS) {
        const int q = i % (TX / 4);
        int r = i / (TX / 4);
        const int ty = r % TY; r /= TY;
        const int tz = r % TZ;
        const int ol = r / TZ;
        const int gx = X0 + q * 4, gy = Y0 + ty, gz = Z0 + tz;
        const int co = cob * WG_CO + ol;
        const bool ok = co < a.Cout && gx < a.W && gy < a.H && gz < a.D;
        const float* src = ok ? a.dy + ((int64_t)n * a.Cout + co) * V + ((int64_t)gz * a.H + gy) * a.W + gx : a.dy;
        cp_async16(sd + ((ol * TZ + tz) * TY + ty) * TX + q * 4, src, ok);
      }
    } else {
      for (int i = threadIdx.x; i < WG_CO * TZ * TY * TX; i += WT_THREADS) {
        const int tx = i % TX;
        int r = i / TX;
        const int ty = r % TY; r /= TY;
        const int tz = r % TZ;
        const int ol = r / TZ;
        const int gx = X0 + tx, gy = Y0 + ty, gz = Z0 + tz;
        const int co = cob * WG_CO + ol;
        const bool ok = co < a.Cout && gx < a.W && gy < a.H && gz < a.D;
        const float* src = ok ? a.dy + ((int64_t)n * a.Cout + co) * V + ((int64_t)gz * a.H + gy) * a.W + gx : a.dy;
        cp_async4(sd + i, src, ok);
      }
    }
  };

  if (t0 < t1) {
    stage_tile(t0, 0);
    cp_async_commit();
  }
  for (int t = t0; t < t1; ++t) {
    const int s = (t - t0) & 1;
    if (t + 1 < t1) {
      stage_tile(t + 1, s ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* sx = smem + s * WT_STAGE;
    const float* sd = sx + WT_SX;
#pragma unroll 1
    for (int it = 0; it < (TZ * TY * TX / 4) / 32; ++it) {
      const int q = it * 32 + lane;
      const int tx4 = q & 7, ty = (q >> 3) & 7, tz = q >> 6;
      float4 d[WG_CO];
#pragma unroll
      for (int o = 0; o < WG_CO; ++o)
        d[o] = *reinterpret_cast<const float4*>(sd + ((o * TZ + tz) * TY + ty) * TX + tx4 * 4);
      if (do_bias) {
#pragma unroll
        for (int o = 0; o < WG_CO; ++o) bacc[o] += (d[o].x + d[o].y) + (d[o].z + d[o].w);
      }
#pragma unroll
      for (int c = 0; c < WG_CI; ++c) {
        const float* row = sx + ((c * HZ + tz + kz) * HY + ty + ky) * HXP + tx4 * 4;
        const float4 p = *reinterpret_cast<const float4*>(row);
        const float2 r2 = *reinterpret_cast<const float2*>(row + 4);
        const float in[6] = {p.x, p.y, p.z, p.w, r2.x, r2.y};
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
          for (int o = 0; o < WG_CO; ++o) {
            float v = acc[(kx * WG_CI + c) * WG_CO + o];
            v = fmaf(in[kx + 0], d[o].x, v);
            v = fmaf(in[kx + 1], d[o].y, v);
            v = fmaf(in[kx + 2], d[o].z, v);
            v = fmaf(in[kx + 3], d[o].w, v);
            acc[(kx * WG_CI + c) * WG_CO + o] = v;
          }
        }
      }
    }
    __syncthreads();
  }

  butterfly_reduce<96>(acc, lane);
  float* pr = a.partials + (int64_t)region * a.region_stride;
#pragma unroll
  for (int gi = 0; gi < 3; ++gi) {
    const int e = gi * 32 + lane;
    const int o = e % WG_CO, c = (e / WG_CO) % WG_CI, kx = e / (WG_CO * WG_CI);
    const int co = cob * WG_CO + o, ci = cib * WG_CI + c;
    if (co < a.Cout && ci < a.C)
      pr[((int64_t)(a.co_off + co) * a.Cin_total + a.ci_off + ci) * 27 + (kz * 3 + ky) * 3 + kx] = acc[gi];
  }
  if (do_bias) {
#pragma unroll
    for (int o = 0; o < WG_CO; ++o) {
      const float b = warp_sum(bacc[o]);
      const int co = cob * WG_CO + o;
      if (lane == 0 && co < a.Cout) a.bias_partials[(int64_t)region * a.Cout + co] = b;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 1x1 weight gradient (the class head: 16 -> n_classes at full resolution).  HBM/L2 streaming:
//   dW[co][ci] = sum_v dY[co][v] * X[ci][v];  block = one (4 ci x 8 co) task over one region of voxels,
// every thread walks float4 voxel-quads (12 LDG.128 per 128 FFMA), block-level shuffle reduce at the end.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv1x1_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                            float* __restrict__ partials, int N, int C, int ci_off,
                                                            int Cin_total, int Cout, int64_t V, int64_t quads_per_region,
                                                            int64_t region_stride) {
  __shared__ float red[8][32];
  const int nCoB = (Cout + WG_CO - 1) / WG_CO;
  const int cob = blockIdx.y % nCoB, cib = blockIdx.y / nCoB;
  const int64_t nq = V / 4;
  const int64_t q0 = (int64_t)blockIdx.x * quads_per_region, q1 = min(nq, q0 + quads_per_region);
  float acc[WG_CI][WG_CO];
#pragma unroll
  for (int c = 0; c < WG_CI; ++c)
#pragma unroll
    for (int o = 0; o < WG_CO; ++o) acc[c][o] = 0.f;
  for (int n = 0; n < N; ++n) {
    const float4* xp[WG_CI];
    const float4* dp[WG_CO];
#pragma unroll
    for (int c = 0; c < WG_CI; ++c) xp[c] = reinterpret_cast<const float4*>(x + ((int64_t)n * C + min(cib * WG_CI + c, C - 1)) * V);
#pragma unroll
    for (int o = 0; o < WG_CO; ++o) dp[o] = reinterpret_cast<const float4*>(dy + ((int64_t)n * Cout + min(cob * WG_CO + o, Cout - 1)) * V);
    for (int64_t q = q0 + threadIdx.x; q < q1; q += 256) {
      float4 xv[WG_CI], dv[WG_CO];
#pragma unroll
      for (int c = 0; c < WG_CI; ++c) xv[c] = __ldg(xp[c] + q);
#pragma unroll
      for (int o = 0; o < WG_CO; ++o) dv[o] = __ldg(dp[o] + q);
#pragma unroll
      for (int c = 0; c < WG_CI; ++c)
#pragma unroll
        for (int o = 0; o < WG_CO; ++o)
          acc[c][o] = fmaf(xv[c].x, dv[o].x, fmaf(xv[c].y, dv[o].y, fmaf(xv[c].z, dv[o].z, fmaf(xv[c].w, dv[o].w, acc[c][o]))));
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < WG_CI; ++c)
#pragma unroll
    for (int o = 0; o < WG_CO; ++o) {
      const float t = warp_sum(acc[c][o]);
      if (lane == 0) red[warp][c * WG_CO + o] = t;
    }
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    const int c = threadIdx.x / WG_CO, o = threadIdx.x % WG_CO;
    const int ci = cib * WG_CI + c, co = cob * WG_CO + o;
    if (ci < C && co < Cout) partials[(int64_t)blockIdx.x * region_stride + (int64_t)co * Cin_total + ci_off + ci] = t;
  }
}

#include "conv3d_tma.inc.cuh"
#include "conv3d_umma.inc.cuh"
#include "conv3d_wgrad_umma.inc.cuh"
#include "conv3d_smallcin.inc.cuh"

// out[i] = sum_r partials[r][i]  (fixed order)
// accumulate: out[i] += the sum (a parameter gradient that already holds an earlier contribution of the same step)
__global__ void reduce_partials_kernel(const float* __restrict__ partials, int nregions, int64_t count,
                                       float* __restrict__ out, int accumulate) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  double acc = accumulate ? (double)out[i] : 0.0;   // <= 148 partials per element: fp64 costs nothing here and keeps cancelling sums clean
  for (int r = 0; r < nregions; ++r) acc += (double)partials[(int64_t)r * count + i];
  out[i] = (float)acc;
}

// small counts (a thread per element leaves the GPU empty and walks up to 592 regions serially): block = 32 consecutive
// elements (lane = element: coalesced 128-byte rows) x 8 warps that deal the regions among themselves; fixed order
__global__ void __launch_bounds__(256) reduce_partials_warp_kernel(const float* __restrict__ partials, int nregions, int64_t count,
                                                                   float* __restrict__ out, int accumulate) {
  __shared__ double red[8][33];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 32 + lane;
  double acc = 0.0;
  if (i < count) {
#pragma unroll 4
    for (int r = wp; r < nregions; r += 8) acc += (double)partials[(int64_t)r * count + i];
  }
  red[wp][lane] = acc;
  __syncthreads();
  if (wp == 0 && i < count) {
    double t = red[0][lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) t += red[w][lane];
    out[i] = (float)(accumulate ? t + (double)out[i] : t);
  }
}

// set by da_conv3d_wgrad_ex for the duration of the call: the gradient outputs are added to instead of overwritten
thread_local int g_grad_accumulate = 0;
struct GradAccumulateScope {
  explicit GradAccumulateScope(int a) { g_grad_accumulate = a; }
  ~GradAccumulateScope() { g_grad_accumulate = 0; }
};

inline void launch_reduce_partials(const float* partials, int nregions, int64_t count, float* out, cudaStream_t stream) {
  if (nregions >= 16 && count <= 32768)
    reduce_partials_warp_kernel<<<(unsigned)da_cdiv(count, 32), 256, 0, stream>>>(partials, nregions, count, out, g_grad_accumulate);
  else
    reduce_partials_kernel<<<(unsigned)da_cdiv(count, 256), 256, 0, stream>>>(partials, nregions, count, out, g_grad_accumulate);
}

// per-channel sum over N and space (bias gradient):  out[c] = sum_{n,v} x[n][c][v]
// two deterministic stages: CS_SPLITS blocks per channel write fp64 partials, then one warp per channel folds them.
constexpr int CS_SPLITS = 64;
__global__ void __launch_bounds__(256) channel_sum_partial_kernel(const float* __restrict__ x, int N, int C, int64_t V,
                                                                  double* __restrict__ part) {
  __shared__ double red[8];
  const int c = blockIdx.x, sp = blockIdx.y;
  const int64_t chunk = ((V + CS_SPLITS - 1) / CS_SPLITS + 3) & ~(int64_t)3;
  const int64_t b = (int64_t)sp * chunk, e = min(V, b + chunk);
  double acc = 0;
  for (int n = 0; n < N; ++n) {
    const float* p = x + ((int64_t)n * C + c) * V;
    float a = 0.f;
    int k = 0;
    if ((((uintptr_t)p) & 15) == 0) {
      const int64_t e4 = b + ((e > b ? e - b : 0) & ~(int64_t)3);
      for (int64_t i = b + 4 * (int64_t)threadIdx.x; i < e4; i += 4 * 256) {
        const float4 v = *reinterpret_cast<const float4*>(p + i);
        a += (v.x + v.y) + (v.z + v.w);
        if (++k == 32) { acc += (double)a; a = 0.f; k = 0; }
      }
      for (int64_t i = e4 + threadIdx.x; i < e; i += 256) a += p[i];
    } else {
      for (int64_t i = b + threadIdx.x; i < e; i += 256) {
        a += p[i];
        if (++k == 64) { acc += (double)a; a = 0.f; k = 0; }
      }
    }
    acc += (double)a;
  }
  const double t = block_sum<double, 8>(acc, red);
  if (threadIdx.x == 0) part[(int64_t)c * CS_SPLITS + sp] = t;
}
__global__ void channel_sum_final_kernel(const double* __restrict__ part, int C, float* __restrict__ out, int accumulate) {
  const int c = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= C) return;
  double a = part[(int64_t)c * CS_SPLITS + lane] + part[(int64_t)c * CS_SPLITS + 32 + lane];
  a = warp_sum(a);
  if (lane == 0) out[c] = (float)(accumulate ? a + (double)out[c] : a);
}
inline int64_t channel_sum_scratch_bytes(int C) { return (int64_t)sizeof(double) * C * CS_SPLITS + 256; }
inline int run_channel_sum(const float* x, int N, int C, int64_t V, float* out, void* scratch, cudaStream_t stream, int accumulate = 0) {
  double* part = (double*)(((uintptr_t)scratch + 15) & ~(uintptr_t)15);
  channel_sum_partial_kernel<<<dim3(C, CS_SPLITS), 256, 0, stream>>>(x, N, C, V, part);
  channel_sum_final_kernel<<<(C + 7) / 8, 256, 0, stream>>>(part, C, out, accumulate);
  return da_check_launch("channel_sum", 2);
}

inline int conv_out(int in, int k, int s, int p) { return (in + 2 * p - k) / s + 1; }
inline int64_t pad_to(int64_t v, int m) { return (v + m - 1) / m * m; }
int g_force_direct = -1;
inline bool force_direct() {
  if (g_force_direct < 0) {
    const char* e = getenv("DA_CONV_IMPL");
    g_force_direct = (e && strcmp(e, "direct") == 0) ? 1 : 0;
  }
  return g_force_direct == 1;
}

template <int CK, int CO>
int launch_tiled(const float* x1, const float* x2, const float* wp, const float* bias, float* out, const ConvGeom& g,
                 cudaStream_t stream) {
  const int tiles_x = (g.Wo + TX - 1) / TX, tiles_y = (g.Ho + TY - 1) / TY, tiles_z = (g.Do + TZ - 1) / TZ;
  const size_t smem = sizeof(float) * (CK * HZ * HY * HXP + CK * 27 * CO);
  static DaPerDeviceOnce configured;
  if (configured.first()) {
    cudaFuncSetAttribute(conv3d_tiled_kernel<CK, CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  dim3 grid(tiles_x * tiles_y * tiles_z, g.Cop / CO, g.N);
  conv3d_tiled_kernel<CK, CO><<<grid, TILED_THREADS, smem, stream>>>(x1, x2, wp, bias, out, g, tiles_x, tiles_y);
  return da_check_launch("conv3d_tiled");
}

template <int KS>
int launch_direct(const float* x1, const float* x2, const float* wp, const float* bias, float* out, const ConvGeom& g,
                  cudaStream_t stream) {
  const int64_t Vo = (int64_t)g.Do * g.Ho * g.Wo;
  const int CO = (g.Cop % 16 == 0) ? 16 : (g.Cop % 8 == 0 ? 8 : 4);
  dim3 grid((unsigned)da_cdiv(Vo, DIRECT_THREADS), g.Cop / CO, g.N);
  if (CO == 16) conv3d_direct_kernel<KS, 16><<<grid, DIRECT_THREADS, 0, stream>>>(x1, x2, wp, bias, out, g);
  else if (CO == 8) conv3d_direct_kernel<KS, 8><<<grid, DIRECT_THREADS, 0, stream>>>(x1, x2, wp, bias, out, g);
  else conv3d_direct_kernel<KS, 4><<<grid, DIRECT_THREADS, 0, stream>>>(x1, x2, wp, bias, out, g);
  return da_check_launch("conv3d_direct");
}

// 1x1x1 convolution, streaming: thread = four consecutive voxels (one float4 per input channel) x CO output channels
// held in registers; the (small) weight block sits in shared memory and is read by broadcast.  HBM bound: the input is
// read once per CO block (L2 serves the repeats), the output written once.  w(o, i) = weight[o*so + i*si] covers the
// forward (so = Cin, si = 1) and the data gradient (so = 1, si = Cin_total, pointer advanced by ci_off).
constexpr int K1_MAXC = 128;
template <int CO>
__global__ void __launch_bounds__(256, 2) conv1x1_stream_kernel(const float* __restrict__ x1, const float* __restrict__ x2, int C1, int C2,
                                                             const float* __restrict__ weight, int so, int si,
                                                             const float* __restrict__ bias, float* __restrict__ out, int Cout,
                                                             int64_t V4, int act, float slope) {
  __shared__ __align__(16) float sw[K1_MAXC * CO];  // [ci][co]
  const int Cin = C1 + C2, co0 = blockIdx.y * CO, n = blockIdx.z;
  for (int i = threadIdx.x; i < Cin * CO; i += 256) {
    const int ci = i / CO, c = i - ci * CO;
    sw[i] = co0 + c < Cout ? weight[(int64_t)(co0 + c) * so + (int64_t)ci * si] : 0.f;
  }
  __syncthreads();
  const int64_t i4 = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i4 >= V4) return;
  float4 acc[CO];
#pragma unroll
  for (int c = 0; c < CO; ++c) {
    const float b = (bias && co0 + c < Cout) ? bias[co0 + c] : 0.f;
    acc[c] = make_float4(b, b, b, b);
  }
  const float4* p1 = reinterpret_cast<const float4*>(x1) + (int64_t)n * C1 * V4 + i4;
  const float4* p2 = x2 ? reinterpret_cast<const float4*>(x2) + (int64_t)n * C2 * V4 + i4 : nullptr;
#pragma unroll 8
  for (int ci = 0; ci < Cin; ++ci) {
    const float4 v = ci < C1 ? __ldcs(p1 + (int64_t)ci * V4) : __ldcs(p2 + (int64_t)(ci - C1) * V4);
    const float4* w4 = reinterpret_cast<const float4*>(sw + ci * CO);
#pragma unroll
    for (int q = 0; q < CO / 4; ++q) {
      const float4 w = w4[q];
      acc[4 * q + 0].x = fmaf(w.x, v.x, acc[4 * q + 0].x); acc[4 * q + 0].y = fmaf(w.x, v.y, acc[4 * q + 0].y);
      acc[4 * q + 0].z = fmaf(w.x, v.z, acc[4 * q + 0].z); acc[4 * q + 0].w = fmaf(w.x, v.w, acc[4 * q + 0].w);
      acc[4 * q + 1].x = fmaf(w.y, v.x, acc[4 * q + 1].x); acc[4 * q + 1].y = fmaf(w.y, v.y, acc[4 * q + 1].y);
      acc[4 * q + 1].z = fmaf(w.y, v.z, acc[4 * q + 1].z); acc[4 * q + 1].w = fmaf(w.y, v.w, acc[4 * q + 1].w);
      acc[4 * q + 2].x = fmaf(w.z, v.x, acc[4 * q + 2].x); acc[4 * q + 2].y = fmaf(w.z, v.y, acc[4 * q + 2].y);
      acc[4 * q + 2].z = fmaf(w.z, v.z, acc[4 * q + 2].z); acc[4 * q + 2].w = fmaf(w.z, v.w, acc[4 * q + 2].w);
      acc[4 * q + 3].x = fmaf(w.w, v.x, acc[4 * q + 3].x); acc[4 * q + 3].y = fmaf(w.w, v.y, acc[4 * q + 3].y);
      acc[4 * q + 3].z = fmaf(w.w, v.z, acc[4 * q + 3].z); acc[4 * q + 3].w = fmaf(w.w, v.w, acc[4 * q + 3].w);
    }
  }
  float4* po = reinterpret_cast<float4*>(out) + ((int64_t)n * Cout + co0) * V4 + i4;
#pragma unroll
  for (int c = 0; c < CO; ++c) {
    if (co0 + c >= Cout) break;
    float4 r = acc[c];
    if (act) {
      r.x = r.x > 0.f ? r.x : r.x * slope; r.y = r.y > 0.f ? r.y : r.y * slope;
      r.z = r.z > 0.f ? r.z : r.z * slope; r.w = r.w > 0.f ? r.w : r.w * slope;
    }
    __stcs(po + (int64_t)c * V4, r);
  }
}

// returns -1 if the streaming kernel does not apply (caller falls back to the direct kernel)
inline int run_conv1x1_stream(const float* x1, const float* x2, int C1, int C2, const float* weight, int so, int si, const float* bias,
                              float* out, int N, int Cout, int64_t V, int act, float slope, cudaStream_t stream) {
  if ((V & 3) != 0 || C1 + C2 > K1_MAXC || V < 4096) return -1;
  if ((((uintptr_t)x1 | (uintptr_t)x2 | (uintptr_t)out) & 15) != 0) return -1;
  const int64_t V4 = V / 4;
  if (Cout > 8) {
    dim3 grid((unsigned)da_cdiv(V4, 256), (Cout + 15) / 16, N);
    conv1x1_stream_kernel<16><<<grid, 256, 0, stream>>>(x1, x2, C1, C2, weight, so, si, bias, out, Cout, V4, act, slope);
  } else {
    dim3 grid((unsigned)da_cdiv(V4, 256), (Cout + 7) / 8, N);
    conv1x1_stream_kernel<8><<<grid, 256, 0, stream>>>(x1, x2, C1, C2, weight, so, si, bias, out, Cout, V4, act, slope);
  }
  return da_check_launch("conv1x1_stream");
}

inline bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }
int g_force_no_tma = -1;
inline bool tma_disabled() {
  if (g_force_no_tma < 0) {
    const char* e = getenv("DA_CONV_TMA");
    g_force_no_tma = (e && strcmp(e, "0") == 0) ? 1 : 0;
  }
  return g_force_no_tma == 1 || da_get_encode_tiled() == nullptr;
}

// k3 s1 p1 forward/dgrad through the TMA-staged kernel?  (geometry already in g; KS == 3)
inline bool fwd_tma_ok(const float* x1, const float* x2, const float* out, const ConvGeom& g) {
  const int Cin = g.C1 + g.C2;
  return !force_direct() && !tma_disabled() && g.stride == 1 && g.pad == 1 && Cin >= 3 && (g.Wi & 3) == 0 && g.Wo >= 16 &&
         (int64_t)g.Do * g.Ho * g.Wo >= 8192 && aligned16(x1) && aligned16(x2) && aligned16(out) &&
         (g.C2 == 0 || g.C1 % TMA_CK == 0);
}

int g_fwd_tiling = -1;  // 2 = thread owns 4 channels x 16/8 voxels (default), 1 = 4 voxels x CO channels (DA_FWD_TILING=1)
template <int CO>
int launch_fwd_tma(const float* x1, const float* x2, const float* wp, const float* bias, float* out, const ConvGeom& g,
                   int cin_pad, cudaStream_t stream) {
  if (g_fwd_tiling < 0) {
    const char* e = getenv("DA_FWD_TILING");
    g_fwd_tiling = (e && strcmp(e, "1") == 0) ? 1 : 2;
  }
  const bool v2 = (CO >= 8) && g_fwd_tiling == 2;
  const int pitch = v2 ? HXW : HXT;
  CUtensorMap m1, m2;
  int rc = da_make_volume_map(&m1, x1, g.N, g.C1, g.Di, g.Hi, g.Wi, pitch, HY, HZ, TMA_CK);
  if (rc) return rc;
  if (g.C2) {
    rc = da_make_volume_map(&m2, x2, g.N, g.C2, g.Di, g.Hi, g.Wi, pitch, HY, HZ, TMA_CK);
    if (rc) return rc;
  } else {
    m2 = m1;
  }
  static DaPerDeviceOnce configured;
  if (configured.first()) {
    cudaFuncSetAttribute(conv3d_fwd_tma_kernel<CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdTmaCfg<CO>::SMEM_BYTES);
    if constexpr (CO >= 8)
      cudaFuncSetAttribute(conv3d_fwd_tma2_kernel<CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fwd2Cfg<CO>::SMEM_BYTES);
  }
  const int tiles_x = (g.Wo + TX - 1) / TX, tiles_y = (g.Ho + TY - 1) / TY, tiles_z = (g.Do + TZ - 1) / TZ;
  dim3 grid(tiles_x * tiles_y * tiles_z, g.Cop / CO, g.N);
  if constexpr (CO >= 8) {
    if (v2) {
      conv3d_fwd_tma2_kernel<CO><<<grid, TILED_THREADS, Fwd2Cfg<CO>::SMEM_BYTES, stream>>>(m1, m2, wp, bias, out, g, tiles_x, tiles_y, cin_pad);
      return da_check_launch("conv3d_fwd_tma2");
    }
  }
  conv3d_fwd_tma_kernel<CO><<<grid, TILED_THREADS, FwdTmaCfg<CO>::SMEM_BYTES, stream>>>(m1, m2, wp, bias, out, g, tiles_x, tiles_y, cin_pad);
  return da_check_launch("conv3d_fwd_tma");
}

// repack (layout [cog][cin_pad][27][CO]) + launch; weight indexing arguments as for repack()
int run_conv_tma(const float* x1, const float* x2, const float* weight, float* wp, const float* bias, float* out, const ConvGeom& g,
                 int d0, int d1, int a_is_dim0, int flip, int b_off, cudaStream_t stream) {
  const int Cin = g.C1 + g.C2;
  const int cin_pad = (Cin + TMA_CK - 1) / TMA_CK * TMA_CK;
  int CO = (g.Cop % 16 == 0) ? 16 : (g.Cop % 8 == 0 ? 8 : 4);
  {  // small volumes (the 20x24x20 level): 8-channel blocks double the CTA count when 16-channel blocks leave SMs idle
    const int64_t tiles = (int64_t)((g.Wo + TX - 1) / TX) * ((g.Ho + TY - 1) / TY) * ((g.Do + TZ - 1) / TZ) * g.N;
    if (CO == 16 && tiles * (g.Cop / 16) < DA_NUM_SMS) CO = 8;
  }
  const int64_t total = (int64_t)cin_pad * 27 * g.Cop;
  int blocks = (int)da_cdiv(total, 256);
  if (blocks > 4096) blocks = 4096;
  repack_weights_tma_kernel<<<blocks, 256, 0, stream>>>(weight, wp, d0, d1, 27, a_is_dim0, flip, Cin, 0, cin_pad, g.Cout, b_off, g.Cop, CO);
  int rc = da_check_launch("repack_weights_tma");
  if (rc) return rc;
  if (CO == 16) return launch_fwd_tma<16>(x1, x2, wp, bias, out, g, cin_pad, stream);
  if (CO == 8) return launch_fwd_tma<8>(x1, x2, wp, bias, out, g, cin_pad, stream);
  return launch_fwd_tma<4>(x1, x2, wp, bias, out, g, cin_pad, stream);
}

// ---- tensor-core (tcgen05 3xTF32) path: the default for k3 s1 p1 forward / dgrad; DA_CONV_UMMA=0 or
// da_set_conv_impl(2) select the exact-FFMA kernels instead ---------------------------------------------------------
int g_use_umma = -1;
inline bool umma_enabled() {
  if (g_use_umma < 0) {
    const char* e = getenv("DA_CONV_UMMA");
    g_use_umma = (e && strcmp(e, "0") == 0) ? 0 : 1;
  }
  return g_use_umma == 1 && g_force_direct != 2;
}
constexpr int64_t UMMA_IMG_BYTES = UMMA_IMG_STRIDE_BYTES;  // 92160
// operand format of the tensor-core kernels: 0 = 3xFP16 (default: fp16 hi/lo pairs of per-tensor scaled operands, 22
// significant bits, twice the channels per MMA), 1 = 3xTF32.  da_set_conv_split() or DA_CONV_SPLIT=tf32|fp16 before
// the first call.
int g_conv_split = -1;
inline bool split_tf32() {
  if (g_conv_split < 0) {
    const char* e = getenv("DA_CONV_SPLIT");
    g_conv_split = (e && strcmp(e, "tf32") == 0) ? 1 : 0;
  }
  return g_conv_split == 1;
}
// da_set_conv_split(2): one MMA per product on the scaled fp16 operands (11 significant bits, fp32 accumulate) -- the
// reduced-precision mode BASELINE config C2 asks for (its "bf16"; fp16 on per-tensor scaled operands has three more
// mantissa bits at the same MMA rate).  Everything around the k3 convolutions stays fp32.
inline int single_pass() { return g_conv_split == 2 ? 1 : 0; }
inline bool fwd_umma_ok(const ConvGeom& g) {
  // structural limits of the kernel: 32-bit element offsets (output tensor, eight input channels), fused activation
  // written as max(r, r * slope)
  const int64_t V = (int64_t)g.Do * g.Ho * g.Wo;
  if ((int64_t)g.N * g.Cout * V >= ((int64_t)1 << 32) || 8 * V >= ((int64_t)1 << 31)) return false;
  if (g.act && !(g.slope >= 0.f && g.slope <= 1.f)) return false;
  if (g_force_direct == 3) return g.stride == 1 && g.pad == 1;  // tests: tensor-core path whatever the size heuristics say
  // (measured, round 2, 32 input channels per launch and all output-channel blocks of a layer in one launch: the tensor
  // path wins from ~8k voxels up whatever the channel counts -- 64 -> 64 @20x24x20: 0.060 ms against 0.111 ms for the FFMA
  // kernel; 768 -> 256 @32^3 (24 accumulating launches): 2.04 ms against 7.31 ms)
  const bool big = V >= 8192;
  return umma_enabled() && !force_direct() && g.stride == 1 && g.pad == 1 && g.C1 + g.C2 >= 8 && g.Wo >= 20 && big;
}
inline int64_t umma_workspace_bytes(int Cin, int Cout) {
  return (int64_t)((Cout + UM_CB - 1) / UM_CB) * ((Cin + 15) / 16) * UMMA_IMG_BYTES + 256;   // + the max-abs slots
}
// max|.| of up to four device arrays (null = absent) into out[slot] (out: 4 floats, zeroed here); one memset + one kernel
int run_absmax(const float* p0, int64_t n0, int s0, const float* p1, int64_t n1, int s1, const float* p2, int64_t n2, int s2,
               const float* p3, int64_t n3, int s3, float* out, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(out, 0, 4 * sizeof(float), stream);
  if (e != cudaSuccess) { da_set_error("absmax: memset failed: %s", cudaGetErrorString(e)); return (int)e; }
  AbsmaxArgs a;
  a.p[0] = p0; a.n[0] = p0 ? n0 : 0; a.slot[0] = s0; a.p[1] = p1; a.n[1] = p1 ? n1 : 0; a.slot[1] = s1;
  a.p[2] = p2; a.n[2] = p2 ? n2 : 0; a.slot[2] = s2; a.p[3] = p3; a.n[3] = p3 ? n3 : 0; a.slot[3] = s3;
  const int64_t total = a.n[0] + a.n[1] + a.n[2] + a.n[3];
  int64_t nb = da_cdiv(total, 256 * 16);
  if (nb > 8 * DA_NUM_SMS) nb = 8 * DA_NUM_SMS;
  if (nb < 1) nb = 1;
  absmax_kernel<<<(unsigned)nb, 256, 0, stream>>>(a, out);
  return da_check_launch("absmax");
}

unsigned long long* g_umma_dbg = nullptr;
inline unsigned long long* umma_dbg_buffer() {
  static int want = -1;
  if (want < 0) {
    const char* e = getenv("DA_UMMA_DEBUG");
    want = (e && strcmp(e, "1") == 0) ? 1 : 0;
    if (want && cudaMalloc(&g_umma_dbg, 16 * sizeof(unsigned long long)) == cudaSuccess) cudaMemset(g_umma_dbg, 0, 16 * sizeof(unsigned long long));
  }
  return g_umma_dbg;
}

template <int MODE>
int launch_umma_mode(const UmmaArgs& a, const CUtensorMap& m1, const CUtensorMap& m2, cudaStream_t stream) {
  static DaPerDeviceOnce configured;
  constexpr int SMEM = UmmaCfg<MODE>::SMEM_BYTES;
  constexpr bool HAS_DBG = MODE == 0 || MODE >= 3;   // cycle counters: the 3xTF32 kernel and the TMA-fed ones
  if (configured.first()) {
    cudaFuncSetAttribute(conv3d_umma_kernel<false, false, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    cudaFuncSetAttribute(conv3d_umma_kernel<false, true, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if constexpr (HAS_DBG) {
      cudaFuncSetAttribute(conv3d_umma_kernel<true, false, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
      cudaFuncSetAttribute(conv3d_umma_kernel<true, true, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    }
  }
  dim3 grid(a.tiles_x * a.tiles_y, (a.D + a.zg - 1) / a.zg, a.N * a.nco);
  if constexpr (HAS_DBG) {
    if (a.dbg) {
      if (a.accumulate) conv3d_umma_kernel<true, true, MODE><<<grid, UM_THREADS, SMEM, stream>>>(m1, m2, a);
      else conv3d_umma_kernel<true, false, MODE><<<grid, UM_THREADS, SMEM, stream>>>(m1, m2, a);
      return da_check_launch("conv3d_umma");
    }
  }
  if (a.accumulate) conv3d_umma_kernel<false, true, MODE><<<grid, UM_THREADS, SMEM, stream>>>(m1, m2, a);
  else conv3d_umma_kernel<false, false, MODE><<<grid, UM_THREADS, SMEM, stream>>>(m1, m2, a);
  return da_check_launch("conv3d_umma");
}

// weight source indexing arguments as for repack(); wp must hold umma_workspace_bytes(Cin, Cout)
// amax_x: optional device slot with an upper bound of max|x1, x2| (valid) or to be filled here (!valid); null: a
// workspace slot is used
int run_conv_umma(const float* x1, const float* x2, const float* weight, int64_t wcount, float* wp, const float* bias, float* out,
                  const ConvGeom& g, int d1, int a_is_dim0, int flip, int b_off, cudaStream_t stream, float* amax_x = nullptr,
                  int amax_x_valid = 0) {
  const int Cin = g.C1 + g.C2;
  constexpr int CB = UM_CB;
  const bool tf32 = split_tf32();
  // channel chunks of a launch: 16 as 3xTF32; 32 as 3xBF16, the last one 16 if that covers the remainder
  const int KC = tf32 ? 16 : 32;
  const int nco = (g.Cout + CB - 1) / CB, nk = (Cin + KC - 1) / KC;
  const int last_nch = (!tf32 && Cin - (nk - 1) * 32 <= 16) ? 2 : 4;
  float* amax = reinterpret_cast<float*>(reinterpret_cast<char*>(wp) + (int64_t)nco * nk * UMMA_IMG_BYTES);   // 4 floats at the end of the images
  int rc;
  if (tf32) {
    umma_prep_weights_kernel<<<dim3(45, nk, nco), 256, 0, stream>>>(weight, wp, d1, a_is_dim0, flip, Cin, KC, g.Cout, b_off, CB);
  } else {
    // 3xFP16: per-tensor power-of-two scales from max|input| (both sources) and max|weight| (the whole tensor: every
    // output-channel block and a data gradient's channel slice share one scale)
    const int64_t V = (int64_t)g.Di * g.Hi * g.Wi;
    const bool have_x = amax_x && amax_x_valid;
    rc = run_absmax(have_x ? nullptr : x1, (int64_t)g.N * g.C1 * V, 0, have_x ? nullptr : x2, (int64_t)g.N * g.C2 * V, 0, weight, wcount, 1,
                    nullptr, 0, 0, amax, stream);
    if (rc) return rc;
    if (amax_x && !amax_x_valid) {   // hand the freshly computed bound to the caller's slot
      cudaError_t e = cudaMemcpyAsync(amax_x, amax, sizeof(float), cudaMemcpyDeviceToDevice, stream);
      if (e != cudaSuccess) { da_set_error("conv3d: amax copy failed: %s", cudaGetErrorString(e)); return (int)e; }
    }
    umma_prep_weights16_kernel<<<dim3(45, nk, nco), 256, 0, stream>>>(weight, reinterpret_cast<uint16_t*>(wp), d1, a_is_dim0, flip, Cin, g.Cout, b_off, last_nch, amax + 1);
  }
  rc = da_check_launch("umma_prep_weights");
  if (rc) return rc;
  UmmaArgs a;
  a.amax_x = (amax_x && amax_x_valid) ? amax_x : amax; a.amax_w = amax + 1;
  a.single_pass = single_pass();
  a.dbg = umma_dbg_buffer();
  { static int fl = -1; if (fl < 0) { const char* e = getenv("DA_UMMA_FLAGS"); fl = e ? atoi(e) : 0; } a.flags = fl; }
  a.x1 = x1; a.x2 = x2; a.C1 = g.C1; a.C2 = g.C2; a.bias = bias; a.out = out;
  a.N = g.N; a.D = g.Do; a.H = g.Ho; a.W = g.Wo; a.Cout = g.Cout;
  a.act = g.act; a.slope = g.slope;
  a.tiles_x = (g.Wo + UM_TX - 1) / UM_TX; a.tiles_y = (g.Ho + UM_TY - 1) / UM_TY;
  int zg = g.Do;
  const int xy = a.tiles_x * a.tiles_y * g.N * nco;
  // planes per CTA: minimise waves x (planes per CTA + pipeline fill); one CTA per SM
  {
    int64_t best = -1;
    for (int groups = 1; groups <= g.Do; ++groups) {
      const int cand = (g.Do + groups - 1) / groups;
      if (cand < 4 && groups > 1) break;
      const int64_t ctas = (int64_t)xy * ((g.Do + cand - 1) / cand);
      const int64_t cost = ((ctas + DA_NUM_SMS - 1) / DA_NUM_SMS) * (cand + 3);
      if (best < 0 || cost < best) { best = cost; zg = cand; }
    }
  }
  a.zg = zg;
  // the channel chunks accumulate through the output tensor (serial launches); the output-channel blocks are
  // independent and share each launch (grid.z)
  a.nco = nco; a.img_stride = (int64_t)nk * (UMMA_IMG_BYTES / 4);
  // TMA-staged input planes (modes 3, 4): 16-byte aligned rows and bases, 8-channel boxes that never straddle the
  // concatenation; the register-staged modes 1, 2 serve everything else
  CUtensorMap m1, m2;
  memset(&m1, 0, sizeof(m1)); memset(&m2, 0, sizeof(m2));
  bool tma_in = !tf32 && !tma_disabled() && (g.Wi & 3) == 0 && aligned16(x1) && aligned16(x2) && (g.C2 == 0 || (g.C1 & 7) == 0);
  if (tma_in) {
    { static int off = -1; if (off < 0) { const char* e = getenv("DA_UMMA_TMA_IN"); off = (e && strcmp(e, "0") == 0) ? 1 : 0; } if (off) tma_in = false; }
  }
  if (tma_in) {
    rc = da_make_volume_map(&m1, x1, g.N, g.C1, g.Di, g.Hi, g.Wi, UM_RAWX, UM_TY + 2, 1, 8);
    if (!rc && g.C2) rc = da_make_volume_map(&m2, x2, g.N, g.C2, g.Di, g.Hi, g.Wi, UM_RAWX, UM_TY + 2, 1, 8);
    if (rc) tma_in = false;   // the encoder refused this extent: register-staged path
  }
  for (int ik = 0; ik < nk; ++ik) {
    a.wimg = wp + (int64_t)ik * (UMMA_IMG_BYTES / 4);
    a.c0 = ik * KC; a.accumulate = ik > 0; a.last = ik == nk - 1;
    const bool narrow = ik == nk - 1 && last_nch == 2;
    if (tf32) rc = launch_umma_mode<0>(a, m1, m2, stream);
    else if (tma_in) rc = narrow ? launch_umma_mode<4>(a, m1, m2, stream) : launch_umma_mode<3>(a, m1, m2, stream);
    else rc = narrow ? launch_umma_mode<2>(a, m1, m2, stream) : launch_umma_mode<1>(a, m1, m2, stream);
    if (rc) return rc;
  }
  return DA_OK;
}

int run_conv(const float* x1, const float* x2, const float* wp, const float* bias, float* out, const ConvGeom& g, int KS,
             cudaStream_t stream) {
  if (KS == 1) return launch_direct<1>(x1, x2, wp, bias, out, g, stream);
  const int Cin = g.C1 + g.C2;
  const bool tiled_ok = !force_direct() && g.stride == 1 && g.pad == 1 && g.Wo >= 16 && (int64_t)g.Do * g.Ho * g.Wo >= 32768;
  if (tiled_ok) {
    if (g.Cop % 16 == 0) {
      if (Cin >= 8) return launch_tiled<8, 16>(x1, x2, wp, bias, out, g, stream);
      if (Cin >= 3) return launch_tiled<4, 16>(x1, x2, wp, bias, out, g, stream);
      if (Cin == 2) return launch_tiled<2, 16>(x1, x2, wp, bias, out, g, stream);
      return launch_tiled<1, 16>(x1, x2, wp, bias, out, g, stream);
    }
    if (g.Cop % 8 == 0) {
      if (Cin >= 8) return launch_tiled<8, 8>(x1, x2, wp, bias, out, g, stream);
      if (Cin >= 3) return launch_tiled<4, 8>(x1, x2, wp, bias, out, g, stream);
      if (Cin == 2) return launch_tiled<2, 8>(x1, x2, wp, bias, out, g, stream);
      return launch_tiled<1, 8>(x1, x2, wp, bias, out, g, stream);
    }
    if (Cin >= 8) return launch_tiled<8, 4>(x1, x2, wp, bias, out, g, stream);
    return launch_tiled<4, 4>(x1, x2, wp, bias, out, g, stream);
  }
  return launch_direct<3>(x1, x2, wp, bias, out, g, stream);
}

inline int repack(const float* src, float* dst, int d0, int d1, int T, int a_is_dim0, int flip, int A, int a_off, int B,
                  int b_off, int Bpad, cudaStream_t stream) {
  const int64_t total = (int64_t)A * T * Bpad;
  int blocks = (int)da_cdiv(total, 256);
  if (blocks > 4096) blocks = 4096;
  repack_weights_kernel<<<blocks, 256, 0, stream>>>(src, dst, d0, d1, T, a_is_dim0, flip, A, a_off, B, b_off, Bpad);
  return da_check_launch("repack_weights");
}

inline int cpad(int c) { return (int)pad_to(c, c >= 16 ? 16 : (c > 4 ? 8 : 4)); }

constexpr int WG_MAX_REGIONS = 148;
inline int wg_region_cap(int64_t count) {
  int64_t r = ((int64_t)32 << 20) / (count > 0 ? count : 1);
  if (r > WG_MAX_REGIONS) r = WG_MAX_REGIONS;
  if (r < 4) r = 4;
  return (int)r;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------------

da_encode_tiled_fn da_get_encode_tiled() {
  static da_encode_tiled_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (da_encode_tiled_fn)p;
    else
      (void)cudaGetLastError();
    tried = true;
  }
  return fn;
}

// Debug aid (DA_UMMA_DEBUG=1): cycle counters accumulated by the MMA-issuing warps of conv3d_umma_kernel since the last
// call: out[0..8] = MMA warp: waiting for a free accumulator, for input planes, issuing, total, plane steps, CTAs;
// epilogue warp 0: waiting for the MMAs, TMEM read + clear, total.
DA_API int da_umma_debug_read(int64_t* out6) {
  DA_REQUIRE(out6, "da_umma_debug_read: null pointer");
  for (int i = 0; i < 11; ++i) out6[i] = 0;
  if (!g_umma_dbg) return DA_OK;
  unsigned long long h[16];
  cudaError_t e = cudaMemcpy(h, g_umma_dbg, sizeof(h), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemset(g_umma_dbg, 0, sizeof(h));
  if (e != cudaSuccess) { da_set_error("da_umma_debug_read: %s", cudaGetErrorString(e)); return (int)e; }
  for (int i = 0; i < 11; ++i) out6[i] = (int64_t)h[i];
  return DA_OK;
}

// Kernel selection for k3 s1 p1 convolutions: 0 = automatic (tcgen05 3xTF32 forward/dgrad + TMA-staged FFMA weight
// gradient where they apply), 1 = always the generic direct kernels, 2 = tiled exact-FFMA kernels only (the parity
// tests cross-check all of them), 3 = tcgen05 forward/dgrad whenever structurally possible (ignores the size heuristics).  Also settable through the
// environment variable DA_CONV_IMPL=direct before the first call.
DA_API int da_set_conv_split(int split) {
  DA_REQUIRE(split >= 0 && split <= 2, "da_set_conv_split: split must be 0 (3xFP16), 1 (3xTF32) or 2 (1xFP16, reduced precision)");
  g_conv_split = split;
  return DA_OK;
}
DA_API int da_set_conv_impl(int impl) {
  DA_REQUIRE(impl >= 0 && impl <= 3, "da_set_conv_impl: impl must be 0 (auto), 1 (direct), 2 (tiled FFMA, no tensor cores) or 3 (tensor cores forced)");
  g_force_direct = impl;
  return DA_OK;
}

// workspace for da_conv3d_fwd / da_conv3d_dgrad: one packed weight copy
DA_API int64_t da_conv3d_pack_bytes(int Cin, int Cout, int ks) {
  const int m = Cin > Cout ? Cin : Cout;
  const int64_t base = (int64_t)sizeof(float) * (int64_t)(m + 3) * ks * ks * ks * cpad(m) + 256;
  const int64_t um = ks == 3 ? umma_workspace_bytes(m, m) : 0;
  return base > um ? base : um;
}
// workspace of da_conv3d_dgrad: the weight image plus, for stride 2, the zero-inserted dy ([N,Cout,Di,Hi,Wi])
DA_API int64_t da_conv3d_dgrad_workspace_bytes(int N, int Cin, int Cout, int Di, int Hi, int Wi, int ks, int stride) {
  const int64_t pack = (da_conv3d_pack_bytes(Cin, Cout, ks) + 255) & ~(int64_t)255;
  return stride == 2 ? pack + (int64_t)sizeof(float) * N * Cout * Di * Hi * Wi : pack;
}
DA_API int64_t da_conv3d_wgrad_workspace_bytes(int Cin, int Cout, int ks) {
  const int64_t count = (int64_t)Cin * Cout * ks * ks * ks;
  const int m = Cin > Cout ? Cin : Cout;
  const int64_t gen = (int64_t)sizeof(float) * ((int64_t)wg_region_cap(count) * count + (int64_t)WG_MAX_REGIONS * m) + 512;
  const int64_t sc = (ks == 3 && Cin <= SC_MAX_CIN) ? (int64_t)sizeof(float) * SC_REGIONS * (count + Cout) + 512 : 0;   // input layers
  return gen > sc ? gen : sc;
}

// Forward.  x1 [N,C1,Di,Hi,Wi], x2 [N,C2,...] or null (C2=0): the conv sees cat(x1,x2) along channels.
// weight: transposed==0 -> nn.Conv3d layout (Cout, C1+C2, k,k,k); transposed==1 -> nn.ConvTranspose3d layout
// (C1+C2, Cout, k,k,k), only k3 s1 p1 (equivalent to a conv with the flipped kernel).
// bias nullable.  act: 0 none, 1 leaky-relu(slope) fused (slope 0 = ReLU).  out [N,Cout,Do,Ho,Wo].
DA_API int da_conv3d_fwd_ex(const float* x1, int C1, const float* x2, int C2, const float* weight, int transposed, const float* bias,
                            float* out, int N, int Di, int Hi, int Wi, int Cout, int ks, int stride, int pad, int act, float slope,
                            void* workspace, int64_t workspace_bytes, cudaStream_t stream, float* amax_x, int amax_x_valid);
DA_API int da_conv3d_dgrad_ex(const float* dy, const float* weight, int transposed, float* dx, int N, int Cin_total, int ci_off, int Cdx,
                              int Cout, int Di, int Hi, int Wi, int ks, int stride, int pad, void* workspace, int64_t workspace_bytes,
                              cudaStream_t stream, float* amax_dy, int amax_dy_valid);
DA_API int da_conv3d_wgrad_ex(const float* x1, int C1, const float* x2, int C2, const float* dy, int transposed, float* grad_weight,
                              float* grad_bias, int N, int Di, int Hi, int Wi, int Cout, int ks, int stride, int pad, void* workspace,
                              int64_t workspace_bytes, cudaStream_t stream, float* amax_x, int amax_x_valid, float* amax_dy,
                              int amax_dy_valid, int accumulate);

namespace {
// fills the caller's max-abs slot when the chosen kernel did not need it itself (the _ex contract: a slot passed with
// valid = 0 always holds the bound afterwards)
inline int fill_amax_slot(float* slot, const float* p1, int64_t n1, const float* p2, int64_t n2, cudaStream_t stream) {
  float* tmp = slot;   // run_absmax zeroes four floats: go through the slot only if it is the head of such a block
  cudaError_t e = cudaMemsetAsync(tmp, 0, sizeof(float), stream);
  if (e != cudaSuccess) { da_set_error("absmax: memset failed: %s", cudaGetErrorString(e)); return (int)e; }
  AbsmaxArgs a;
  a.p[0] = p1; a.n[0] = p1 ? n1 : 0; a.slot[0] = 0; a.p[1] = p2; a.n[1] = p2 ? n2 : 0; a.slot[1] = 0;
  a.p[2] = nullptr; a.n[2] = 0; a.slot[2] = 0; a.p[3] = nullptr; a.n[3] = 0; a.slot[3] = 0;
  int64_t nb = da_cdiv(a.n[0] + a.n[1], 256 * 16);
  if (nb > 8 * DA_NUM_SMS) nb = 8 * DA_NUM_SMS;
  if (nb < 1) nb = 1;
  absmax_kernel<<<(unsigned)nb, 256, 0, stream>>>(a, tmp);
  return da_check_launch("absmax");
}
}  // namespace

DA_API int da_absmax(const float* x1, int64_t n1, const float* x2, int64_t n2, float* out, cudaStream_t stream) {
  DA_REQUIRE(x1 && out, "da_absmax: null pointer");
  return fill_amax_slot(out, x1, n1, x2, n2, stream);
}

DA_API int da_conv3d_fwd(const float* x1, int C1, const float* x2, int C2, const float* weight, int transposed,
                         const float* bias, float* out, int N, int Di, int Hi, int Wi, int Cout, int ks, int stride,
                         int pad, int act, float slope, void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  return da_conv3d_fwd_ex(x1, C1, x2, C2, weight, transposed, bias, out, N, Di, Hi, Wi, Cout, ks, stride, pad, act, slope, workspace,
                          workspace_bytes, stream, nullptr, 0);
}

// The same with a caller-owned max-abs slot for the input (one device float: an upper bound of max|x1, x2|).
// amax_x_valid = 1: the slot already holds the bound (a producer kernel or an earlier call computed it) and the input
// pass of the tensor-core path is skipped; 0: the slot is filled by this call, whatever kernel runs.
DA_API int da_conv3d_fwd_ex(const float* x1, int C1, const float* x2, int C2, const float* weight, int transposed,
                            const float* bias, float* out, int N, int Di, int Hi, int Wi, int Cout, int ks, int stride,
                            int pad, int act, float slope, void* workspace, int64_t workspace_bytes, cudaStream_t stream,
                            float* amax_x, int amax_x_valid) {
  DA_REQUIRE(x1 && weight && out && workspace, "da_conv3d_fwd: null pointer");
  DA_REQUIRE(ks == 1 || ks == 3, "da_conv3d_fwd: unsupported kernel size %d (1 or 3)", ks);
  DA_REQUIRE(stride == 1 || stride == 2, "da_conv3d_fwd: unsupported stride %d", stride);
  DA_REQUIRE(!transposed || (ks == 3 && stride == 1 && pad == 1), "da_conv3d_fwd: transposed only for k3 s1 p1");
  DA_REQUIRE((C2 == 0) == (x2 == nullptr), "da_conv3d_fwd: x2/C2 mismatch");
  const int Cin = C1 + C2, T = ks * ks * ks;
  if (workspace_bytes < da_conv3d_pack_bytes(Cin, Cout, ks)) { da_set_error("da_conv3d_fwd: workspace too small"); return DA_ERR_WORKSPACE; }
  ConvGeom g{N, C1, C2, Di, Hi, Wi, conv_out(Di, ks, stride, pad), conv_out(Hi, ks, stride, pad), conv_out(Wi, ks, stride, pad),
             Cout, cpad(Cout), stride, pad, act, slope};
  float* wp = (float*)workspace;
  if (ks == 1 && stride == 1 && pad == 0 && !transposed && !force_direct()) {
    const int rc1 = run_conv1x1_stream(x1, x2, C1, C2, weight, Cin, 1, bias, out, N, Cout, (int64_t)Di * Hi * Wi, act, slope, stream);
    if (rc1 >= 0) return rc1;
  }
  if (ks == 3 && aligned16(wp) && fwd_umma_ok(g) && !(split_tf32() && amax_x && !amax_x_valid))
    return transposed ? run_conv_umma(x1, x2, weight, (int64_t)Cin * Cout * 27, wp, bias, out, g, Cout, 1, 1, 0, stream, amax_x, amax_x_valid)
                      : run_conv_umma(x1, x2, weight, (int64_t)Cin * Cout * 27, wp, bias, out, g, Cin, 0, 0, 0, stream, amax_x, amax_x_valid);
  if (amax_x && !amax_x_valid) {
    const int64_t Vi = (int64_t)Di * Hi * Wi;
    const int rca = fill_amax_slot(amax_x, x1, (int64_t)N * C1 * Vi, x2, (int64_t)N * C2 * Vi, stream);
    if (rca) return rca;
  }
  if (ks == 3 && aligned16(wp) && fwd_umma_ok(g))
    return transposed ? run_conv_umma(x1, x2, weight, (int64_t)Cin * Cout * 27, wp, bias, out, g, Cout, 1, 1, 0, stream, amax_x, 1)
                      : run_conv_umma(x1, x2, weight, (int64_t)Cin * Cout * 27, wp, bias, out, g, Cin, 0, 0, 0, stream, amax_x, 1);
  if (ks == 3 && aligned16(wp) && fwd_tma_ok(x1, x2, out, g))
    return transposed ? run_conv_tma(x1, x2, weight, wp, bias, out, g, Cin, Cout, 1, 1, 0, stream)
                      : run_conv_tma(x1, x2, weight, wp, bias, out, g, Cout, Cin, 0, 0, 0, stream);
  int rc = transposed ? repack(weight, wp, Cin, Cout, T, 1, 1, Cin, 0, Cout, 0, g.Cop, stream)
                      : repack(weight, wp, Cout, Cin, T, 0, 0, Cin, 0, Cout, 0, g.Cop, stream);
  if (rc) return rc;
  return run_conv(x1, x2, wp, bias, out, g, ks, stream);
}

namespace {
// dyz[c][2z][2y][2x] = dy[c][z][y][x], zero elsewhere (extent 2Do x 2Ho x 2Wo): turns the data gradient of a stride-2
// convolution into the stride-1 data gradient the tensor-core kernel computes.  One thread = four x-positions of dyz.
__global__ void __launch_bounds__(256) zero_insert2_kernel(const float* __restrict__ dy, float* __restrict__ dyz, int64_t planes,
                                                           int Do, int Ho, int Wo) {
  const int W4 = Wo / 2;  // float4 groups per dyz row (2*Wo / 4); Wo is even here
  const int64_t total = planes * (2 * Do) * (2 * Ho) * W4;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int xg = (int)(i % W4);
    int64_t t = i / W4;
    const int y = (int)(t % (2 * Ho)); t /= 2 * Ho;
    const int z = (int)(t % (2 * Do));
    const int64_t c = t / (2 * Do);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (((y | z) & 1) == 0) {
      const float2 d = __ldg(reinterpret_cast<const float2*>(dy + ((c * Do + z / 2) * Ho + y / 2) * Wo) + xg);
      v.x = d.x; v.z = d.y;
    }
    reinterpret_cast<float4*>(dyz)[i] = v;
  }
}

}  // namespace

// Data gradient for one source: dx [N,Cdx,Di,Hi,Wi] = gradient w.r.t. channels [ci_off, ci_off+Cdx) of the conv input.
// dy [N,Cout,Do,Ho,Wo].  Same weight/transposed convention as forward (Cin_total = all input channels of the layer).
DA_API int da_conv3d_dgrad(const float* dy, const float* weight, int transposed, float* dx, int N, int Cin_total,
                           int ci_off, int Cdx, int Cout, int Di, int Hi, int Wi, int ks, int stride, int pad,
                           void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  return da_conv3d_dgrad_ex(dy, weight, transposed, dx, N, Cin_total, ci_off, Cdx, Cout, Di, Hi, Wi, ks, stride, pad, workspace,
                            workspace_bytes, stream, nullptr, 0);
}

// amax_dy: caller-owned max-abs slot of dy, as amax_x of da_conv3d_fwd_ex
DA_API int da_conv3d_dgrad_ex(const float* dy, const float* weight, int transposed, float* dx, int N, int Cin_total,
                              int ci_off, int Cdx, int Cout, int Di, int Hi, int Wi, int ks, int stride, int pad,
                              void* workspace, int64_t workspace_bytes, cudaStream_t stream, float* amax_dy, int amax_dy_valid) {
  DA_REQUIRE(dy && weight && dx && workspace, "da_conv3d_dgrad: null pointer");
  DA_REQUIRE(ks == 1 || ks == 3, "da_conv3d_dgrad: unsupported kernel size %d", ks);
  DA_REQUIRE(stride == 1 || (stride == 2 && ks == 3 && pad == 1 && !transposed), "da_conv3d_dgrad: unsupported stride/kernel");
  const int T = ks * ks * ks;
  if (workspace_bytes < da_conv3d_pack_bytes(Cin_total, Cout, ks)) { da_set_error("da_conv3d_dgrad: workspace too small"); return DA_ERR_WORKSPACE; }
  const int Do = conv_out(Di, ks, stride, pad), Ho = conv_out(Hi, ks, stride, pad), Wo = conv_out(Wi, ks, stride, pad);
  float* wp = (float*)workspace;
  const int Cp = cpad(Cdx);
  if (amax_dy && !amax_dy_valid) {   // one pass over dy itself (the stride-2 path below would otherwise scan its zero-inserted copy)
    const int rca = fill_amax_slot(amax_dy, dy, (int64_t)N * Cout * Do * Ho * Wo, nullptr, 0, stream);
    if (rca) return rca;
    amax_dy_valid = 1;
  }
  if (stride == 1) {
    // dgrad = conv of dy (Cout channels) with [co][flip tap][ci]; for a transposed layer: no flip, dims swapped
    ConvGeom g{N, Cout, 0, Do, Ho, Wo, Di, Hi, Wi, Cdx, Cp, 1, ks == 3 ? 1 : 0, 0, 0.f};
    DA_REQUIRE(ks == 1 || pad == 1, "da_conv3d_dgrad: k3 needs pad 1");
    if (ks == 1 && pad == 0 && !transposed && !force_direct()) {
      const int rc1 = run_conv1x1_stream(dy, nullptr, Cout, 0, weight + ci_off, 1, Cin_total, nullptr, dx, N, Cdx, (int64_t)Di * Hi * Wi, 0, 0.f, stream);
      if (rc1 >= 0) return rc1;
    }
    if (ks == 3 && aligned16(wp) && fwd_umma_ok(g))
      return transposed ? run_conv_umma(dy, nullptr, weight, (int64_t)Cin_total * Cout * 27, wp, nullptr, dx, g, Cout, 0, 0, ci_off, stream, amax_dy, amax_dy_valid)
                        : run_conv_umma(dy, nullptr, weight, (int64_t)Cin_total * Cout * 27, wp, nullptr, dx, g, Cin_total, 1, 1, ci_off, stream, amax_dy, amax_dy_valid);
    if (ks == 3 && aligned16(wp) && fwd_tma_ok(dy, nullptr, dx, g))
      return transposed ? run_conv_tma(dy, nullptr, weight, wp, nullptr, dx, g, Cin_total, Cout, 0, 0, ci_off, stream)
                        : run_conv_tma(dy, nullptr, weight, wp, nullptr, dx, g, Cout, Cin_total, 1, 1, ci_off, stream);
    int rc = transposed ? repack(weight, wp, Cin_total, Cout, T, 0, 0, Cout, 0, Cdx, ci_off, Cp, stream)
                        : repack(weight, wp, Cout, Cin_total, T, 1, 1, Cout, 0, Cdx, ci_off, Cp, stream);
    if (rc) return rc;
    return run_conv(dy, nullptr, wp, nullptr, dx, g, ks, stream);
  }
  {
    // stride 2 on the tensor cores: zero-insert dy to the input extent, then the stride-1 data gradient (7/8 of the
    // MMAs multiply zeros and it is still ~2x faster than the FFMA kernel below).  Needs the larger workspace.
    ConvGeom g{N, Cout, 0, Di, Hi, Wi, Di, Hi, Wi, Cdx, Cp, 1, 1, 0, 0.f};
    const int64_t pack = da_conv3d_pack_bytes(Cin_total, Cout, ks);
    const int64_t zbytes = (int64_t)sizeof(float) * N * Cout * Di * Hi * Wi;
    float* dyz = (float*)((char*)workspace + ((pack + 255) & ~(int64_t)255));
    if (((Di | Hi | Wi) & 1) == 0 && (Wo & 1) == 0 && workspace_bytes >= ((pack + 255) & ~(int64_t)255) + zbytes && aligned16(wp) &&
        aligned16(dyz) && aligned16(dy) && fwd_umma_ok(g)) {
      const int64_t groups = (int64_t)N * Cout * Di * Hi * (Wi / 4);
      int64_t nb = da_cdiv(groups, 256 * 4);
      if (nb > (int64_t)DA_NUM_SMS * 32) nb = (int64_t)DA_NUM_SMS * 32;
      zero_insert2_kernel<<<(unsigned)nb, 256, 0, stream>>>(dy, dyz, (int64_t)N * Cout, Do, Ho, Wo);
      int rc = da_check_launch("conv3d_dgrad_s2/zero_insert");
      if (rc) return rc;
      if (!amax_dy && !split_tf32()) {   // no caller slot: bound dy (not its 8x larger zero-inserted copy) into a spare workspace float
        amax_dy = reinterpret_cast<float*>(reinterpret_cast<char*>(wp) + pack - 64);
        const int rca = fill_amax_slot(amax_dy, dy, (int64_t)N * Cout * Do * Ho * Wo, nullptr, 0, stream);
        if (rca) return rca;
        amax_dy_valid = 1;
      }
      return run_conv_umma(dyz, nullptr, weight, (int64_t)Cin_total * Cout * 27, wp, nullptr, dx, g, Cin_total, 1, 1, ci_off, stream, amax_dy, amax_dy_valid);
    }
  }
  int rc = repack(weight, wp, Cout, Cin_total, T, 1, 0, Cout, 0, Cdx, ci_off, Cp, stream);
  if (rc) return rc;
  const int CO = (Cp % 16 == 0) ? 16 : (Cp % 8 == 0 ? 8 : 4);
  dim3 grid((unsigned)da_cdiv((int64_t)Di * Hi * Wi, DIRECT_THREADS), Cp / CO, N);
  if (CO == 16) conv3d_dgrad_s2_kernel<16><<<grid, DIRECT_THREADS, 0, stream>>>(dy, wp, dx, N, Cout, Cdx, Cp, Do, Ho, Wo, Di, Hi, Wi);
  else if (CO == 8) conv3d_dgrad_s2_kernel<8><<<grid, DIRECT_THREADS, 0, stream>>>(dy, wp, dx, N, Cout, Cdx, Cp, Do, Ho, Wo, Di, Hi, Wi);
  else conv3d_dgrad_s2_kernel<4><<<grid, DIRECT_THREADS, 0, stream>>>(dy, wp, dx, N, Cout, Cdx, Cp, Do, Ho, Wo, Di, Hi, Wi);
  return da_check_launch("conv3d_dgrad_s2");
}

// Weight gradient (+ optional bias gradient).  grad_weight has the layer's own layout
// ((Cout,Cin,k^3), or (Cin,Cout,k^3) when transposed).  x2 may be null.
// tcgen05 weight gradient (conv3d_wgrad_umma_kernel): k3 s1 p1 layers whose volume amortises the per-CTA TMEM read-out
inline bool wgrad_umma_ok(int N, int D, int H, int W, int Cin, int Cout) {
  if (g_force_direct == 3) return true;
  const char* e = getenv("DA_WGRAD_UMMA");
  if (e && strcmp(e, "0") == 0) return false;
  const int64_t V = (int64_t)N * D * H * W;
  (void)Cin; (void)Cout;
  const bool big = V >= 8192;   // as fwd_umma_ok
  return umma_enabled() && !force_direct() && W >= 16 && big;
}

int run_wgrad_umma(const float* x1, int C1, const float* x2, int C2, const float* dy, int transposed, float* grad_weight,
                   float* grad_bias, int N, int Di, int Hi, int Wi, int Cout, float* partials, int cap, cudaStream_t stream,
                   float* amax_x, int amax_x_valid, float* amax_dy, int amax_dy_valid) {
  const int Cin = C1 + C2;
  const int64_t count = (int64_t)Cin * Cout * 27;
  const int tiles_x = (Wi + WU_XT - 1) / WU_XT, tiles_y = (Hi + WU_YT - 1) / WU_YT;
  static DaPerDeviceOnce configured;
  if (configured.first()) {
    cudaFuncSetAttribute(conv3d_wgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WU_SMEM_BYTES);
    cudaFuncSetAttribute(conv3d_wgrad_umma_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WV_SMEM_BYTES);
    cudaFuncSetAttribute(conv3d_wgrad_umma16_kernel<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WbCfg<16>::SMEM_BYTES);
    cudaFuncSetAttribute(conv3d_wgrad_umma16_kernel<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WbCfg<32>::SMEM_BYTES);
    cudaFuncSetAttribute(conv3d_wgrad_umma16_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WbCfg<16>::SMEM_BYTES);
    cudaFuncSetAttribute(conv3d_wgrad_umma16_kernel<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WbCfg<32>::SMEM_BYTES);
  }
  int rc, nregions;
  float* bias_partials = nullptr;  // set when the kernel folds the bias gradient in
  const bool use_tma = !tma_disabled() && (Wi & 3) == 0 && Wi >= 24 && Hi >= 6 && aligned16(x1) && aligned16(x2) && aligned16(dy);
  if (use_tma) {
    // one launch: halo-side blocks x plain-side blocks x one wave of CTAs.  Work units = (column, z segment) dealt
    // round-robin, so CTAs that run together walk neighbouring columns in lockstep (DRAM pages, L2 lines shared);
    // the segment count minimises rounds x (planes per unit + pipeline fill).
    WgUmmaTmaArgs a;
    a.single_pass = 0;
    const float* h1 = transposed ? dy : x1; const float* h2 = transposed ? nullptr : x2;
    const float* p1 = transposed ? x1 : dy; const float* p2 = transposed ? x2 : nullptr;
    a.H1 = transposed ? Cout : C1; a.H2 = transposed ? 0 : C2;
    a.P1 = transposed ? C1 : Cout; a.P2 = transposed ? C2 : 0;
    // 3xFP16 (default): halo-side blocks of 32 channels when a halo tensor has more than 16 (96 of 128 MMA rows useful)
    const bool bf16 = !split_tf32() && Wi >= WB_XBOX;
    const int cib = (bf16 && (a.H1 > 16 || a.H2 > 16)) ? 32 : 16;
    a.nH1 = (a.H1 + cib - 1) / cib; a.nP1 = (a.P1 + 15) / 16;
    const int nHB = a.nH1 + (a.H2 + cib - 1) / cib;
    a.nPB = a.nP1 + (a.P2 + 15) / 16;
    const int groups = nHB * a.nPB;
    nregions = DA_NUM_SMS / groups;
    if (nregions > cap) nregions = cap;
    if (nregions < 1) nregions = 1;
    a.ncols = N * tiles_y * tiles_x; a.zlen = Di; a.nunits = a.ncols;
    int64_t best = -1;
    for (int segs = 1; segs <= Di; ++segs) {
      const int cand = (Di + segs - 1) / segs;
      if (cand < 8 && segs > 1) break;
      const int64_t units = (int64_t)a.ncols * ((Di + cand - 1) / cand);
      const int64_t cost = ((units + nregions - 1) / nregions) * (cand + 3);
      if (best < 0 || cost < best) { best = cost; a.zlen = cand; a.nunits = (int)units; }
    }
    if (nregions > a.nunits) nregions = a.nunits;
    a.partials = partials; a.region_stride = count; a.D = Di; a.tiles_x = tiles_x; a.tiles_y = tiles_y;
    bias_partials = (grad_bias && !transposed) ? partials + (int64_t)nregions * count : nullptr;
    a.bias_partials = bias_partials;
    a.dbg = nullptr;
    CUtensorMap mh1, mh2, mp1, mp2;
    if (bf16) {
      // per-tensor scales: slot 0 = halo-side tensors, slot 1 = plain-side tensors; the slots sit behind the partials
      // (caller-owned slots with valid bounds replace the passes; slots passed as not valid are filled)
      float* amax = partials + (int64_t)cap * count + (int64_t)WG_MAX_REGIONS * (Cin > Cout ? Cin : Cout);
      const int64_t V = (int64_t)Di * Hi * Wi;
      float* amax_hs = transposed ? amax_dy : amax_x; const int hs_valid = transposed ? amax_dy_valid : amax_x_valid;
      float* amax_ps = transposed ? amax_x : amax_dy; const int ps_valid = transposed ? amax_x_valid : amax_dy_valid;
      const bool have_h = amax_hs && hs_valid, have_p = amax_ps && ps_valid;
      if (!have_h || !have_p) {
        rc = run_absmax(have_h ? nullptr : h1, (int64_t)N * a.H1 * V, 0, have_h ? nullptr : h2, (int64_t)N * a.H2 * V, 0,
                        have_p ? nullptr : p1, (int64_t)N * a.P1 * V, 1, have_p ? nullptr : p2, (int64_t)N * a.P2 * V, 1, amax, stream);
        if (rc) return rc;
        cudaError_t e = cudaSuccess;
        if (amax_hs && !hs_valid) e = cudaMemcpyAsync(amax_hs, amax, sizeof(float), cudaMemcpyDeviceToDevice, stream);
        if (e == cudaSuccess && amax_ps && !ps_valid) e = cudaMemcpyAsync(amax_ps, amax + 1, sizeof(float), cudaMemcpyDeviceToDevice, stream);
        if (e != cudaSuccess) { da_set_error("conv3d_wgrad: amax copy failed: %s", cudaGetErrorString(e)); return (int)e; }
      }
      a.amax_h = have_h ? amax_hs : amax; a.amax_p = have_p ? amax_ps : amax + 1;
      a.single_pass = single_pass();
      rc = da_make_volume_map_xcy(&mh1, h1, N, a.H1, Di, Hi, Wi, WB_XBOX, cib, 4);
      if (!rc) rc = a.H2 ? da_make_volume_map_xcy(&mh2, h2, N, a.H2, Di, Hi, Wi, WB_XBOX, cib, 4) : (mh2 = mh1, 0);
      if (!rc) rc = da_make_volume_map_xcy(&mp1, p1, N, a.P1, Di, Hi, Wi, 16, 16, 6);
      if (!rc) rc = a.P2 ? da_make_volume_map_xcy(&mp2, p2, N, a.P2, Di, Hi, Wi, 16, 16, 6) : (mp2 = mp1, 0);
      if (rc) return rc;
      a.dbg = umma_dbg_buffer();
      const dim3 grid(groups, nregions);
      if (a.dbg) {
        if (cib == 32) conv3d_wgrad_umma16_kernel<32, true><<<grid, WbCfg<32>::THREADS, WbCfg<32>::SMEM_BYTES, stream>>>(mh1, mh2, mp1, mp2, a);
        else conv3d_wgrad_umma16_kernel<16, true><<<grid, WbCfg<16>::THREADS, WbCfg<16>::SMEM_BYTES, stream>>>(mh1, mh2, mp1, mp2, a);
      } else {
        if (cib == 32) conv3d_wgrad_umma16_kernel<32, false><<<grid, WbCfg<32>::THREADS, WbCfg<32>::SMEM_BYTES, stream>>>(mh1, mh2, mp1, mp2, a);
        else conv3d_wgrad_umma16_kernel<16, false><<<grid, WbCfg<16>::THREADS, WbCfg<16>::SMEM_BYTES, stream>>>(mh1, mh2, mp1, mp2, a);
      }
    } else {
      rc = da_make_volume_map(&mh1, h1, N, a.H1, Di, Hi, Wi, 24, 4, 1, 16);
      if (!rc) rc = a.H2 ? da_make_volume_map(&mh2, h2, N, a.H2, Di, Hi, Wi, 24, 4, 1, 16) : (mh2 = mh1, 0);
      if (!rc) rc = da_make_volume_map(&mp1, p1, N, a.P1, Di, Hi, Wi, 16, 6, 1, 16);
      if (!rc) rc = a.P2 ? da_make_volume_map(&mp2, p2, N, a.P2, Di, Hi, Wi, 16, 6, 1, 16) : (mp2 = mp1, 0);
      if (rc) return rc;
      conv3d_wgrad_umma_tma_kernel<<<dim3(groups, nregions), WU_THREADS, WV_SMEM_BYTES, stream>>>(mh1, mh2, mp1, mp2, a);
    }
    rc = da_check_launch("conv3d_wgrad_umma_tma");
  } else {
    const int ntiles = N * Di * tiles_y * tiles_x;
    const int a_ch = transposed ? Cout : Cin, b_ch = transposed ? Cin : Cout;
    const int groups = ((a_ch + 15) / 16) * ((b_ch + 15) / 16);
    nregions = (2 * DA_NUM_SMS + groups / 2) / groups;
    if (nregions > cap) nregions = cap;
    if (nregions > ntiles) nregions = ntiles;
    if (nregions < 1) nregions = 1;
    const int tpr = (ntiles + nregions - 1) / nregions;
    nregions = (ntiles + tpr - 1) / tpr;
    auto launch = [&](const float* xin, int C, int ci_off, int Cin_total_, const float* gout, int Cout_, int co_off) -> int {
      WgUmmaArgs a;
      a.x = xin; a.dy = gout; a.partials = partials;
      a.N = N; a.C = C; a.ci_off = ci_off; a.Cin_total = Cin_total_; a.Cout = Cout_; a.co_off = co_off;
      a.region_stride = count; a.D = Di; a.H = Hi; a.W = Wi;
      a.tiles_x = tiles_x; a.tiles_y = tiles_y; a.tiles_per_region = tpr; a.ntiles = ntiles;
      a.nCoB = (Cout_ + 15) / 16;
      dim3 grid(((C + 15) / 16) * a.nCoB, nregions);
      conv3d_wgrad_umma_kernel<<<grid, WU_THREADS, WU_SMEM_BYTES, stream>>>(a);
      return da_check_launch("conv3d_wgrad_umma");
    };
    if (!transposed) {
      rc = launch(x1, C1, 0, Cin, dy, Cout, 0);
      if (!rc && C2) rc = launch(x2, C2, C1, Cin, dy, Cout, 0);
    } else {
      rc = launch(dy, Cout, 0, Cout, x1, C1, 0);
      if (!rc && C2) rc = launch(dy, Cout, 0, Cout, x2, C2, C1);
    }
  }
  if (rc) return rc;
  launch_reduce_partials(partials, nregions, count, grad_weight, stream);
  rc = da_check_launch("conv3d_wgrad_umma/reduce");
  if (!rc && grad_bias) {
    if (bias_partials) {
      launch_reduce_partials(bias_partials, nregions, Cout, grad_bias, stream);
      rc = da_check_launch("conv3d_wgrad_umma/bias-reduce");
    } else {
      rc = run_channel_sum(dy, N, Cout, (int64_t)Di * Hi * Wi, grad_bias, partials + (int64_t)nregions * count, stream, g_grad_accumulate);
    }
  }
  return rc;
}

DA_API int da_conv3d_wgrad(const float* x1, int C1, const float* x2, int C2, const float* dy, int transposed,
                           float* grad_weight, float* grad_bias, int N, int Di, int Hi, int Wi, int Cout, int ks,
                           int stride, int pad, void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  return da_conv3d_wgrad_ex(x1, C1, x2, C2, dy, transposed, grad_weight, grad_bias, N, Di, Hi, Wi, Cout, ks, stride, pad, workspace,
                            workspace_bytes, stream, nullptr, 0, nullptr, 0, 0);
}

// amax_x / amax_dy: caller-owned max-abs slots of cat(x1, x2) and of dy, as in da_conv3d_fwd_ex.  Here a slot passed as
// not valid is filled only if the tensor-core kernel runs (nothing downstream of a weight gradient reuses it).
// accumulate = 1: the results are ADDED to grad_weight / grad_bias (a parameter used twice in a step, or a gradient
// bucket that already holds this step's earlier contributions) -- folded into the fixed-order region reduce.
DA_API int da_conv3d_wgrad_ex(const float* x1, int C1, const float* x2, int C2, const float* dy, int transposed,
                              float* grad_weight, float* grad_bias, int N, int Di, int Hi, int Wi, int Cout, int ks,
                              int stride, int pad, void* workspace, int64_t workspace_bytes, cudaStream_t stream,
                              float* amax_x, int amax_x_valid, float* amax_dy, int amax_dy_valid, int accumulate) {
  DA_REQUIRE(x1 && dy && grad_weight && workspace, "da_conv3d_wgrad: null pointer");
  GradAccumulateScope accumulate_scope(accumulate ? 1 : 0);   // grad_weight / grad_bias are added to, not overwritten
  DA_REQUIRE(ks == 1 || ks == 3, "da_conv3d_wgrad: unsupported kernel size %d", ks);
  DA_REQUIRE(!transposed || (ks == 3 && stride == 1 && pad == 1), "da_conv3d_wgrad: transposed only for k3 s1 p1");
  const int Cin = C1 + C2, T = ks * ks * ks;
  if (workspace_bytes < da_conv3d_wgrad_workspace_bytes(Cin, Cout, ks)) { da_set_error("da_conv3d_wgrad: workspace too small"); return DA_ERR_WORKSPACE; }
  const int Do = conv_out(Di, ks, stride, pad), Ho = conv_out(Hi, ks, stride, pad), Wo = conv_out(Wi, ks, stride, pad);
  float* partials = (float*)workspace;
  const int64_t count = (int64_t)Cin * Cout * T;
  const int cap = wg_region_cap(count);
  if (ks == 3 && stride == 1 && pad == 1 && !transposed && Cin <= SC_MAX_CIN && g_force_direct <= 0 &&
      (int64_t)N * Di * Hi * ((Wi + 31) / 32) < ((int64_t)1 << 30)) {
    // input layers (automatic selection only): exact-FFMA kernel, one partial row per block
    const int nxc = (Wi + 31) / 32;
    const int64_t units = (int64_t)N * Di * ((Hi + SC_YSEG - 1) / SC_YSEG) * nxc;
    const int nregions = (int)(units < SC_REGIONS ? units : SC_REGIONS);
    float* bias_partials = grad_bias ? partials + (int64_t)nregions * count : nullptr;
    conv3d_wgrad_smallcin_kernel<<<dim3(nregions, (Cout + SC_COB - 1) / SC_COB), 96 * Cin, 0, stream>>>(
        x1, x2, C1, C2, dy, partials, bias_partials, N, Di, Hi, Wi, Cout, count);
    int rc = da_check_launch("conv3d_wgrad_smallcin");
    if (rc) return rc;
    launch_reduce_partials(partials, nregions, count, grad_weight, stream);
    rc = da_check_launch("conv3d_wgrad_smallcin/reduce");
    if (rc || !grad_bias) return rc;
    launch_reduce_partials(bias_partials, nregions, Cout, grad_bias, stream);
    return da_check_launch("conv3d_wgrad_smallcin/bias-reduce");
  }
  if (ks == 3 && stride == 1 && pad == 1 && !force_direct()) {
    // tiled kernel: regions of 4x8x32 tiles; bias gradient folded in (non-transposed layers)
    const int tiles_x = (Wi + TX - 1) / TX, tiles_y = (Hi + TY - 1) / TY, tiles_z = (Di + TZ - 1) / TZ;
    const int ntiles = N * tiles_x * tiles_y * tiles_z;
    const int a_ch = transposed ? Cout : Cin, b_ch = transposed ? Cin : Cout;   // halo-side / plain-side channels
    const int groups = ((a_ch + WG_CI - 1) / WG_CI) * ((b_ch + WG_CO - 1) / WG_CO);
    int nregions = (4 * DA_NUM_SMS + groups / 2) / groups;
    if (nregions > cap) nregions = cap;
    if (nregions > ntiles) nregions = ntiles;
    if (nregions < 1) nregions = 1;
    const int tpr = (ntiles + nregions - 1) / nregions;
    nregions = (ntiles + tpr - 1) / tpr;
    if (wgrad_umma_ok(N, Di, Hi, Wi, Cin, Cout))
      return run_wgrad_umma(x1, C1, x2, C2, dy, transposed, grad_weight, grad_bias, N, Di, Hi, Wi, Cout, partials, cap, stream, amax_x,
                            amax_x_valid, amax_dy, amax_dy_valid);
    float* bias_partials = (grad_bias && !transposed) ? partials + (int64_t)nregions * count : nullptr;
    static DaPerDeviceOnce configured;
    if (configured.first()) {
      cudaFuncSetAttribute(conv3d_wgrad_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WT_SMEM);
      cudaFuncSetAttribute(conv3d_wgrad_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WTM_SMEM_BYTES);
    }
    const bool use_tma = !tma_disabled() && (Wi & 3) == 0;
    auto launch_t = [&](const float* xin, int C, int ci_off, int Cin_total_, const float* gout, int Cout_, int co_off,
                        float* bp) -> int {
      WgTiledArgs a;
      a.x = xin; a.dy = gout; a.partials = partials; a.bias_partials = bp;
      a.N = N; a.C = C; a.ci_off = ci_off; a.Cin_total = Cin_total_; a.Cout = Cout_; a.co_off = co_off;
      a.region_stride = count; a.D = Di; a.H = Hi; a.W = Wi;
      a.tiles_x = tiles_x; a.tiles_y = tiles_y; a.tiles_z = tiles_z; a.tiles_per_region = tpr; a.ntiles = ntiles;
      a.nCoB = (Cout_ + WG_CO - 1) / WG_CO;
      dim3 grid(((C + WG_CI - 1) / WG_CI) * a.nCoB, nregions);
      if (use_tma && aligned16(xin) && aligned16(gout)) {
        CUtensorMap mx, mdy;
        int r = da_make_volume_map(&mx, xin, N, C, Di, Hi, Wi, HXT, HY, HZ, WG_CI);
        if (!r) r = da_make_volume_map(&mdy, gout, N, Cout_, Di, Hi, Wi, TX, TY, TZ, WG_CO);
        if (r) return r;
        conv3d_wgrad_tma_kernel<<<grid, WTM_THREADS, WTM_SMEM_BYTES, stream>>>(mx, mdy, a);
        return da_check_launch("conv3d_wgrad_tma");
      }
      conv3d_wgrad_tiled_kernel<<<grid, WT_THREADS, WT_SMEM, stream>>>(a);
      return da_check_launch("conv3d_wgrad_tiled");
    };
    int rc;
    if (!transposed) {
      rc = launch_t(x1, C1, 0, Cin, dy, Cout, 0, bias_partials);
      if (!rc && C2) rc = launch_t(x2, C2, C1, Cin, dy, Cout, 0, nullptr);
    } else {
      rc = launch_t(dy, Cout, 0, Cout, x1, C1, 0, nullptr);
      if (!rc && C2) rc = launch_t(dy, Cout, 0, Cout, x2, C2, C1, nullptr);
    }
    if (rc) return rc;
    launch_reduce_partials(partials, nregions, count, grad_weight, stream);
    rc = da_check_launch("conv3d_wgrad/reduce");
    if (rc) return rc;
    if (grad_bias) {
      if (bias_partials) {
        launch_reduce_partials(bias_partials, nregions, Cout, grad_bias, stream);
        rc = da_check_launch("conv3d_wgrad/bias-reduce");
      } else {
        rc = run_channel_sum(dy, N, Cout, (int64_t)Do * Ho * Wo, grad_bias, partials + (int64_t)nregions * count, stream, g_grad_accumulate);
      }
    }
    return rc;
  }
  if (ks == 3 && stride == 2 && pad == 1 && !transposed && !force_direct() && !tma_disabled() && (Wi & 3) == 0 && (Wo & 3) == 0 &&
      aligned16(x1) && aligned16(x2) && aligned16(dy)) {
    // stride-2 encoder convolutions: TMA-staged tiled kernel over 2x4x32 OUTPUT tiles
    const int tiles_x = (Wo + S2_TX - 1) / S2_TX, tiles_y = (Ho + S2_TY - 1) / S2_TY, tiles_z = (Do + S2_TZ - 1) / S2_TZ;
    const int ntiles = N * tiles_x * tiles_y * tiles_z;
    const int groups = ((Cin + WG_CI - 1) / WG_CI) * ((Cout + WG_CO - 1) / WG_CO);
    int nregions = (4 * DA_NUM_SMS + groups / 2) / groups;
    if (nregions > cap) nregions = cap;
    if (nregions > ntiles) nregions = ntiles;
    if (nregions < 1) nregions = 1;
    const int tpr = (ntiles + nregions - 1) / nregions;
    nregions = (ntiles + tpr - 1) / tpr;
    float* bias_partials = grad_bias ? partials + (int64_t)nregions * count : nullptr;
    static DaPerDeviceOnce configured_s2;
    if (configured_s2.first()) {
      cudaFuncSetAttribute(conv3d_wgrad_s2_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S2_SMEM_BYTES);
    }
    auto launch_s2 = [&](const float* xin, int C, int ci_off, float* bp) -> int {
      WgTiledArgs a;
      a.x = xin; a.dy = dy; a.partials = partials; a.bias_partials = bp;
      a.N = N; a.C = C; a.ci_off = ci_off; a.Cin_total = Cin; a.Cout = Cout; a.co_off = 0;
      a.region_stride = count; a.D = Do; a.H = Ho; a.W = Wo;
      a.tiles_x = tiles_x; a.tiles_y = tiles_y; a.tiles_z = tiles_z; a.tiles_per_region = tpr; a.ntiles = ntiles;
      a.nCoB = (Cout + WG_CO - 1) / WG_CO;
      CUtensorMap mx, mdy;
      int r = da_make_volume_map(&mx, xin, N, C, Di, Hi, Wi, S2_HX, S2_HY, S2_HZ, WG_CI);
      if (!r) r = da_make_volume_map(&mdy, dy, N, Cout, Do, Ho, Wo, S2_TX, S2_TY, S2_TZ, WG_CO);
      if (r) return r;
      dim3 grid(((C + WG_CI - 1) / WG_CI) * a.nCoB, nregions);
      conv3d_wgrad_s2_tma_kernel<<<grid, WTM_THREADS, S2_SMEM_BYTES, stream>>>(mx, mdy, a);
      return da_check_launch("conv3d_wgrad_s2_tma");
    };
    int rc = launch_s2(x1, C1, 0, bias_partials);
    if (!rc && C2) rc = launch_s2(x2, C2, C1, nullptr);
    if (rc) return rc;
    launch_reduce_partials(partials, nregions, count, grad_weight, stream);
    rc = da_check_launch("conv3d_wgrad_s2/reduce");
    if (!rc && grad_bias) {
      launch_reduce_partials(bias_partials, nregions, Cout, grad_bias, stream);
      rc = da_check_launch("conv3d_wgrad_s2/bias-reduce");
    }
    return rc;
  }
  const int64_t Vk1 = (int64_t)Di * Hi * Wi;
  if (ks == 1 && stride == 1 && pad == 0 && (Vk1 & 3) == 0 && aligned16(x1) && aligned16(x2) && aligned16(dy) && !force_direct()) {
    const int64_t nq = Vk1 / 4;
    int nregions = cap < 64 ? cap : 64;
    if (nregions > nq) nregions = (int)nq;
    const int64_t qpr = da_cdiv(nq, nregions);
    nregions = (int)da_cdiv(nq, qpr);
    auto launch1 = [&](const float* xin, int C, int ci_off) -> int {
      dim3 grid(nregions, ((C + WG_CI - 1) / WG_CI) * ((Cout + WG_CO - 1) / WG_CO));
      conv1x1_wgrad_kernel<<<grid, 256, 0, stream>>>(xin, dy, partials, N, C, ci_off, Cin, Cout, Vk1, qpr, count);
      return da_check_launch("conv1x1_wgrad");
    };
    int rc = launch1(x1, C1, 0);
    if (!rc && C2) rc = launch1(x2, C2, C1);
    if (rc) return rc;
    launch_reduce_partials(partials, nregions, count, grad_weight, stream);
    rc = da_check_launch("conv1x1_wgrad/reduce");
    if (!rc && grad_bias) rc = run_channel_sum(dy, N, Cout, Vk1, grad_bias, partials + (int64_t)nregions * count, stream, g_grad_accumulate);
    return rc;
  }
  const int64_t total_rows = (int64_t)N * Do * Ho * ((Wo + 31) / 32);
  int nregions = (int)(total_rows < cap ? total_rows : cap);
  if (nregions < 1) nregions = 1;
  const int64_t rpr = da_cdiv(total_rows, nregions);
  nregions = (int)da_cdiv(total_rows, rpr);
  // kernel computes  P[co_off+o][ci_off+i][k] = sum_p xin[i][p*s+k-pad] * gout[o][p]
  auto launch = [&](const float* xin, int C, int ci_off, int Cin_total_, const float* gout, int Cout_, int co_off) -> int {
    const int ntasks = ks * ks * ((C + WG_CI - 1) / WG_CI) * ((Cout_ + WG_CO - 1) / WG_CO);
    dim3 grid(nregions, (ntasks + WG_THREADS / 32 - 1) / (WG_THREADS / 32));
    if (ks == 3)
      conv3d_wgrad_kernel<3><<<grid, WG_THREADS, 0, stream>>>(xin, gout, partials, N, C, ci_off, Cin_total_, Cout_, co_off, count, Di, Hi, Wi, Do, Ho, Wo, stride, pad, rpr, total_rows);
    else
      conv3d_wgrad_kernel<1><<<grid, WG_THREADS, 0, stream>>>(xin, gout, partials, N, C, ci_off, Cin_total_, Cout_, co_off, count, Di, Hi, Wi, Do, Ho, Wo, stride, pad, rpr, total_rows);
    return da_check_launch("conv3d_wgrad");
  };
  int rc;
  if (!transposed) {
    rc = launch(x1, C1, 0, Cin, dy, Cout, 0);
    if (!rc && C2) rc = launch(x2, C2, C1, Cin, dy, Cout, 0);
  } else {
    // transposed (k3 s1 p1): dWt[ci][co][k] = sum_q x[ci][q] * dy[co][q+k-1]: the same kernel with the operands
    // swapped ("input" = dy with Cout channels, "outgrad" = x_k); source 2 rows start at ci = C1.
    rc = launch(dy, Cout, 0, Cout, x1, C1, 0);
    if (!rc && C2) rc = launch(dy, Cout, 0, Cout, x2, C2, C1);
  }
  if (rc) return rc;
  launch_reduce_partials((const float*)workspace, nregions, count, grad_weight, stream);
  rc = da_check_launch("conv3d_wgrad/reduce");
  if (rc) return rc;
  if (grad_bias) rc = run_channel_sum(dy, N, Cout, (int64_t)Do * Ho * Wo, grad_bias, partials + (int64_t)nregions * count, stream, g_grad_accumulate);
  return rc;
}

// out[c] = sum over batch and space of x[n][c][:]; workspace of da_channel_sum_workspace_bytes(C)
DA_API int64_t da_channel_sum_workspace_bytes(int C) { return channel_sum_scratch_bytes(C); }
DA_API int da_channel_sum(const float* x, int N, int C, int64_t V, float* out, void* workspace, int64_t workspace_bytes,
                          cudaStream_t stream) {
  DA_REQUIRE(x && out && workspace, "da_channel_sum: null pointer");
  if (workspace_bytes < channel_sum_scratch_bytes(C)) { da_set_error("da_channel_sum: workspace too small"); return DA_ERR_WORKSPACE; }
  return run_channel_sum(x, N, C, V, out, workspace, stream);
}
