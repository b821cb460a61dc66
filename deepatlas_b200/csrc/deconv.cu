// deconv_k2s2: nn.ConvTranspose3d(kernel 2, stride 2) of unets.deconvBlock
// (lib/network_factory/unets.py:42-58, used at :240-241 and UNet.dc9/dc6/dc3 :88,91,94).
//
//   out[co][2z+a][2y+b][2x+c] = bias[co] + sum_ci x[ci][z][y][x] * W[ci][co][a][b][c]
// i.e. eight independent 1x1 convolutions interleaved in space.  The output is 8x the input, so the op is
// write-bound (AI 14-28 FLOP/B).  One thread owns two adjacent input voxels along W
// (= four adjacent output voxels -> float4 stores) and four output channels; the weight layout
// (Cin,Cout,2,2,2) is already [ci][co][pos], so no repack is needed.
#include "common.cuh"
#include <stdlib.h>
#include "tma.cuh"

namespace {

constexpr int DC_THREADS = 128;
constexpr int DC_CC = 32;  // channels staged per smem chunk

// grid (ceil(D*H*ceil(W/2)/128), ceil(Cout/4), N)
__global__ void __launch_bounds__(DC_THREADS) deconv_k2s2_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                     const float* __restrict__ bias, float* __restrict__ out,
                                                                     int Cin, int Cout, int D, int H, int W) {
  __shared__ __align__(16) float sw[DC_CC * 32];  // [ci_local][co 4][pos 8]
  const int n = blockIdx.z, co0 = blockIdx.y * 4;
  const int W2 = (W + 1) / 2;
  const int64_t pairs = (int64_t)D * H * W2, V = (int64_t)D * H * W;
  const int64_t p = (int64_t)blockIdx.x * DC_THREADS + threadIdx.x;
  const bool live = p < pairs;
  const int xt = (int)(p % W2), y = (int)((p / W2) % H), z = (int)(p / ((int64_t)W2 * H));
  const int x0 = 2 * xt;
  const bool two = x0 + 1 < W;
  const int64_t vin = ((int64_t)z * H + y) * W + x0;
  float acc[2][4][8];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[i][c][k] = 0.f;
  for (int c0 = 0; c0 < Cin; c0 += DC_CC) {
    const int cc = min(DC_CC, Cin - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < cc * 32; i += DC_THREADS) {
      const int cl = i / 32, r = i % 32, co = co0 + r / 8;
      sw[i] = co < Cout ? w[((int64_t)(c0 + cl) * Cout + co) * 8 + (r % 8)] : 0.f;
    }
    __syncthreads();
    if (!live) continue;
    for (int cl = 0; cl < cc; ++cl) {
      const float* px = x + ((int64_t)n * Cin + c0 + cl) * V + vin;
      const float xa = __ldg(px), xb = two ? __ldg(px + 1) : 0.f;
      const float4* w4 = reinterpret_cast<const float4*>(sw + cl * 32);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 wv = w4[q];
        const int c = q >> 1, k = (q & 1) * 4;
        acc[0][c][k + 0] = fmaf(xa, wv.x, acc[0][c][k + 0]); acc[1][c][k + 0] = fmaf(xb, wv.x, acc[1][c][k + 0]);
        acc[0][c][k + 1] = fmaf(xa, wv.y, acc[0][c][k + 1]); acc[1][c][k + 1] = fmaf(xb, wv.y, acc[1][c][k + 1]);
        acc[0][c][k + 2] = fmaf(xa, wv.z, acc[0][c][k + 2]); acc[1][c][k + 2] = fmaf(xb, wv.z, acc[1][c][k + 2]);
        acc[0][c][k + 3] = fmaf(xa, wv.w, acc[0][c][k + 3]); acc[1][c][k + 3] = fmaf(xb, wv.w, acc[1][c][k + 3]);
      }
    }
  }
  if (!live) return;
  const int Ho = 2 * H, Wo = 2 * W;
  const int64_t Vo = 8 * V;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int co = co0 + c;
    if (co >= Cout) break;
    const float bv = bias ? bias[co] : 0.f;
#pragma unroll
    for (int ab = 0; ab < 4; ++ab) {
      const int a = ab >> 1, b = ab & 1;
      float* o = out + ((int64_t)n * Cout + co) * Vo + ((int64_t)(2 * z + a) * Ho + (2 * y + b)) * Wo + 2 * x0;
      const float r0 = acc[0][c][ab * 2] + bv, r1 = acc[0][c][ab * 2 + 1] + bv;
      const float r2 = acc[1][c][ab * 2] + bv, r3 = acc[1][c][ab * 2 + 1] + bv;
      if (two && (W & 1) == 0) {
        *reinterpret_cast<float4*>(o) = make_float4(r0, r1, r2, r3);
      } else {
        o[0] = r0; o[1] = r1;
        if (two) { o[2] = r2; o[3] = r3; }
      }
    }
  }
}

// dX[ci][v] = sum_co sum_pos dY[co][2v+pos] * W[ci][co][pos];  grid (pairs/128, ceil(Cin/CI), N).
// dY (8x the voxels of dX) is read once per CI-block of input channels: CI = 16 keeps the re-reads of a tensor that does
// not fit L2 at Cin/16 (the first version used 4 channels per block and was DRAM bound on the repeats).
template <int CI>
__global__ void __launch_bounds__(DC_THREADS) deconv_k2s2_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                                       float* __restrict__ dx, int Cin, int Cout, int D, int H,
                                                                       int W) {
  constexpr int CC = 16;                             // output channels staged per chunk
  __shared__ __align__(16) float sw[CC * CI * 8];    // [co_local][ci][pos 8]
  const int n = blockIdx.z, ci0 = blockIdx.y * CI;
  const int W2 = (W + 1) / 2;
  const int64_t pairs = (int64_t)D * H * W2, V = (int64_t)D * H * W;
  const int64_t p = (int64_t)blockIdx.x * DC_THREADS + threadIdx.x;
  const bool live = p < pairs;
  const int xt = (int)(p % W2), y = (int)((p / W2) % H), z = (int)(p / ((int64_t)W2 * H));
  const int x0 = 2 * xt;
  const bool two = x0 + 1 < W;
  const int Ho = 2 * H, Wo = 2 * W;
  const int64_t Vo = 8 * V;
  float acc[2][CI];
#pragma unroll
  for (int c = 0; c < CI; ++c) acc[0][c] = acc[1][c] = 0.f;
  for (int c0 = 0; c0 < Cout; c0 += CC) {
    const int cc = min(CC, Cout - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < cc * CI * 8; i += DC_THREADS) {
      const int cl = i / (CI * 8), r = i % (CI * 8), ci = ci0 + r / 8;
      sw[i] = ci < Cin ? w[((int64_t)ci * Cout + c0 + cl) * 8 + (r % 8)] : 0.f;
    }
    __syncthreads();
    if (!live) continue;
    for (int cl = 0; cl < cc; ++cl) {
      const float* pg = dy + ((int64_t)n * Cout + c0 + cl) * Vo;
      float g[2][8];
#pragma unroll
      for (int ab = 0; ab < 4; ++ab) {
        const float* q = pg + ((int64_t)(2 * z + (ab >> 1)) * Ho + (2 * y + (ab & 1))) * Wo + 2 * x0;
        if (two && (W & 1) == 0) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(q));
          g[0][ab * 2] = v.x; g[0][ab * 2 + 1] = v.y; g[1][ab * 2] = v.z; g[1][ab * 2 + 1] = v.w;
        } else {
          g[0][ab * 2] = q[0]; g[0][ab * 2 + 1] = q[1];
          g[1][ab * 2] = two ? q[2] : 0.f; g[1][ab * 2 + 1] = two ? q[3] : 0.f;
        }
      }
      const float4* ws = reinterpret_cast<const float4*>(sw + cl * CI * 8);
#pragma unroll
      for (int c = 0; c < CI; ++c) {
        const float4 wa = ws[2 * c], wb = ws[2 * c + 1];
        acc[0][c] = fmaf(g[0][0], wa.x, acc[0][c]); acc[1][c] = fmaf(g[1][0], wa.x, acc[1][c]);
        acc[0][c] = fmaf(g[0][1], wa.y, acc[0][c]); acc[1][c] = fmaf(g[1][1], wa.y, acc[1][c]);
        acc[0][c] = fmaf(g[0][2], wa.z, acc[0][c]); acc[1][c] = fmaf(g[1][2], wa.z, acc[1][c]);
        acc[0][c] = fmaf(g[0][3], wa.w, acc[0][c]); acc[1][c] = fmaf(g[1][3], wa.w, acc[1][c]);
        acc[0][c] = fmaf(g[0][4], wb.x, acc[0][c]); acc[1][c] = fmaf(g[1][4], wb.x, acc[1][c]);
        acc[0][c] = fmaf(g[0][5], wb.y, acc[0][c]); acc[1][c] = fmaf(g[1][5], wb.y, acc[1][c]);
        acc[0][c] = fmaf(g[0][6], wb.z, acc[0][c]); acc[1][c] = fmaf(g[1][6], wb.z, acc[1][c]);
        acc[0][c] = fmaf(g[0][7], wb.w, acc[0][c]); acc[1][c] = fmaf(g[1][7], wb.w, acc[1][c]);
      }
    }
  }
  if (!live) return;
  const int64_t vin = ((int64_t)z * H + y) * W + x0;
#pragma unroll
  for (int c = 0; c < CI; ++c) {
    const int ci = ci0 + c;
    if (ci >= Cin) break;
    float* o = dx + ((int64_t)n * Cin + ci) * V + vin;
    o[0] = acc[0][c];
    if (two) o[1] = acc[1][c];
  }
}

// dW[ci][co][pos] = sum_{n,v} x[ci][v] * dY[co][2v+pos].  lane = input voxel along W; warp task = (4 ci, 2 co).
constexpr int DW_THREADS = 256;
__global__ void __launch_bounds__(DW_THREADS) deconv_k2s2_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                       float* __restrict__ partials, int N, int Cin, int Cout,
                                                                       int D, int H, int W, int64_t rows_per_region,
                                                                       int64_t total_rows) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nCiB = (Cin + 3) / 4, nCoB = (Cout + 1) / 2;
  const int task = blockIdx.y * (DW_THREADS / 32) + warp;
  if (task >= nCiB * nCoB) return;
  const int cob = task % nCoB, cib = task / nCoB;
  const int xb = (W + 31) / 32;
  const int Ho = 2 * H, Wo = 2 * W;
  const int64_t V = (int64_t)D * H * W, Vo = 8 * V;
  float acc[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) acc[i] = 0.f;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_region, r1 = min(total_rows, r0 + rows_per_region);
  for (int64_t rb = r0; rb < r1; ++rb) {
    const int bxi = (int)(rb % xb);
    int64_t t = rb / xb;
    const int y = (int)(t % H); t /= H;
    const int z = (int)(t % D);
    const int n = (int)(t / D);
    const int xi = bxi * 32 + lane;
    const bool live = xi < W;
    float xv[4], g[2][8];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int ci = cib * 4 + c;
      xv[c] = (live && ci < Cin) ? __ldg(x + ((int64_t)n * Cin + ci) * V + ((int64_t)z * H + y) * W + xi) : 0.f;
    }
#pragma unroll
    for (int o = 0; o < 2; ++o) {
      const int co = cob * 2 + o;
#pragma unroll
      for (int ab = 0; ab < 4; ++ab) {
        float2 v = make_float2(0.f, 0.f);
        if (live && co < Cout)
          v = __ldg(reinterpret_cast<const float2*>(dy + ((int64_t)n * Cout + co) * Vo +
                                                    ((int64_t)(2 * z + (ab >> 1)) * Ho + (2 * y + (ab & 1))) * Wo + 2 * xi));
        g[o][ab * 2] = v.x; g[o][ab * 2 + 1] = v.y;
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int o = 0; o < 2; ++o)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[(c * 2 + o) * 8 + k] = fmaf(xv[c], g[o][k], acc[(c * 2 + o) * 8 + k]);
  }
  // halving butterfly: 2 groups of 32
#pragma unroll
  for (int gI = 0; gI < 2; ++gI) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
      const bool up = (lane & s) != 0;
#pragma unroll
      for (int i = 0; i < s; ++i) {
        const float keep = up ? acc[gI * 32 + i + s] : acc[gI * 32 + i];
        const float send = up ? acc[gI * 32 + i] : acc[gI * 32 + i + s];
        acc[gI * 32 + i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
      }
    }
  }
  float* pr = partials + (int64_t)blockIdx.x * Cin * Cout * 8;
#pragma unroll
  for (int gI = 0; gI < 2; ++gI) {
    const int e = gI * 32 + lane;
    const int k = e % 8, o = (e / 8) % 2, c = e / 16;
    const int ci = cib * 4 + c, co = cob * 2 + o;
    if (ci < Cin && co < Cout) pr[((int64_t)ci * Cout + co) * 8 + k] = acc[gI * 32];
  }
}

// TMA-staged tiled variant (W % 4 == 0, 16-byte aligned tensors): a block owns 8 ci x 8 co and walks a region of
// 1 x 4 x 32 input-voxel tiles; the dY tile [8 co][2][8][64] (32 KB) and the x tile [8 ci][4][32] are box-loaded once and
// shared by the four warps, warp = (a, b) output-row parity, both c parities in registers: 128 accumulators, 512 FFMA
// per 24 LDS.128.  dY is read (Cin/8) times from L2 instead of (Cin/4) times through L1.
constexpr int DT_TY = 4, DT_TX = 32, DT_CI = 8, DT_CO = 8;
constexpr int DT_SX_BYTES = DT_CI * DT_TY * DT_TX * 4;            // 4096
constexpr int DT_SD_BYTES = DT_CO * 2 * (2 * DT_TY) * (2 * DT_TX) * 4;  // 32768
constexpr int DT_STAGE_BYTES = DT_SX_BYTES + DT_SD_BYTES;
constexpr int DT_SMEM_BYTES = 2 * DT_STAGE_BYTES + 128;
constexpr int DT_THREADS = 128;

struct DeconvWgArgs {
  float* partials;
  int Cin, Cout, N, D, H, W;
  int tiles_x, tiles_y, tiles_per_region, ntiles, nCoB;
  int64_t region_stride;
};

__global__ void __launch_bounds__(DT_THREADS, 2)
    deconv_k2s2_wgrad_tma_kernel(const __grid_constant__ CUtensorMap mx, const __grid_constant__ CUtensorMap mdy, DeconvWgArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[2];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pa = warp >> 1, pb = warp & 1;  // output parity along z and y handled by this warp
  const int cob = blockIdx.x % a.nCoB, cib = blockIdx.x / a.nCoB;
  const int region = blockIdx.y;
  if (threadIdx.x == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_fence_init();
    tma_prefetch_desc(&mx);
    tma_prefetch_desc(&mdy);
  }
  __syncthreads();
  float acc[DT_CI][DT_CO][2];
#pragma unroll
  for (int c = 0; c < DT_CI; ++c)
#pragma unroll
    for (int o = 0; o < DT_CO; ++o) acc[c][o][0] = acc[c][o][1] = 0.f;
  const int t0 = region * a.tiles_per_region, t1 = min(a.ntiles, t0 + a.tiles_per_region);
  auto issue = [&](int t, int s) {
    int tb = t;
    const int bx = tb % a.tiles_x; tb /= a.tiles_x;
    const int by = tb % a.tiles_y; tb /= a.tiles_y;
    const int z = tb % a.D;
    const int n = tb / a.D;
    uint8_t* st = smem + s * DT_STAGE_BYTES;
    mbar_expect_tx(&full[s], DT_STAGE_BYTES);
    tma_load_5d(st, &mx, &full[s], bx * DT_TX, by * DT_TY, z, cib * DT_CI, n);
    tma_load_5d(st + DT_SX_BYTES, &mdy, &full[s], 2 * bx * DT_TX, 2 * by * DT_TY, 2 * z, cob * DT_CO, n);
  };
  if (threadIdx.x == 0 && t0 < t1) issue(t0, 0);
  const int tx4 = lane & 7, ty = lane >> 3;  // one iteration: 4 rows x 8 quads of 4 input voxels
  for (int t = t0; t < t1; ++t) {
    const int k = t - t0, s = k & 1;
    if (threadIdx.x == 0 && t + 1 < t1) issue(t + 1, s ^ 1);
    mbar_wait(&full[s], (k >> 1) & 1);
    const float* sx = reinterpret_cast<const float*>(smem + s * DT_STAGE_BYTES);
    const float* sd = reinterpret_cast<const float*>(smem + s * DT_STAGE_BYTES + DT_SX_BYTES);
    float4 xv[DT_CI];
#pragma unroll
    for (int c = 0; c < DT_CI; ++c) xv[c] = *reinterpret_cast<const float4*>(sx + (c * DT_TY + ty) * DT_TX + tx4 * 4);
#pragma unroll
    for (int o = 0; o < DT_CO; ++o) {
      const float4* dr = reinterpret_cast<const float4*>(sd + ((o * 2 + pa) * (2 * DT_TY) + 2 * ty + pb) * (2 * DT_TX) + tx4 * 8);
      const float4 d0 = dr[0], d1 = dr[1];  // dY at output x = 8*tx4 .. 8*tx4+7: even = parity 0, odd = parity 1
#pragma unroll
      for (int c = 0; c < DT_CI; ++c) {
        acc[c][o][0] = fmaf(xv[c].x, d0.x, fmaf(xv[c].y, d0.z, fmaf(xv[c].z, d1.x, fmaf(xv[c].w, d1.z, acc[c][o][0]))));
        acc[c][o][1] = fmaf(xv[c].x, d0.y, fmaf(xv[c].y, d0.w, fmaf(xv[c].z, d1.y, fmaf(xv[c].w, d1.w, acc[c][o][1]))));
      }
    }
    __syncthreads();
  }
  // fold the 32 lanes (xor butterfly, fixed order) and write this region's partial sums
  float* pr = a.partials + (int64_t)region * a.region_stride;
#pragma unroll
  for (int c = 0; c < DT_CI; ++c)
#pragma unroll
    for (int o = 0; o < DT_CO; ++o)
#pragma unroll
      for (int pc = 0; pc < 2; ++pc) {
        const float v = warp_sum(acc[c][o][pc]);
        const int ci = cib * DT_CI + c, co = cob * DT_CO + o;
        if (lane == 0 && ci < a.Cin && co < a.Cout) pr[((int64_t)ci * a.Cout + co) * 8 + (pa * 2 + pb) * 2 + pc] = v;
      }
}

// out[i] (+)= sum_r partials[r][i], fixed order: block = 32 consecutive elements (coalesced rows) x 8 warps that deal the
// regions among themselves.  accumulate: a parameter gradient that already holds an earlier contribution of the same step.
__global__ void __launch_bounds__(256) dc_reduce_partials_kernel(const float* __restrict__ partials, int nregions, int64_t count,
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
__global__ void dc_add_kernel(float* __restrict__ out, const float* __restrict__ v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] += v[i];
}

constexpr int DC_MAX_REGIONS = 444;   // three blocks per SM of the tensor-core weight gradient
inline int dc_region_cap(int64_t count) {
  int64_t r = ((int64_t)16 << 20) / (count > 0 ? count : 1);
  if (r > DC_MAX_REGIONS) r = DC_MAX_REGIONS;
  if (r < 4) r = 4;
  return (int)r;
}

#include "deconv_mma.inc.cuh"

inline bool dm_aligned8(const void* p) { return (((uintptr_t)p) & 7) == 0; }
// tensor-core kernels: 32 or 64 input channels (the U-Net's up-samplers); everything else keeps the FFMA kernels.
// DA_DECONV_MMA=0 switches them off (A/B timing).
inline bool dm_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("DA_DECONV_MMA"); on = (e && e[0] == '0') ? 0 : 1; }
  return on == 1;
}

}  // namespace

extern "C" int64_t da_channel_sum_workspace_bytes(int C);
extern "C" int da_channel_sum(const float* x, int N, int C, int64_t V, float* out, void* workspace, int64_t workspace_bytes,
                              cudaStream_t stream);

DA_API int64_t da_deconv_k2s2_wgrad_workspace_bytes(int Cin, int Cout) {
  const int64_t count = (int64_t)Cin * Cout * 8;
  return (int64_t)sizeof(float) * ((int64_t)dc_region_cap(count) * count + (int64_t)DC_MAX_REGIONS * Cout) + 256 +
         da_channel_sum_workspace_bytes(Cout);
}

// x [N,Cin,D,H,W]; weight (Cin,Cout,2,2,2); out [N,Cout,2D,2H,2W]
DA_API int da_deconv_k2s2_fwd(const float* x, const float* weight, const float* bias, float* out, int N, int Cin, int Cout,
                              int D, int H, int W, cudaStream_t stream) {
  DA_REQUIRE(x && weight && out, "da_deconv_k2s2_fwd: null pointer");
  if (dm_enabled() && (Cin == 32 || Cin == 64) && Cout % 32 == 0 && dm_aligned8(out)) {
    static DaPerDeviceOnce configured;
    if (configured.first()) {
      cudaFuncSetAttribute(deconv_k2s2_fwd_mma_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 264 * 4);
      cudaFuncSetAttribute(deconv_k2s2_fwd_mma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 264 * 4);
    }
    const int64_t nt = da_cdiv((int64_t)D * H * W, 32);
    int64_t nbx = da_cdiv(nt, DM_FWD_THREADS / 32);
    if (nbx > 3 * DA_NUM_SMS) nbx = 3 * DA_NUM_SMS;
    dim3 grid((unsigned)nbx, Cout / 32, N);
    if (Cin == 32) deconv_k2s2_fwd_mma_kernel<32><<<grid, DM_FWD_THREADS, 32 * 264 * 4, stream>>>(x, weight, bias, out, Cout, D, H, W);
    else deconv_k2s2_fwd_mma_kernel<64><<<grid, DM_FWD_THREADS, 64 * 264 * 4, stream>>>(x, weight, bias, out, Cout, D, H, W);
    return da_check_launch("da_deconv_k2s2_fwd/mma");
  }
  const int64_t pairs = (int64_t)D * H * ((W + 1) / 2);
  dim3 grid((unsigned)da_cdiv(pairs, DC_THREADS), (Cout + 3) / 4, N);
  deconv_k2s2_fwd_kernel<<<grid, DC_THREADS, 0, stream>>>(x, weight, bias, out, Cin, Cout, D, H, W);
  return da_check_launch("da_deconv_k2s2_fwd");
}

DA_API int da_deconv_k2s2_dgrad(const float* dy, const float* weight, float* dx, int N, int Cin, int Cout, int D, int H, int W,
                                cudaStream_t stream) {
  DA_REQUIRE(dy && weight && dx, "da_deconv_k2s2_dgrad: null pointer");
  const int64_t dm_smem = (int64_t)Cin * (8 * Cout + 4) * 4;
  if (dm_enabled() && (Cin == 32 || Cin == 64) && Cout % 4 == 0 && dm_smem <= 200 * 1024) {
    static DaPerDeviceOnce configured;
    if (configured.first()) {
      cudaFuncSetAttribute(deconv_k2s2_dgrad_mma_kernel<32, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      cudaFuncSetAttribute(deconv_k2s2_dgrad_mma_kernel<64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    }
    const int mt = Cin == 32 ? 2 : 1;
    int64_t nbx = da_cdiv(da_cdiv((int64_t)D * H * W, 16 * mt), DM_DGRAD_THREADS / 32);
    const int per_sm = dm_smem > 100 * 1024 ? 1 : 2;
    if (nbx > per_sm * DA_NUM_SMS) nbx = per_sm * DA_NUM_SMS;
    dim3 grid((unsigned)nbx, 1, N);
    if (Cin == 32) deconv_k2s2_dgrad_mma_kernel<32, 2><<<grid, DM_DGRAD_THREADS, dm_smem, stream>>>(dy, weight, dx, Cout, D, H, W);
    else deconv_k2s2_dgrad_mma_kernel<64, 1><<<grid, DM_DGRAD_THREADS, dm_smem, stream>>>(dy, weight, dx, Cout, D, H, W);
    return da_check_launch("da_deconv_k2s2_dgrad/mma");
  }
  const int64_t pairs = (int64_t)D * H * ((W + 1) / 2);
  // few channels per block only where that is needed to fill the GPU (small volumes)
  const int64_t nbx = da_cdiv(pairs, DC_THREADS);
  if (Cin > 8 && nbx * ((Cin + 15) / 16) * N >= DA_NUM_SMS) {
    dim3 grid((unsigned)nbx, (Cin + 15) / 16, N);
    deconv_k2s2_dgrad_kernel<16><<<grid, DC_THREADS, 0, stream>>>(dy, weight, dx, Cin, Cout, D, H, W);
  } else {
    dim3 grid((unsigned)nbx, (Cin + 3) / 4, N);
    deconv_k2s2_dgrad_kernel<4><<<grid, DC_THREADS, 0, stream>>>(dy, weight, dx, Cin, Cout, D, H, W);
  }
  return da_check_launch("da_deconv_k2s2_dgrad");
}

namespace {
// bias gradient of the FFMA paths: channel sums of dY (the weight partials at the head of the workspace have been consumed
// by the reduce on the same stream); accumulating: sums into the workspace tail, then added
int dc_bias_sum(const float* dy, int N, int Cout, int64_t Vo, float* grad_bias, int accumulate, void* workspace, int64_t workspace_bytes,
                cudaStream_t stream) {
  if (!accumulate) return da_channel_sum(dy, N, Cout, Vo, grad_bias, workspace, workspace_bytes, stream);
  float* tmp = (float*)workspace;   // [Cout] ahead of the channel-sum scratch
  const int64_t head = 256 * (((int64_t)Cout * 4 + 255) / 256);
  int rc = da_channel_sum(dy, N, Cout, Vo, tmp, (char*)workspace + head, workspace_bytes - head, stream);
  if (rc) return rc;
  dc_add_kernel<<<(Cout + 127) / 128, 128, 0, stream>>>(grad_bias, tmp, Cout);
  return da_check_launch("da_deconv_k2s2_wgrad/bias-add");
}
}  // namespace

DA_API int da_deconv_k2s2_wgrad_ex(const float* x, const float* dy, float* grad_weight, float* grad_bias, int N, int Cin, int Cout,
                                   int D, int H, int W, int accumulate, void* workspace, int64_t workspace_bytes, cudaStream_t stream);
DA_API int da_deconv_k2s2_wgrad(const float* x, const float* dy, float* grad_weight, float* grad_bias, int N, int Cin, int Cout,
                                int D, int H, int W, void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  return da_deconv_k2s2_wgrad_ex(x, dy, grad_weight, grad_bias, N, Cin, Cout, D, H, W, 0, workspace, workspace_bytes, stream);
}

// accumulate = 1: grad_weight / grad_bias += the result (a gradient bucket that already holds earlier contributions)
DA_API int da_deconv_k2s2_wgrad_ex(const float* x, const float* dy, float* grad_weight, float* grad_bias, int N, int Cin, int Cout,
                                   int D, int H, int W, int accumulate, void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  DA_REQUIRE(x && dy && grad_weight && workspace, "da_deconv_k2s2_wgrad: null pointer");
  if (workspace_bytes < da_deconv_k2s2_wgrad_workspace_bytes(Cin, Cout)) { da_set_error("da_deconv_k2s2_wgrad: workspace too small"); return DA_ERR_WORKSPACE; }
  const int64_t count = (int64_t)Cin * Cout * 8;
  const int cap = dc_region_cap(count);
  if (dm_enabled() && (Cin == 32 || Cin == 64) && Cout % 32 == 0 && W % 8 == 0) {
    const int64_t nsteps = (int64_t)N * D * H * (W / 8);
    const int gy = Cout / 32;
    int64_t nregions = (Cin == 32 ? 3 : 1) * DA_NUM_SMS / gy;   // 128-thread blocks: three per SM; 256-thread: one
    if (nregions > cap) nregions = cap;
    if (nregions > nsteps) nregions = nsteps;
    if (nregions < 1) nregions = 1;
    const int64_t spr = da_cdiv(nsteps, nregions);
    nregions = da_cdiv(nsteps, spr);
    float* partials = (float*)workspace;
    float* bias_partials = grad_bias ? partials + (int64_t)cap * count : nullptr;
    dim3 grid((unsigned)nregions, gy);
    if (Cin == 32) deconv_k2s2_wgrad_mma_kernel<32><<<grid, 128, 0, stream>>>(x, dy, partials, bias_partials, N, Cout, D, H, W, (int)spr);
    else deconv_k2s2_wgrad_mma_kernel<64><<<grid, 256, 0, stream>>>(x, dy, partials, bias_partials, N, Cout, D, H, W, (int)spr);
    int rc = da_check_launch("da_deconv_k2s2_wgrad/mma");
    if (rc) return rc;
    dc_reduce_partials_kernel<<<(unsigned)da_cdiv(count, 32), 256, 0, stream>>>(partials, (int)nregions, count, grad_weight, accumulate);
    rc = da_check_launch("da_deconv_k2s2_wgrad/reduce");
    if (rc || !grad_bias) return rc;
    dc_reduce_partials_kernel<<<(unsigned)da_cdiv(Cout, 32), 256, 0, stream>>>(bias_partials, (int)nregions, Cout, grad_bias, accumulate);
    return da_check_launch("da_deconv_k2s2_wgrad/bias-reduce");
  }
  if ((W & 3) == 0 && ((((uintptr_t)x) | ((uintptr_t)dy)) & 15) == 0 && da_get_encode_tiled() != nullptr) {
    DeconvWgArgs a;
    a.partials = (float*)workspace; a.Cin = Cin; a.Cout = Cout; a.N = N; a.D = D; a.H = H; a.W = W;
    a.tiles_x = (W + DT_TX - 1) / DT_TX; a.tiles_y = (H + DT_TY - 1) / DT_TY;
    a.ntiles = N * D * a.tiles_x * a.tiles_y;
    a.nCoB = (Cout + DT_CO - 1) / DT_CO;
    const int groups = ((Cin + DT_CI - 1) / DT_CI) * a.nCoB;
    int nregions = (6 * DA_NUM_SMS + groups / 2) / groups;
    if (nregions > cap) nregions = cap;
    if (nregions > a.ntiles) nregions = a.ntiles;
    if (nregions < 1) nregions = 1;
    a.tiles_per_region = (a.ntiles + nregions - 1) / nregions;
    nregions = (a.ntiles + a.tiles_per_region - 1) / a.tiles_per_region;
    a.region_stride = count;
    CUtensorMap mx, mdy;
    int r = da_make_volume_map(&mx, x, N, Cin, D, H, W, DT_TX, DT_TY, 1, DT_CI);
    if (!r) r = da_make_volume_map(&mdy, dy, N, Cout, 2 * D, 2 * H, 2 * W, 2 * DT_TX, 2 * DT_TY, 2, DT_CO);
    if (r) return r;
    static DaPerDeviceOnce configured;
    if (configured.first()) {
      cudaFuncSetAttribute(deconv_k2s2_wgrad_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DT_SMEM_BYTES);
    }
    deconv_k2s2_wgrad_tma_kernel<<<dim3(groups, nregions), DT_THREADS, DT_SMEM_BYTES, stream>>>(mx, mdy, a);
    int rc = da_check_launch("da_deconv_k2s2_wgrad_tma");
    if (rc) return rc;
    dc_reduce_partials_kernel<<<(unsigned)da_cdiv(count, 32), 256, 0, stream>>>((const float*)workspace, nregions, count, grad_weight, accumulate);
    rc = da_check_launch("da_deconv_k2s2_wgrad/reduce");
    if (rc || !grad_bias) return rc;
    return dc_bias_sum(dy, N, Cout, (int64_t)8 * D * H * W, grad_bias, accumulate, workspace, workspace_bytes, stream);
  }
  const int64_t total_rows = (int64_t)N * D * H * ((W + 31) / 32);
  int nregions = (int)(total_rows < cap ? total_rows : cap);
  const int64_t rpr = da_cdiv(total_rows, nregions);
  nregions = (int)da_cdiv(total_rows, rpr);
  const int ntasks = ((Cin + 3) / 4) * ((Cout + 1) / 2);
  dim3 grid(nregions, (ntasks + DW_THREADS / 32 - 1) / (DW_THREADS / 32));
  deconv_k2s2_wgrad_kernel<<<grid, DW_THREADS, 0, stream>>>(x, dy, (float*)workspace, N, Cin, Cout, D, H, W, rpr, total_rows);
  int rc = da_check_launch("da_deconv_k2s2_wgrad");
  if (rc) return rc;
  dc_reduce_partials_kernel<<<(unsigned)da_cdiv(count, 32), 256, 0, stream>>>((const float*)workspace, nregions, count, grad_weight, accumulate);
  rc = da_check_launch("da_deconv_k2s2_wgrad/reduce");
  if (rc || !grad_bias) return rc;
  // the weight partials have been consumed by the reduce above (same stream): reuse the workspace head
  return dc_bias_sum(dy, N, Cout, (int64_t)8 * D * H * W, grad_bias, accumulate, workspace, workspace_bytes, stream);
}
