// TMA-staged variants of the k3 s1 p1 convolution kernels (included by conv3d.cu inside its anonymous namespace).
//
// Same arithmetic and thread mappings as conv3d_tiled_kernel / conv3d_wgrad_tiled_kernel, but the shared-memory
// tiles are filled by the TMA engine (cp.async.bulk.tensor, rank-5 map {W,H,D,C,N}) instead of by the compute
// threads: one elected thread arms an mbarrier with the stage's byte count and issues the box loads; halo
// positions outside the volume and channels beyond C arrive as zeros (hardware OOB fill), so the kernels carry
// no boundary predicates and no staging index arithmetic.  Two stages: the loads of chunk/tile k+1 fly while the
// FFMA loop consumes k.  Weights of the forward kernel come in as one 1-D bulk copy per chunk from a
// [cout-group][cin_pad][27][CO] repack.
//
// Eligibility (checked on the host, otherwise the non-TMA kernels run): W % 4 == 0 (16-byte global strides),
// 16-byte aligned base pointers, and for two-source layers C1 % CK == 0.

constexpr int TMA_CK = 4;  // input channels per forward stage
// The TMA unit requires the innermost box coordinate to be 16-byte aligned (measured: x0 % 4 != 0 raises an illegal-
// instruction fault on sm_100a, tools/probes/tma_probe.cu), so the halo box starts at X0-4 instead of X0-1 and is 40 wide;
// element x-1 of a thread's first voxel sits at float 3 of three aligned LDS.128.
constexpr int HXT = 40;    // halo row pitch of the TMA kernels (floats)
constexpr int HX0 = 4;     // halo starts HX0 voxels left of the tile

__host__ __device__ constexpr int round128(int bytes) { return (bytes + 127) / 128 * 128; }

template <int CO>
struct FwdTmaCfg {
  static constexpr int SX_BYTES = round128(TMA_CK * HZ * HY * HXT * 4);  // 38400
  static constexpr int SW_BYTES = round128(TMA_CK * 27 * CO * 4);
  static constexpr int STAGE_BYTES = SX_BYTES + SW_BYTES;
  static constexpr int SMEM_BYTES = 2 * STAGE_BYTES + 128;  // + alignment slack
  static constexpr uint32_t TX_BYTES = TMA_CK * HZ * HY * HXT * 4 + TMA_CK * 27 * CO * 4;
};

template <int CO>
__global__ void __launch_bounds__(TILED_THREADS, 2)
    conv3d_fwd_tma_kernel(const __grid_constant__ CUtensorMap mx1, const __grid_constant__ CUtensorMap mx2,
                          const float* __restrict__ wp, const float* __restrict__ bias, float* __restrict__ out, ConvGeom g,
                          int tiles_x, int tiles_y, int cin_pad) {
  using Cfg = FwdTmaCfg<CO>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[2];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  const int n = blockIdx.z, cog = blockIdx.y;
  int tb = blockIdx.x;
  const int bx = tb % tiles_x; tb /= tiles_x;
  const int by = tb % tiles_y;
  const int bz = tb / tiles_y;
  const int X0 = bx * TX, Y0 = by * TY, Z0 = bz * TZ;
  const int tx = threadIdx.x % (TX / VX), ty = (threadIdx.x / (TX / VX)) % TY, tz = threadIdx.x / ((TX / VX) * TY);
  const int nchunks = cin_pad / TMA_CK;
  const int chunks1 = (g.C1 + TMA_CK - 1) / TMA_CK;  // chunks served by source 1 (C1 % CK == 0 whenever C2 > 0)

  if (threadIdx.x == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_fence_init();
    tma_prefetch_desc(&mx1);
    if (g.C2) tma_prefetch_desc(&mx2);
  }
  __syncthreads();

  auto issue = [&](int c, int s) {
    uint8_t* st = smem + s * Cfg::STAGE_BYTES;
    mbar_expect_tx(&full[s], Cfg::TX_BYTES);
    if (c < chunks1)
      tma_load_5d(st, &mx1, &full[s], X0 - HX0, Y0 - 1, Z0 - 1, c * TMA_CK, n);
    else
      tma_load_5d(st, &mx2, &full[s], X0 - HX0, Y0 - 1, Z0 - 1, (c - chunks1) * TMA_CK, n);
    bulk_load_1d(st + Cfg::SX_BYTES, wp + ((int64_t)cog * cin_pad + (int64_t)c * TMA_CK) * 27 * CO, TMA_CK * 27 * CO * 4, &full[s]);
  };

  float acc[VX][CO];
#pragma unroll
  for (int i = 0; i < VX; ++i)
#pragma unroll
    for (int c = 0; c < CO; ++c) acc[i][c] = 0.f;

  if (threadIdx.x == 0) issue(0, 0);
  for (int c = 0; c < nchunks; ++c) {
    const int s = c & 1;
    if (threadIdx.x == 0 && c + 1 < nchunks) issue(c + 1, s ^ 1);  // stage s^1 was released by the barrier ending chunk c-1
    mbar_wait(&full[s], (c >> 1) & 1);
    const float* sx = reinterpret_cast<const float*>(smem + s * Cfg::STAGE_BYTES);
    const float* sw = reinterpret_cast<const float*>(smem + s * Cfg::STAGE_BYTES + Cfg::SX_BYTES);
#pragma unroll 1
    for (int cl = 0; cl < TMA_CK; ++cl) {
#pragma unroll
      for (int kz = 0; kz < 3; ++kz) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const float* row = sx + ((cl * HZ + tz + kz) * HY + ty + ky) * HXT + tx * VX;
          const float4 a = *reinterpret_cast<const float4*>(row);
          const float4 b = *reinterpret_cast<const float4*>(row + 4);
          const float4 e = *reinterpret_cast<const float4*>(row + 8);
          const float in[6] = {a.w, b.x, b.y, b.z, b.w, e.x};
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
#ifdef DA_W_LDG
            const float4* w4 = reinterpret_cast<const float4*>(wp + (((int64_t)cog * cin_pad + (int64_t)c * TMA_CK + cl) * 27 + (kz * 3 + ky) * 3 + kx) * CO);
#else
            const float4* w4 = reinterpret_cast<const float4*>(sw + (cl * 27 + (kz * 3 + ky) * 3 + kx) * CO);
#endif
#pragma unroll
            for (int q = 0; q < CO / 4; ++q) {
#ifdef DA_W_LDG
              const float4 w = __ldg(w4 + q);
#else
              const float4 w = w4[q];
#endif
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
    __syncthreads();  // every thread is done reading stage s -> it may be refilled
  }

  const int z = Z0 + tz, y = Y0 + ty, x = X0 + tx * VX;
  if (z >= g.Do || y >= g.Ho || x >= g.Wo) return;
  const int64_t Vo = (int64_t)g.Do * g.Ho * g.Wo;
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
    // W % 4 == 0 on this path: the 4 voxels are all inside and 16-byte aligned
    *reinterpret_cast<float4*>(out + ((int64_t)n * g.Cout + co) * Vo + ((int64_t)z * g.Ho + y) * g.Wo + x) =
        make_float4(r[0], r[1], r[2], r[3]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Forward / dgrad, second register tiling: a thread owns 4 output channels x VT consecutive voxels of one row
// (VT = 16 for 16-channel blocks, 8 for 8-channel blocks).  Per (ci, kz, ky) it reads its input row segment with
// aligned LDS.128 (distinct per lane, conflict-free at a 44-float pitch) and three broadcast LDS.128 of weights:
// 192 FFMA per 36 shared-memory wavefronts, versus 60 for the 4-voxel x 16-channel tiling above (whose broadcast
// weight loads cost a full 4 wavefronts each and held the LSU at 86 % while the FMA pipe sat at 70 %, ncu r11).
// ---------------------------------------------------------------------------------------------------------
constexpr int HXW = 44;  // halo row pitch (floats) of this kernel: X0-4 .. X0+39

template <int CO_BLK>
struct Fwd2Cfg {
  static constexpr int NCG = CO_BLK / 4;            // channel groups of 4 per block
  static constexpr int NVS = 8 / NCG;               // voxel sets (warps per channel group)
  static constexpr int VT = 32 / NVS;               // voxels per thread along x
  static constexpr int SEGS = TX / VT;              // threads per row
  static constexpr int NLD = (VT + 2 + 3 + 3) / 4;  // aligned float4 loads covering halo idx seg*VT+3 .. seg*VT+VT+4
  static constexpr int SX_BYTES = round128(TMA_CK * HZ * HY * HXW * 4);
  static constexpr int SW_BYTES = round128(TMA_CK * 27 * CO_BLK * 4);
  static constexpr int STAGE_BYTES = SX_BYTES + SW_BYTES;
  static constexpr int SMEM_BYTES = 2 * STAGE_BYTES + 128;
  static constexpr uint32_t TX_BYTES = TMA_CK * HZ * HY * HXW * 4 + TMA_CK * 27 * CO_BLK * 4;
};

template <int CO_BLK>
__global__ void __launch_bounds__(TILED_THREADS, 2)
    conv3d_fwd_tma2_kernel(const __grid_constant__ CUtensorMap mx1, const __grid_constant__ CUtensorMap mx2,
                           const float* __restrict__ wp, const float* __restrict__ bias, float* __restrict__ out, ConvGeom g,
                           int tiles_x, int tiles_y, int cin_pad) {
  using Cfg = Fwd2Cfg<CO_BLK>;
  constexpr int VT = Cfg::VT, NLD = Cfg::NLD;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[2];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  const int n = blockIdx.z, cog = blockIdx.y;
  int tb = blockIdx.x;
  const int bx = tb % tiles_x; tb /= tiles_x;
  const int by = tb % tiles_y;
  const int bz = tb / tiles_y;
  const int X0 = bx * TX, Y0 = by * TY, Z0 = bz * TZ;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cg = warp % Cfg::NCG, vs = warp / Cfg::NCG;
  const int seg = lane % Cfg::SEGS;
  const int row = vs * (32 / Cfg::SEGS) + lane / Cfg::SEGS;  // 0..31 = tz*8 + ty
  const int tz = row >> 3, ty = row & 7;
  const int nchunks = cin_pad / TMA_CK;
  const int chunks1 = (g.C1 + TMA_CK - 1) / TMA_CK;

  if (threadIdx.x == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_fence_init();
    tma_prefetch_desc(&mx1);
    if (g.C2) tma_prefetch_desc(&mx2);
  }
  __syncthreads();

  auto issue = [&](int c, int s) {
    uint8_t* st = smem + s * Cfg::STAGE_BYTES;
    mbar_expect_tx(&full[s], Cfg::TX_BYTES);
    if (c < chunks1)
      tma_load_5d(st, &mx1, &full[s], X0 - HX0, Y0 - 1, Z0 - 1, c * TMA_CK, n);
    else
      tma_load_5d(st, &mx2, &full[s], X0 - HX0, Y0 - 1, Z0 - 1, (c - chunks1) * TMA_CK, n);
    bulk_load_1d(st + Cfg::SX_BYTES, wp + ((int64_t)cog * cin_pad + (int64_t)c * TMA_CK) * 27 * CO_BLK, TMA_CK * 27 * CO_BLK * 4,
                 &full[s]);
  };

  float acc[4][VT];
#pragma unroll
  for (int o = 0; o < 4; ++o)
#pragma unroll
    for (int i = 0; i < VT; ++i) acc[o][i] = 0.f;

  if (threadIdx.x == 0) issue(0, 0);
  for (int c = 0; c < nchunks; ++c) {
    const int s = c & 1;
    if (threadIdx.x == 0 && c + 1 < nchunks) issue(c + 1, s ^ 1);
    mbar_wait(&full[s], (c >> 1) & 1);
    const float* sx = reinterpret_cast<const float*>(smem + s * Cfg::STAGE_BYTES);
    const float* sw = reinterpret_cast<const float*>(smem + s * Cfg::STAGE_BYTES + Cfg::SX_BYTES) + cg * 4;
#pragma unroll 1
    for (int cl = 0; cl < TMA_CK; ++cl) {
#pragma unroll
      for (int kz = 0; kz < 3; ++kz) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const float4* rowp = reinterpret_cast<const float4*>(sx + ((cl * HZ + tz + kz) * HY + ty + ky) * HXW + seg * VT);
          float in[NLD * 4];
#pragma unroll
          for (int j = 0; j < NLD; ++j) {
            const float4 v = rowp[j];
            in[4 * j + 0] = v.x; in[4 * j + 1] = v.y; in[4 * j + 2] = v.z; in[4 * j + 3] = v.w;
          }
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float4 w = *reinterpret_cast<const float4*>(sw + (cl * 27 + (kz * 3 + ky) * 3 + kx) * CO_BLK);
#pragma unroll
            for (int i = 0; i < VT; ++i) {
              const float xv = in[i + kx + 3];  // halo idx 3 = voxel x-1
              acc[0][i] = fmaf(xv, w.x, acc[0][i]);
              acc[1][i] = fmaf(xv, w.y, acc[1][i]);
              acc[2][i] = fmaf(xv, w.z, acc[2][i]);
              acc[3][i] = fmaf(xv, w.w, acc[3][i]);
            }
          }
        }
      }
    }
    __syncthreads();
  }

  const int z = Z0 + tz, y = Y0 + ty, x = X0 + seg * VT;
  if (z >= g.Do || y >= g.Ho) return;
  const int64_t Vo = (int64_t)g.Do * g.Ho * g.Wo;
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    const int co = cog * CO_BLK + cg * 4 + o;
    if (co >= g.Cout) break;
    const float bv = bias ? bias[co] : 0.f;
    float* op = out + ((int64_t)n * g.Cout + co) * Vo + ((int64_t)z * g.Ho + y) * g.Wo + x;
#pragma unroll
    for (int q = 0; q < VT / 4; ++q) {
      if (x + 4 * q >= g.Wo) break;  // W % 4 == 0: quads are all-in or all-out
      float r[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        r[i] = acc[o][4 * q + i] + bv;
        if (g.act) r[i] = r[i] > 0.f ? r[i] : r[i] * g.slope;
      }
      *reinterpret_cast<float4*>(op + 4 * q) = make_float4(r[0], r[1], r[2], r[3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// weight gradient, TMA staged (see conv3d_wgrad_tiled_kernel for the mapping)
// ---------------------------------------------------------------------------------------------------------
constexpr int WTM_SX_BYTES = WG_CI * HZ * HY * HXT * 4;  // 38400
constexpr int WTM_SD_BYTES = WG_CO * TZ * TY * TX * 4;   // 32768
constexpr int WTM_STAGE_BYTES = WTM_SX_BYTES + WTM_SD_BYTES;
constexpr int WTM_SMEM_BYTES = 2 * WTM_STAGE_BYTES + 128;
static_assert(WTM_SX_BYTES % 128 == 0 && WTM_SD_BYTES % 128 == 0, "TMA destinations must stay 128-byte aligned");

constexpr int WTM_WARPS = 12, WTM_THREADS = WTM_WARPS * 32;

// Warp w owns (kz = w / 4, ci = w % 4): all nine (ky,kx) taps x 8 output channels = 72 accumulators.  Twelve warps
// are three per SM sub-partition (the nine-warp (kz,ky) split of the cp.async kernel leaves one sub-partition with
// three warps and three with two: 25 % of the FMA issue slots idle at every tile barrier, ncu r11).
__global__ void __launch_bounds__(WTM_THREADS, 1)
    conv3d_wgrad_tma_kernel(const __grid_constant__ CUtensorMap mx, const __grid_constant__ CUtensorMap mdy, WgTiledArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[2];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int kz = warp >> 2, cl = warp & 3;
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

  float acc[96];  // 72 used: [(ky*3+kx)*8 + o]; padded to 96 for the butterfly
#pragma unroll
  for (int i = 0; i < 96; ++i) acc[i] = 0.f;
  float bacc[WG_CO];
#pragma unroll
  for (int o = 0; o < WG_CO; ++o) bacc[o] = 0.f;
  const bool do_bias = a.bias_partials != nullptr && cib == 0 && warp == 4;

  const int t0 = region * a.tiles_per_region;
  const int t1 = min(a.ntiles, t0 + a.tiles_per_region);

  auto issue = [&](int t, int s) {
    int tb = t;
    const int bx = tb % a.tiles_x; tb /= a.tiles_x;
    const int by = tb % a.tiles_y; tb /= a.tiles_y;
    const int bz = tb % a.tiles_z;
    const int n = tb / a.tiles_z;
    uint8_t* st = smem + s * WTM_STAGE_BYTES;
    mbar_expect_tx(&full[s], WTM_STAGE_BYTES);
    tma_load_5d(st, &mx, &full[s], bx * TX - HX0, by * TY - 1, bz * TZ - 1, cib * WG_CI, n);
    tma_load_5d(st + WTM_SX_BYTES, &mdy, &full[s], bx * TX, by * TY, bz * TZ, cob * WG_CO, n);
  };

  if (threadIdx.x == 0 && t0 < t1) issue(t0, 0);
  for (int t = t0; t < t1; ++t) {
    const int k = t - t0, s = k & 1;
    if (threadIdx.x == 0 && t + 1 < t1) issue(t + 1, s ^ 1);
    mbar_wait(&full[s], (k >> 1) & 1);
    const float* sx = reinterpret_cast<const float*>(smem + s * WTM_STAGE_BYTES) + cl * HZ * HY * HXT;
    const float* sd = reinterpret_cast<const float*>(smem + s * WTM_STAGE_BYTES + WTM_SX_BYTES);
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
      for (int ky = 0; ky < 3; ++ky) {
        const float* row = sx + ((tz + kz) * HY + ty + ky) * HXT + tx4 * 4;
        const float4 p = *reinterpret_cast<const float4*>(row);
        const float4 p1 = *reinterpret_cast<const float4*>(row + 4);
        const float4 p2 = *reinterpret_cast<const float4*>(row + 8);
        const float in[6] = {p.w, p1.x, p1.y, p1.z, p1.w, p2.x};
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
          for (int o = 0; o < WG_CO; ++o) {
            float v = acc[(ky * 3 + kx) * WG_CO + o];
            v = fmaf(in[kx + 0], d[o].x, v);
            v = fmaf(in[kx + 1], d[o].y, v);
            v = fmaf(in[kx + 2], d[o].z, v);
            v = fmaf(in[kx + 3], d[o].w, v);
            acc[(ky * 3 + kx) * WG_CO + o] = v;
          }
        }
      }
    }
    __syncthreads();
  }

  butterfly_reduce<96>(acc, lane);
  float* pr = a.partials + (int64_t)region * a.region_stride;
  const int ci = cib * WG_CI + cl;
#pragma unroll
  for (int gi = 0; gi < 3; ++gi) {
    const int e = gi * 32 + lane;
    if (e >= 72) continue;
    const int o = e % WG_CO, kyx = e / WG_CO;
    const int co = cob * WG_CO + o;
    if (co < a.Cout && ci < a.C) pr[((int64_t)(a.co_off + co) * a.Cin_total + a.ci_off + ci) * 27 + kz * 9 + kyx] = acc[gi];
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

// ---------------------------------------------------------------------------------------------------------
// weight gradient of the stride-2 (k3, pad 1) convolutions of the registration encoder (voxel_morph.py:46), TMA staged.
//   dW[co][ci][kz][ky][kx] = sum_o dY[co][o] * X[ci][2*o + k - 1]
// Same warp mapping as conv3d_wgrad_tma_kernel (12 warps = (kz, ci), 72 accumulators = 9 (ky,kx) x 8 co); the output
// tile is 2 x 4 x 32 so that its 5 x 9 x 72 input box (x from 2*X0-4: 16-byte aligned) fits two pipeline stages.
// ---------------------------------------------------------------------------------------------------------
constexpr int S2_TZ = 2, S2_TY = 4, S2_TX = 32;
constexpr int S2_HZ = 2 * S2_TZ + 1, S2_HY = 2 * S2_TY + 1, S2_HX = 72;
constexpr int S2_SX_BYTES = WG_CI * S2_HZ * S2_HY * S2_HX * 4;  // 51840
constexpr int S2_SD_BYTES = WG_CO * S2_TZ * S2_TY * S2_TX * 4;  // 8192
constexpr int S2_STAGE_BYTES = round128(S2_SX_BYTES) + S2_SD_BYTES;
constexpr int S2_SMEM_BYTES = 2 * S2_STAGE_BYTES + 128;

__global__ void __launch_bounds__(WTM_THREADS, 1)
    conv3d_wgrad_s2_tma_kernel(const __grid_constant__ CUtensorMap mx, const __grid_constant__ CUtensorMap mdy, WgTiledArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[2];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int kz = warp >> 2, cl = warp & 3;
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

  float acc[96];
#pragma unroll
  for (int i = 0; i < 96; ++i) acc[i] = 0.f;
  float bacc[WG_CO];
#pragma unroll
  for (int o = 0; o < WG_CO; ++o) bacc[o] = 0.f;
  const bool do_bias = a.bias_partials != nullptr && cib == 0 && warp == 4;

  const int t0 = region * a.tiles_per_region;
  const int t1 = min(a.ntiles, t0 + a.tiles_per_region);

  auto issue = [&](int t, int s) {   // tiles enumerate the OUTPUT volume (a.tiles_* / a.D,H,W are output extents here)
    int tb = t;
    const int bx = tb % a.tiles_x; tb /= a.tiles_x;
    const int by = tb % a.tiles_y; tb /= a.tiles_y;
    const int bz = tb % a.tiles_z;
    const int n = tb / a.tiles_z;
    uint8_t* st = smem + s * S2_STAGE_BYTES;
    mbar_expect_tx(&full[s], S2_SX_BYTES + S2_SD_BYTES);
    tma_load_5d(st, &mx, &full[s], 2 * bx * S2_TX - 4, 2 * by * S2_TY - 1, 2 * bz * S2_TZ - 1, cib * WG_CI, n);
    tma_load_5d(st + round128(S2_SX_BYTES), &mdy, &full[s], bx * S2_TX, by * S2_TY, bz * S2_TZ, cob * WG_CO, n);
  };

  if (threadIdx.x == 0 && t0 < t1) issue(t0, 0);
  for (int t = t0; t < t1; ++t) {
    const int k = t - t0, s = k & 1;
    if (threadIdx.x == 0 && t + 1 < t1) issue(t + 1, s ^ 1);
    mbar_wait(&full[s], (k >> 1) & 1);
    const float* sx = reinterpret_cast<const float*>(smem + s * S2_STAGE_BYTES) + cl * S2_HZ * S2_HY * S2_HX;
    const float* sd = reinterpret_cast<const float*>(smem + s * S2_STAGE_BYTES + round128(S2_SX_BYTES));
#pragma unroll 1
    for (int it = 0; it < (S2_TZ * S2_TY * S2_TX / 4) / 32; ++it) {
      const int q = it * 32 + lane;
      const int tx4 = q & 7, ty = (q >> 3) & 3, tz = q >> 5;
      float4 d[WG_CO];
#pragma unroll
      for (int o = 0; o < WG_CO; ++o)
        d[o] = *reinterpret_cast<const float4*>(sd + ((o * S2_TZ + tz) * S2_TY + ty) * S2_TX + tx4 * 4);
      if (do_bias) {
#pragma unroll
        for (int o = 0; o < WG_CO; ++o) bacc[o] += (d[o].x + d[o].y) + (d[o].z + d[o].w);
      }
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const float4* row = reinterpret_cast<const float4*>(sx + ((2 * tz + kz) * S2_HY + 2 * ty + ky) * S2_HX + tx4 * 8);
        const float4 p0 = row[0], p1 = row[1], p2 = row[2];
        // box x starts at 2*X0-4: input x = 2*(4*tx4+i) + kx - 1  <->  float 8*tx4 + 2*i + kx + 3
        const float in[9] = {p0.w, p1.x, p1.y, p1.z, p1.w, p2.x, p2.y, p2.z, p2.w};
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
          for (int o = 0; o < WG_CO; ++o) {
            float v = acc[(ky * 3 + kx) * WG_CO + o];
            v = fmaf(in[kx + 0], d[o].x, v);
            v = fmaf(in[kx + 2], d[o].y, v);
            v = fmaf(in[kx + 4], d[o].z, v);
            v = fmaf(in[kx + 6], d[o].w, v);
            acc[(ky * 3 + kx) * WG_CO + o] = v;
          }
        }
      }
    }
    __syncthreads();
  }

  butterfly_reduce<96>(acc, lane);
  float* pr = a.partials + (int64_t)region * a.region_stride;
  const int ci = cib * WG_CI + cl;
#pragma unroll
  for (int gi = 0; gi < 3; ++gi) {
    const int e = gi * 32 + lane;
    if (e >= 72) continue;
    const int o = e % WG_CO, kyx = e / WG_CO;
    const int co = cob * WG_CO + o;
    if (co < a.Cout && ci < a.C) pr[((int64_t)(a.co_off + co) * a.Cin_total + a.ci_off + ci) * 27 + kz * 9 + kyx] = acc[gi];
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

// weight repack for the TMA forward kernel: dst[cog][a (cin_pad)][tap'][CO], zero padded in both channel dims
__global__ void repack_weights_tma_kernel(const float* __restrict__ src, float* __restrict__ dst, int d0, int d1, int T,
                                          int a_is_dim0, int flip, int A, int a_off, int Apad, int B, int b_off, int Bpad,
                                          int CO) {
  const int64_t total = (int64_t)Apad * T * Bpad;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int bl = (int)(i % CO);
    int64_t r = i / CO;
    const int t = (int)(r % T); r /= T;
    const int a = (int)(r % Apad);
    const int cog = (int)(r / Apad);
    const int b = cog * CO + bl;
    float v = 0.f;
    if (a < A && b < B) {
      const int ts = flip ? (T - 1 - t) : t;
      const int i0 = a_is_dim0 ? (a + a_off) : (b + b_off);
      const int i1 = a_is_dim0 ? (b + b_off) : (a + a_off);
      v = src[((int64_t)i0 * d1 + i1) * T + ts];
    }
    dst[i] = v;
  }
}
