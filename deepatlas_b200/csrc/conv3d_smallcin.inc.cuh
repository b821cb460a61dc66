// Weight gradient of the k3 s1 p1 input layers (1 -> 8 of the segmentation net, cat(1, 1) -> 16 of the registration
// net, unets.py:225 / voxel_morph.py:39): with Cin <= 4 a tensor-core tile carries 3..12 useful rows of 128, and the
// arithmetic (27 * Cin * Cout FMA per voxel) is small against reading dY once.  Exact FFMA: warp = (kz, ci) of one block
// of 8 output channels, lane = x position of a 32-voxel row segment, 9 (ky, kx) x 8 co accumulators per lane resident for
// the block's whole walk over its units (32 x positions times 32 rows of one plane, dealt round-robin so that concurrently
// running blocks read neighbouring x chunks of the same rows), lanes folded once at the end, one partial row per block,
// fixed-order reduce.
// Included by conv3d.cu.

constexpr int SC_COB = 8;
constexpr int SC_MAX_CIN = 4;
constexpr int SC_REGIONS = 4 * DA_NUM_SMS;
constexpr int SC_YSEG = 32;    // rows per unit

// grid (regions, ceil(Cout / 8)), block 96 * Cin threads.  partials [region][Cout][Cin][27]; bias_partials nullable
// [region][Cout]
__global__ void __launch_bounds__(96 * SC_MAX_CIN) conv3d_wgrad_smallcin_kernel(const float* __restrict__ x1, const float* __restrict__ x2,
                                                                               int C1, int C2, const float* __restrict__ dy,
                                                                               float* __restrict__ partials, float* __restrict__ bias_partials,
                                                                               int N, int D, int H, int W, int Cout, int64_t region_stride) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int kz = warp % 3, ci = warp / 3, Cin = C1 + C2;
  const int co0 = blockIdx.y * SC_COB;
  const int64_t V = (int64_t)D * H * W;
  const bool from1 = ci < C1;
  const float* xs = from1 ? x1 + (int64_t)ci * V : x2 + (int64_t)(ci - C1) * V;
  const int64_t xn_stride = (int64_t)(from1 ? C1 : C2) * V;
  const bool do_bias = bias_partials != nullptr && warp == 1;   // (kz = 1, ci = 0)
  float acc[9][SC_COB], bacc[SC_COB];
#pragma unroll
  for (int o = 0; o < SC_COB; ++o) {
    bacc[o] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t][o] = 0.f;
  }
  // unit = 32 x positions times a segment of SC_YSEG rows of one (n, z) plane, walked down y: the three input rows of
  // a step are kept in registers and rotated (one new row per step instead of three), and the loads of step y + 1 are
  // issued before the arithmetic of step y.  The first version (one row per unit: 17 loads behind three branches, four
  // divisions and 64-bit address arithmetic for 72 FMA) ran at 4.4k cycles per unit and warp.
  const int nxc = (W + 31) / 32, nys = (H + SC_YSEG - 1) / SC_YSEG;
  const int units = N * D * nys * nxc;   // < 2^31 (checked by the host)
  for (int u = blockIdx.x; u < units; u += gridDim.x) {
    const int xc = u % nxc;
    int r = u / nxc;
    const int ys = r % nys; r /= nys;
    const int z = r % D;
    const int n = r / D;
    const int zz = z + kz - 1;
    if (zz < 0 || zz >= D) continue;   // warp-uniform; the bias warp has kz = 1
    const int xx = xc * 32 + lane;
    const bool live = xx < W, lm = live && xx >= 1, lp = xx + 1 < W;
    const int y0 = ys * SC_YSEG, y1 = min(H, y0 + SC_YSEG);
    const float* pg = dy + ((int64_t)n * Cout + co0) * V + ((int64_t)z * H + y0) * W + xx;
    const float* px = xs + (int64_t)n * xn_stride + ((int64_t)zz * H + y0) * W + xx;
    float xr[3][3], g[SC_COB], gn[SC_COB], xn[3];
    auto load_row = [&](const float* row, bool ok, float (&q)[3]) {
      q[0] = (ok && lm) ? __ldg(row - 1) : 0.f;
      q[1] = (ok && live) ? __ldg(row) : 0.f;
      q[2] = (ok && lp) ? __ldg(row + 1) : 0.f;
    };
    auto load_g = [&](const float* p, bool ok, float (&q)[SC_COB]) {
#pragma unroll
      for (int o = 0; o < SC_COB; ++o) q[o] = (ok && live && co0 + o < Cout) ? __ldg(p + (int64_t)o * V) : 0.f;
    };
    load_row(px - W, y0 >= 1, xr[0]);
    load_row(px, true, xr[1]);
    load_row(px + W, y0 + 1 < H, xr[2]);
    load_g(pg, true, g);
    for (int y = y0; y < y1; ++y) {
      pg += W;
      px += W;
      load_g(pg, y + 1 < y1, gn);             // next step's operands
      load_row(px + W, y + 2 < H && y + 1 < y1, xn);
      if (do_bias) {
#pragma unroll
        for (int o = 0; o < SC_COB; ++o) bacc[o] += g[o];
      }
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int o = 0; o < SC_COB; ++o) {
          acc[ky * 3 + 0][o] = fmaf(xr[ky][0], g[o], acc[ky * 3 + 0][o]);
          acc[ky * 3 + 1][o] = fmaf(xr[ky][1], g[o], acc[ky * 3 + 1][o]);
          acc[ky * 3 + 2][o] = fmaf(xr[ky][2], g[o], acc[ky * 3 + 2][o]);
        }
#pragma unroll
      for (int k = 0; k < 3; ++k) { xr[0][k] = xr[1][k]; xr[1][k] = xr[2][k]; xr[2][k] = xn[k]; }
#pragma unroll
      for (int o = 0; o < SC_COB; ++o) g[o] = gn[o];
    }
  }
  float* pr = partials + (int64_t)blockIdx.x * region_stride;
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int o = 0; o < SC_COB; ++o) {
      const float s = warp_sum(acc[t][o]);
      if (lane == 0 && co0 + o < Cout) pr[((int64_t)(co0 + o) * Cin + ci) * 27 + kz * 9 + t] = s;
    }
  if (do_bias) {
#pragma unroll
    for (int o = 0; o < SC_COB; ++o) {
      const float s = warp_sum(bacc[o]);
      if (lane == 0 && co0 + o < Cout) bias_partials[(int64_t)blockIdx.x * Cout + co0 + o] = s;
    }
  }
}
