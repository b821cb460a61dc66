// Weight gradient of the k3 s1 p1 input layers (1 -> 8 of the segmentation net, cat(1, 1) -> 16 of the registration
// net, unets.py:225 / voxel_morph.py:39): with Cin <= 4 a tensor-core tile carries 3..12 useful rows of 128, and the
// arithmetic (27 * Cin * Cout FMA per voxel) is small against reading dY once.  Exact FFMA: warp = (kz, ci) of one block
// of 8 output channels, lane = x position of a 32-voxel row segment, 9 (ky, kx) x 8 co accumulators per lane resident for
// the block's whole walk over (n, z, y, x-chunk) units (dealt round-robin, so that concurrently running blocks read
// neighbouring rows), lanes folded once at the end, one partial row per block, fixed-order reduce.
// Included by conv3d.cu.

constexpr int SC_COB = 8;
constexpr int SC_MAX_CIN = 4;
constexpr int SC_REGIONS = 4 * DA_NUM_SMS;

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
  const int nxc = (W + 31) / 32;
  const int units = N * D * H * nxc;   // < 2^31 (checked by the host)
  // One unit's operands: 8 dY values + the 3 x 3 input values of this warp's kz plane.  All 17 loads are issued
  // back to back (predicated, no branches) and two units ahead of the arithmetic: with branches per ky the loads of a unit
  // formed four dependent round trips to memory (measured: 4.4k cycles per unit and warp).
  struct Ops { float g[SC_COB]; float x[3][3]; };
  auto load_unit = [&](int u, Ops& q) {
    const int xc = u % nxc;
    int r = u / nxc;
    const int y = r % H; r /= H;
    const int z = r % D;
    const int n = r / D;
    const int xx = xc * 32 + lane;
    const bool live = u < units && xx < W;
    const float* pg = dy + ((int64_t)n * Cout + co0) * V + ((int64_t)z * H + y) * W + xx;
#pragma unroll
    for (int o = 0; o < SC_COB; ++o) q.g[o] = (live && co0 + o < Cout) ? __ldg(pg + (int64_t)o * V) : 0.f;
    const int zz = z + kz - 1;
    const bool zok = live && zz >= 0 && zz < D;
    const float* pz = xs + (int64_t)n * xn_stride + (int64_t)zz * H * W;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y + ky - 1;
      const bool ok = zok && yy >= 0 && yy < H;
      const float* row = pz + (int64_t)yy * W + xx;
      q.x[ky][0] = (ok && xx >= 1) ? __ldg(row - 1) : 0.f;
      q.x[ky][1] = ok ? __ldg(row) : 0.f;
      q.x[ky][2] = (ok && xx + 1 < W) ? __ldg(row + 1) : 0.f;
    }
  };
  auto fma_unit = [&](const Ops& q) {
    if (do_bias) {
#pragma unroll
      for (int o = 0; o < SC_COB; ++o) bacc[o] += q.g[o];
    }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int o = 0; o < SC_COB; ++o) {
        acc[ky * 3 + 0][o] = fmaf(q.x[ky][0], q.g[o], acc[ky * 3 + 0][o]);
        acc[ky * 3 + 1][o] = fmaf(q.x[ky][1], q.g[o], acc[ky * 3 + 1][o]);
        acc[ky * 3 + 2][o] = fmaf(q.x[ky][2], q.g[o], acc[ky * 3 + 2][o]);
      }
  };
  // three rotating operand sets: the loads run two units ahead (a unit past the end loads nothing and adds zeros)
  Ops qa, qb, qc;
  const int S = gridDim.x;
  load_unit(blockIdx.x, qa);
  load_unit(blockIdx.x + S, qb);
  for (int u = blockIdx.x; u < units; u += 3 * S) {
    load_unit(u + 2 * S, qc);
    fma_unit(qa);
    load_unit(u + 3 * S, qa);
    fma_unit(qb);
    load_unit(u + 4 * S, qb);
    fma_unit(qc);
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
