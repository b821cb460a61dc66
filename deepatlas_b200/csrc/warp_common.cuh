// Shared device helpers of the warp kernels: identity grid, un-normalisation and the 8 trilinear corners with the
// reference's (ATen grid_sampler, align_corners=True, zeros padding) conventions.  See warp3d.cu for the citations.
#pragma once
#include "common.cuh"

namespace {

struct WarpGeom {
  int N, C, D, H, W;      // source extent
  int Do, Ho, Wo;         // output / field extent
};

__device__ __forceinline__ float ident_coord(int k, int n) {
  // torch.arange(0,n).float() / (n-1) * 2.0 - 1   -- every op rounded to fp32, no contraction
  return __fadd_rn(__fmul_rn(__fdiv_rn((float)k, (float)(n - 1)), 2.0f), -1.0f);
}
__device__ __forceinline__ float unnormalize(float g, int size) {
  // ATen grid_sampler_unnormalize, align_corners=True: ((g + 1) / 2) * (size - 1)
  return __fmul_rn(__fdiv_rn(__fadd_rn(g, 1.0f), 2.0f), (float)(size - 1));
}

template <bool ADD_ID>
__device__ __forceinline__ void load_phi(const float* __restrict__ field, int64_t Vo, int64_t v, int x,
                                         int y, int z, const WarpGeom& g, float& px, float& py,
                                         float& pz) {
  px = field[v];
  py = field[Vo + v];
  pz = field[2 * Vo + v];
  if (ADD_ID) {
    px = __fadd_rn(px, ident_coord(x, g.Wo));
    py = __fadd_rn(py, ident_coord(y, g.Ho));
    pz = __fadd_rn(pz, ident_coord(z, g.Do));
  }
}

struct Corners {
  int64_t off[8];
  float w[8];
  bool ok[8];
  float fx[2], fy[2], fz[2];
};

__device__ __forceinline__ void make_corners(float ix, float iy, float iz, const WarpGeom& g, Corners& c) {
  const float x0f = floorf(ix), y0f = floorf(iy), z0f = floorf(iz);
  const int x0 = (int)x0f, y0 = (int)y0f, z0 = (int)z0f;
  // weight of the "0" corner is (x1 - ix); of the "1" corner (ix - x0)  (ATen GridSampler naming:
  // tnw = (ix_bse-ix)(iy_bse-iy)(iz_bse-iz), ...)
  c.fx[0] = (x0f + 1.0f) - ix; c.fx[1] = ix - x0f;
  c.fy[0] = (y0f + 1.0f) - iy; c.fy[1] = iy - y0f;
  c.fz[0] = (z0f + 1.0f) - iz; c.fz[1] = iz - z0f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {  // order tnw,tne,tsw,tse,bnw,bne,bsw,bse = (dz,dy,dx) binary count
    const int dx = k & 1, dy = (k >> 1) & 1, dz = k >> 2;
    const int xi = x0 + dx, yi = y0 + dy, zi = z0 + dz;
    c.ok[k] = (xi >= 0) & (xi < g.W) & (yi >= 0) & (yi < g.H) & (zi >= 0) & (zi < g.D);
    c.off[k] = ((int64_t)zi * g.H + yi) * g.W + xi;
    c.w[k] = c.fx[dx] * c.fy[dy] * c.fz[dz];
  }
}


}  // namespace
