// Kernels for the UNet_generator variants and the device-side input stage (row 4):
//   * trilinear x2 up-sampling, nn.Upsample(scale_factor=2, mode='trilinear') (lib/network_factory/unets.py:236)
//   * residual add with channel broadcast, `enc(x) + x` / `dec(...) + x` (unets.py:264,275) and its channel reduction
//   * crop + clip of the image and crop of the uint8 label map (lib/transforms.py:79-80 SitkToTensor clip to [0,1],
//     :124-158 CropTensor), done on the device after one H2D copy of the raw volume.
// (The stride-2 k2 down-sampling convolution of maxpool=False, unets.py:231, needs no kernel of its own: it is the
// adjoint of the k2 s2 deconvolution, see deepatlas_b200/ops.py:ConvK2S2Function.)  All HBM-bound, planar NCDHW.
#include "common.cuh"

namespace {

inline int ew_grid(int64_t total, int per_block = 256) {
  int64_t b = da_cdiv(total, per_block);
  const int64_t cap = (int64_t)DA_NUM_SMS * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// PyTorch's area_pixel_compute_source_index for align_corners=False with the given scale factor 2 (scale = 0.5):
// src = 0.5*(dst + 0.5) - 0.5, clamped at 0; i0 = floor(src), i1 = min(i0 + 1, n - 1), l1 = src - i0, l0 = 1 - l1.
__device__ __forceinline__ void tri_src(int dst, int n, int& i0, int& i1, float& l0, float& l1) {
  float s = 0.5f * ((float)dst + 0.5f) - 0.5f;
  s = s < 0.f ? 0.f : s;
  i0 = (int)s;
  i1 = i0 + (i0 < n - 1 ? 1 : 0);
  l1 = s - (float)i0;
  l0 = 1.f - l1;
}

__global__ void __launch_bounds__(256) upsample_tri2_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t NC,
                                                                int D, int H, int W) {
  const int Do = 2 * D, Ho = 2 * H, Wo = 2 * W;
  const int64_t total = NC * Do * Ho * Wo;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int xo = (int)(i % Wo);
    int64_t t = i / Wo;
    const int yo = (int)(t % Ho); t /= Ho;
    const int zo = (int)(t % Do);
    const int64_t nc = t / Do;
    int z0, z1, y0, y1, x0, x1;
    float a0, a1, b0, b1, c0, c1;
    tri_src(zo, D, z0, z1, a0, a1);
    tri_src(yo, H, y0, y1, b0, b1);
    tri_src(xo, W, x0, x1, c0, c1);
    const float* p = x + nc * D * H * W;
    auto at = [&](int z, int yy, int xx) { return __ldg(p + ((int64_t)z * H + yy) * W + xx); };
    // summation order of ATen's upsample_trilinear3d
    y[i] = a0 * (b0 * (c0 * at(z0, y0, x0) + c1 * at(z0, y0, x1)) + b1 * (c0 * at(z0, y1, x0) + c1 * at(z0, y1, x1))) +
           a1 * (b0 * (c0 * at(z1, y0, x0) + c1 * at(z1, y0, x1)) + b1 * (c0 * at(z1, y1, x0) + c1 * at(z1, y1, x1)));
  }
}

// weight with which output index dst (extent 2n) reads input index k; 0 outside the range
__device__ __forceinline__ float tri_w(int dst, int k, int n) {
  if (dst < 0 || dst >= 2 * n) return 0.f;
  int i0, i1;
  float l0, l1;
  tri_src(dst, n, i0, i1, l0, l1);
  return (i0 == k ? l0 : 0.f) + (i1 == k ? l1 : 0.f);
}

// gather form of the backward: every input voxel collects from the <= 4x4x4 outputs that read it (deterministic)
__global__ void __launch_bounds__(256) upsample_tri2_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int64_t NC,
                                                                int D, int H, int W) {
  const int Ho = 2 * H, Wo = 2 * W;
  const int64_t total = NC * D * H * W;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int xs = (int)(i % W);
    int64_t t = i / W;
    const int ys = (int)(t % H); t /= H;
    const int zs = (int)(t % D);
    const int64_t nc = t / D;
    float wz[4], wy[4], wx[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      wz[k] = tri_w(2 * zs - 1 + k, zs, D);
      wy[k] = tri_w(2 * ys - 1 + k, ys, H);
      wx[k] = tri_w(2 * xs - 1 + k, xs, W);
    }
    const float* p = dy + nc * 8 * D * H * W;
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      if (wz[a] == 0.f) continue;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        if (wy[b] == 0.f) continue;
        const float* row = p + ((int64_t)(2 * zs - 1 + a) * Ho + (2 * ys - 1 + b)) * Wo + (2 * xs - 1);
        float r = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (wx[c] != 0.f) r += wx[c] * __ldg(row + c);
        acc += wz[a] * wy[b] * r;
      }
    }
    dx[i] = acc;
  }
}

// out[n][c][v] = a[n][c][v] + b[n][Cb == 1 ? 0 : c][v]
__global__ void __launch_bounds__(256) add_bcast_kernel(const float* __restrict__ a, const float* __restrict__ b, int N, int Ca,
                                                        int Cb, int64_t V, float* __restrict__ out) {
  const int64_t total = (int64_t)N * Ca * V;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    int64_t j = i;
    if (Cb != Ca) {
      const int64_t n = i / (Ca * V), v = i % V;
      j = n * V + v;
    }
    out[i] = __ldg(a + i) + __ldg(b + j);
  }
}

// out[n][v] = sum_c g[n][c][v]   (gradient of the broadcast operand)
__global__ void __launch_bounds__(256) channel_reduce_kernel(const float* __restrict__ g, int N, int C, int64_t V,
                                                             float* __restrict__ out) {
  const int64_t total = (int64_t)N * V;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t n = i / V, v = i - n * V;
    const float* p = g + n * C * V + v;
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc += __ldg(p + (int64_t)c * V);
    out[i] = acc;
  }
}

struct CropGeo { int D, H, W, z0, y0, x0, Do, Ho, Wo; };

template <typename T, bool CLIP>
__global__ void __launch_bounds__(256) crop_kernel(const T* __restrict__ src, T* __restrict__ dst, int64_t NC, CropGeo g, float lo,
                                                   float hi) {
  const int64_t total = NC * g.Do * g.Ho * g.Wo;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int xo = (int)(i % g.Wo);
    int64_t t = i / g.Wo;
    const int yo = (int)(t % g.Ho); t /= g.Ho;
    const int zo = (int)(t % g.Do);
    const int64_t nc = t / g.Do;
    T v = src[((nc * g.D + zo + g.z0) * g.H + yo + g.y0) * g.W + xo + g.x0];
    if (CLIP) {
      // img_np[img_np > 1] = 1; img_np[img_np < 0] = 0 (lib/transforms.py:79-80): NaN stays NaN, as there
      float f = (float)v;
      f = f > hi ? hi : f;
      f = f < lo ? lo : f;
      v = (T)f;
    }
    dst[i] = v;
  }
}

int crop_check(const char* who, int D, int H, int W, int z0, int y0, int x0, int Do, int Ho, int Wo) {
  DA_REQUIRE(z0 >= 0 && y0 >= 0 && x0 >= 0 && Do >= 1 && Ho >= 1 && Wo >= 1 && z0 + Do <= D && y0 + Ho <= H && x0 + Wo <= W,
             "%s: crop window (%d,%d,%d)+(%d,%d,%d) leaves the volume (%d,%d,%d)", who, z0, y0, x0, Do, Ho, Wo, D, H, W);
  return DA_OK;
}

}  // namespace

// x [NC,D,H,W] -> y [NC,2D,2H,2W], trilinear, align_corners=False
DA_API int da_upsample_trilinear2_fwd(const float* x, float* y, int64_t NC, int D, int H, int W, cudaStream_t stream) {
  DA_REQUIRE(x && y, "da_upsample_trilinear2_fwd: null pointer");
  DA_REQUIRE(NC >= 1 && D >= 1 && H >= 1 && W >= 1, "da_upsample_trilinear2_fwd: bad extents");
  upsample_tri2_fwd_kernel<<<ew_grid(NC * 8 * D * H * W), 256, 0, stream>>>(x, y, NC, D, H, W);
  return da_check_launch("da_upsample_trilinear2_fwd");
}

// dy [NC,2D,2H,2W] -> dx [NC,D,H,W]
DA_API int da_upsample_trilinear2_bwd(const float* dy, float* dx, int64_t NC, int D, int H, int W, cudaStream_t stream) {
  DA_REQUIRE(dy && dx, "da_upsample_trilinear2_bwd: null pointer");
  DA_REQUIRE(NC >= 1 && D >= 1 && H >= 1 && W >= 1, "da_upsample_trilinear2_bwd: bad extents");
  upsample_tri2_bwd_kernel<<<ew_grid(NC * D * H * W), 256, 0, stream>>>(dy, dx, NC, D, H, W);
  return da_check_launch("da_upsample_trilinear2_bwd");
}

// out [N,Ca,V] = a [N,Ca,V] + b [N,Cb,V] with Cb == Ca or Cb == 1 (broadcast over channels)
DA_API int da_add_bcast(const float* a, const float* b, int N, int Ca, int Cb, int64_t V, float* out, cudaStream_t stream) {
  DA_REQUIRE(a && b && out, "da_add_bcast: null pointer");
  DA_REQUIRE(Cb == Ca || Cb == 1, "da_add_bcast: channel counts %d and %d do not broadcast", Ca, Cb);
  add_bcast_kernel<<<ew_grid((int64_t)N * Ca * V), 256, 0, stream>>>(a, b, N, Ca, Cb, V, out);
  return da_check_launch("da_add_bcast");
}

// out [N,1,V] = sum over channels of g [N,C,V]
DA_API int da_channel_reduce(const float* g, int N, int C, int64_t V, float* out, cudaStream_t stream) {
  DA_REQUIRE(g && out, "da_channel_reduce: null pointer");
  channel_reduce_kernel<<<ew_grid((int64_t)N * V), 256, 0, stream>>>(g, N, C, V, out);
  return da_check_launch("da_channel_reduce");
}

// src [NC,D,H,W] fp32 -> dst [NC,Do,Ho,Wo] = clip(src[z0:z0+Do, y0:y0+Ho, x0:x0+Wo], lo, hi)
DA_API int da_crop_clip_f32(const float* src, float* dst, int64_t NC, int D, int H, int W, int z0, int y0, int x0, int Do, int Ho,
                            int Wo, float lo, float hi, cudaStream_t stream) {
  DA_REQUIRE(src && dst, "da_crop_clip_f32: null pointer");
  int rc = crop_check("da_crop_clip_f32", D, H, W, z0, y0, x0, Do, Ho, Wo);
  if (rc) return rc;
  CropGeo g{D, H, W, z0, y0, x0, Do, Ho, Wo};
  crop_kernel<float, true><<<ew_grid(NC * Do * Ho * Wo), 256, 0, stream>>>(src, dst, NC, g, lo, hi);
  return da_check_launch("da_crop_clip_f32");
}

// uint8 label maps: src [NC,D,H,W] -> dst [NC,Do,Ho,Wo]
DA_API int da_crop_u8(const uint8_t* src, uint8_t* dst, int64_t NC, int D, int H, int W, int z0, int y0, int x0, int Do, int Ho,
                      int Wo, cudaStream_t stream) {
  DA_REQUIRE(src && dst, "da_crop_u8: null pointer");
  int rc = crop_check("da_crop_u8", D, H, W, z0, y0, x0, Do, Ho, Wo);
  if (rc) return rc;
  CropGeo g{D, H, W, z0, y0, x0, Do, Ho, Wo};
  crop_kernel<uint8_t, false><<<ew_grid(NC * Do * Ho * Wo), 256, 0, stream>>>(src, dst, NC, g, 0.f, 0.f);
  return da_check_launch("da_crop_u8");
}
