// warp3d: the deformation-field spatial transformer of the registration branch.
//
// Replaces, in ONE pass each way, the reference sequence (lib/network_factory/voxel_morph.py:85-91)
//   id    = get_identity_transform(...)                (lib/utils.py:89-102, materialised 3xDxHxW)
//   phi   = disp + id                                  (:88)
//   out   = F.grid_sample(src, phi.permute(0,2,3,4,1), 'bilinear', 'zeros', align_corners=True)  (:90-91)
// The identity grid is recomputed from the voxel index with the reference's own fp32 rounding
// sequence (k/(n-1)*2-1, each step rounded), never stored unless the caller asks for phi.
// Layout: planar NCDHW fp32 (the reference's), one thread per output voxel looping over channels so
// the 8 corner offsets / weights are computed once; consecutive threads are consecutive along W.
// HBM-bound: image warp fwd+bwd 13*V*4 B, C-channel (5C+9)*V*4 B.
#include "common.cuh"
#include "warp_common.cuh"

namespace {

template <bool ADD_ID>
__global__ void __launch_bounds__(256) warp3d_fwd_kernel(const float* __restrict__ src,
                                                         const float* __restrict__ field,
                                                         float* __restrict__ out,
                                                         float* __restrict__ phi_out, WarpGeom g) {
  const int64_t Vo = (int64_t)g.Do * g.Ho * g.Wo, Vs = (int64_t)g.D * g.H * g.W;
  const int64_t total = (int64_t)g.N * Vo;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / Vo);
    const int64_t v = i - (int64_t)n * Vo;
    const int x = (int)(v % g.Wo), y = (int)((v / g.Wo) % g.Ho), z = (int)(v / ((int64_t)g.Wo * g.Ho));
    float px, py, pz;
    load_phi<ADD_ID>(field + (int64_t)n * 3 * Vo, Vo, v, x, y, z, g, px, py, pz);
    if (phi_out) {
      float* p = phi_out + (int64_t)n * 3 * Vo;
      p[v] = px; p[Vo + v] = py; p[2 * Vo + v] = pz;
    }
    Corners c;
    make_corners(unnormalize(px, g.W), unnormalize(py, g.H), unnormalize(pz, g.D), g, c);
    const float* s = src + (int64_t)n * g.C * Vs;
    float* o = out + (int64_t)n * g.C * Vo + v;
    for (int ch = 0; ch < g.C; ++ch, s += Vs, o += Vo) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (c.ok[k]) acc += __ldg(s + c.off[k]) * c.w[k];
      *o = acc;
    }
  }
}

// grad_src must be zero-filled by the caller-facing entry point (done below with a memset).
template <bool ADD_ID>
__global__ void __launch_bounds__(256) warp3d_bwd_kernel(const float* __restrict__ gout,
                                                         const float* __restrict__ src,
                                                         const float* __restrict__ field,
                                                         float* __restrict__ gsrc,
                                                         float* __restrict__ gfield, WarpGeom g) {
  const int64_t Vo = (int64_t)g.Do * g.Ho * g.Wo, Vs = (int64_t)g.D * g.H * g.W;
  const int64_t total = (int64_t)g.N * Vo;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / Vo);
    const int64_t v = i - (int64_t)n * Vo;
    const int x = (int)(v % g.Wo), y = (int)((v / g.Wo) % g.Ho), z = (int)(v / ((int64_t)g.Wo * g.Ho));
    float px, py, pz;
    load_phi<ADD_ID>(field + (int64_t)n * 3 * Vo, Vo, v, x, y, z, g, px, py, pz);
    Corners c;
    make_corners(unnormalize(px, g.W), unnormalize(py, g.H), unnormalize(pz, g.D), g, c);
    const float* s = src + (int64_t)n * g.C * Vs;
    float* gs = gsrc ? gsrc + (int64_t)n * g.C * Vs : nullptr;
    const float* go = gout + (int64_t)n * g.C * Vo + v;
    float gx = 0.f, gy = 0.f, gz = 0.f;
    for (int ch = 0; ch < g.C; ++ch, s += Vs, go += Vo) {
      const float gO = *go;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (!c.ok[k]) continue;
        const int dx = k & 1, dy = (k >> 1) & 1, dz = k >> 2;
        if (gs) atomicAdd(gs + (int64_t)ch * Vs + c.off[k], c.w[k] * gO);
        if (gfield) {
          const float val = __ldg(s + c.off[k]) * gO;
          gx += (dx ? val : -val) * c.fy[dy] * c.fz[dz];
          gy += (dy ? val : -val) * c.fx[dx] * c.fz[dz];
          gz += (dz ? val : -val) * c.fx[dx] * c.fy[dy];
        }
      }
    }
    if (gfield) {
      float* gf = gfield + (int64_t)n * 3 * Vo;
      gf[v] = gx * (0.5f * (float)(g.W - 1));
      gf[Vo + v] = gy * (0.5f * (float)(g.H - 1));
      gf[2 * Vo + v] = gz * (0.5f * (float)(g.D - 1));
    }
  }
}

// Channels-last scatter for multi-channel sources (the 32-class probability warp of the anatomy loss):
// grad_src_cl is [N][Vs][C] so that the C contributions to one source voxel are contiguous and four of them go out
// in ONE vector reduction (red.global.add.v4.f32, sm_90+): 8*C/4 reductions per output voxel instead of 8*C.
// The caller hands the buffer back to autograd as a permuted (N,C,D,H,W) view.
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <bool ADD_ID>
__global__ void __launch_bounds__(256) warp3d_bwd_cl_kernel(const float* __restrict__ gout, const float* __restrict__ src,
                                                            const float* __restrict__ field, float* __restrict__ gsrc_cl,
                                                            float* __restrict__ gfield, WarpGeom g) {
  const int64_t Vo = (int64_t)g.Do * g.Ho * g.Wo, Vs = (int64_t)g.D * g.H * g.W;
  const int64_t total = (int64_t)g.N * Vo;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / Vo);
    const int64_t v = i - (int64_t)n * Vo;
    const int x = (int)(v % g.Wo), y = (int)((v / g.Wo) % g.Ho), z = (int)(v / ((int64_t)g.Wo * g.Ho));
    float px, py, pz;
    load_phi<ADD_ID>(field + (int64_t)n * 3 * Vo, Vo, v, x, y, z, g, px, py, pz);
    Corners c;
    make_corners(unnormalize(px, g.W), unnormalize(py, g.H), unnormalize(pz, g.D), g, c);
    const float* s = src + (int64_t)n * g.C * Vs;
    float* gs = gsrc_cl + (int64_t)n * Vs * g.C;
    const float* go = gout + (int64_t)n * g.C * Vo + v;
    float gx = 0.f, gy = 0.f, gz = 0.f;
    for (int ch = 0; ch < g.C; ch += 4, go += 4 * Vo) {
      const float g0 = go[0], g1 = go[Vo], g2 = go[2 * Vo], g3 = go[3 * Vo];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (!c.ok[k]) continue;
        const int dx = k & 1, dy = (k >> 1) & 1, dz = k >> 2;
        const float w = c.w[k];
        red_add_v4(gs + c.off[k] * g.C + ch, w * g0, w * g1, w * g2, w * g3);
        if (gfield) {
          const float* sp = s + (int64_t)ch * Vs + c.off[k];
          const float val = __ldg(sp) * g0 + __ldg(sp + Vs) * g1 + __ldg(sp + 2 * Vs) * g2 + __ldg(sp + 3 * Vs) * g3;
          gx += (dx ? val : -val) * c.fy[dy] * c.fz[dz];
          gy += (dy ? val : -val) * c.fx[dx] * c.fz[dz];
          gz += (dz ? val : -val) * c.fx[dx] * c.fy[dy];
        }
      }
    }
    if (gfield) {
      float* gf = gfield + (int64_t)n * 3 * Vo;
      gf[v] = gx * (0.5f * (float)(g.W - 1));
      gf[Vo + v] = gy * (0.5f * (float)(g.H - 1));
      gf[2 * Vo + v] = gz * (0.5f * (float)(g.D - 1));
    }
  }
}

inline int grid_for(int64_t total, int threads) {
  int64_t b = da_cdiv(total, threads);
  const int64_t cap = (int64_t)DA_NUM_SMS * 16;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace

// src [N,C,D,H,W]; field [N,3,Do,Ho,Wo] (channel 0 = x/W, 1 = y/H, 2 = z/D, normalised [-1,1]);
// out [N,C,Do,Ho,Wo]; phi_out (nullable) [N,3,Do,Ho,Wo] receives field (+ identity).
DA_API int da_warp3d_fwd(const float* src, const float* field, int add_identity, float* out,
                         float* phi_out, int N, int C, int D, int H, int W, int Do, int Ho, int Wo,
                         cudaStream_t stream) {
  DA_REQUIRE(src && field && out, "da_warp3d_fwd: null pointer");
  DA_REQUIRE(N > 0 && C > 0 && D > 0 && H > 0 && W > 0 && Do > 0 && Ho > 0 && Wo > 0,
             "da_warp3d_fwd: bad extent");
  WarpGeom g{N, C, D, H, W, Do, Ho, Wo};
  const int64_t total = (int64_t)N * Do * Ho * Wo;
  const int grid = grid_for(total, 256);
  if (add_identity)
    warp3d_fwd_kernel<true><<<grid, 256, 0, stream>>>(src, field, out, phi_out, g);
  else
    warp3d_fwd_kernel<false><<<grid, 256, 0, stream>>>(src, field, out, phi_out, g);
  return da_check_launch("da_warp3d_fwd");
}

// grad_src (nullable) [N,C,D,H,W] is zeroed here then scatter-added (fp32 atomics: the only
// non-deterministic reduction in the library, as in ATen's grid_sampler_3d_backward);
// grad_field (nullable) [N,3,Do,Ho,Wo].
DA_API int da_warp3d_bwd(const float* grad_out, const float* src, const float* field, int add_identity,
                         float* grad_src, float* grad_field, int N, int C, int D, int H, int W, int Do,
                         int Ho, int Wo, cudaStream_t stream) {
  DA_REQUIRE(grad_out && src && field, "da_warp3d_bwd: null pointer");
  DA_REQUIRE(grad_src || grad_field, "da_warp3d_bwd: nothing to compute");
  WarpGeom g{N, C, D, H, W, Do, Ho, Wo};
  if (grad_src) {
    cudaError_t e = cudaMemsetAsync(grad_src, 0, sizeof(float) * (size_t)N * C * D * H * W, stream);
    if (e != cudaSuccess) { da_set_error("da_warp3d_bwd memset: %s", cudaGetErrorString(e)); return (int)e; }
  }
  const int64_t total = (int64_t)N * Do * Ho * Wo;
  const int grid = grid_for(total, 256);
  if (add_identity)
    warp3d_bwd_kernel<true><<<grid, 256, 0, stream>>>(grad_out, src, field, grad_src, grad_field, g);
  else
    warp3d_bwd_kernel<false><<<grid, 256, 0, stream>>>(grad_out, src, field, grad_src, grad_field, g);
  return da_check_launch("da_warp3d_bwd");
}

// Same as da_warp3d_bwd, but grad_src is written CHANNELS-LAST: grad_src_cl [N, D*H*W, C] (zeroed here).  Requires
// C % 4 == 0 and a 16-byte aligned buffer; used for the multi-channel (probability-map) warp where the vectorised
// reductions cut the scatter cost about four-fold.
DA_API int da_warp3d_bwd_cl(const float* grad_out, const float* src, const float* field, int add_identity,
                            float* grad_src_cl, float* grad_field, int N, int C, int D, int H, int W, int Do, int Ho,
                            int Wo, cudaStream_t stream) {
  DA_REQUIRE(grad_out && src && field && grad_src_cl, "da_warp3d_bwd_cl: null pointer");
  DA_REQUIRE((C & 3) == 0 && (((uintptr_t)grad_src_cl) & 15) == 0, "da_warp3d_bwd_cl: needs C %% 4 == 0 and a 16-byte aligned buffer");
  WarpGeom g{N, C, D, H, W, Do, Ho, Wo};
  cudaError_t e = cudaMemsetAsync(grad_src_cl, 0, sizeof(float) * (size_t)N * C * D * H * W, stream);
  if (e != cudaSuccess) { da_set_error("da_warp3d_bwd_cl memset: %s", cudaGetErrorString(e)); return (int)e; }
  const int64_t total = (int64_t)N * Do * Ho * Wo;
  const int grid = grid_for(total, 256);
  if (add_identity)
    warp3d_bwd_cl_kernel<true><<<grid, 256, 0, stream>>>(grad_out, src, field, grad_src_cl, grad_field, g);
  else
    warp3d_bwd_cl_kernel<false><<<grid, 256, 0, stream>>>(grad_out, src, field, grad_src_cl, grad_field, g);
  return da_check_launch("da_warp3d_bwd_cl");
}
