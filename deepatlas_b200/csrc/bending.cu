// bending_energy: BendingEnergyLoss (lib/loss.py:674-730), L2 form.
//
// The reference builds six strided-slice second-difference expressions (loss.py:702-718), each
// materialising temporaries, then abs -> square -> mean (~40 tiny ATen kernels).  Here one stencil
// pass over the 3-channel displacement produces the 18 sums  S[term][channel] = sum_interior r^2
// (term order: ddD, ddH, ddW, dDdH, dHdW, dDdW -- the reference's "x" is the D axis) through a
// shuffle tree + fixed-order second stage.  The per-channel scale factors (the reference's
// spatial_dims quirk, loss.py:722-727) are applied by the host-side mirror on the 18 numbers.
// Backward: residual fields r[term][c] then a gather with the transposed stencils (deterministic).
// Algorithmic bytes fwd+bwd: 9*V*4.
#include "common.cuh"

namespace {

constexpr int BE_THREADS = 256;
constexpr int BE_BLOCKS = DA_NUM_SMS * 4;

struct Geo { int D, H, W; int64_t sH, sD, V; };

__device__ __forceinline__ void residuals(const float* __restrict__ u, const Geo& g, int64_t i, float (&r)[6]) {
  const float c = u[i];
  const float dp = u[i + g.sD], dm = u[i - g.sD], hp = u[i + g.sH], hm = u[i - g.sH], wp = u[i + 1], wm = u[i - 1];
  r[0] = dp + dm - 2.f * c;
  r[1] = hp + hm - 2.f * c;
  r[2] = wp + wm - 2.f * c;
  r[3] = u[i + g.sD + g.sH] + u[i - g.sD - g.sH] - u[i + g.sD - g.sH] - u[i - g.sD + g.sH];
  r[4] = u[i + g.sH + 1] + u[i - g.sH - 1] - u[i + g.sH - 1] - u[i - g.sH + 1];
  r[5] = u[i + g.sD + 1] + u[i - g.sD - 1] - u[i + g.sD - 1] - u[i - g.sD + 1];
}

// partials [gridDim.x][N*3*6] ; index ((n*3+c)*6+term)
// l1: sums of |r| (the reference's norm != 'L2' path, lib/loss.py:729 without the squares) instead of r^2
__global__ void __launch_bounds__(BE_THREADS) bending_fwd_kernel(const float* __restrict__ u, int N, Geo g,
                                                                 double* __restrict__ partials, int l1) {
  __shared__ double red[BE_THREADS / 32];
  const int64_t interior = (int64_t)(g.D - 2) * (g.H - 2) * (g.W - 2);
  for (int nc = 0; nc < N * 3; ++nc) {
    const float* uc = u + (int64_t)nc * g.V;
    float acc[6] = {0, 0, 0, 0, 0, 0};
    for (int64_t j = (int64_t)blockIdx.x * BE_THREADS + threadIdx.x; j < interior;
         j += (int64_t)gridDim.x * BE_THREADS) {
      const int x = (int)(j % (g.W - 2)) + 1;
      const int64_t t = j / (g.W - 2);
      const int y = (int)(t % (g.H - 2)) + 1, z = (int)(t / (g.H - 2)) + 1;
      float r[6];
      residuals(uc, g, (int64_t)z * g.sD + (int64_t)y * g.sH + x, r);
#pragma unroll
      for (int k = 0; k < 6; ++k) acc[k] += l1 ? fabsf(r[k]) : r[k] * r[k];
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const double b = block_sum<double, BE_THREADS / 32>((double)acc[k], red);
      if (threadIdx.x == 0) partials[(int64_t)blockIdx.x * (N * 18) + nc * 6 + k] = b;
    }
  }
}

__global__ void bending_finalize_kernel(const double* __restrict__ partials, int nb, int n18, float* __restrict__ sums) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n18) return;
  double acc = 0;
  for (int b = 0; b < nb; ++b) acc += partials[(int64_t)b * n18 + i];
  sums[i] = (float)acc;
}

// residual fields rf [N*3][6][V] (zero outside the interior); l1: their signs (d|r|/dr, 0 at 0 as torch.abs)
__global__ void __launch_bounds__(BE_THREADS) bending_resid_kernel(const float* __restrict__ u, int N, Geo g,
                                                                   float* __restrict__ rf, int l1) {
  const int64_t total = (int64_t)N * 3 * g.V;
  for (int64_t i = (int64_t)blockIdx.x * BE_THREADS + threadIdx.x; i < total; i += (int64_t)gridDim.x * BE_THREADS) {
    const int64_t nc = i / g.V, v = i - nc * g.V;
    const int x = (int)(v % g.W), y = (int)((v / g.W) % g.H), z = (int)(v / g.sD);
    float r[6] = {0, 0, 0, 0, 0, 0};
    if (x >= 1 && x < g.W - 1 && y >= 1 && y < g.H - 1 && z >= 1 && z < g.D - 1) residuals(u + nc * g.V, g, v, r);
    float* o = rf + nc * 6 * g.V + v;
#pragma unroll
    for (int k = 0; k < 6; ++k) o[(int64_t)k * g.V] = l1 ? (float)((r[k] > 0.f) - (r[k] < 0.f)) : r[k];
  }
}

// grad[q] = sum_term 2*gs[term] * sum_k a_k * r_term[q - o_k]; gsums [N*3*6] upstream grads of the sums
__global__ void __launch_bounds__(BE_THREADS) bending_gather_kernel(const float* __restrict__ rf,
                                                                    const float* __restrict__ gsums, int N, Geo g,
                                                                    float* __restrict__ grad, float outer) {
  const int64_t total = (int64_t)N * 3 * g.V;
  for (int64_t i = (int64_t)blockIdx.x * BE_THREADS + threadIdx.x; i < total; i += (int64_t)gridDim.x * BE_THREADS) {
    const int64_t nc = i / g.V, v = i - nc * g.V;
    const int x = (int)(v % g.W), y = (int)((v / g.W) % g.H), z = (int)(v / g.sD);
    const float* r = rf + nc * 6 * g.V;
    const float* gs = gsums + nc * 6;
    auto at = [&](int term, int dz, int dy, int dx) -> float {
      const int zz = z + dz, yy = y + dy, xx = x + dx;
      if (zz < 0 || zz >= g.D || yy < 0 || yy >= g.H || xx < 0 || xx >= g.W) return 0.f;
      return r[(int64_t)term * g.V + (int64_t)zz * g.sD + (int64_t)yy * g.sH + xx];
    };
    float acc = 0.f;
    // r(p) = u(p+e) + u(p-e) - 2u(p)   =>  d/du(q): r(q-e) + r(q+e) - 2 r(q)
    acc += gs[0] * (at(0, -1, 0, 0) + at(0, 1, 0, 0) - 2.f * at(0, 0, 0, 0));
    acc += gs[1] * (at(1, 0, -1, 0) + at(1, 0, 1, 0) - 2.f * at(1, 0, 0, 0));
    acc += gs[2] * (at(2, 0, 0, -1) + at(2, 0, 0, 1) - 2.f * at(2, 0, 0, 0));
    // r(p) = u(p+a+b) + u(p-a-b) - u(p+a-b) - u(p-a+b)  =>  r(q-a-b) + r(q+a+b) - r(q-a+b) - r(q+a-b)
    acc += gs[3] * (at(3, -1, -1, 0) + at(3, 1, 1, 0) - at(3, -1, 1, 0) - at(3, 1, -1, 0));
    acc += gs[4] * (at(4, 0, -1, -1) + at(4, 0, 1, 1) - at(4, 0, -1, 1) - at(4, 0, 1, -1));
    acc += gs[5] * (at(5, -1, 0, -1) + at(5, 1, 0, 1) - at(5, -1, 0, 1) - at(5, 1, 0, -1));
    grad[i] = outer * acc;   // 2 for the squares, 1 for the absolute values
  }
}

inline Geo make_geo(int D, int H, int W) { return Geo{D, H, W, (int64_t)W, (int64_t)H * W, (int64_t)D * H * W}; }

}  // namespace

DA_API int64_t da_bending_fwd_workspace_bytes(int N) { return (int64_t)sizeof(double) * BE_BLOCKS * N * 18; }
DA_API int64_t da_bending_bwd_workspace_bytes(int N, int D, int H, int W) {
  return (int64_t)sizeof(float) * N * 18 * D * H * W;
}

// u [N,3,D,H,W]; sums [N,3,6] (per channel, term order ddD,ddH,ddW,dDdH,dHdW,dDdW) = sum over the interior of r^2
// (norm_l1 = 0) or of |r| (norm_l1 = 1: BendingEnergyLoss with norm != 'L2', lib/loss.py:696-730)
DA_API int da_bending_fwd_ex(const float* u, int N, int D, int H, int W, int norm_l1, float* sums, void* workspace,
                             int64_t workspace_bytes, cudaStream_t stream) {
  DA_REQUIRE(u && sums && workspace, "da_bending_fwd: null pointer");
  DA_REQUIRE(D >= 3 && H >= 3 && W >= 3, "da_bending_fwd: extent must be >= 3 per axis");
  if (workspace_bytes < da_bending_fwd_workspace_bytes(N)) { da_set_error("da_bending_fwd: workspace too small"); return DA_ERR_WORKSPACE; }
  Geo g = make_geo(D, H, W);
  bending_fwd_kernel<<<BE_BLOCKS, BE_THREADS, 0, stream>>>(u, N, g, (double*)workspace, norm_l1 ? 1 : 0);
  bending_finalize_kernel<<<(N * 18 + 63) / 64, 64, 0, stream>>>((const double*)workspace, BE_BLOCKS, N * 18, sums);
  return da_check_launch("da_bending_fwd", 2);
}
DA_API int da_bending_fwd(const float* u, int N, int D, int H, int W, float* sums, void* workspace,
                          int64_t workspace_bytes, cudaStream_t stream) {
  return da_bending_fwd_ex(u, N, D, H, W, 0, sums, workspace, workspace_bytes, stream);
}

// grad_sums [N,3,6] upstream; grad_u [N,3,D,H,W]
DA_API int da_bending_bwd_ex(const float* u, const float* grad_sums, int N, int D, int H, int W, int norm_l1, float* grad_u,
                             void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  DA_REQUIRE(u && grad_sums && grad_u && workspace, "da_bending_bwd: null pointer");
  if (workspace_bytes < da_bending_bwd_workspace_bytes(N, D, H, W)) { da_set_error("da_bending_bwd: workspace too small"); return DA_ERR_WORKSPACE; }
  Geo g = make_geo(D, H, W);
  const int64_t total = (int64_t)N * 3 * g.V;
  int64_t b = da_cdiv(total, BE_THREADS);
  const int grid = (int)(b > (int64_t)DA_NUM_SMS * 16 ? (int64_t)DA_NUM_SMS * 16 : b);
  bending_resid_kernel<<<grid, BE_THREADS, 0, stream>>>(u, N, g, (float*)workspace, norm_l1 ? 1 : 0);
  bending_gather_kernel<<<grid, BE_THREADS, 0, stream>>>((const float*)workspace, grad_sums, N, g, grad_u, norm_l1 ? 1.f : 2.f);
  return da_check_launch("da_bending_bwd", 2);
}
DA_API int da_bending_bwd(const float* u, const float* grad_sums, int N, int D, int H, int W, float* grad_u,
                          void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  return da_bending_bwd_ex(u, grad_sums, N, D, H, W, 0, grad_u, workspace, workspace_bytes, stream);
}
