// lncc: local normalised cross-correlation of VoxelMorphLNCC (lib/loss.py:589-617).
//
// The reference runs five dense 729-tap F.conv3d box filters (loss.py:602-606; the top cost of its
// registration step) plus a useless filter gradient.  Here the window sums are separable running
// sums (x, then y, then z) and the backward is the transposed ("full") box filter of three
// coefficient fields.  All window arithmetic is fp64: the reference's variance-by-cancellation
// (loss.py:611-613) loses up to 1.6e-3 of the gradient in fp32 on smooth images;
// fp64 sums put this implementation on the fp64-truth side of the parity ladder at negligible
// cost (the op is HBM-bound on its temporaries, not FLOP-bound).
// Algorithmic bytes fwd+bwd: 5*V*4.
#include "common.cuh"

namespace {

// x pass: I,J fp32 [NB][L] rows -> five fp64 fields [5][NB][Lv]
__global__ void __launch_bounds__(256) lncc_xsum_kernel(const float* __restrict__ I, const float* __restrict__ J,
                                                        double* __restrict__ out, int64_t rows, int W, int Wv,
                                                        int win) {
  const int64_t total = rows * Wv;
  const int64_t fs = total;  // field stride
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / Wv;
    const int xo = (int)(i - r * Wv);
    const float* pi = I + r * W + xo;
    const float* pj = J + r * W + xo;
    double a = 0, b = 0, c = 0, d = 0, e = 0;
    for (int k = 0; k < win; ++k) {
      const double vi = (double)pi[k], vj = (double)pj[k];
      a += vi; b += vj; c += vi * vi; d += vj * vj; e += vi * vj;
    }
    out[i] = a; out[fs + i] = b; out[2 * fs + i] = c; out[3 * fs + i] = d; out[4 * fs + i] = e;
  }
}

// generic box pass along the middle axis of [outer][L][inner];
// FULL=false: valid sums, Lout = Lin-win+1, out[l] = sum_{k<win} in[l+k]
// FULL=true : transposed,  Lout = Lin+win-1, out[p] = sum_{w=max(0,p-win+1)}^{min(p,Lin-1)} in[w]
template <bool FULL>
__global__ void __launch_bounds__(256) box_pass_kernel(const double* __restrict__ in, double* __restrict__ out,
                                                       int64_t outer, int Lin, int Lout, int64_t inner, int win) {
  const int64_t total = outer * Lout * inner;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t in_i = i % inner;
    const int64_t t = i / inner;
    const int l = (int)(t % Lout);
    const int64_t o = t / Lout;
    int lo, hi;
    if (FULL) { lo = l - win + 1 < 0 ? 0 : l - win + 1; hi = l < Lin - 1 ? l : Lin - 1; }
    else { lo = l; hi = l + win - 1; }
    const double* p = in + (o * Lin) * inner + in_i;
    double acc = 0;
    for (int k = lo; k <= hi; ++k) acc += p[(int64_t)k * inner];
    out[i] = acc;
  }
}

// z pass (valid) over the five fields + NCC + reduction + (optional) coefficient fields.
// in: [5][N][D][Hv*Wv]; coef: [ncoef][N][Dv][Hv*Wv] with ncoef = 3*popcount(need)
__global__ void __launch_bounds__(256) lncc_zsum_cc_kernel(const double* __restrict__ in, int N, int D, int Dv,
                                                           int64_t plane, int win, double eps, int need,
                                                           double* __restrict__ coef, double* __restrict__ partials) {
  const int64_t total = (int64_t)N * Dv * plane;
  const int64_t fs_in = (int64_t)N * D * plane;
  const double nwin = (double)win * win * win;
  double local = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pi = i % plane;
    const int64_t t = i / plane;
    const int z = (int)(t % Dv);
    const int64_t n = t / Dv;
    const double* p = in + ((n * D + z) * plane) + pi;
    double s[5] = {0, 0, 0, 0, 0};
    for (int k = 0; k < win; ++k) {
#pragma unroll
      for (int f = 0; f < 5; ++f) s[f] += p[f * fs_in + (int64_t)k * plane];
    }
    const double Is = s[0], Js = s[1], I2 = s[2], J2 = s[3], IJ = s[4];
    const double Im = Is / nwin, Jm = Js / nwin;
    const double cross = IJ - Im * Js - Jm * Is + Im * Jm * nwin;
    const double Iv = I2 - 2 * Im * Is + Im * Im * nwin;
    const double Jv = J2 - 2 * Jm * Js + Jm * Jm * nwin;
    const double den = Iv * Jv + eps;
    local += cross * cross / den;
    if (need) {
      const double c2d2 = cross * cross / (den * den);
      int slot = 0;
      if (need & 1) {  // d cc / d(Is, I2s, IJs)
        coef[(slot + 0) * total + i] = 2 * cross * (-Js / nwin) / den + c2d2 * Jv * (2 * Is / nwin);
        coef[(slot + 1) * total + i] = -c2d2 * Jv;
        coef[(slot + 2) * total + i] = 2 * cross / den;
        slot += 3;
      }
      if (need & 2) {  // d cc / d(Js, J2s, IJs)
        coef[(slot + 0) * total + i] = 2 * cross * (-Is / nwin) / den + c2d2 * Iv * (2 * Js / nwin);
        coef[(slot + 1) * total + i] = -c2d2 * Iv;
        coef[(slot + 2) * total + i] = 2 * cross / den;
      }
    }
  }
  __shared__ double red[8];
  const double b = block_sum<double, 8>(local, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = b;
}

__global__ void lncc_finalize_kernel(const double* __restrict__ partials, int nb, double count,
                                     float* __restrict__ loss_out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double acc = 0;
    for (int b = 0; b < nb; ++b) acc += partials[b];
    *loss_out = (float)(1.0 - acc / count);
  }
}

// final backward pass: transposed box along x over three fields [3][rows][Wv] + combine:
// grad[p] = scale * (A + 2*X_p*B + Y_p*C)   with X the differentiated image, Y the other one
__global__ void __launch_bounds__(256) lncc_bwd_x_kernel(const double* __restrict__ f, const float* __restrict__ X,
                                                         const float* __restrict__ Y, const float* __restrict__ gout,
                                                         double neg_inv_count, float* __restrict__ grad, int64_t rows,
                                                         int W, int Wv, int win) {
  const int64_t total = rows * W;
  const int64_t fs = rows * Wv;
  const double scale = (double)(*gout) * neg_inv_count;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / W;
    const int x = (int)(i - r * W);
    const int lo = x - win + 1 < 0 ? 0 : x - win + 1, hi = x < Wv - 1 ? x : Wv - 1;
    const double* p = f + r * Wv;
    double a = 0, b = 0, c = 0;
    for (int k = lo; k <= hi; ++k) { a += p[k]; b += p[fs + k]; c += p[2 * fs + k]; }
    grad[i] = (float)(scale * (a + 2.0 * (double)X[i] * b + (double)Y[i] * c));
  }
}

inline int gs_grid(int64_t total) {
  int64_t b = da_cdiv(total, 256);
  const int64_t cap = (int64_t)DA_NUM_SMS * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}
constexpr int LNCC_RED_BLOCKS = DA_NUM_SMS * 4;

}  // namespace

// need_grad: bit0 = I, bit1 = J.  coef (saved for backward) holds 3 fp64 fields per requested grad.
DA_API int64_t da_lncc_coef_bytes(int N, int D, int H, int W, int win, int need_grad) {
  const int64_t nw = (int64_t)N * (D - win + 1) * (H - win + 1) * (W - win + 1);
  const int nc = 3 * ((need_grad & 1) + ((need_grad >> 1) & 1));
  return (int64_t)sizeof(double) * nc * nw;
}
DA_API int64_t da_lncc_fwd_workspace_bytes(int N, int D, int H, int W, int win) {
  const int64_t Wv = W - win + 1, Hv = H - win + 1;
  return (int64_t)sizeof(double) * (5 * (int64_t)N * D * H * Wv + 5 * (int64_t)N * D * Hv * Wv + LNCC_RED_BLOCKS) + 512;
}
DA_API int64_t da_lncc_bwd_workspace_bytes(int N, int D, int H, int W, int win) {
  const int64_t Wv = W - win + 1, Hv = H - win + 1;
  return (int64_t)sizeof(double) * (3 * (int64_t)N * D * Hv * Wv + 3 * (int64_t)N * D * H * Wv) + 512;
}

// I, J: [N,1,D,H,W] fp32.  loss_out: 1 float (device).  coef: da_lncc_coef_bytes (may be null if need_grad==0).
DA_API int da_lncc_fwd(const float* I, const float* J, int N, int D, int H, int W, int win, double eps,
                       int need_grad, float* loss_out, void* coef, void* workspace, int64_t workspace_bytes,
                       cudaStream_t stream) {
  DA_REQUIRE(I && J && loss_out && workspace, "da_lncc_fwd: null pointer");
  DA_REQUIRE(win >= 1 && D >= win && H >= win && W >= win, "da_lncc_fwd: volume %dx%dx%d smaller than window %d", D, H, W, win);
  DA_REQUIRE(need_grad == 0 || coef, "da_lncc_fwd: coef buffer required when need_grad != 0");
  if (workspace_bytes < da_lncc_fwd_workspace_bytes(N, D, H, W, win)) {
    da_set_error("da_lncc_fwd: workspace too small");
    return DA_ERR_WORKSPACE;
  }
  const int Wv = W - win + 1, Hv = H - win + 1, Dv = D - win + 1;
  double* t1 = (double*)workspace;
  double* t2 = t1 + 5 * (int64_t)N * D * H * Wv;
  double* partials = t2 + 5 * (int64_t)N * D * Hv * Wv;
  const int64_t rows = (int64_t)N * D * H;
  lncc_xsum_kernel<<<gs_grid(rows * Wv), 256, 0, stream>>>(I, J, t1, rows, W, Wv, win);
  box_pass_kernel<false><<<gs_grid(5 * (int64_t)N * D * Hv * Wv), 256, 0, stream>>>(t1, t2, 5 * (int64_t)N * D, H, Hv, Wv, win);
  lncc_zsum_cc_kernel<<<LNCC_RED_BLOCKS, 256, 0, stream>>>(t2, N, D, Dv, (int64_t)Hv * Wv, win, eps, need_grad,
                                                            (double*)coef, partials);
  lncc_finalize_kernel<<<1, 32, 0, stream>>>(partials, LNCC_RED_BLOCKS, (double)N * Dv * Hv * Wv, loss_out);
  return da_check_launch("da_lncc_fwd", 4);
}

// grad_out: 1 float on device (upstream gradient of the scalar loss).  which: 0 -> grad wrt I, 1 -> grad wrt J
// (coef_slot selects the 3-field group inside coef: 0 or 1).
DA_API int da_lncc_bwd(const float* I, const float* J, const float* grad_out, const void* coef, int coef_slot,
                       int which, int N, int D, int H, int W, int win, float* grad, void* workspace,
                       int64_t workspace_bytes, cudaStream_t stream) {
  DA_REQUIRE(I && J && grad_out && coef && grad && workspace, "da_lncc_bwd: null pointer");
  if (workspace_bytes < da_lncc_bwd_workspace_bytes(N, D, H, W, win)) {
    da_set_error("da_lncc_bwd: workspace too small");
    return DA_ERR_WORKSPACE;
  }
  const int Wv = W - win + 1, Hv = H - win + 1, Dv = D - win + 1;
  const int64_t nw = (int64_t)N * Dv * Hv * Wv;
  const double* c = (const double*)coef + (int64_t)coef_slot * 3 * nw;
  double* t1 = (double*)workspace;                      // [3][N][D][Hv][Wv]
  double* t2 = t1 + 3 * (int64_t)N * D * Hv * Wv;         // [3][N][D][H][Wv]
  box_pass_kernel<true><<<gs_grid(3 * (int64_t)N * D * Hv * Wv), 256, 0, stream>>>(c, t1, 3 * (int64_t)N, Dv, D, (int64_t)Hv * Wv, win);
  box_pass_kernel<true><<<gs_grid(3 * (int64_t)N * D * H * Wv), 256, 0, stream>>>(t1, t2, 3 * (int64_t)N * D, Hv, H, Wv, win);
  const int64_t rows = (int64_t)N * D * H;
  lncc_bwd_x_kernel<<<gs_grid(rows * W), 256, 0, stream>>>(t2, which == 0 ? I : J, which == 0 ? J : I, grad_out,
                                                           -1.0 / (double)nw, grad, rows, W, Wv, win);
  return da_check_launch("da_lncc_bwd", 3);
}
