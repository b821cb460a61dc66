/* deepatlas_b200.h -- C ABI of libdeepatlas_b200.so (sm_100a).
 *
 * The reference (uncbiag/DeepAtlas) has no FFI of its own: its hot path sits behind two Python dict
 * registries, `network_dic` (lib/network_factory/__init__.py:9-16) and `loss_dict` (lib/loss.py:739-750).
 * The entry points below are what a binding for that path attaches to: one forward / backward pair per
 * library op the reference's modules call, taking raw DEVICE pointers, extents, scalar
 * hyper-parameters, a caller-owned workspace and a CUDA stream.  No torch types cross this boundary.
 *
 * Conventions
 *   - all tensors are dense planar NCDHW fp32 (the reference's layout) unless stated otherwise;
 *   - every function returns 0 on success, a positive cudaError_t, or a negative library code
 *     (-1 bad argument, -2 unsupported configuration, -3 workspace too small); the message is
 *     available from da_last_error() (thread-local);
 *   - the library never allocates or frees device memory; `*_bytes` queries size the workspaces;
 *   - calls are asynchronous on `stream` and re-entrant (the autograd engine calls backward from its
 *     own thread).
 */
#ifndef DEEPATLAS_B200_H
#define DEEPATLAS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* da_stream_t; /* == cudaStream_t */

int da_version(void);
const char* da_last_error(void);
int da_memset_zero(void* dst, int64_t bytes, da_stream_t stream);
int64_t da_launch_count(void); /* kernels launched by this library so far in this process */

/* ---- warp3d: identity grid + displacement + trilinear grid_sample in one pass ------------------------
 * replaces lib/utils.py:89-102 (get_identity_transform), lib/network_factory/voxel_morph.py:88 (disp + id)
 * and voxel_morph.py:90-91 (F.grid_sample bilinear / zeros / align_corners=True) and their autograd.
 * src [N,C,D,H,W]; field [N,3,Do,Ho,Wo] (channel 0 = x/W, 1 = y/H, 2 = z/D, normalised [-1,1]);
 * out [N,C,Do,Ho,Wo]; phi_out (nullable) receives field (+ identity). */
int da_warp3d_fwd(const float* src, const float* field, int add_identity, float* out, float* phi_out,
                  int N, int C, int D, int H, int W, int Do, int Ho, int Wo, da_stream_t stream);
int da_warp3d_bwd(const float* grad_out, const float* src, const float* field, int add_identity,
                  float* grad_src, float* grad_field, int N, int C, int D, int H, int W, int Do, int Ho,
                  int Wo, da_stream_t stream);

/* grad_src written channels-last [N, D*H*W, C] with vector reductions (C % 4 == 0); same arguments otherwise */
int da_warp3d_bwd_cl(const float* grad_out, const float* src, const float* field, int add_identity,
                     float* grad_src_cl, float* grad_field, int N, int C, int D, int H, int W, int Do, int Ho,
                     int Wo, da_stream_t stream);

/* ---- softmax + Dice sums -----------------------------------------------------------------------------
 * replaces F.softmax (lib/loss.py:427), mask_to_one_hot (lib/transforms.py:675-689) and the three
 * spatial sums of DiceLossMultiClass.forward (lib/loss.py:449-450,472).
 * source [N,C,V]; target_kind 0 = uint8 labels [N,V], 1 = int64 labels, 3 = int32 labels,
 * 2 = soft target fp32 [N,C,V].  sums [N,3,C] = (S = sum p, T = sum t, I = sum p*t). */
int64_t da_dice_workspace_bytes(int N, int C, int64_t V);
int da_dice_sums_fwd(const float* source, const void* target, int target_kind, int apply_softmax, int N,
                     int C, int64_t V, float* sums, void* workspace, int64_t workspace_bytes,
                     da_stream_t stream);
int da_dice_sums_bwd(const float* source, const void* target, int target_kind, int apply_softmax, int N,
                     int C, int64_t V, const float* gS, const float* gT, const float* gI,
                     float* grad_source, float* grad_target, da_stream_t stream);

/* ---- warped Dice sums: the anatomy term dice(grid_sample(P, phi), onehot(S_t)) of the joint step, fused ------
 * (F.grid_sample as at voxel_morph.py:90-91 + DiceLossMultiClass label-target sums, lib/loss.py:433-450,472).
 * prob [N,C,D,H,W]; field [N,3,Do,Ho,Wo]; labels [N,Do,Ho,Wo] (kind 0 uint8, 1 int64, 3 int32); sums [N,3,C].
 * The warped map is never materialised.  wsum (nullable, [N,D,H,W]) selects the scatter formulation: S_c = <P[c], Wsum>
 * with Wsum the trilinear weights scattered once (8 atomics + 8 gathers per voxel instead of 8*C gathers); the
 * forward leaves Wsum there and the backward takes it as wsum_fwd (null = recompute).  wsum = null runs the gather
 * kernel (fixed summation order).  Backward: 8 atomics + 16 gathers per voxel + dense passes (csrc/warp_dice.cu). */
int64_t da_warp_dice_fwd_workspace_bytes(int N, int C, int64_t Vo, int64_t Vs);
int64_t da_warp_dice_bwd_workspace_bytes(int N, int64_t Vs);
int da_warp_dice_sums_fwd(const float* prob, const float* field, int add_identity, const void* labels, int label_kind,
                          int N, int C, int D, int H, int W, int Do, int Ho, int Wo, float* sums, float* wsum,
                          void* workspace, int64_t workspace_bytes, da_stream_t stream);
int da_warp_dice_sums_bwd(const float* prob, const float* field, int add_identity, const void* labels, int label_kind,
                          const float* gS, const float* gI, const float* wsum_fwd, int N, int C, int D, int H, int W,
                          int Do, int Ho, int Wo, float* grad_prob, float* grad_field, void* workspace,
                          int64_t workspace_bytes, da_stream_t stream);

/* channel softmax (F.softmax(dim=1)) materialised for the anatomy branch, where probabilities are warped */
int da_softmax_fwd(const float* x, float* y, int N, int C, int64_t V, da_stream_t stream);
int da_softmax_bwd(const float* y, const float* dy, float* dx, int N, int C, int64_t V, da_stream_t stream);

/* ---- validation: label argmax + per-class overlap counts (models/segmentation.py:188-194, lib/evalMetrics.py:58-68) ----
 * logits [N,C,V]; truth (nullable) [N,V] labels; counts [N,3,C] int64 = (#argmax==c, #truth==c, #both); pred (nullable)
 * [N,V] uint8 label map (first maximum wins, as torch.max).  Bit-exact (integer atomics). */
int da_argmax_counts(const float* logits, const void* truth, int truth_kind, int N, int C, int64_t V, int64_t* counts,
                     uint8_t* pred, da_stream_t stream);

/* ---- local NCC (VoxelMorphLNCC.forward, lib/loss.py:597-617) --------------------------------------------
 * I,J [N,1,D,H,W]; loss_out: one device float = 1 - mean(cc).  need_grad bit0 = I, bit1 = J; `coef`
 * (da_lncc_coef_bytes) is saved by forward for backward. */
int64_t da_lncc_coef_bytes(int N, int D, int H, int W, int win, int need_grad);
int64_t da_lncc_fwd_workspace_bytes(int N, int D, int H, int W, int win);
int64_t da_lncc_bwd_workspace_bytes(int N, int D, int H, int W, int win);
int da_lncc_fwd(const float* I, const float* J, int N, int D, int H, int W, int win, double eps,
                int need_grad, float* loss_out, void* coef, void* workspace, int64_t workspace_bytes,
                da_stream_t stream);
int da_lncc_bwd(const float* I, const float* J, const float* grad_out, const void* coef, int coef_slot,
                int which, int N, int D, int H, int W, int win, float* grad, void* workspace,
                int64_t workspace_bytes, da_stream_t stream);

/* ---- bending energy (BendingEnergyLoss.forward, lib/loss.py:687-730) -----------------------------------
 * u [N,3,D,H,W]; sums [N,3,6]: per channel, sum over the interior of the squared second differences in
 * the order ddD, ddH, ddW, dDdH, dHdW, dDdW. */
int64_t da_bending_fwd_workspace_bytes(int N);
int64_t da_bending_bwd_workspace_bytes(int N, int D, int H, int W);
int da_bending_fwd(const float* u, int N, int D, int H, int W, float* sums, void* workspace,
                   int64_t workspace_bytes, da_stream_t stream);
int da_bending_bwd(const float* u, const float* grad_sums, int N, int D, int H, int W, float* grad_u,
                   void* workspace, int64_t workspace_bytes, da_stream_t stream);
/* norm_l1 = 1: sums of |r| instead of r^2 -- BendingEnergyLoss with norm != 'L2' (lib/loss.py:696-730, no scale factors) */
int da_bending_fwd_ex(const float* u, int N, int D, int H, int W, int norm_l1, float* sums, void* workspace,
                      int64_t workspace_bytes, da_stream_t stream);
int da_bending_bwd_ex(const float* u, const float* grad_sums, int N, int D, int H, int W, int norm_l1, float* grad_u,
                      void* workspace, int64_t workspace_bytes, da_stream_t stream);

/* ---- 3D convolution --------------------------------------------------------------------------------------
 * nn.Conv3d k3 (stride 1/2, pad 1) and k1: lib/network_factory/unets.py:30,36,98,115,120,250,
 * lib/network_factory/modules.py:48, lib/network_factory/voxel_morph.py:57; nn.ConvTranspose3d k3 s1 p1
 * (`transposed` = 1, weight (Cin,Cout,3,3,3)): unets.py:128,134 as used by UNet.dc8/dc7/dc5/dc4/dc2/dc1.
 * The conv input is cat(x1, x2) along channels (torch.cat at unets.py:275,157-171, voxel_morph.py:65-82);
 * x2 may be NULL with C2 = 0.  act: 0 none, 1 fused leaky-ReLU(slope) (slope 0 = ReLU). */
int da_umma_debug_read(int64_t* out11); /* debug: cycle counters of the tensor-core conv's MMA warps (DA_UMMA_DEBUG=1) */
int da_set_conv_impl(int impl); /* 0 = auto (tcgen05 fwd/dgrad/wgrad where they apply), 1 = generic direct kernels, 2 = tiled FFMA only, 3 = tcgen05 forced */
int da_set_conv_split(int split); /* operand format of the tcgen05 kernels: 0 = 3xFP16 (default: fp16 hi/lo pairs of operands scaled per tensor by a power of two, 22 significant bits, fp32 accumulate), 1 = 3xTF32 (half the MMA rate), 2 = 1xFP16 (one MMA per product on the scaled fp16 operands, 11 significant bits, fp32 accumulate: the reduced-precision mode of BASELINE config C2; parity tolerance 5e-3 instead of 1e-4) */
int64_t da_conv3d_pack_bytes(int Cin, int Cout, int ks);
int64_t da_conv3d_dgrad_workspace_bytes(int N, int Cin, int Cout, int Di, int Hi, int Wi, int ks, int stride); /* >= pack_bytes; stride 2 adds the zero-inserted dy that lets the tcgen05 path run */
int64_t da_conv3d_wgrad_workspace_bytes(int Cin, int Cout, int ks);
int da_conv3d_fwd(const float* x1, int C1, const float* x2, int C2, const float* weight, int transposed,
                  const float* bias, float* out, int N, int Di, int Hi, int Wi, int Cout, int ks, int stride,
                  int pad, int act, float slope, void* workspace, int64_t workspace_bytes,
                  da_stream_t stream);
int da_conv3d_dgrad(const float* dy, const float* weight, int transposed, float* dx, int N, int Cin_total,
                    int ci_off, int Cdx, int Cout, int Di, int Hi, int Wi, int ks, int stride, int pad,
                    void* workspace, int64_t workspace_bytes, da_stream_t stream);
int da_conv3d_wgrad(const float* x1, int C1, const float* x2, int C2, const float* dy, int transposed,
                    float* grad_weight, float* grad_bias, int N, int Di, int Hi, int Wi, int Cout, int ks,
                    int stride, int pad, void* workspace, int64_t workspace_bytes, da_stream_t stream);
/* The three calls above with caller-owned max-abs slots (one device float each, an upper bound of max|tensor|): the
 * tensor-core kernels scale their fp16 operand pairs per tensor by a power of two derived from it.  valid = 1: the slot
 * already holds the bound (da_absmax, an earlier _ex call on the same tensor, or a producer kernel); valid = 0: this
 * call fills it (da_conv3d_wgrad_ex: only when its tensor-core kernel runs).  NULL slots = the plain calls. */
int da_absmax(const float* x1, int64_t n1, const float* x2, int64_t n2, float* out, da_stream_t stream); /* out[0] = max(|x1|, |x2|); x2 may be NULL */
int da_conv3d_fwd_ex(const float* x1, int C1, const float* x2, int C2, const float* weight, int transposed,
                     const float* bias, float* out, int N, int Di, int Hi, int Wi, int Cout, int ks, int stride,
                     int pad, int act, float slope, void* workspace, int64_t workspace_bytes,
                     da_stream_t stream, float* amax_x, int amax_x_valid);
int da_conv3d_dgrad_ex(const float* dy, const float* weight, int transposed, float* dx, int N, int Cin_total,
                       int ci_off, int Cdx, int Cout, int Di, int Hi, int Wi, int ks, int stride, int pad,
                       void* workspace, int64_t workspace_bytes, da_stream_t stream, float* amax_dy, int amax_dy_valid);
int da_conv3d_wgrad_ex(const float* x1, int C1, const float* x2, int C2, const float* dy, int transposed,
                       float* grad_weight, float* grad_bias, int N, int Di, int Hi, int Wi, int Cout, int ks,
                       int stride, int pad, void* workspace, int64_t workspace_bytes, da_stream_t stream,
                       float* amax_x, int amax_x_valid, float* amax_dy, int amax_dy_valid, int accumulate);
/* (accumulate = 1: grad_weight / grad_bias += the result, inside the fixed-order region reduce.) */
int64_t da_channel_sum_workspace_bytes(int C);
int da_channel_sum(const float* x, int N, int C, int64_t V, float* out, void* workspace, int64_t workspace_bytes,
                   da_stream_t stream);

/* ---- batch norm (train mode) + activation; max-pool; nearest upsampling ---------------------------------
 * nn.BatchNorm3d (unets.py:31,51,116,130), nn.LeakyReLU / nn.ReLU (unets.py:32, modules.py:58),
 * nn.MaxPool3d(2) (unets.py:84-86,230), F.interpolate(size=) nearest (voxel_morph.py:72-80). */
int64_t da_bn_workspace_bytes(int C);
int da_bn_stats(const float* x, int N, int C, int64_t V, float eps, float momentum, float* mean,
                float* invstd, float* running_mean, float* running_var, void* workspace,
                int64_t workspace_bytes, da_stream_t stream);
int da_bn_act_fwd(const float* x, const float* mean, const float* invstd, const float* gamma,
                  const float* beta, int N, int C, int64_t V, int act, float slope, float* y,
                  da_stream_t stream);
int da_bn_act_bwd(const float* dy, const float* x, const float* mean, const float* invstd,
                  const float* gamma, const float* beta, int N, int C, int64_t V, int training, int act,
                  float slope, float* dx, float* dgamma, float* dbeta, void* workspace,
                  int64_t workspace_bytes, da_stream_t stream);
/* da_bn_stats / da_bn_act_bwd with one more output: an upper bound of max|.| of the tensor the layer is about to write
 * (amax_y: the activated output of da_bn_act_fwd with the same gamma / beta; amax_dx: the input gradient), derived from
 * per-channel value ranges gathered by the statistics passes.  One device float each, nullable; a consumer convolution
 * passes it to da_conv3d_*_ex as a valid max-abs slot and skips its own pass over the tensor. */
int da_bn_stats_ex(const float* x, int N, int C, int64_t V, float eps, float momentum, float* mean, float* invstd,
                   float* running_mean, float* running_var, const float* gamma, const float* beta, int act, float slope,
                   float* amax_y, void* workspace, int64_t workspace_bytes, da_stream_t stream);
int da_bn_act_bwd_ex(const float* dy, const float* x, const float* mean, const float* invstd, const float* gamma,
                     const float* beta, int N, int C, int64_t V, int training, int act, float slope, float* dx,
                     float* dgamma, float* dbeta, float* amax_dx, int accumulate, void* workspace, int64_t workspace_bytes,
                     da_stream_t stream); /* accumulate = 1: dgamma / dbeta += */
int da_act_bwd(const float* dy, const float* y, float slope, int64_t total, float* dx, da_stream_t stream);
int da_maxpool2_fwd(const float* x, float* y, int64_t NC, int D, int H, int W, da_stream_t stream);
int da_maxpool2_bwd(const float* dy, const float* x, float* dx, int64_t NC, int D, int H, int W,
                    da_stream_t stream);
int da_upsample_nearest_fwd(const float* x, float* y, int64_t NC, int D, int H, int W, int Do, int Ho,
                            int Wo, da_stream_t stream);
int da_upsample_nearest_bwd(const float* dy, float* dx, int64_t NC, int D, int H, int W, int Do, int Ho,
                            int Wo, da_stream_t stream);

/* ---- ConvTranspose3d kernel 2 stride 2 (unets.deconvBlock, unets.py:42-58,240-241; UNet.dc9/dc6/dc3) ---
 * x [N,Cin,D,H,W]; weight (Cin,Cout,2,2,2); out [N,Cout,2D,2H,2W]. */
int64_t da_deconv_k2s2_wgrad_workspace_bytes(int Cin, int Cout);
int da_deconv_k2s2_fwd(const float* x, const float* weight, const float* bias, float* out, int N, int Cin,
                       int Cout, int D, int H, int W, da_stream_t stream);
int da_deconv_k2s2_dgrad(const float* dy, const float* weight, float* dx, int N, int Cin, int Cout, int D,
                         int H, int W, da_stream_t stream);
int da_deconv_k2s2_wgrad(const float* x, const float* dy, float* grad_weight, float* grad_bias, int N,
                         int Cin, int Cout, int D, int H, int W, void* workspace, int64_t workspace_bytes,
                         da_stream_t stream);
int da_deconv_k2s2_wgrad_ex(const float* x, const float* dy, float* grad_weight, float* grad_bias, int N, int Cin, int Cout,
                            int D, int H, int W, int accumulate, void* workspace, int64_t workspace_bytes,
                            da_stream_t stream); /* accumulate = 1: grad_weight / grad_bias += */

/* ---- remaining entries of the loss registry (lib/loss.py:739-750) ------------------
 * pair moments: replaces the ATen passes of NormalizedCrossCorrelationLoss (lib/loss.py:494-501), nn.MSELoss
 * ('mse') and L2Loss (lib/loss.py:733-736).  a, b [N,V] (b nullable); out [N,9] per sample =
 * (sum a, sum b, sum a^2, sum b^2, sum a*b, sum (a-b)^2, sum (a-ma)^2, sum (b-mb)^2, sum (a-ma)(b-mb)), accumulated
 * in fp64.  The gradient of any function of those moments is affine in (a, b, 1): da_affine2 with per-sample
 * coefficients coef [N,3] on the device. */
int64_t da_pair_moments_workspace_bytes(int N);
int da_pair_moments_fwd(const float* a, const float* b, int N, int64_t V, float* out, void* workspace,
                        int64_t workspace_bytes, da_stream_t stream);
int da_affine2(const float* a, const float* b, const float* coef, int N, int64_t V, float* out, da_stream_t stream);

/* gradientLoss (lib/loss.py:625-671): u [N,C,D,H,W]; sums [N,C,3] = sum f(u[d+2]-u[d]), sum f(u[h+2]+u[h]),
 * sum f(u[w+2]+u[w]) (the reference adds along H and W -- kept); f = square (norm_l1 0) or abs (1). */
int64_t da_gradient_loss_workspace_bytes(int N, int C);
int da_gradient_loss_fwd(const float* u, int N, int C, int D, int H, int W, int norm_l1, float* sums, void* workspace,
                         int64_t workspace_bytes, da_stream_t stream);
int da_gradient_loss_bwd(const float* u, const float* grad_sums, int N, int C, int D, int H, int W, int norm_l1,
                         float* grad_u, da_stream_t stream);

/* channel log-softmax terms: nn.CrossEntropyLoss (mode 0; registry 'cross_entropy', lib/loss.py:748), FocalLoss
 * (mode 1; lib/loss.py:149-186, including probs = nll_loss(P) = -P[t]), SoftCrossEntropy with softmax (mode 2,
 * lib/loss.py:114) and without (mode 3, lib/loss.py:116).  x [N,C,V]; target: labels [N,V] (kind 0 uint8, 1 int64,
 * 3 int32) in modes 0-1, fp32 [N,C,V] (kind 2) in modes 2-3; class_weight (nullable, [C]) = CrossEntropyLoss weight /
 * FocalLoss alpha; out2 [2] = (sum of terms, sum of weights (mode 0) or voxel count).  Backward: grad_scale is a DEVICE
 * scalar d loss / d out2[0]; grad_target (nullable) only in the soft modes. */
int64_t da_xent_workspace_bytes(int N);
int da_xent_fwd(const float* x, const void* target, int target_kind, int mode, int N, int C, int64_t V,
                const float* class_weight, float gamma, int focal_softmax, int64_t ignore_index, float* out2,
                void* workspace, int64_t workspace_bytes, da_stream_t stream);
int da_xent_bwd(const float* x, const void* target, int target_kind, int mode, int N, int C, int64_t V,
                const float* class_weight, float gamma, int focal_softmax, int64_t ignore_index,
                const float* grad_scale, float* grad_x, float* grad_target, da_stream_t stream);

/* ---- UNet_generator variants (lib/network_factory/unets.py:230-241,264,275) ------------
 * nn.Upsample(scale_factor=2, mode='trilinear') (align_corners=False): x [NC,D,H,W] -> y [NC,2D,2H,2W]. */
int da_upsample_trilinear2_fwd(const float* x, float* y, int64_t NC, int D, int H, int W, da_stream_t stream);
int da_upsample_trilinear2_bwd(const float* dy, float* dx, int64_t NC, int D, int H, int W, da_stream_t stream);
/* residual add `enc(x) + x`: out [N,Ca,V] = a + b, b [N,Cb,V] with Cb == Ca or 1 (broadcast); and the gradient of the
 * broadcast operand, out [N,1,V] = sum over channels of g [N,C,V]. */
int da_add_bcast(const float* a, const float* b, int N, int Ca, int Cb, int64_t V, float* out, da_stream_t stream);
int da_channel_reduce(const float* g, int N, int C, int64_t V, float* out, da_stream_t stream);

/* ---- device-side input stage (lib/transforms.py:79-80 clip to [0,1], :124-158 CropTensor)
 * src [NC,D,H,W] -> dst [NC,Do,Ho,Wo] = src[z0:z0+Do, y0:y0+Ho, x0:x0+Wo] (fp32 clipped to [lo,hi]; uint8 labels). */
int da_crop_clip_f32(const float* src, float* dst, int64_t NC, int D, int H, int W, int z0, int y0, int x0, int Do,
                     int Ho, int Wo, float lo, float hi, da_stream_t stream);
int da_crop_u8(const uint8_t* src, uint8_t* dst, int64_t NC, int D, int H, int W, int z0, int y0, int x0, int Do,
               int Ho, int Wo, da_stream_t stream);

/* DiceLossOnLabel (lib/loss.py:348-391), integer half: a, b [N,V] label maps (kind 0 uint8, 1 int64, 3 int32, 4 fp32
 * truncated as mask.long()); counts [N,3,bins] int64 = (#[a == c], #[b == c], #[a == b == c]); bins <= 256. */
int da_label_overlap_counts(const void* a, int kind_a, const void* b, int kind_b, int N, int bins, int64_t V,
                            int64_t* counts, da_stream_t stream);

/* Multi-scale LNCCLoss (lib/loss.py:512-586), one scale per call: k^3 ones filter with dilation `dil` and stride `stride`
 * (F.conv3d, padding 0) over I, J, I^2, J^2, I*J [N,1,D,H,W]; out_sum [1] = sum over windows of
 * cross^2 / (Ivar*Jvar + 1e-5); coef (nullable, da_lncc_ms_coef_bytes) keeps the per-window derivatives for the backward.
 * Backward: grad (+)= grad_scale[0] (device scalar) * scale * d out_sum / d I (which 0) or J (which 1). */
int64_t da_lncc_ms_workspace_bytes(void);
int64_t da_lncc_ms_coef_bytes(int N, int D, int H, int W, int k, int dil, int stride);
int da_lncc_ms_fwd(const float* I, const float* J, int N, int D, int H, int W, int k, int dil, int stride, float* out_sum,
                   float* coef, void* workspace, int64_t workspace_bytes, da_stream_t stream);
int da_lncc_ms_bwd(const float* I, const float* J, const float* coef, int which, const float* grad_scale, float scale, int N,
                   int D, int H, int W, int k, int dil, int stride, int accumulate, float* grad, da_stream_t stream);

/* softmax + Dice sums + the probabilities in one pass over the logits, for a prediction that feeds the supervised Dice
 * term and, as probabilities, a second consumer (the anatomy term of the joint step): replaces da_dice_sums_fwd
 * (apply_softmax) + da_softmax_fwd, and in the backward da_dice_sums_bwd + da_softmax_bwd + the sum of their results.
 * probs, grad_probs (nullable), grad_logits [N,C,V]; targets as for da_dice_sums_fwd. */
int da_softmax_dice_fwd(const float* logits, const void* target, int target_kind, int N, int C, int64_t V, float* sums,
                        float* probs, void* workspace, int64_t workspace_bytes, da_stream_t stream);
int da_softmax_dice_bwd(const float* logits, const void* target, int target_kind, int N, int C, int64_t V, const float* gS,
                        const float* gT, const float* gI, const float* grad_probs, float* grad_logits, da_stream_t stream);
/* The U-Net's 1x1x1 class head (lib/network_factory/unets.py:250, Conv3d(16, C, 1)) fused with softmax + the Dice sums
 * (lib/loss.py:427-476): the logits and their gradient never reach HBM.  feat [N,16,V]; weight [C,16], bias [C]
 * nullable; label target (kinds 0, 1, 3).  Forward: sums [N,3,C], probs [N,C,V] nullable (a second consumer of the
 * probabilities).  Backward: gS, gI [N,C]; grad_probs nullable; grad_feat [N,16,V], grad_weight [C,16], grad_bias [C]
 * nullable.  da_head_dice_supported: K == 16, C <= 32, V even (callers take the unfused calls otherwise). */
int64_t da_head_dice_supported(int K, int C, int64_t V); /* 1 or 0 */
int64_t da_head_dice_workspace_bytes(int N, int C, int64_t V);
int da_head_dice_fwd(const float* feat, const float* weight, const float* bias, const void* target, int target_kind, int N,
                     int K, int C, int64_t V, float* sums, float* probs, void* workspace, int64_t workspace_bytes,
                     da_stream_t stream);
int da_head_dice_bwd(const float* feat, const float* weight, const float* bias, const void* target, int target_kind, int N,
                     int K, int C, int64_t V, const float* gS, const float* gI, const float* grad_probs, float* grad_feat,
                     float* grad_weight, float* grad_bias, int accumulate, void* workspace, int64_t workspace_bytes,
                     da_stream_t stream); /* accumulate = 1: grad_weight / grad_bias += */

#ifdef __cplusplus
}
#endif
#endif /* DEEPATLAS_B200_H */
