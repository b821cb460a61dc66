"""torch.autograd.Function wrappers over the C-ABI kernels (one per library op the reference calls).

PyTorch is plumbing here: it owns device memory (outputs, saved tensors, workspaces come from the
caching allocator), the CUDA stream, and the autograd graph.  All arithmetic happens in
``libdeepatlas_b200.so``.  Inputs must be CUDA fp32 tensors -- there is no CPU or eager fallback, a
CPU tensor raises.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import torch

from . import _lib


def _p(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"deepatlas_b200: '{name}' must be a CUDA tensor (no CPU fallback exists)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"deepatlas_b200: '{name}' must be float32, got {t.dtype}")
    _check_device(t, name)
    return t.contiguous()


# Max-abs bounds travel with the tensors they describe: a producer that knows an upper bound of max|t| (batch norm, from
# the value ranges its statistics pass sees anyway; max-pooling and nearest up-sampling, which only select values)
# attaches it as ``t._da_amax`` (a one-element device tensor), and the convolutions' tensor-core path, which scales its
# fp16 operand pairs by such a bound, then skips its own pass over the tensor.  Any op that makes a new tensor drops it.
def _get_amax(t):
    a = getattr(t, "_da_amax", None)
    if a is None or getattr(t, "_da_amax_version", None) != t._version:
        return None
    return a


def _set_amax(t, a):
    if a is not None:
        t._da_amax = a
        t._da_amax_version = t._version
    return t


# Parameter gradients written straight into their destination.  A gradient tensor marked ``_da_inplace`` (FlatGradBucket
# marks its views: ``p.grad`` slices of one flat all-reduce buffer that exist before backward starts) is added to inside
# the producing backward's own fixed-order reduce kernel (``accumulate = 1``), and autograd is handed ``None`` -- instead of
# a fresh tensor that autograd then adds to ``p.grad`` with one more kernel per parameter and use (187 launches per joint
# step).  Opt-in per parameter, because ``torch.autograd.grad`` and parameter hooks expect the tensors.
# A second set of gradient slots: two passes of ONE network that run on different streams (the joint step's two
# segmentation passes) must not add into the same memory concurrently.  Ops created under ``grad_slot(1)`` remember the
# slot and their backward adds into the bucket's alternate views (``FlatGradBucket.enable_alt``), which the bucket folds
# into the primary ones before the all-reduce.
_GRAD_SLOT = 0


class grad_slot:
    def __init__(self, slot: int):
        self.slot, self.prev = int(slot), 0

    def __enter__(self):
        global _GRAD_SLOT
        self.prev, _GRAD_SLOT = _GRAD_SLOT, self.slot
        return self

    def __exit__(self, *exc):
        global _GRAD_SLOT
        _GRAD_SLOT = self.prev


def _grad_dst(p, wanted: bool, slot: int = 0):
    """``p.grad`` (or its alternate view for ``slot`` 1) if the gradient of parameter ``p`` is to be accumulated in
    place, else None."""
    if not wanted or p is None or not p.is_leaf:
        return None
    g = p.grad
    if g is None or not getattr(g, "_da_inplace", False):
        return None
    if slot:
        g = getattr(g, "_da_alt", None)
        if g is None:
            return None
    if g.dtype != torch.float32 or g.device != p.device or not g.is_contiguous() or g.shape != p.shape:
        return None
    return g


# Weight gradients on a side stream.  A convolution's weight gradient has no consumer before the optimizer step, so (when
# it is accumulated in place, i.e. nothing is handed back to autograd) it can leave the backward chain: it is launched on a
# per-device side stream as soon as dy exists, and the chain (data gradient -> batch-norm backward -> next data gradient)
# continues on the calling stream.  The tcgen05 kernels own a whole SM each, so two of them never share one, but the
# HBM-bound elementwise kernels of the chain run next to the shared-memory-bound weight-gradient kernels.
# ``join_wgrad_stream()`` must be called after backward() and before the gradients are read.
_WGRAD_OVERLAP = False
_WGRAD_SIDE = {}
_WGRAD_PENDING = set()   # devices whose side stream holds weight gradients that have not been joined yet


def set_wgrad_overlap(on: bool):
    global _WGRAD_OVERLAP
    _WGRAD_OVERLAP = bool(on)


def _wgrad_side(device):
    st = _WGRAD_SIDE.get(device.index)
    if st is None:
        st = _WGRAD_SIDE[device.index] = torch.cuda.Stream(device=device)
    return st


def join_wgrad_stream():
    """The current stream waits for every weight gradient launched on the side stream of the current device since the
    last join.  (Only then: waiting for a side stream that took no part in an ongoing graph capture would tie the capture
    to uncaptured work -- cudaErrorStreamCaptureIsolation.)"""
    if not torch.cuda.is_available():
        return
    dev = torch.cuda.current_device()
    if dev in _WGRAD_PENDING:
        _WGRAD_PENDING.discard(dev)
        torch.cuda.current_stream().wait_stream(_WGRAD_SIDE[dev])


def _check_device(t: torch.Tensor, name: str):
    """Kernels launch on the CURRENT device's current stream (``_stream()``): a tensor living on another GPU would be
    dereferenced in the wrong context.  Fail loudly instead; ``with torch.cuda.device(t.device):`` is the fix."""
    cur = torch.cuda.current_device()
    if t.device.index != cur:
        raise RuntimeError(f"deepatlas_b200: '{name}' lives on {t.device} but the current CUDA device is cuda:{cur}; "
                           f"wrap the call in `with torch.cuda.device({t.device.index}):` (one process per GPU is the supported layout)")


# ------------------------------------------------------------------------------------------------------
# warp3d  (lib/network_factory/voxel_morph.py:85-91)
# ------------------------------------------------------------------------------------------------------
class Warp3dFunction(torch.autograd.Function):
    """out, phi = warp3d(src, field, add_identity): phi = field (+ identity grid), out = trilinear
    sample of src at phi (zeros padding, align_corners=True)."""

    @staticmethod
    def forward(ctx, src, field, add_identity: bool, want_phi: bool):
        src, field = _f32(src, "src"), _f32(field, "field")
        N, C, D, H, W = src.shape
        if field.dim() != 5 or field.shape[0] != N or field.shape[1] != 3:
            raise ValueError(f"warp3d: field must be (N,3,Do,Ho,Wo), got {tuple(field.shape)}")
        Do, Ho, Wo = field.shape[2:]
        out = torch.empty((N, C, Do, Ho, Wo), dtype=torch.float32, device=src.device)
        phi = torch.empty_like(field) if want_phi else None
        _lib.call("da_warp3d_fwd", _p(src), _p(field), int(add_identity), _p(out), _p(phi), N, C, D, H, W,
                  Do, Ho, Wo, _stream())
        ctx.save_for_backward(src, field)
        ctx.add_identity = bool(add_identity)
        if want_phi:
            return out, phi
        ctx.mark_non_differentiable()
        return out, None

    @staticmethod
    def backward(ctx, g_out, g_phi):
        src, field = ctx.saved_tensors
        N, C, D, H, W = src.shape
        Do, Ho, Wo = field.shape[2:]
        need_src, need_field = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        g_src = g_field = None
        if g_out is not None and (need_src or need_field):
            g_out = _f32(g_out, "grad_out")
            g_field = torch.empty_like(field) if need_field else None
            if need_src and C % 4 == 0:
                # multi-channel scatter: channels-last buffer + vector reductions, returned as a permuted view
                g_cl = torch.empty((N, D, H, W, C), dtype=torch.float32, device=src.device)
                _lib.call("da_warp3d_bwd_cl", _p(g_out), _p(src), _p(field), int(ctx.add_identity), _p(g_cl),
                          _p(g_field), N, C, D, H, W, Do, Ho, Wo, _stream())
                g_src = g_cl.permute(0, 4, 1, 2, 3)
            else:
                g_src = torch.empty_like(src) if need_src else None
                _lib.call("da_warp3d_bwd", _p(g_out), _p(src), _p(field), int(ctx.add_identity), _p(g_src),
                          _p(g_field), N, C, D, H, W, Do, Ho, Wo, _stream())
        if need_field and g_phi is not None:
            g_field = g_phi if g_field is None else g_field + g_phi
        return g_src, g_field, None, None


def warp3d(src, field, add_identity=False, want_phi=False):
    out, phi = Warp3dFunction.apply(src, field, add_identity, want_phi)
    return (out, phi) if want_phi else out


# ------------------------------------------------------------------------------------------------------
# softmax + Dice sums  (lib/loss.py:427-472, lib/transforms.py:675-689)
# ------------------------------------------------------------------------------------------------------
_KIND = {torch.uint8: 0, torch.int64: 1, torch.int32: 3}


class DiceSumsFunction(torch.autograd.Function):
    """sums[N,3,C] = (sum p, sum t, sum p*t) with p = softmax(source) if apply_softmax else source and
    t = one-hot(labels) (never materialised) or a soft target."""

    @staticmethod
    def forward(ctx, source, target, apply_softmax: bool):
        source = _f32(source, "source")
        N, C = source.shape[:2]
        V = source[0, 0].numel()
        if not target.is_cuda:
            raise RuntimeError("deepatlas_b200: 'target' must be a CUDA tensor")
        if target.is_floating_point():
            target = _f32(target, "target")
            if target.numel() != source.numel():
                raise ValueError("dice: soft target must have the shape of source")
            kind = 2
        else:
            if target.dtype not in _KIND:
                target = target.long()
            target = target.contiguous()
            if target.numel() != N * V:
                raise ValueError("dice: label target must have N*D*H*W elements")
            kind = _KIND[target.dtype]
        sums = torch.empty((N, 3, C), dtype=torch.float32, device=source.device)
        nb = _lib.size("da_dice_workspace_bytes", N, C, V)
        ws = _ws(nb, source.device)
        _lib.call("da_dice_sums_fwd", _p(source), _p(target), kind, int(apply_softmax), N, C, V, _p(sums),
                  _p(ws), nb, _stream())
        ctx.save_for_backward(source, target)
        ctx.kind, ctx.apply_softmax = kind, bool(apply_softmax)
        return sums

    @staticmethod
    def backward(ctx, g):
        source, target = ctx.saved_tensors
        N, C = source.shape[:2]
        V = source[0, 0].numel()
        g = _f32(g, "grad_sums")
        gS, gT, gI = g[:, 0].contiguous(), g[:, 1].contiguous(), g[:, 2].contiguous()
        need_t = ctx.kind == 2 and ctx.needs_input_grad[1]
        g_src = torch.empty_like(source) if ctx.needs_input_grad[0] else None
        g_tgt = torch.empty_like(target) if need_t else None
        if g_src is not None or g_tgt is not None:
            _lib.call("da_dice_sums_bwd", _p(source), _p(target), ctx.kind, int(ctx.apply_softmax), N, C, V,
                      _p(gS), _p(gT), _p(gI), _p(g_src), _p(g_tgt), _stream())
        return g_src, g_tgt, None


def dice_sums(source, target, apply_softmax=False):
    return DiceSumsFunction.apply(source, target, apply_softmax)


class SoftmaxDiceFunction(torch.autograd.Function):
    """(sums[N,3,C], probs[N,C,...]) = Dice sums of softmax(logits) against ``target`` and the softmax itself, one pass;
    the backward folds the gradient arriving at ``probs`` into the same pass (no second softmax backward, no sum of
    two full-size gradients)."""

    @staticmethod
    def forward(ctx, logits, target):
        logits = _f32(logits, "source")
        N, C = logits.shape[:2]
        V = logits[0, 0].numel()
        if not target.is_cuda:
            raise RuntimeError("deepatlas_b200: 'target' must be a CUDA tensor")
        if target.is_floating_point():
            target = _f32(target, "target")
            if target.numel() != logits.numel():
                raise ValueError("dice: soft target must have the shape of source")
            kind = 2
        else:
            if target.dtype not in _KIND:
                target = target.long()
            target = target.contiguous()
            if target.numel() != N * V:
                raise ValueError("dice: label target must have N*D*H*W elements")
            kind = _KIND[target.dtype]
        sums = torch.empty((N, 3, C), dtype=torch.float32, device=logits.device)
        probs = torch.empty_like(logits)
        nb = _lib.size("da_dice_workspace_bytes", N, C, V)
        ws = _ws(nb, logits.device)
        _lib.call("da_softmax_dice_fwd", _p(logits), _p(target), kind, N, C, V, _p(sums), _p(probs), _p(ws), nb, _stream())
        if kind == 2 and ctx.needs_input_grad[1]:
            # the fused backward forms only the logits gradient; DiceSumsFunction handles differentiable soft targets
            raise RuntimeError("deepatlas_b200: softmax_dice does not differentiate a soft target; "
                               "use dice_sums(source, target, apply_softmax=True) for a target that requires grad")
        ctx.save_for_backward(logits, target)
        ctx.kind = kind
        return sums, probs

    @staticmethod
    def backward(ctx, g, gp):
        logits, target = ctx.saved_tensors
        if not ctx.needs_input_grad[0]:
            return None, None
        N, C = logits.shape[:2]
        V = logits[0, 0].numel()
        g = torch.zeros((N, 3, C), device=logits.device) if g is None else _f32(g, "grad_sums")
        gS, gT, gI = g[:, 0].contiguous(), g[:, 1].contiguous(), g[:, 2].contiguous()
        gp = _f32(gp, "grad_probs") if gp is not None else None
        gx = torch.empty_like(logits)
        _lib.call("da_softmax_dice_bwd", _p(logits), _p(target), ctx.kind, N, C, V, _p(gS), _p(gT), _p(gI), _p(gp), _p(gx),
                  _stream())
        return gx, None


def softmax_dice(logits, target):
    return SoftmaxDiceFunction.apply(logits, target)


def head_dice_supported(K: int, C: int, V: int) -> bool:
    return bool(_lib.size("da_head_dice_supported", int(K), int(C), int(V)))


class HeadSoftmaxDiceFunction(torch.autograd.Function):
    """(sums[N,3,C], probs or None) of softmax(conv1x1(feat, weight, bias)) against a label ``target``: the U-Net's
    1x1x1 class head (unets.py:250) fused with softmax and the Dice sums (loss.py:427-476), so that neither the logits
    nor their gradient are written to HBM.  The backward recomputes the logits from the 16 features and forms the
    feature, weight and bias gradients in one pass."""

    @staticmethod
    def forward(ctx, feat, weight, bias, target, want_probs: bool):
        feat = _f32(feat, "features")
        weight = _f32(weight, "weight")
        bias = _f32(bias, "bias") if bias is not None else None
        N, K = feat.shape[:2]
        C = weight.shape[0]
        V = feat[0, 0].numel()
        if weight.numel() != C * K:
            raise ValueError("head_dice: weight must be (C, K, 1, 1, 1)")
        if not target.is_cuda or target.is_floating_point():
            raise RuntimeError("deepatlas_b200: head_dice needs a CUDA label target")
        if target.dtype not in _KIND:
            target = target.long()
        target = target.contiguous()
        if target.numel() != N * V:
            raise ValueError("dice: label target must have N*D*H*W elements")
        kind = _KIND[target.dtype]
        sums = torch.empty((N, 3, C), dtype=torch.float32, device=feat.device)
        probs = torch.empty((N, C) + tuple(feat.shape[2:]), dtype=torch.float32, device=feat.device) if want_probs else None
        nb = _lib.size("da_head_dice_workspace_bytes", N, C, V)
        ws = _ws(nb, feat.device)
        _lib.call("da_head_dice_fwd", _p(feat), _p(weight), _p(bias), _p(target), kind, N, K, C, V, _p(sums), _p(probs), _p(ws),
                  nb, _stream())
        ctx.save_for_backward(feat, weight, bias, target)
        ctx.kind = kind
        ctx.grad_slot = _GRAD_SLOT
        if probs is None:
            ctx.mark_non_differentiable()
            return sums, None
        return sums, probs

    @staticmethod
    def backward(ctx, g, gp):
        feat, weight, bias, target = ctx.saved_tensors
        N, K = feat.shape[:2]
        C = weight.shape[0]
        V = feat[0, 0].numel()
        g = torch.zeros((N, 3, C), device=feat.device) if g is None else _f32(g, "grad_sums")
        gS, gI = g[:, 0].contiguous(), g[:, 2].contiguous()
        gp = _f32(gp, "grad_probs") if gp is not None else None
        gfeat = torch.empty_like(feat)
        gw_dst = _grad_dst(weight, ctx.needs_input_grad[1], ctx.grad_slot)
        gb_dst = _grad_dst(bias, ctx.needs_input_grad[2], ctx.grad_slot) if bias is not None else None
        inplace = gw_dst is not None and (bias is None or gb_dst is not None)
        gw = gw_dst if inplace else torch.empty_like(weight)
        gb = (gb_dst if inplace else torch.empty_like(bias)) if bias is not None else None
        nb = _lib.size("da_head_dice_workspace_bytes", N, C, V)
        ws = _ws(nb, feat.device)
        _lib.call("da_head_dice_bwd", _p(feat), _p(weight), _p(bias), _p(target), ctx.kind, N, K, C, V, _p(gS), _p(gI), _p(gp),
                  _p(gfeat), _p(gw), _p(gb), int(inplace), _p(ws), nb, _stream())
        if inplace:
            gw = gb = None
        return gfeat, gw, gb, None, None


def head_softmax_dice(feat, weight, bias, target, want_probs=False):
    return HeadSoftmaxDiceFunction.apply(feat, weight, bias, target, want_probs)


class WarpedDiceSumsFunction(torch.autograd.Function):
    """sums[N,3,C] of dice(grid_sample(prob, phi), onehot(labels)) without materialising the warped map: the anatomy
    term of the joint step; see csrc/warp_dice.cu for the structure of the backward."""

    @staticmethod
    def forward(ctx, prob, phi, labels, add_identity: bool, deterministic: bool = False):
        prob, phi = _f32(prob, "prob"), _f32(phi, "phi")
        N, C, D, H, W = prob.shape
        if phi.dim() != 5 or phi.shape[0] != N or phi.shape[1] != 3:
            raise ValueError(f"warped_dice_sums: phi must be (N,3,Do,Ho,Wo), got {tuple(phi.shape)}")
        Do, Ho, Wo = phi.shape[2:]
        if not labels.is_cuda or labels.is_floating_point():
            raise RuntimeError("deepatlas_b200: 'labels' must be an integer CUDA tensor")
        if labels.dtype not in _KIND:
            labels = labels.long()
        labels = labels.contiguous()
        if labels.numel() != N * Do * Ho * Wo:
            raise ValueError("warped_dice_sums: labels must have N*Do*Ho*Wo elements")
        sums = torch.empty((N, 3, C), dtype=torch.float32, device=prob.device)
        nb = _lib.size("da_warp_dice_fwd_workspace_bytes", N, C, Do * Ho * Wo, D * H * W)
        ws = _ws(nb, prob.device)
        # Wsum: the scattered trilinear weights; produced here, consumed again by the backward
        wsum = None if deterministic else torch.empty((N, D, H, W), dtype=torch.float32, device=prob.device)
        _lib.call("da_warp_dice_sums_fwd", _p(prob), _p(phi), int(add_identity), _p(labels), _KIND[labels.dtype], N, C,
                  D, H, W, Do, Ho, Wo, _p(sums), _p(wsum), _p(ws), nb, _stream())
        ctx.save_for_backward(prob, phi, labels, wsum)
        ctx.add_identity = bool(add_identity)
        return sums

    @staticmethod
    def backward(ctx, g):
        prob, phi, labels, wsum = ctx.saved_tensors
        N, C, D, H, W = prob.shape
        Do, Ho, Wo = phi.shape[2:]
        g = _f32(g, "grad_sums")
        gS, gI = g[:, 0].contiguous(), g[:, 2].contiguous()
        g_prob = torch.empty_like(prob) if ctx.needs_input_grad[0] else None
        g_phi = torch.empty_like(phi) if ctx.needs_input_grad[1] else None
        if g_prob is None and g_phi is None:
            return None, None, None, None, None
        nb = _lib.size("da_warp_dice_bwd_workspace_bytes", N, D * H * W)
        ws = _ws(nb, prob.device)
        _lib.call("da_warp_dice_sums_bwd", _p(prob), _p(phi), int(ctx.add_identity), _p(labels), _KIND[labels.dtype],
                  _p(gS), _p(gI), _p(wsum), N, C, D, H, W, Do, Ho, Wo, _p(g_prob), _p(g_phi), _p(ws), nb, _stream())
        return g_prob, g_phi, None, None, None


def warped_dice_sums(prob, phi, labels, add_identity=False, deterministic=False):
    """deterministic=True: gather kernel with a fixed summation order for S (about 5x slower forward)."""
    return WarpedDiceSumsFunction.apply(prob, phi, labels, add_identity, deterministic)


class SoftmaxFunction(torch.autograd.Function):
    """Channel softmax (F.softmax(dim=1)) for the anatomy branch, where the probabilities get warped."""

    @staticmethod
    def forward(ctx, x):
        x = _f32(x, "x")
        N, C = x.shape[:2]
        y = torch.empty_like(x)
        _lib.call("da_softmax_fwd", _p(x), _p(y), N, C, x[0, 0].numel(), _stream())
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = _f32(dy, "grad_out")
        N, C = y.shape[:2]
        dx = torch.empty_like(y)
        _lib.call("da_softmax_bwd", _p(y), _p(dy), _p(dx), N, C, y[0, 0].numel(), _stream())
        return dx


def softmax(x):
    return SoftmaxFunction.apply(x)


# ------------------------------------------------------------------------------------------------------
# local NCC  (lib/loss.py:597-617)
# ------------------------------------------------------------------------------------------------------
class LnccFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, I, J, win: int, eps: float):
        I, J = _f32(I, "I"), _f32(J, "J")
        if I.shape != J.shape or I.dim() != 5 or I.shape[1] != 1:
            raise ValueError(f"lncc: I and J must both be (N,1,D,H,W), got {tuple(I.shape)} / {tuple(J.shape)}")
        N, _, D, H, W = I.shape
        need = (1 if ctx.needs_input_grad[0] else 0) | (2 if ctx.needs_input_grad[1] else 0)
        loss = torch.empty((), dtype=torch.float32, device=I.device)
        coef = _ws(_lib.size("da_lncc_coef_bytes", N, D, H, W, win, need), I.device) if need else None
        nb = _lib.size("da_lncc_fwd_workspace_bytes", N, D, H, W, win)
        ws = _ws(nb, I.device)
        _lib.call("da_lncc_fwd", _p(I), _p(J), N, D, H, W, win, float(eps), need, _p(loss), _p(coef), _p(ws),
                  nb, _stream())
        ctx.save_for_backward(I, J, coef if coef is not None else torch.empty(0, device=I.device))
        ctx.win, ctx.need = win, need
        return loss

    @staticmethod
    def backward(ctx, g):
        I, J, coef = ctx.saved_tensors
        N, _, D, H, W = I.shape
        g = _f32(g, "grad_out").reshape(1)
        nb = _lib.size("da_lncc_bwd_workspace_bytes", N, D, H, W, ctx.win)
        ws = _ws(nb, I.device)
        gI = gJ = None
        slot = 0
        if ctx.need & 1:
            gI = torch.empty_like(I)
            _lib.call("da_lncc_bwd", _p(I), _p(J), _p(g), _p(coef), slot, 0, N, D, H, W, ctx.win, _p(gI),
                      _p(ws), nb, _stream())
            slot += 1
        if ctx.need & 2:
            gJ = torch.empty_like(J)
            _lib.call("da_lncc_bwd", _p(I), _p(J), _p(g), _p(coef), slot, 1, N, D, H, W, ctx.win, _p(gJ),
                      _p(ws), nb, _stream())
        return gI, gJ, None, None


def lncc(I, J, win=9, eps=1e-6):
    return LnccFunction.apply(I, J, win, eps)


# ------------------------------------------------------------------------------------------------------
# bending energy sums  (lib/loss.py:702-718)
# ------------------------------------------------------------------------------------------------------
class BendingSumsFunction(torch.autograd.Function):
    """sums[N,3,6]: per channel the interior sum of squared (``l1``: absolute) (ddD, ddH, ddW, dDdH, dHdW, dDdW)."""

    @staticmethod
    def forward(ctx, u, l1: bool):
        u = _f32(u, "input")
        if u.dim() != 5 or u.shape[1] != 3:
            raise ValueError(f"bending: input must be (N,3,D,H,W), got {tuple(u.shape)}")
        N, _, D, H, W = u.shape
        sums = torch.empty((N, 3, 6), dtype=torch.float32, device=u.device)
        nb = _lib.size("da_bending_fwd_workspace_bytes", N)
        ws = _ws(nb, u.device)
        _lib.call("da_bending_fwd_ex", _p(u), N, D, H, W, int(l1), _p(sums), _p(ws), nb, _stream())
        ctx.save_for_backward(u)
        ctx.l1 = bool(l1)
        return sums

    @staticmethod
    def backward(ctx, g):
        (u,) = ctx.saved_tensors
        N, _, D, H, W = u.shape
        g = _f32(g, "grad_sums")
        nb = _lib.size("da_bending_bwd_workspace_bytes", N, D, H, W)
        ws = _ws(nb, u.device)
        gu = torch.empty_like(u)
        _lib.call("da_bending_bwd_ex", _p(u), _p(g), N, D, H, W, int(ctx.l1), _p(gu), _p(ws), nb, _stream())
        return gu, None


def bending_sums(u, l1=False):
    return BendingSumsFunction.apply(u, l1)


# ------------------------------------------------------------------------------------------------------
# conv3d  (lib/network_factory/unets.py:30,36,98,250; modules.py:48; voxel_morph.py:57)
# ------------------------------------------------------------------------------------------------------
def _conv_out(n, k, s, p):
    return (n + 2 * p - k) // s + 1


class Conv3dFunction(torch.autograd.Function):
    """out = act(conv3d(cat(x1, x2), weight) + bias); weight in nn.Conv3d layout, or nn.ConvTranspose3d
    layout when ``transposed`` (k3 s1 p1 only).  act: None or the leaky slope (0.0 = ReLU)."""

    @staticmethod
    def forward(ctx, x1, x2, weight, bias, transposed: bool, stride: int, pad: int, slope):
        x1, weight = _f32(x1, "x1"), _f32(weight, "weight")
        x2 = _f32(x2, "x2") if x2 is not None else None
        bias = _f32(bias, "bias") if bias is not None else None
        N, C1, Di, Hi, Wi = x1.shape
        C2 = x2.shape[1] if x2 is not None else 0
        if x2 is not None and (x2.shape[0] != N or tuple(x2.shape[2:]) != (Di, Hi, Wi)):
            raise ValueError("conv3d: x1 and x2 must share batch and spatial extents")
        ks = weight.shape[2]
        Cout, Cin = (weight.shape[1], weight.shape[0]) if transposed else (weight.shape[0], weight.shape[1])
        if Cin != C1 + C2:
            raise ValueError(f"conv3d: weight expects {Cin} input channels, got {C1}+{C2}")
        Do, Ho, Wo = (_conv_out(n, ks, stride, pad) for n in (Di, Hi, Wi))
        out = torch.empty((N, Cout, Do, Ho, Wo), dtype=torch.float32, device=x1.device)
        nb = _lib.size("da_conv3d_pack_bytes", Cin, Cout, ks)
        ws = _ws(nb, x1.device)
        # max|input|: the tensor-core kernels scale their fp16 operand pairs by it; computed once here (the call fills
        # the slot) and reused by the weight gradient in backward
        a1 = _get_amax(x1)
        a2 = _get_amax(x2) if x2 is not None else None
        if a1 is not None and (x2 is None or a2 is not None):
            amax_x, x_valid = (a1 if x2 is None else torch.maximum(a1, a2)), 1   # a producer's bound: no pass over the input
        else:
            amax_x, x_valid = torch.empty((1,), dtype=torch.float32, device=x1.device), 0
        _lib.call("da_conv3d_fwd_ex", _p(x1), C1, _p(x2), C2, _p(weight), int(transposed), _p(bias), _p(out), N, Di,
                  Hi, Wi, Cout, ks, stride, pad, 0 if slope is None else 1, 0.0 if slope is None else float(slope),
                  _p(ws), nb, _stream(), _p(amax_x), x_valid)
        ctx.amax_x = amax_x
        ctx.save_for_backward(x1, x2, weight, out if slope is not None else None)
        ctx.bias_ref = bias   # (only its .grad is looked at in backward)
        ctx.grad_slot = _GRAD_SLOT
        ctx.cfg = (bool(transposed), ks, stride, pad, slope, bias is not None, Cout)
        return out

    @staticmethod
    def backward(ctx, dy):
        x1, x2, weight, out = ctx.saved_tensors
        transposed, ks, stride, pad, slope, has_bias, Cout = ctx.cfg
        dy = _f32(dy, "grad_out")
        N, C1, Di, Hi, Wi = x1.shape
        C2 = x2.shape[1] if x2 is not None else 0
        Cin = C1 + C2
        st = _stream()
        if slope is not None:
            g = torch.empty_like(dy)
            _lib.call("da_act_bwd", _p(dy), _p(out), float(slope), dy.numel(), _p(g), st)
            dy = g
        dx1 = dx2 = dw = db = None
        nb = _lib.size("da_conv3d_dgrad_workspace_bytes", N, Cin, Cout, Di, Hi, Wi, ks, stride)
        ws = _ws(nb, dy.device)
        amax_dy = _get_amax(dy)   # the producer's bound (batch-norm backward), else filled by the first call that takes it
        dy_valid = 1
        if amax_dy is None:
            amax_dy, dy_valid = torch.empty((1,), dtype=torch.float32, device=dy.device), 0
        want_w = ctx.needs_input_grad[2] or (has_bias and ctx.needs_input_grad[3])
        gw_dst = gb_dst = None
        inplace = False
        if want_w:
            bias = ctx.bias_ref
            gw_dst = _grad_dst(weight, ctx.needs_input_grad[2], ctx.grad_slot)
            gb_dst = _grad_dst(bias, has_bias and ctx.needs_input_grad[3], ctx.grad_slot) if has_bias else None
            inplace = gw_dst is not None and (not has_bias or gb_dst is not None)

        def wgrad(stream, valid):
            dw_ = gw_dst if inplace else torch.empty_like(weight)
            db_ = (gb_dst if inplace else torch.empty((Cout,), dtype=torch.float32, device=dy.device)) if has_bias else None
            nbw = _lib.size("da_conv3d_wgrad_workspace_bytes", Cin, Cout, ks)
            wsw = _ws(nbw, dy.device)
            _lib.call("da_conv3d_wgrad_ex", _p(x1), C1, _p(x2), C2, _p(dy), int(transposed), _p(dw_), _p(db_), N, Di, Hi,
                      Wi, Cout, ks, stride, pad, _p(wsw), nbw, stream, _p(ctx.amax_x), 1, _p(amax_dy), valid, int(inplace))
            return (None, None) if inplace else (dw_, db_)   # in place: already added to weight.grad / bias.grad

        side_wgrad = want_w and inplace and _WGRAD_OVERLAP
        if side_wgrad:
            # off the chain: the weight gradient starts as soon as dy (and its max-abs bound) exist
            if not dy_valid:
                _lib.call("da_absmax", _p(dy), dy.numel(), None, 0, _p(amax_dy), st)
                dy_valid = 1
            cur, side = torch.cuda.current_stream(dy.device), _wgrad_side(dy.device)
            ev = torch.cuda.Event()
            ev.record(cur)
            side.wait_event(ev)
            with torch.cuda.stream(side):
                wgrad(_stream(), 1)
            _WGRAD_PENDING.add(dy.device.index)
            for t in (x1, x2, dy, ctx.amax_x, amax_dy):
                if t is not None:
                    t.record_stream(side)
        if ctx.needs_input_grad[0]:
            dx1 = torch.empty_like(x1)
            _lib.call("da_conv3d_dgrad_ex", _p(dy), _p(weight), int(transposed), _p(dx1), N, Cin, 0, C1, Cout, Di, Hi,
                      Wi, ks, stride, pad, _p(ws), nb, st, _p(amax_dy), dy_valid)
            dy_valid = 1
        if x2 is not None and ctx.needs_input_grad[1]:
            dx2 = torch.empty_like(x2)
            _lib.call("da_conv3d_dgrad_ex", _p(dy), _p(weight), int(transposed), _p(dx2), N, Cin, C1, C2, Cout, Di, Hi,
                      Wi, ks, stride, pad, _p(ws), nb, st, _p(amax_dy), dy_valid)
            dy_valid = 1
        if want_w and not side_wgrad:
            dw, db = wgrad(st, dy_valid)
        return dx1, dx2, dw, db, None, None, None, None


def conv3d(x1, weight, bias=None, x2=None, transposed=False, stride=1, pad=1, slope=None):
    return Conv3dFunction.apply(x1, x2, weight, bias, transposed, stride, pad, slope)


# ------------------------------------------------------------------------------------------------------
# batch norm (+ activation)  (lib/network_factory/unets.py:31-32,51)
# ------------------------------------------------------------------------------------------------------
class BnActFunction(torch.autograd.Function):
    """y = act(batch_norm(x)); training mode uses batch statistics and updates the running buffers in
    place (momentum, unbiased variance) exactly as nn.BatchNorm3d; eval mode uses the running buffers."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, training: bool, momentum: float, eps: float,
                slope):
        x = _f32(x, "x")
        N, C = x.shape[:2]
        V = x[0, 0].numel()
        dev = x.device
        st = _stream()
        nb = _lib.size("da_bn_workspace_bytes", C)
        if training:
            mean = torch.empty((C,), dtype=torch.float32, device=dev)
            invstd = torch.empty((C,), dtype=torch.float32, device=dev)
            ws = _ws(nb, dev)
            amax_y = torch.empty((1,), dtype=torch.float32, device=dev)
            _lib.call("da_bn_stats_ex", _p(x), N, C, V, float(eps), float(momentum), _p(mean), _p(invstd),
                      _p(running_mean), _p(running_var), _p(gamma), _p(beta), 0 if slope is None else 1,
                      0.0 if slope is None else float(slope), _p(amax_y), _p(ws), nb, st)
        else:
            amax_y = None
            mean = running_mean.detach().float().contiguous()
            invstd = torch.rsqrt(running_var.detach().float() + eps).contiguous()
        y = torch.empty_like(x)
        _lib.call("da_bn_act_fwd", _p(x), _p(mean), _p(invstd), _p(gamma), _p(beta), N, C, V,
                  0 if slope is None else 1, 0.0 if slope is None else float(slope), _p(y), st)
        ctx.save_for_backward(x, mean, invstd, gamma, beta)
        ctx.cfg = (bool(training), slope)
        ctx.grad_slot = _GRAD_SLOT
        # the bound is attached to y by bn_act() (attributes set here do not reach the tensor apply() returns); the batch
        # statistics go out too, for callers that apply the running-statistics update themselves (deferred_bn_updates)
        stats_out = (mean, invstd) if training else (None, None)
        ctx.mark_non_differentiable(*[t for t in (amax_y,) + stats_out if t is not None])
        return (y, amax_y) + stats_out

    @staticmethod
    def backward(ctx, dy, _g_amax=None, _g_mean=None, _g_invstd=None):
        x, mean, invstd, gamma, beta = ctx.saved_tensors
        training, slope = ctx.cfg
        dy = _f32(dy, "grad_out")
        N, C = x.shape[:2]
        V = x[0, 0].numel()
        dx = torch.empty_like(x)
        g_dst = _grad_dst(gamma, gamma is not None and ctx.needs_input_grad[1], ctx.grad_slot)
        b_dst = _grad_dst(beta, beta is not None and ctx.needs_input_grad[2], ctx.grad_slot)
        inplace = g_dst is not None and b_dst is not None
        dg = g_dst if inplace else torch.empty_like(mean)
        db = b_dst if inplace else torch.empty_like(mean)
        nb = _lib.size("da_bn_workspace_bytes", C)
        ws = _ws(nb, x.device)
        amax_dx = torch.empty((1,), dtype=torch.float32, device=x.device)
        _lib.call("da_bn_act_bwd_ex", _p(dy), _p(x), _p(mean), _p(invstd), _p(gamma), _p(beta), N, C, V, int(training),
                  0 if slope is None else 1, 0.0 if slope is None else float(slope), _p(dx), _p(dg), _p(db), _p(amax_dx),
                  int(inplace), _p(ws), nb, _stream())
        _set_amax(dx, amax_dx)
        if inplace:
            dg = db = None   # already added to gamma.grad / beta.grad
        return dx, (dg if gamma is not None else None), (db if beta is not None else None), None, None, None, None, None, None


def bn_act(x, gamma, beta, running_mean, running_var, training=True, momentum=0.1, eps=1e-5, slope=None, return_stats=False):
    """``return_stats``: also return the batch mean and 1/sqrt(var + eps) (training mode), for a caller that passed
    ``running_mean = running_var = None`` and applies the running-statistics update itself."""
    y, amax, mean, invstd = BnActFunction.apply(x, gamma, beta, running_mean, running_var, training, momentum, eps, slope)
    y = _set_amax(y, amax)
    return (y, mean, invstd) if return_stats else y


# ------------------------------------------------------------------------------------------------------
# max-pool 2, nearest upsample, deconv k2 s2
# ------------------------------------------------------------------------------------------------------
class MaxPool2Function(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _f32(x, "x")
        N, C, D, H, W = x.shape
        y = torch.empty((N, C, D // 2, H // 2, W // 2), dtype=torch.float32, device=x.device)
        _lib.call("da_maxpool2_fwd", _p(x), _p(y), N * C, D, H, W, _stream())
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        N, C, D, H, W = x.shape
        dy = _f32(dy, "grad_out")
        dx = torch.empty_like(x)
        _lib.call("da_maxpool2_bwd", _p(dy), _p(x), _p(dx), N * C, D, H, W, _stream())
        return _set_amax(dx, _get_amax(dy))   # dy values or zeros


def maxpool2(x):
    return _set_amax(MaxPool2Function.apply(x), _get_amax(x))   # a maximum of input values: the input's bound holds


class UpsampleNearestFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, size: Tuple[int, int, int]):
        x = _f32(x, "x")
        N, C, D, H, W = x.shape
        Do, Ho, Wo = (int(s) for s in size)
        y = torch.empty((N, C, Do, Ho, Wo), dtype=torch.float32, device=x.device)
        _lib.call("da_upsample_nearest_fwd", _p(x), _p(y), N * C, D, H, W, Do, Ho, Wo, _stream())
        ctx.shape = (N, C, D, H, W, Do, Ho, Wo)
        return y

    @staticmethod
    def backward(ctx, dy):
        N, C, D, H, W, Do, Ho, Wo = ctx.shape
        dy = _f32(dy, "grad_out")
        dx = torch.empty((N, C, D, H, W), dtype=torch.float32, device=dy.device)
        _lib.call("da_upsample_nearest_bwd", _p(dy), _p(dx), N * C, D, H, W, Do, Ho, Wo, _stream())
        return dx, None


def upsample_nearest(x, size: Sequence[int]):
    if tuple(x.shape[2:]) == tuple(size):
        return x
    return _set_amax(UpsampleNearestFunction.apply(x, tuple(size)), _get_amax(x))   # copies of input values


class DeconvK2S2Function(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        x, weight = _f32(x, "x"), _f32(weight, "weight")
        bias = _f32(bias, "bias") if bias is not None else None
        N, Cin, D, H, W = x.shape
        if weight.shape[0] != Cin or tuple(weight.shape[2:]) != (2, 2, 2):
            raise ValueError(f"deconv_k2s2: weight must be ({Cin},Cout,2,2,2), got {tuple(weight.shape)}")
        Cout = weight.shape[1]
        out = torch.empty((N, Cout, 2 * D, 2 * H, 2 * W), dtype=torch.float32, device=x.device)
        _lib.call("da_deconv_k2s2_fwd", _p(x), _p(weight), _p(bias), _p(out), N, Cin, Cout, D, H, W, _stream())
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        ctx.bias_ref = bias
        ctx.grad_slot = _GRAD_SLOT
        return out

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = _f32(dy, "grad_out")
        N, Cin, D, H, W = x.shape
        Cout = weight.shape[1]
        st = _stream()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            _lib.call("da_deconv_k2s2_dgrad", _p(dy), _p(weight), _p(dx), N, Cin, Cout, D, H, W, st)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            gw_dst = _grad_dst(weight, ctx.needs_input_grad[1], ctx.grad_slot)
            gb_dst = _grad_dst(ctx.bias_ref, ctx.has_bias and ctx.needs_input_grad[2], ctx.grad_slot) if ctx.has_bias else None
            inplace = gw_dst is not None and (not ctx.has_bias or gb_dst is not None)
            dw = gw_dst if inplace else torch.empty_like(weight)
            db = (gb_dst if inplace else torch.empty((Cout,), dtype=torch.float32, device=dy.device)) if ctx.has_bias else None
            nb = _lib.size("da_deconv_k2s2_wgrad_workspace_bytes", Cin, Cout)
            ws = _ws(nb, dy.device)
            _lib.call("da_deconv_k2s2_wgrad_ex", _p(x), _p(dy), _p(dw), _p(db), N, Cin, Cout, D, H, W, int(inplace), _p(ws), nb, st)
            if inplace:
                dw = db = None
        return dx, dw, db


def deconv_k2s2(x, weight, bias=None):
    return DeconvK2S2Function.apply(x, weight, bias)


# ------------------------------------------------------------------------------------------------------
# Conv3d kernel 2 stride 2 (maxpool=False down-sampler, lib/network_factory/unets.py:231): the adjoint of
# the k2 s2 deconvolution, so its three passes are the deconvolution's three kernels with roles swapped.
# ------------------------------------------------------------------------------------------------------
class ConvK2S2Function(torch.autograd.Function):
    """y = conv3d(x, weight (Cout,Cin,2,2,2), stride 2, padding 0) + bias; even extents only."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x, weight = _f32(x, "x"), _f32(weight, "weight")
        bias = _f32(bias, "bias") if bias is not None else None
        N, Cin, D, H, W = x.shape
        if weight.shape[1] != Cin or tuple(weight.shape[2:]) != (2, 2, 2):
            raise ValueError(f"conv_k2s2: weight must be (Cout,{Cin},2,2,2), got {tuple(weight.shape)}")
        if D % 2 or H % 2 or W % 2:
            raise RuntimeError("deepatlas_b200: conv k2 s2 is built for even extents")
        Cout = weight.shape[0]
        st = _stream()
        y = torch.empty((N, Cout, D // 2, H // 2, W // 2), dtype=torch.float32, device=x.device)
        # deconv data gradient with (Cin_deconv, Cout_deconv) = (Cout, Cin): y[co] = sum x[ci, 2z+a, ..] w[co, ci, a, ..]
        _lib.call("da_deconv_k2s2_dgrad", _p(x), _p(weight), _p(y), N, Cout, Cin, D // 2, H // 2, W // 2, st)
        if bias is not None:  # y = (y - 0) * 1 + bias through the normalisation kernel, in place
            zero = torch.zeros(Cout, device=x.device)
            one = torch.ones(Cout, device=x.device)
            _lib.call("da_bn_act_fwd", _p(y), _p(zero), _p(one), None, _p(bias), N, Cout, y[0, 0].numel(), 0, 0.0, _p(y), st)
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = _f32(dy, "grad_out")
        N, Cin, D, H, W = x.shape
        Cout = weight.shape[0]
        st = _stream()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            _lib.call("da_deconv_k2s2_fwd", _p(dy), _p(weight), None, _p(dx), N, Cout, Cin, D // 2, H // 2, W // 2, st)
        if ctx.needs_input_grad[1]:
            dw = torch.empty_like(weight)
            nb = _lib.size("da_deconv_k2s2_wgrad_workspace_bytes", Cout, Cin)
            ws = _ws(nb, dy.device)
            _lib.call("da_deconv_k2s2_wgrad", _p(dy), _p(x), _p(dw), None, N, Cout, Cin, D // 2, H // 2, W // 2, _p(ws), nb, st)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = torch.empty((Cout,), dtype=torch.float32, device=dy.device)
            nb = _lib.size("da_channel_sum_workspace_bytes", Cout)
            ws = _ws(nb, dy.device)
            _lib.call("da_channel_sum", _p(dy), N, Cout, dy[0, 0].numel(), _p(db), _p(ws), nb, st)
        return dx, dw, db


def conv_k2s2(x, weight, bias=None):
    return ConvK2S2Function.apply(x, weight, bias)


# ------------------------------------------------------------------------------------------------------
# trilinear x2 up-sampling, residual add  (lib/network_factory/unets.py:236,264,275)
# ------------------------------------------------------------------------------------------------------
class UpsampleTrilinear2Function(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _f32(x, "x")
        N, C, D, H, W = x.shape
        y = torch.empty((N, C, 2 * D, 2 * H, 2 * W), dtype=torch.float32, device=x.device)
        _lib.call("da_upsample_trilinear2_fwd", _p(x), _p(y), N * C, D, H, W, _stream())
        ctx.shape = (N, C, D, H, W)
        return y

    @staticmethod
    def backward(ctx, dy):
        N, C, D, H, W = ctx.shape
        dy = _f32(dy, "grad_out")
        dx = torch.empty((N, C, D, H, W), dtype=torch.float32, device=dy.device)
        _lib.call("da_upsample_trilinear2_bwd", _p(dy), _p(dx), N * C, D, H, W, _stream())
        return dx


def upsample_trilinear2(x):
    return UpsampleTrilinear2Function.apply(x)


class AddFunction(torch.autograd.Function):
    """a + b for a [N,Ca,...] and b [N,Cb,...] with Cb == Ca or Cb == 1 (channel broadcast, as torch's `+`)."""

    @staticmethod
    def forward(ctx, a, b):
        a, b = _f32(a, "a"), _f32(b, "b")
        if a.dim() != b.dim() or a.shape[0] != b.shape[0] or tuple(a.shape[2:]) != tuple(b.shape[2:]) or \
                b.shape[1] not in (a.shape[1], 1):
            raise RuntimeError(f"The size of tensor a {tuple(a.shape)} must match the size of tensor b "
                               f"{tuple(b.shape)} (equal, or b with a single channel)")
        N, Ca = a.shape[:2]
        V = a[0, 0].numel()
        out = torch.empty_like(a)
        _lib.call("da_add_bcast", _p(a), _p(b), N, Ca, b.shape[1], V, _p(out), _stream())
        ctx.cfg = (N, Ca, b.shape[1], V, tuple(b.shape))
        return out

    @staticmethod
    def backward(ctx, g):
        N, Ca, Cb, V, bshape = ctx.cfg
        ga = g if ctx.needs_input_grad[0] else None
        gb = None
        if ctx.needs_input_grad[1]:
            if Cb == Ca:
                gb = g
            else:
                g = _f32(g, "grad_out")
                gb = torch.empty(bshape, dtype=torch.float32, device=g.device)
                _lib.call("da_channel_reduce", _p(g), N, Ca, V, _p(gb), _stream())
        return ga, gb


def add(a, b):
    """``a + b`` of the residual variants; the operand with fewer channels is broadcast."""
    if b.shape[1] > a.shape[1]:
        a, b = b, a
    return AddFunction.apply(a, b)


# ------------------------------------------------------------------------------------------------------
# remaining registry losses  (lib/loss.py:96-186, 485-501, 625-671, 733-736)
# ------------------------------------------------------------------------------------------------------
class PairMomentsFunction(torch.autograd.Function):
    """m[N,9] = (sum a, sum b, sum a^2, sum b^2, sum ab, sum (a-b)^2, sum (a-ma)^2, sum (b-mb)^2, sum (a-ma)(b-mb))
    per sample; b may be None.  The backward is one affine pass per input."""

    @staticmethod
    def forward(ctx, a, b):
        a = _f32(a, "input")
        if b is not None:
            b = _f32(b, "target")
            if b.shape != a.shape:
                raise RuntimeError(f"The size of tensor a {tuple(a.shape)} must match the size of tensor b {tuple(b.shape)}")
        N = a.shape[0]
        V = a[0].numel()
        out = torch.empty((N, 9), dtype=torch.float32, device=a.device)
        nb = _lib.size("da_pair_moments_workspace_bytes", N)
        ws = _ws(nb, a.device)
        _lib.call("da_pair_moments_fwd", _p(a), _p(b), N, V, _p(out), _p(ws), nb, _stream())
        ctx.save_for_backward(a, b, out)
        return out

    @staticmethod
    def backward(ctx, g):
        a, b, m = ctx.saved_tensors
        N = a.shape[0]
        V = a[0].numel()
        coef_a, coef_b = pair_moment_coefs(g.float(), m, V)
        st = _stream()
        ga = gb = None
        if ctx.needs_input_grad[0]:
            ga = torch.empty_like(a)
            _lib.call("da_affine2", _p(a), _p(b), _p(coef_a), N, V, _p(ga), st)
        if b is not None and ctx.needs_input_grad[1]:
            gb = torch.empty_like(b)
            _lib.call("da_affine2", _p(b), _p(a), _p(coef_b), N, V, _p(gb), st)
        return ga, gb


def pair_moment_coefs(g, m, V):
    """Per-sample coefficients of the affine gradients of the nine moments: grad_a = ca[:,0]*a + ca[:,1]*b + ca[:,2] and
    grad_b = cb[:,0]*b + cb[:,1]*a + cb[:,2], for upstream gradients g [N,9] and moments m [N,9].
    d/da_i: Sa 1 | Saa 2a | Sab b | Sdd 2(a-b) | cAA 2(a-ma) | cAB (b-mb); symmetrically for b."""
    ma, mb = m[:, 0] / V, m[:, 1] / V
    ca = torch.stack([2 * g[:, 2] + 2 * g[:, 5] + 2 * g[:, 6],
                      g[:, 4] - 2 * g[:, 5] + g[:, 8],
                      g[:, 0] - 2 * g[:, 6] * ma - g[:, 8] * mb], dim=1).contiguous()
    cb = torch.stack([2 * g[:, 3] + 2 * g[:, 5] + 2 * g[:, 7],
                      g[:, 4] - 2 * g[:, 5] + g[:, 8],
                      g[:, 1] - 2 * g[:, 7] * mb - g[:, 8] * ma], dim=1).contiguous()
    return ca, cb


def pair_moments(a, b=None):
    return PairMomentsFunction.apply(a, b)


class GradientSumsFunction(torch.autograd.Function):
    """sums[N,C,3] of gradientLoss (lib/loss.py:655-659): f(u[d+2]-u[d]), f(u[h+2]+u[h]), f(u[w+2]+u[w])."""

    @staticmethod
    def forward(ctx, u, l1: bool):
        u = _f32(u, "input")
        if u.dim() != 5:
            raise ValueError(f"gradient loss: input must be (N,C,D,H,W), got {tuple(u.shape)}")
        N, C, D, H, W = u.shape
        sums = torch.empty((N, C, 3), dtype=torch.float32, device=u.device)
        nb = _lib.size("da_gradient_loss_workspace_bytes", N, C)
        ws = _ws(nb, u.device)
        _lib.call("da_gradient_loss_fwd", _p(u), N, C, D, H, W, int(l1), _p(sums), _p(ws), nb, _stream())
        ctx.save_for_backward(u)
        ctx.l1 = bool(l1)
        return sums

    @staticmethod
    def backward(ctx, g):
        (u,) = ctx.saved_tensors
        N, C, D, H, W = u.shape
        g = _f32(g, "grad_sums")
        gu = torch.empty_like(u)
        _lib.call("da_gradient_loss_bwd", _p(u), _p(g), N, C, D, H, W, int(ctx.l1), _p(gu), _stream())
        return gu, None


def gradient_sums(u, l1=False):
    return GradientSumsFunction.apply(u, l1)


class XentFunction(torch.autograd.Function):
    """out2 = (sum of per-voxel terms, sum of weights) of the channel log-softmax losses; mode 0 cross entropy,
    1 focal, 2 soft cross entropy (log_softmax), 3 soft cross entropy (log of clamped probabilities)."""

    @staticmethod
    def forward(ctx, x, target, mode: int, class_weight, gamma: float, focal_softmax: bool, ignore_index: int):
        x = _f32(x, "input")
        N, C = x.shape[:2]
        V = x[0, 0].numel()
        if not target.is_cuda:
            raise RuntimeError("deepatlas_b200: 'target' must be a CUDA tensor")
        if mode >= 2:
            target = _f32(target, "target")
            if target.shape != x.shape:
                raise ValueError("soft cross entropy: target must have the shape of the prediction")
            kind = 2
        else:
            if target.is_floating_point():
                raise RuntimeError("deepatlas_b200: class-index target expected (integer dtype)")
            if target.dtype not in _KIND:
                target = target.long()
            target = target.contiguous()
            if target.numel() != N * V:
                raise ValueError("Expected target of N*D*H*W class indices")
            kind = _KIND[target.dtype]
        cw = _f32(class_weight, "class weight").reshape(-1) if class_weight is not None else None
        if cw is not None and cw.numel() != C:
            raise ValueError("class weight must have one entry per class")
        out2 = torch.empty((2,), dtype=torch.float32, device=x.device)
        nb = _lib.size("da_xent_workspace_bytes", N)
        ws = _ws(nb, x.device)
        _lib.call("da_xent_fwd", _p(x), _p(target), kind, mode, N, C, V, _p(cw), float(gamma), int(focal_softmax),
                  int(ignore_index), _p(out2), _p(ws), nb, _stream())
        ctx.save_for_backward(x, target, cw)
        ctx.cfg = (kind, mode, float(gamma), int(focal_softmax), int(ignore_index))
        return out2

    @staticmethod
    def backward(ctx, g):
        x, target, cw = ctx.saved_tensors
        kind, mode, gamma, fsm, ign = ctx.cfg
        N, C = x.shape[:2]
        V = x[0, 0].numel()
        gscale = g[0:1].float().contiguous()
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gt = torch.empty_like(target) if (mode >= 2 and ctx.needs_input_grad[1]) else None
        if gx is None and gt is None:
            return (None,) * 7
        if mode < 2 and gx is None:
            return (None,) * 7
        _lib.call("da_xent_bwd", _p(x), _p(target), kind, mode, N, C, V, _p(cw), gamma, fsm, ign, _p(gscale), _p(gx), _p(gt),
                  _stream())
        return gx, gt, None, None, None, None, None


def xent_sums(x, target, mode, class_weight=None, gamma=0.0, focal_softmax=True, ignore_index=-100):
    return XentFunction.apply(x, target, mode, class_weight, gamma, focal_softmax, ignore_index)


class LnccMsSumFunction(torch.autograd.Function):
    """One scale of the multi-scale LNCCLoss (lib/loss.py:545-586): sum over all windows (k^3 ones filter, dilation,
    stride, no padding) of cross^2 / (Ivar * Jvar + 1e-5), as a 1-element tensor."""

    @staticmethod
    def forward(ctx, I, J, k: int, dil: int, stride: int):
        I, J = _f32(I, "input"), _f32(J, "target")
        if I.shape != J.shape or I.dim() != 5 or I.shape[1] != 1:
            raise ValueError(f"LNCCLoss: input and target must both be (N,1,D,H,W), got {tuple(I.shape)} / {tuple(J.shape)}")
        N, _, D, H, W = I.shape
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        out = torch.empty((1,), dtype=torch.float32, device=I.device)
        coef = _ws(_lib.size("da_lncc_ms_coef_bytes", N, D, H, W, k, dil, stride), I.device) if need else None
        nb = _lib.size("da_lncc_ms_workspace_bytes")
        ws = _ws(nb, I.device)
        _lib.call("da_lncc_ms_fwd", _p(I), _p(J), N, D, H, W, k, dil, stride, _p(out), _p(coef), _p(ws), nb, _stream())
        ctx.save_for_backward(I, J, coef if coef is not None else torch.empty(0, device=I.device))
        ctx.cfg = (k, dil, stride)
        return out

    @staticmethod
    def backward(ctx, g):
        I, J, coef = ctx.saved_tensors
        k, dil, stride = ctx.cfg
        N, _, D, H, W = I.shape
        g = _f32(g, "grad_out").reshape(1)
        gI = gJ = None
        if ctx.needs_input_grad[0]:
            gI = torch.empty_like(I)
            _lib.call("da_lncc_ms_bwd", _p(I), _p(J), _p(coef), 0, _p(g), 1.0, N, D, H, W, k, dil, stride, 0, _p(gI), _stream())
        if ctx.needs_input_grad[1]:
            gJ = torch.empty_like(J)
            _lib.call("da_lncc_ms_bwd", _p(I), _p(J), _p(coef), 1, _p(g), 1.0, N, D, H, W, k, dil, stride, 0, _p(gJ), _stream())
        return gI, gJ, None, None, None


def lncc_ms_sum(I, J, k, dil, stride):
    return LnccMsSumFunction.apply(I, J, int(k), int(dil), int(stride))


# ------------------------------------------------------------------------------------------------------
# device-side input stage  (lib/transforms.py:79-80, 124-158)
# ------------------------------------------------------------------------------------------------------
def crop_clip(image, lo_corner, size, clip=(0.0, 1.0)):
    """image [..., D, H, W] fp32 on the device -> clip(image[..., z0:z0+Do, y0:y0+Ho, x0:x0+Wo], *clip)."""
    image = _f32(image, "image")
    D, H, W = image.shape[-3:]
    NC = image.numel() // (D * H * W)
    out = torch.empty(tuple(image.shape[:-3]) + tuple(size), dtype=torch.float32, device=image.device)
    _lib.call("da_crop_clip_f32", _p(image), _p(out), NC, D, H, W, *[int(v) for v in lo_corner], *[int(v) for v in size],
              float(clip[0]), float(clip[1]), _stream())
    return out


def crop_labels(labels, lo_corner, size):
    """uint8 label map [..., D, H, W] on the device -> the cropped window, still uint8 (read directly by the losses)."""
    if not labels.is_cuda or labels.dtype != torch.uint8:
        raise RuntimeError("deepatlas_b200: 'labels' must be a CUDA uint8 tensor")
    labels = labels.contiguous()
    D, H, W = labels.shape[-3:]
    NC = labels.numel() // (D * H * W)
    out = torch.empty(tuple(labels.shape[:-3]) + tuple(size), dtype=torch.uint8, device=labels.device)
    _lib.call("da_crop_u8", _p(labels), _p(out), NC, D, H, W, *[int(v) for v in lo_corner], *[int(v) for v in size], _stream())
    return out
