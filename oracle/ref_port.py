"""CPU restatement ("port") of the DeepAtlas volumetric hot path.  TEST INFRASTRUCTURE ONLY.

This file is the parity oracle: it restates, on the CPU and on top of the same third-party
arithmetic the reference itself calls (PyTorch ATen: conv3d / conv_transpose3d / batch_norm /
max_pool3d / interpolate / grid_sample / softmax), what each reference function on the hot path
computes.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import it.  The product (``deepatlas_b200``) never does.

Pinning: the reference has no tests or golden vectors of its own, so this port
is pinned by (i) ``tests/test_oracle_vs_reference.py`` which imports the real reference modules
from /root/reference (build container only) and demands bit-identical results from this port, and
(ii) ``tests/golden/*.npz`` produced by ``oracle/make_golden.py`` from those same reference
modules under torch 2.11.0+cu128 (CPU), which travel to the GPU box.

Every function is functional (weights come in as a ``state_dict`` with the reference's key names)
and dtype-generic: call it with float64 tensors for the "truth" rung of the precision ladder.
All citations are file:line under /root/reference.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

# --------------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------------


def _act(x: torch.Tensor, act: str) -> torch.Tensor:
    """nn.ReLU / nn.LeakyReLU(0.01) as selected at lib/network_factory/unets.py:5-6,32."""
    if act == "ReLU":
        return F.relu(x)
    if act == "LeakyReLU":
        return F.leaky_relu(x, 0.01)
    raise ValueError(act)


def _bn_train(x, sd: SD, prefix: str, training: bool, stats_out: Optional[dict]):
    """nn.BatchNorm3d (lib/network_factory/unets.py:31,51): train mode normalises with the batch
    mean / biased variance (eps 1e-5) and moves running stats with momentum 0.1 using the unbiased
    variance.  ``stats_out`` receives the updated running buffers (state is not mutated)."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    rm, rv = sd[prefix + ".running_mean"].clone(), sd[prefix + ".running_var"].clone()
    y = F.batch_norm(x, rm, rv, w, b, training=training, momentum=0.1, eps=1e-5)
    if stats_out is not None:
        stats_out[prefix + ".running_mean"], stats_out[prefix + ".running_var"] = rm, rv
    return y


def unet_conv_block(x, sd: SD, prefix: str, bn: bool, act: str, training=True, stats_out=None,
                    named=True):
    """unets.convBlock (lib/network_factory/unets.py:24-39): Conv3d k3 s1 p1 -> [BN] -> act.
    ``named`` selects the child naming: 'conv'/'BN' (generator) or '0'/'1' (UNet.encoder :108-122)."""
    c, b_ = ("conv", "BN") if named else ("0", "1")
    y = F.conv3d(x, sd[f"{prefix}.{c}.weight"], sd.get(f"{prefix}.{c}.bias"), stride=1, padding=1)
    if bn:
        y = _bn_train(y, sd, f"{prefix}.{b_}", training, stats_out)
    return _act(y, act)


def unet_deconv_block(x, sd: SD, prefix: str, bn: bool, act: str, kernel: int, stride: int,
                      padding: int, training=True, stats_out=None, named=True):
    """unets.deconvBlock (lib/network_factory/unets.py:42-58) / UNet.decoder (:124-137):
    ConvTranspose3d -> [BN] -> act."""
    c, b_ = ("deconv", "BN") if named else ("0", "1")
    y = F.conv_transpose3d(x, sd[f"{prefix}.{c}.weight"], sd.get(f"{prefix}.{c}.bias"),
                           stride=stride, padding=padding)
    if bn:
        y = _bn_train(y, sd, f"{prefix}.{b_}", training, stats_out)
    return _act(y, act)


# --------------------------------------------------------------------------------------------
# networks
# --------------------------------------------------------------------------------------------

UNET_LIGHT_CFG = dict(  # lib/network_factory/__init__.py:12-15
    encoders=[(8, 16), (16, 16, 32), (32, 32, 64), (64, 64, 64)],
    decoders=[(64, 64, 64), (64, 32, 32), (32, 16, 16)],
    act="LeakyReLU")


def unet_generator_forward(x, sd: SD, in_channel: int, bn: bool, cfg=UNET_LIGHT_CFG, training=True,
                           stats_out=None):
    """UNetTemplate.forward (lib/network_factory/unets.py:259-278; construction :222-252) for every
    variant: cfg['maxpool'] (default True; False = Conv3d k2 s2 down-samplers, :231), cfg['upsample']
    (default False; True = nn.Upsample(scale_factor=2, mode='trilinear'), :236) and cfg['res']
    (default False; True = ``enc(x) + x`` / ``dec(cat) + x``, :264,275)."""
    encs, decs, act = cfg["encoders"], cfg["decoders"], cfg["act"]
    maxpool, upsample, res = cfg.get("maxpool", True), cfg.get("upsample", False), cfg.get("res", False)
    skips = []
    for i, enc in enumerate(encs):
        chans = ((in_channel,) + tuple(enc)) if i == 0 else tuple(enc)
        y = x
        for k in range(len(chans) - 1):
            y = unet_conv_block(y, sd, f"encoders.{i}.{k}", bn, act, training, stats_out)
        x = (y + x) if res else y
        if i < len(encs) - 1:
            skips.append(x)
            if maxpool:
                x = F.max_pool3d(x, 2)
            else:
                x = F.conv3d(x, sd[f"down_samplers.{i}.weight"], sd.get(f"down_samplers.{i}.bias"), stride=2)
    n_inner = len(tuple(encs[-1])) - 1  # unets.py:247 re-uses the encoder loop variable
    for j, dec in enumerate(decs):
        if upsample:
            x = F.interpolate(x, scale_factor=2, mode="trilinear")
        else:
            x = unet_deconv_block(x, sd, f"up_samplers.{j}", bn, act, 2, 2, 0, training, stats_out)
        y = torch.cat((x, skips.pop()), dim=1)
        for k in range(n_inner):
            y = unet_conv_block(y, sd, f"decoders.decBlock{j}.{k}", bn, act, training, stats_out)
        if j == len(decs) - 1:
            p = f"decoders.decBlock{j}.{n_inner}"
            y = F.conv3d(y, sd[p + ".weight"], sd.get(p + ".bias"))
        x = (y + x) if res else y
    return x


def unet_forward(x, sd: SD, bn: bool, training=True, stats_out=None):
    """UNet.forward (lib/network_factory/unets.py:139-179).  Child Sequentials are unnamed, so keys
    are ec0.0.weight / ec0.1.* (BN).  dc8/dc7/dc5/dc4/dc2/dc1 are ConvTranspose3d k3 s1 p1."""
    def ec(name, t):
        return unet_conv_block(t, sd, name, bn, "ReLU", training, stats_out, named=False)

    def dc(name, t, k, s, p):
        return unet_deconv_block(t, sd, name, bn, "ReLU", k, s, p, training, stats_out, named=False)

    syn0 = ec("ec1", ec("ec0", x))
    syn1 = ec("ec3", ec("ec2", F.max_pool3d(syn0, 2)))
    syn2 = ec("ec5", ec("ec4", F.max_pool3d(syn1, 2)))
    e7 = ec("ec7", ec("ec6", F.max_pool3d(syn2, 2)))
    d7 = dc("dc7", dc("dc8", torch.cat((dc("dc9", e7, 2, 2, 0), syn2), 1), 3, 1, 1), 3, 1, 1)
    d4 = dc("dc4", dc("dc5", torch.cat((dc("dc6", d7, 2, 2, 0), syn1), 1), 3, 1, 1), 3, 1, 1)
    d1 = dc("dc1", dc("dc2", torch.cat((dc("dc3", d4, 2, 2, 0), syn0), 1), 3, 1, 1), 3, 1, 1)
    return F.conv3d(d1, sd["dc0.weight"], sd.get("dc0.bias"))


def identity_transform(size: Sequence[int], dtype=torch.float32, device=None) -> torch.Tensor:
    """get_identity_transform (lib/utils.py:89-102): 3xDxHxW, channel 0 = x (W axis), 1 = y (H),
    2 = z (D); values k/(n-1)*2-1.  The reference builds it in float32 (arange().float())."""
    D, H, W = size
    ax = [torch.arange(0, n, device=device).float() / (n - 1) * 2.0 - 1 for n in (D, H, W)]
    zz, yy, xx = torch.meshgrid(ax, indexing="ij")
    return torch.stack([xx, yy, zz]).to(dtype)


def warp(source, deform_field):
    """The spatial transformer call at lib/network_factory/voxel_morph.py:90-91."""
    return F.grid_sample(source, deform_field.permute(0, 2, 3, 4, 1), mode="bilinear",
                         padding_mode="zeros", align_corners=True)


def _vm_block(x, sd: SD, prefix: str, stride: int):
    """modules.convBlock as VoxelMorph uses it (lib/network_factory/modules.py:28-62,
    voxel_morph.py:44-55): Conv3d k3 p1 bias -> ReLU, no BN, no residual."""
    return _act(F.conv3d(x, sd[prefix + ".conv.weight"], sd.get(prefix + ".conv.bias"),
                         stride=stride, padding=1), "ReLU")


def voxelmorph_forward(source, target, sd: SD):
    """VoxelMorphCVPR2018.forward (lib/network_factory/voxel_morph.py:62-92) with the default
    filters.  F.interpolate(size=) is the default 'nearest' mode."""
    x1 = _vm_block(torch.cat((source, target), 1), sd, "encoders.0", 1)
    x2 = _vm_block(x1, sd, "encoders.1", 2)
    x3 = _vm_block(x2, sd, "encoders.2", 2)
    x4 = _vm_block(x3, sd, "encoders.3", 2)
    x5 = _vm_block(x4, sd, "encoders.4", 2)
    d1 = _vm_block(F.interpolate(x5, size=x4.shape[2:]), sd, "decoders.0", 1)
    d2 = _vm_block(F.interpolate(torch.cat((d1, x4), 1), size=x3.shape[2:]), sd, "decoders.1", 1)
    d3 = _vm_block(F.interpolate(torch.cat((d2, x3), 1), size=x2.shape[2:]), sd, "decoders.2", 1)
    d4 = _vm_block(torch.cat((d3, x2), 1), sd, "decoders.3", 1)
    d5 = _vm_block(F.interpolate(d4, size=x1.shape[2:]), sd, "decoders.4", 1)
    disp = F.conv3d(torch.cat((d5, x1), 1), sd["flow.weight"], sd["flow.bias"], padding=1)
    deform = disp + identity_transform(source.shape[2:], dtype=disp.dtype, device=disp.device)
    return disp, warp(source, deform), deform


# --------------------------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------------------------


def mask_to_one_hot(mask, n_classes: int, dtype=torch.float32):
    """lib/transforms.py:675-689 (float32 zeros + scatter_ of ones along dim 1)."""
    shape = list(mask.shape)
    shape[1] = n_classes
    return torch.zeros(shape, dtype=dtype, device=mask.device).scatter_(1, mask.long(), 1)


def dice_multiclass(source, target, n_class: int, weight_type="Simple", no_bg=False,
                    softmax=False, eps=1e-7):
    """DiceLossMultiClass.forward (lib/loss.py:410-476)."""
    B, C = source.shape[:2]
    if softmax:
        source = F.softmax(source, dim=1)
    s = source.reshape(B, C, -1)
    if target.dim() == source.dim() - 1:
        t = mask_to_one_hot(target.reshape(B, 1, -1), n_class, dtype=source.dtype)
    elif target.shape[1] == C:
        t = target.reshape(B, C, -1)
    else:
        raise ValueError("Incorrect size of target tensor")
    if no_bg:
        s, t = s[:, 1:], t[:, 1:]
    sv, tv = s.sum(2), t.sum(2)
    if weight_type == "Simple":
        w = (tv ** (1.0 / 3.0) + eps).reciprocal()
    elif weight_type == "Volume":
        w = (tv + eps).reciprocal()
        tmp = torch.where(torch.isinf(w), torch.ones_like(w), w)
        w = torch.where(torch.isinf(w), torch.ones_like(w) * tmp.max(dim=1, keepdim=True)[0], w)
    elif weight_type == "Uniform":
        w = torch.ones(B, C - int(no_bg), dtype=source.dtype, device=source.device)
    else:
        raise ValueError(weight_type)
    w = w / w.max()
    inter = (s * t).sum(2)
    scores = (2.0 * inter + eps) / ((sv + tv) + 2 * eps)
    return 1 - (w * scores).sum() / w.sum()


def lncc(I, J, filter_size=9, eps=1e-6):
    """VoxelMorphLNCC.forward (lib/loss.py:597-617): five valid box-filter sums via F.conv3d with
    a ones kernel, then the reference's cancellation-form cross/variance expressions."""
    n = filter_size ** 3
    k = torch.ones(1, 1, filter_size, filter_size, filter_size, dtype=I.dtype, device=I.device)
    Is, Js = F.conv3d(I, k), F.conv3d(J, k)
    I2s, J2s, IJs = F.conv3d(I * I, k), F.conv3d(J * J, k), F.conv3d(I * J, k)
    Im, Jm = Is / n, Js / n
    cross = IJs - Im * Js - Jm * Is + Im * Jm * n
    Iv = I2s - 2 * Im * Is + Im ** 2 * n
    Jv = J2s - 2 * Jm * Js + Jm ** 2 * n
    return 1 - ((cross ** 2) / (Iv * Jv + eps)).mean()


def bending_energy(u, spacing=(1.0, 1.0, 1.0), normalize=True, norm="L2"):
    """BendingEnergyLoss.forward (lib/loss.py:687-730).  norm='L2': note the per-CHANNEL scale: ``spatial_dims``
    (D,H,W)/min multiplies the (B,3) per-channel means (reference quirk, kept).  Any other norm: the `if` at :721 is
    skipped and :729 takes plain means of the absolute second differences."""
    if norm != "L2":
        B, C = u.shape[:2]
        c = u[:, :, 1:-1, 1:-1, 1:-1]
        a = lambda t: t.abs().reshape(B, C, -1)  # noqa: E731
        ddx = a(u[:, :, 2:, 1:-1, 1:-1] + u[:, :, :-2, 1:-1, 1:-1] - 2 * c)
        ddy = a(u[:, :, 1:-1, 2:, 1:-1] + u[:, :, 1:-1, :-2, 1:-1] - 2 * c)
        ddz = a(u[:, :, 1:-1, 1:-1, 2:] + u[:, :, 1:-1, 1:-1, :-2] - 2 * c)
        dxdy = a(u[:, :, 2:, 2:, 1:-1] + u[:, :, :-2, :-2, 1:-1] - u[:, :, 2:, :-2, 1:-1] - u[:, :, :-2, 2:, 1:-1])
        dydz = a(u[:, :, 1:-1, 2:, 2:] + u[:, :, 1:-1, :-2, :-2] - u[:, :, 1:-1, 2:, :-2] - u[:, :, 1:-1, :-2, 2:])
        dxdz = a(u[:, :, 2:, 1:-1, 2:] + u[:, :, :-2, 1:-1, :-2] - u[:, :, 2:, 1:-1, :-2] - u[:, :, :-2, 1:-1, 2:])
        return (ddx.mean() + ddy.mean() + ddz.mean() + 2 * dxdy.mean() + 2 * dydz.mean() + 2 * dxdz.mean()) / 9.0
    sp = torch.tensor(spacing, dtype=torch.float32, device=u.device)
    if normalize:
        sp = sp / sp.min()
    sp = sp.to(u.dtype)
    dims = torch.tensor(u.shape[2:], dtype=torch.float32, device=u.device)
    if normalize:
        dims = dims / dims.min()
    dims = dims.to(u.dtype)
    B, C = u.shape[:2]
    c = u[:, :, 1:-1, 1:-1, 1:-1]

    def m(t):
        return (t.abs().reshape(B, C, -1) ** 2).mean(2)

    ddx = m(u[:, :, 2:, 1:-1, 1:-1] + u[:, :, :-2, 1:-1, 1:-1] - 2 * c) * (dims * sp / sp[0] ** 2) ** 2
    ddy = m(u[:, :, 1:-1, 2:, 1:-1] + u[:, :, 1:-1, :-2, 1:-1] - 2 * c) * (dims * sp / sp[1] ** 2) ** 2
    ddz = m(u[:, :, 1:-1, 1:-1, 2:] + u[:, :, 1:-1, 1:-1, :-2] - 2 * c) * (dims * sp / sp[2] ** 2) ** 2
    dxdy = m(u[:, :, 2:, 2:, 1:-1] + u[:, :, :-2, :-2, 1:-1] - u[:, :, 2:, :-2, 1:-1]
             - u[:, :, :-2, 2:, 1:-1]) * (dims * sp / (sp[0] * sp[1])) ** 2
    dydz = m(u[:, :, 1:-1, 2:, 2:] + u[:, :, 1:-1, :-2, :-2] - u[:, :, 1:-1, 2:, :-2]
             - u[:, :, 1:-1, :-2, 2:]) * (dims * sp / (sp[1] * sp[2])) ** 2
    dxdz = m(u[:, :, 2:, 1:-1, 2:] + u[:, :, :-2, 1:-1, :-2] - u[:, :, 2:, 1:-1, :-2]
             - u[:, :, :-2, 1:-1, 2:]) * (dims * sp / (sp[2] * sp[0])) ** 2
    return (ddx.mean() + ddy.mean() + ddz.mean() + 2 * dxdy.mean() + 2 * dydz.mean()
            + 2 * dxdz.mean()) / 9.0


def dice_on_label(source, target, n_class=None, eps=10e-6, weight_type="Uniform"):
    """DiceLossOnLabel.forward (lib/loss.py:358-391): two label maps (B,1,D,M,N), background dropped."""
    if n_class is None:
        n_class = int(max(torch.unique(target).max(), torch.unique(source).max()).long().item()) + 1
    B, C1 = target.shape[:2]
    s1 = mask_to_one_hot(source.reshape(B, C1, -1), n_class)[:, 1:, :]
    t1 = mask_to_one_hot(target.reshape(B, C1, -1), n_class)[:, 1:, :]
    sv, tv = s1.sum(2), t1.sum(2)
    if weight_type == "Simple":
        w = tv.float().reciprocal()
        w = torch.where(torch.isinf(w), torch.ones_like(w), w)
    else:
        w = torch.ones(B, C1)
    inter = s1 * t1
    scores = (2.0 * inter.sum(2).float() * w) / (w * (sv.float() + tv.float()) + eps)
    return 1 - scores.mean()


def lncc_multiscale(I, J):
    """LNCCLoss.forward (lib/loss.py:512-586): the scale schedule of __stepup, F.conv3d box filters with dilation and
    stride, the reference's cancellation-form cross / variance expressions, eps 1e-5."""
    ms = min(I.shape[2:])
    if ms > 128:
        scale, weight, dilation = [int(ms / 16), int(ms / 8), int(ms / 4)], [0.1, 0.3, 0.6], [2, 2, 2]
    elif ms > 64:
        scale, weight, dilation = [int(ms / 4), int(ms / 2)], [0.3, 0.7], [2, 2]
    else:
        scale, weight, dilation = [int(ms / 2)], [1.0], [1]
    total = 0.0
    for k, w, d in zip(scale, weight, dilation):
        step = max(int((k + 1) / 4), 1)
        f = torch.ones(1, 1, k, k, k, dtype=I.dtype)
        conv = lambda t: F.conv3d(t, f, padding=0, dilation=d, stride=step).view(I.shape[0], -1)  # noqa: E731
        Is, Js, I2s, J2s, IJs = conv(I), conv(J), conv(I ** 2), conv(J ** 2), conv(I * J)
        n = float(k ** 3)
        Im, Jm = Is / n, Js / n
        cross = IJs - Jm * Is - Im * Js + Jm * Im * n
        Iv = I2s - 2 * Im * Is + Im ** 2 * n
        Jv = J2s - 2 * Jm * Js + Jm ** 2 * n
        total = total + (1 - (cross * cross / (Iv * Jv + 1e-5)).mean()) * w
    return total


def ncc_loss(a, b):
    """NormalizedCrossCorrelationLoss.forward (lib/loss.py:493-501)."""
    a = a.reshape(a.shape[0], -1)
    b = b.reshape(b.shape[0], -1)
    am = a - torch.mean(a, 1, keepdim=True)
    bm = b - torch.mean(b, 1, keepdim=True)
    ncc = (am * bm).mean(1) / (torch.sqrt((am ** 2).mean(1)) * torch.sqrt((bm ** 2).mean(1)))
    return 1 - ncc.mean()


def mse_loss(a, b):
    """nn.MSELoss() (registry 'mse', lib/loss.py:742) == lib/loss.py:504-509."""
    return ((a - b) ** 2).mean()


def l2_loss(a):
    """L2Loss.forward (lib/loss.py:733-736)."""
    return (a ** 2).mean()


def gradient_loss(u, norm="L2", spacing=(1.0, 1.0, 1.0), normalize=True):
    """gradientLoss.forward (lib/loss.py:636-671), with the `+` of the H and W differences (:657,659) and the
    per-channel scale (:663-665) as they are."""
    sp = torch.tensor(spacing, dtype=torch.float32)
    if normalize:
        sp = sp / sp.min()
    sp = sp.to(u.dtype)
    dims = torch.tensor(u.shape[2:], dtype=torch.float32)
    if normalize:
        dims = dims / dims.min()
    dims = dims.to(u.dtype)
    B, C = u.shape[:2]
    dx = torch.abs(u[:, :, 2:, :, :] - u[:, :, :-2, :, :]).reshape(B, C, -1)
    dy = torch.abs(u[:, :, :, 2:, :] + u[:, :, :, :-2, :]).reshape(B, C, -1)
    dz = torch.abs(u[:, :, :, :, 2:] + u[:, :, :, :, :-2]).reshape(B, C, -1)
    if norm == "L2":
        dx = (dx ** 2).mean(2) * (dims * sp / sp[0]) ** 2
        dy = (dy ** 2).mean(2) * (dims * sp / sp[1]) ** 2
        dz = (dz ** 2).mean(2) * (dims * sp / sp[2]) ** 2
    return (dx.mean() + dy.mean() + dz.mean()) / 3.0


def focal_loss(inputs, targets, alpha=None, gamma=2, size_average=True, soft_max=True):
    """FocalLoss.forward (lib/loss.py:149-186).  ``F.nll_loss(P, t)`` returns -P[t], hence (1 + P[t])**gamma."""
    C = inputs.shape[1]
    # lib/loss.py:167-169: five axes are permuted to (voxels, classes); (observations, classes) inputs pass as they are
    x = inputs.permute(0, 2, 3, 4, 1).contiguous().view(-1, C) if (inputs.dim() > 2 and targets.dim() > 1) else inputs
    t = targets.reshape(-1).long()
    P = F.softmax(x, dim=1) if soft_max else x
    a = torch.ones(C, 1, dtype=inputs.dtype) if alpha is None else alpha.to(inputs.dtype)
    a = a[t].view(-1)
    log_p = -F.cross_entropy(x, t, reduction="none")
    probs = F.nll_loss(P, t, reduction="none")
    batch_loss = -a * (torch.pow((1 - probs), gamma)) * log_p
    return batch_loss.mean() if size_average else batch_loss.sum()


def cross_entropy(x, t, weight=None, ignore_index=-100, reduction="mean"):
    """nn.CrossEntropyLoss (registry 'cross_entropy', lib/loss.py:748)."""
    return F.cross_entropy(x, t if t.is_floating_point() else t.long(), weight=weight, ignore_index=ignore_index,
                           reduction=reduction)


def soft_cross_entropy(pred, target, softmax=False):
    """SoftCrossEntropy.forward (lib/loss.py:112-116) for a class-probability target of the shape of pred.
    (The reference clamps ``pred`` in place when softmax is False; the value is the same.)"""
    if softmax:
        return torch.mean(torch.sum(-target * F.log_softmax(pred, 1), 1))
    return torch.mean(torch.sum(-target * torch.log(pred.clamp(min=1e-8)), 1))


def upsample_trilinear2_closed_form(x):
    """nn.Upsample(scale_factor=2, mode='trilinear') written out (align_corners=False, scale 1/2):
    src = max(0.5*(dst+0.5)-0.5, 0), i0 = floor(src), i1 = min(i0+1, n-1), weights (1-frac, frac) per axis."""
    def axis(n):
        dst = torch.arange(2 * n, dtype=torch.float64)
        src = (0.5 * (dst + 0.5) - 0.5).clamp(min=0)
        i0 = src.floor().long()
        i1 = (i0 + 1).clamp(max=n - 1)
        l1 = (src - i0).to(x.dtype)
        return i0, i1, 1 - l1, l1
    D, H, W = x.shape[2:]
    out = x
    for dim, n in ((2, D), (3, H), (4, W)):
        i0, i1, l0, l1 = axis(n)
        shape = [1] * 5
        shape[dim] = -1
        out = out.index_select(dim, i0) * l0.view(shape) + out.index_select(dim, i1) * l1.view(shape)
    return out


# --------------------------------------------------------------------------------------------
# independent closed forms (numpy-free, loop-free torch) used to cross-check the ATen calls above
# --------------------------------------------------------------------------------------------


def warp_closed_form(source, deform_field):
    """Trilinear sampling written out: ix=(x+1)/2*(W-1) etc., 8 corners,
    out-of-range corners contribute zero.  Independent of F.grid_sample."""
    B, C, D, H, W = source.shape
    g = deform_field
    ix = (g[:, 0] + 1) / 2 * (W - 1)
    iy = (g[:, 1] + 1) / 2 * (H - 1)
    iz = (g[:, 2] + 1) / 2 * (D - 1)
    x0, y0, z0 = ix.floor(), iy.floor(), iz.floor()
    out = torch.zeros(B, C, *g.shape[2:], dtype=source.dtype)
    flat = source.reshape(B, C, -1)
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                xi, yi, zi = x0 + dx, y0 + dy, z0 + dz
                wgt = ((1 - (ix - xi).abs()) * (1 - (iy - yi).abs()) * (1 - (iz - zi).abs()))
                ok = (xi >= 0) & (xi <= W - 1) & (yi >= 0) & (yi <= H - 1) & (zi >= 0) & (zi <= D - 1)
                idx = (zi.clamp(0, D - 1) * H + yi.clamp(0, H - 1)) * W + xi.clamp(0, W - 1)
                val = torch.gather(flat, 2, idx.long().reshape(B, 1, -1).expand(B, C, -1))
                out += val.reshape(out.shape) * (wgt * ok).unsqueeze(1)
    return out


def nearest_index(dst: int, n_in: int, n_out: int) -> int:
    """F.interpolate default 'nearest' source index rule: min(floor(dst*in/out), in-1)."""
    return min(int(math.floor(dst * (n_in / n_out))), n_in - 1)


# --------------------------------------------------------------------------------------------
# parameter initialisation in the reference's order (for the CPU baseline / parity runs)
# --------------------------------------------------------------------------------------------


def xavier_state_dict(shapes: Dict[str, Tuple[int, ...]], seed: int = 230) -> SD:
    """weights_init (lib/network_factory/unets.py:61-67, voxel_morph.py:94-101): xavier-normal conv
    weights, zero conv bias; BatchNorm keeps its constructor defaults (weight 1, bias 0,
    running_mean 0, running_var 1)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in shapes.items():
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros((), dtype=torch.long)
        elif len(shp) == 5:
            rf = shp[2] * shp[3] * shp[4]
            std = math.sqrt(2.0 / ((shp[0] + shp[1]) * rf))
            sd[k] = torch.randn(shp, generator=g) * std
        elif k.endswith("running_var") or (".BN.weight" in k) or k.endswith(".1.weight"):
            sd[k] = torch.ones(shp)
        else:
            sd[k] = torch.zeros(shp)
    return sd


# --------------------------------------------------------------------------------------------
# the joint seg+reg step, assembled from the restated reference functions above
# --------------------------------------------------------------------------------------------

DEFAULT_LAMBDAS = dict(sim=1.0, reg=1000.0, ana=1.0, sup=1.0)


def joint_loss(seg_sd: SD, reg_sd: SD, batch, n_classes: int, lambdas=None, dtype=torch.float32):
    """P_m=seg(I_m); P_t=seg(I_t); disp,I_w,phi=reg(I_m,I_t); S_w=warp(softmax(P_m),phi);
    L = l_sim*lncc(I_w,I_t) + l_reg*bending(disp) + l_ana*dice(S_w,onehot(S_t)) + l_sup*(dice(P_m,S_m)+dice(P_t,S_t)).
    seg = UNet_light(1, C, bias=True, BN=True) in train mode, reg = VoxelMorphCVPR2018()."""
    lam = dict(DEFAULT_LAMBDAS, **(lambdas or {}))
    I_m, S_m, I_t, S_t = batch
    I_m, I_t = I_m.to(dtype), I_t.to(dtype)
    C = n_classes
    P_m = unet_generator_forward(I_m, seg_sd, 1, True)
    P_t = unet_generator_forward(I_t, seg_sd, 1, True)
    disp, I_w, phi = voxelmorph_forward(I_m, I_t, reg_sd)
    S_w = warp(torch.softmax(P_m, 1), phi)
    onehot = mask_to_one_hot(S_t.reshape(1, 1, *S_t.shape[1:]), C, dtype=dtype)
    return (lam["sim"] * lncc(I_w, I_t) + lam["reg"] * bending_energy(disp)
            + lam["ana"] * dice_multiclass(S_w, onehot, C, "Uniform", False, False, 1e-6)
            + lam["sup"] * (dice_multiclass(P_m, S_m.long(), C, "Uniform", False, True, 1e-6)
                            + dice_multiclass(P_t, S_t.long(), C, "Uniform", False, True, 1e-6)))


# --------------------------------------------------------------------------------------------
# evaluation step
# --------------------------------------------------------------------------------------------


def eval_dice_per_class(pred_logits, truths, n_classes: int):
    """Inner loop of SegmentationExperiment.eval (models/segmentation.py:188-194) with metricEval('dice', ...,
    num_labels=2) = 1 - scipy.spatial.distance.dice on boolean arrays (lib/evalMetrics.py:58-68), one volume at a time.
    Returns (N, n_classes-1) float64 and the argmax label map."""
    import numpy as np
    import scipy.spatial.distance
    labels = torch.max(pred_logits, 1)[1]
    out = np.zeros((pred_logits.shape[0], n_classes - 1))
    for n in range(pred_logits.shape[0]):
        for c in range(1, n_classes):
            a, b = labels[n].numpy().reshape(-1) == c, truths[n].numpy().reshape(-1) == c
            with np.errstate(invalid="ignore", divide="ignore"):
                out[n, c - 1] = 1.0 - scipy.spatial.distance.dice(a, b)
    return torch.from_numpy(out), labels
