"""Host-side mirror of the reference's loss registry (lib/loss.py:739-750), all ten entries.

Same class names, constructor arguments, ``forward`` signatures and error behaviour as
``DiceLossMultiClass`` (lib/loss.py:397-476), ``VoxelMorphLNCC`` (:589-617), ``BendingEnergyLoss``
(:674-730) -- the hot path -- and ``NormalizedCrossCorrelationLoss`` (:485-501), ``gradientLoss`` (:625-671),
``L2Loss`` (:733-736), ``FocalLoss`` (:120-186), ``SoftCrossEntropy`` (:96-117), ``nn.MSELoss`` and
``nn.CrossEntropyLoss``.  The full-volume arithmetic runs in the CUDA library; only the closing formulas on the
handful of per-class / per-term sums are evaluated with torch ops (on device, no host sync).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class DiceLossMultiClass(nn.Module):
    """Dice loss between a probability map / logits and a label mask or probability map."""

    def __init__(self, n_class=None, weight_type="Simple", no_bg=False, softmax=False, eps=1e-7):
        super().__init__()
        self.weight_type, self.n_class, self.eps, self.no_bg, self.softmax = weight_type, n_class, eps, no_bg, softmax

    def forward(self, source, target):
        assert source.shape[0] == target.shape[0]
        shape = list(source.shape)
        if target.dim() == source.dim() - 1:
            pass  # label mask B x D x M x N (any integer dtype; uint8 is read directly)
        elif target.dim() == source.dim() and target.shape[1] == shape[1] and target.is_floating_point():
            pass  # class-wise probabilities B x C x D x M x N
        else:
            raise ValueError("Incorrect size of target tensor: {}, should be {} or []".format(
                target.shape, shape, shape[:1] + [1, ] + shape[2:]))
        assert tuple(source.shape[-3:]) == tuple(target.shape[-3:])
        if self.weight_type not in ("Simple", "Volume", "Uniform"):
            raise ValueError("Class weighting type {} does not exists!".format(self.weight_type))
        if self.n_class is None:
            self.n_class = shape[1]
        sums = ops.dice_sums(source, target, apply_softmax=self.softmax)  # (B, 3, C)
        return self._closing(sums)

    def forward_with_probs(self, source, target):
        """``(forward(source, target), softmax(source))`` for a softmax=True loss, from one pass over the logits: for a
        prediction whose probabilities have a second consumer (the joint step warps them for the anatomy term)."""
        if not self.softmax:
            raise ValueError("forward_with_probs needs softmax=True (the source must be logits)")
        assert source.shape[0] == target.shape[0] and tuple(source.shape[-3:]) == tuple(target.shape[-3:])
        if self.weight_type not in ("Simple", "Volume", "Uniform"):
            raise ValueError("Class weighting type {} does not exists!".format(self.weight_type))
        if self.n_class is None:
            self.n_class = source.shape[1]
        sums, probs = ops.softmax_dice(source, target)
        return self._closing(sums), probs

    def forward_head(self, features, head, target, want_probs=False):
        """``forward(head(features), target)`` (and ``softmax(head(features))`` with ``want_probs``) for a softmax=True
        loss, a 1x1x1 convolution ``head`` (unets.py:250) and a label-mask target, without materialising the logits:
        head, softmax and Dice sums run as one kernel each way.  Falls back to the separate calls where the fused kernel
        does not apply (head inputs != 16, more than 32 classes, an odd voxel count, a soft target)."""
        if not self.softmax:
            raise ValueError("forward_head needs softmax=True (the head produces logits)")
        N, K = features.shape[:2]
        V = features[0, 0].numel()
        C = head.weight.shape[0]
        fused = (not target.is_floating_point() and target.dim() == features.dim() - 1 and tuple(head.weight.shape[2:]) == (1, 1, 1)
                 and ops.head_dice_supported(K, C, V))
        if not fused:
            logits = head(features)
            if want_probs:
                return self.forward_with_probs(logits, target)
            return self.forward(logits, target), None
        assert features.shape[0] == target.shape[0] and tuple(features.shape[-3:]) == tuple(target.shape[-3:])
        if self.weight_type not in ("Simple", "Volume", "Uniform"):
            raise ValueError("Class weighting type {} does not exists!".format(self.weight_type))
        if self.n_class is None:
            self.n_class = C
        sums, probs = ops.head_softmax_dice(features, head.weight, head.bias, target, want_probs)
        return self._closing(sums), probs

    def forward_warped(self, source, deform_field, target):
        """``forward(grid_sample(source, deform_field), target)`` for a label-mask ``target`` and probability ``source``
        (the anatomy term of the joint step, same F.grid_sample call as voxel_morph.py:90-91) without materialising the
        warped C-channel map.  Falls back to warp + forward when the fused kernel does not apply."""
        if self.softmax or target.is_floating_point() or target.dim() != source.dim() - 1 or source.shape[1] > 32:
            return self.forward(ops.warp3d(source, deform_field, add_identity=False), target)
        if self.weight_type not in ("Simple", "Volume", "Uniform"):
            raise ValueError("Class weighting type {} does not exists!".format(self.weight_type))
        assert source.shape[0] == target.shape[0] and tuple(deform_field.shape[-3:]) == tuple(target.shape[-3:])
        if self.n_class is None:
            self.n_class = source.shape[1]
        return self._closing(ops.warped_dice_sums(source, deform_field, target, add_identity=False))

    def _closing(self, sums):
        """lib/loss.py:444-476 on the per-class sums (B, 3, C) = (source volume, target volume, intersection)."""
        if self.no_bg:
            sums = sums[:, :, 1:]
        source_volume, target_volume, intersection = sums[:, 0], sums[:, 1], sums[:, 2]
        if self.weight_type == "Simple":
            weights = (target_volume ** (1.0 / 3.0) + self.eps).reciprocal()
        elif self.weight_type == "Volume":
            weights = (target_volume + self.eps).reciprocal()
            temp = torch.where(torch.isinf(weights), torch.ones_like(weights), weights)
            max_w = temp.max(dim=1, keepdim=True)[0]
            weights = torch.where(torch.isinf(weights), torch.ones_like(weights) * max_w, weights)
        else:
            weights = torch.ones_like(target_volume)
        weights = weights / weights.max()
        scores = (2.0 * intersection + self.eps) / ((source_volume + target_volume) + 2 * self.eps)
        return 1 - (weights * scores).sum() / weights.sum()


class LNCCLoss(nn.Module):
    """Multi-scale local NCC (lib/loss.py:512-586; in the file but not in the registry).  The scale schedule of
    ``__stepup`` is reproduced: windows min(size)/16, /8, /4 (weights 0.1, 0.3, 0.6, dilation 2) above 128 voxels,
    /4 and /2 (0.3, 0.7, dilation 2) above 64, else /2 (dilation 1); stride max(int((k + 1) / 4), 1)."""

    def initialize(self, kernel_sz=[9, 9, 9], voxel_weights=None):
        pass

    @staticmethod
    def schedule(img_sz):
        max_scale = min(img_sz)
        if max_scale > 128:
            scale, weight, dilation = [int(max_scale / 16), int(max_scale / 8), int(max_scale / 4)], [0.1, 0.3, 0.6], [2, 2, 2]
        elif max_scale > 64:
            scale, weight, dilation = [int(max_scale / 4), int(max_scale / 2)], [0.3, 0.7], [2, 2]
        else:
            scale, weight, dilation = [int(max_scale / 2)], [1.0], [1]
        step = [max(int((k + 1) / 4), 1) for k in scale]
        return list(zip(scale, dilation, step, weight))

    def forward(self, input, target, inst_weights=None, train=None):
        N, _, D, H, W = input.shape
        total = 0.0
        for k, dil, step, weight in self.schedule(list(input.shape[2:])):
            count = N
            for n in (D, H, W):
                count *= (n - dil * (k - 1) - 1) // step + 1
            s = ops.lncc_ms_sum(input, target, k, dil, step)
            total = total + (1 - s[0] / count) * weight
        return total


class VoxelMorphLNCC(nn.Module):
    def __init__(self, filter_size=9, eps=1e-6):
        super().__init__()
        self.filter_size = filter_size
        self.win_numel = filter_size ** 3
        # kept for state_dict compatibility with the reference (lib/loss.py:594); the box filter itself is
        # a separable running sum inside the kernel and no (useless) filter gradient is produced.
        self.filter = nn.Parameter(torch.ones(1, 1, filter_size, filter_size, filter_size), requires_grad=False)
        self.eps = eps

    def forward(self, I, J):
        return ops.lncc(I, J, self.filter_size, self.eps)


class BendingEnergyLoss(nn.Module):
    """Bending energy of a 3-D displacement field: norm 'L2' (squared second differences with the reference's per-channel
    scale factors) or anything else (lib/loss.py:721 falls through: plain means of the absolute second differences)."""

    def __init__(self, norm="L2", spacing=(1, 1, 1), normalize=True):
        super().__init__()
        self.norm = norm
        self.spacing = torch.tensor(spacing).float()
        self.normalize = normalize
        if self.normalize:
            self.spacing = self.spacing / self.spacing.min()

    def _coef(self, shape, device):
        # cached per (shape, device): built on the host, and a pageable host-to-device copy per step would be a host
        # synchronisation (and is not capturable in a CUDA graph)
        key = (tuple(shape), str(device))
        cache = self.__dict__.setdefault("_coef_cache", {})
        if key not in cache:
            cache[key] = self._coef_host(shape).to(device)
        return cache[key]

    def _coef_host(self, shape):
        # lib/loss.py:694-696,721-729: scale_term[c] = (dims[c]*spacing[c]/denom_term)^2 (per CHANNEL, the
        # reference's quirk), mean over space, mean over (B,3), weights (1,1,1,2,2,2)/9
        B, _, D, H, W = shape
        sp = self.spacing.double()
        dims = torch.tensor([D, H, W], dtype=torch.float32)
        if self.normalize:
            dims = dims / dims.min()
        dims = dims.double()
        den = torch.stack([sp[0] ** 2, sp[1] ** 2, sp[2] ** 2, sp[0] * sp[1], sp[1] * sp[2], sp[2] * sp[0]])
        wt = torch.tensor([1.0, 1.0, 1.0, 2.0, 2.0, 2.0], dtype=torch.float64)
        interior = float((D - 2) * (H - 2) * (W - 2))
        scale = ((dims * sp)[:, None] / den[None, :]) ** 2          # (3 channels, 6 terms)
        coef = scale * wt[None, :] / (9.0 * 3.0 * B * interior)
        return coef.float()

    def forward(self, input):
        if self.norm != "L2":
            # lib/loss.py:696-718,729: |second differences|, mean over (B, 3, interior), weights (1,1,1,2,2,2) / 9
            B, _, D, H, W = input.shape
            sums = ops.bending_sums(input, l1=True)                  # (B, 3, 6)
            per_term = sums.sum(dim=(0, 1)) / float(3 * B * (D - 2) * (H - 2) * (W - 2))
            return (per_term[:3].sum() + 2.0 * per_term[3:].sum()) / 9.0
        sums = ops.bending_sums(input)                               # (B, 3, 6)
        return (sums * self._coef(input.shape, input.device)[None]).sum()


class DiceLossOnLabel(nn.Module):
    """Dice between two label maps (lib/loss.py:348-391; not in the registry, not differentiable): background excluded,
    ``scores = 2 I w / (w (S + T) + eps)``, ``1 - scores.mean()``.  Exact integer counts from one CUDA pass; as in the
    reference ``n_class`` is taken from the data on first use (one host sync) and then kept."""

    def __init__(self, n_class=None, eps=10e-6):
        super().__init__()
        self.n_class, self.eps = n_class, eps

    def forward(self, source, target, weight_type="Uniform", average=True):
        from .evaluation import label_overlap_counts
        assert source.shape == target.shape
        counts = label_overlap_counts(source, target)                        # (B, 3, 256)
        if self.n_class is None:
            present = (counts[:, :2].sum((0, 1)) > 0).nonzero()
            self.n_class = int(present.max().item()) + 1
        c = counts[:, :, 1:self.n_class].float()
        source_volume, target_volume, intersection = c[:, 0], c[:, 1], c[:, 2]
        if weight_type == "Simple":
            weights = target_volume.reciprocal()
            weights = torch.where(torch.isinf(weights), torch.ones_like(weights), weights)
        elif weight_type == "Uniform":
            weights = torch.ones(source.shape[0], source.shape[1], device=c.device)
        else:
            raise UnboundLocalError("cannot access local variable 'weights' where it is not associated with a value")
        scores = (2.0 * intersection * weights) / (weights * (source_volume + target_volume) + self.eps)
        return 1 - scores.mean()


class NormalizedCrossCorrelationLoss(nn.Module):
    """1 - NCC (lib/loss.py:485-501): per sample cov(a,b) / (std(a) std(b)) over all voxels, mean over the batch."""

    def forward(self, input, target):
        m = ops.pair_moments(input.reshape(input.shape[0], -1), target.reshape(target.shape[0], -1))
        V = float(input[0].numel())
        ncc = (m[:, 8] / V) / (torch.sqrt(m[:, 6] / V) * torch.sqrt(m[:, 7] / V))
        return 1 - ncc.mean()


class MSELoss(nn.Module):
    """nn.MSELoss (registry 'mse', lib/loss.py:742) and the reference's own MSELoss (lib/loss.py:504-509)."""

    def __init__(self, size_average=None, reduce=None, reduction="mean"):
        super().__init__()
        if size_average is not None or reduce is not None:
            reduction = "mean" if (size_average is None or size_average) and (reduce is None or reduce) else (
                "sum" if reduce is None or reduce else "none")
        if reduction not in ("mean", "sum"):
            raise NotImplementedError("deepatlas_b200: MSELoss is built for reduction 'mean' / 'sum'")
        self.reduction = reduction

    def forward(self, input, target):
        if input.shape != target.shape:
            raise RuntimeError(f"The size of tensor a {tuple(input.shape)} must match the size of tensor b "
                               f"{tuple(target.shape)}")
        total = ops.pair_moments(input.reshape(1, -1), target.reshape(1, -1))[0, 5]
        return total / input.numel() if self.reduction == "mean" else total


class L2Loss(nn.Module):
    """lib/loss.py:733-736: mean of squares."""

    def forward(self, input):
        return ops.pair_moments(input.reshape(1, -1))[0, 2] / input.numel()


class gradientLoss(nn.Module):
    """Spatial-gradient regulariser of a displacement field (lib/loss.py:625-671), including the reference's
    `+` in the H and W differences (lib/loss.py:657,659) and its per-CHANNEL scale factors (:663-665)."""

    def __init__(self, norm="L2", spacing=(1, 1, 1), normalize=True):
        super().__init__()
        self.norm = norm
        self.spacing = torch.tensor(spacing).float()
        self.normalize = normalize
        if self.normalize:
            self.spacing = self.spacing / self.spacing.min()

    def forward(self, input):
        N, C, D, H, W = input.shape
        sums = ops.gradient_sums(input, l1=self.norm != "L2")                     # (N, C, 3)
        counts = torch.tensor([(D - 2) * H * W, D * (H - 2) * W, D * H * (W - 2)], dtype=torch.float32)
        sp = self.spacing
        if self.norm == "L2":
            dims = torch.tensor([D, H, W], dtype=torch.float32)
            if self.normalize:
                dims = dims / dims.min()
            # (dx**2).mean(2) * (spatial_dims * spacing / spacing[k])**2 broadcasts the 3-vector over channels
            scale = torch.stack([(dims * sp / sp[k]) ** 2 for k in range(3)], dim=1)  # (channel, term)
            if C != 3:
                raise RuntimeError(f"The size of tensor a ({C}) must match the size of tensor b (3) at non-singleton "
                                   "dimension 1")
            coef = scale / counts[None, :] / (N * C)
        else:
            coef = (1.0 / counts / (N * C))[None, :].expand(C, 3)
        return (sums * coef.to(input.device)[None]).sum() / 3.0


class CrossEntropyLoss(nn.Module):
    """nn.CrossEntropyLoss (registry 'cross_entropy', lib/loss.py:748) for (N,C,D,H,W) scores and class-index
    targets (N,D,H,W): per-class ``weight``, ``ignore_index`` and reduction 'mean' / 'sum'."""

    def __init__(self, weight=None, size_average=None, ignore_index=-100, reduce=None, reduction="mean",
                 label_smoothing=0.0):
        super().__init__()
        if size_average is not None or reduce is not None:
            reduction = "mean" if (size_average is None or size_average) and (reduce is None or reduce) else (
                "sum" if reduce is None or reduce else "none")
        if reduction not in ("mean", "sum") or label_smoothing != 0.0:
            raise NotImplementedError("deepatlas_b200: CrossEntropyLoss is built for reduction 'mean' / 'sum' without "
                                      "label smoothing")
        self.register_buffer("weight", weight)
        self.ignore_index, self.reduction, self.label_smoothing = ignore_index, reduction, label_smoothing

    def forward(self, input, target):
        if target.is_floating_point():   # class probabilities: -sum_c t_c log_softmax(x)_c, mean over N * voxels
            if self.weight is not None:
                raise NotImplementedError("deepatlas_b200: probability targets with class weights")
            s = ops.xent_sums(input, target, 2)
            return s[0] / s[1] if self.reduction == "mean" else s[0]
        s = ops.xent_sums(input, target, 0, class_weight=self.weight, ignore_index=self.ignore_index)
        return s[0] / s[1] if self.reduction == "mean" else s[0]


class FocalLoss(nn.Module):
    """lib/loss.py:120-186.  ``probs = F.nll_loss(P, targets)`` is -P[target] there, so the modulating factor is
    (1 + P[target])**gamma; reproduced as is (parity with the reference, not with the paper)."""

    def __init__(self, class_num, alpha=None, gamma=2, size_average=True, soft_max=True):
        super().__init__()
        self.alpha = torch.ones(class_num, 1) if alpha is None else alpha
        self.gamma, self.class_num, self.size_average, self.soft_max = gamma, class_num, size_average, soft_max

    def forward(self, inputs, targets):
        if inputs.dim() == 2:
            # (observations, classes) with one label per row -- the only other shape the reference accepts
            # (lib/loss.py:167 permutes five axes or nothing): the rows become the voxels of one volume
            n, c = inputs.shape
            inputs = inputs.t().reshape(1, c, 1, 1, n)
            targets = targets.reshape(1, 1, 1, n)
        elif inputs.dim() != 5:
            raise NotImplementedError("deepatlas_b200: FocalLoss takes (B,C,X,Y,Z) or (N,C) inputs, as the reference does")
        if inputs.is_cuda and not self.alpha.is_cuda:
            self.alpha = self.alpha.cuda()
        s = ops.xent_sums(inputs, targets, 1, class_weight=self.alpha.float().reshape(-1), gamma=float(self.gamma),
                          focal_softmax=bool(self.soft_max))
        return s[0] / s[1] if self.size_average else s[0]


class SoftCrossEntropy(nn.Module):
    """lib/loss.py:96-117 for class-wise probability targets (B,C,D,M,N).  With a label-mask target the reference
    multiplies the raw label VALUES into the log-probabilities (it uses ``target``, not the one-hot it just built,
    lib/loss.py:114-116, and negates a uint8): that path is not reproduced.  ``softmax=False`` clamps the
    prediction at 1e-8 (in place in the reference; here the argument is left untouched)."""

    def __init__(self, n_class=None, weight_type="Simple", no_bg=False, softmax=False):
        super().__init__()
        self.weight_type, self.n_class, self.no_bg, self.softmax = weight_type, n_class, no_bg, softmax

    def forward(self, pred, target):
        shape = list(pred.shape)
        if len(target.shape) == len(shape) - 1:
            raise NotImplementedError("deepatlas_b200: SoftCrossEntropy with a label-mask target (see class docstring)")
        if target.shape[1] != shape[1]:
            raise ValueError("Incorrect size of target tensor: {}, should be {} or []".format(
                target.shape, shape, shape[:1] + [1, ] + shape[2:]))
        s = ops.xent_sums(pred, target, 2 if self.softmax else 3)
        return s[0] / s[1]


loss_dict = {
    "ncc": NormalizedCrossCorrelationLoss,
    "lncc": VoxelMorphLNCC,
    "mse": MSELoss,
    "gradient": gradientLoss,
    "bendingEnergy": BendingEnergyLoss,
    "dice": DiceLossMultiClass,
    "L2": L2Loss,
    "focal": FocalLoss,
    "cross_entropy": CrossEntropyLoss,
    "soft_cross_entropy": SoftCrossEntropy,
}


def get_available_losses():
    return loss_dict.keys()


def get_loss_function(loss_name):
    """lib/loss.py:753-757: KeyError on an unknown name."""
    if loss_name in get_available_losses():
        return loss_dict[loss_name]
    raise KeyError("Network {} is not avaiable!\n Choose from: {}".format(loss_name, get_available_losses()))
