"""Host-side mirror of the reference's loss registry entries on the hot path (lib/loss.py).

Same class names, constructor arguments, ``forward`` signatures and error behaviour as
``DiceLossMultiClass`` (lib/loss.py:397-476), ``VoxelMorphLNCC`` (:589-617) and ``BendingEnergyLoss``
(:674-730).  The full-volume arithmetic runs in the CUDA library; only the closing formulas on the
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
    """Bending energy of a 3-D displacement field (L2 form)."""

    def __init__(self, norm="L2", spacing=(1, 1, 1), normalize=True):
        super().__init__()
        if norm != "L2":
            raise NotImplementedError("deepatlas_b200: BendingEnergyLoss is built for norm='L2'")
        self.norm = norm
        self.spacing = torch.tensor(spacing).float()
        self.normalize = normalize
        if self.normalize:
            self.spacing = self.spacing / self.spacing.min()

    def _coef(self, shape, device):
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
        return coef.float().to(device)

    def forward(self, input):
        sums = ops.bending_sums(input)                               # (B, 3, 6)
        return (sums * self._coef(input.shape, input.device)[None]).sum()


loss_dict = {
    "lncc": VoxelMorphLNCC,
    "bendingEnergy": BendingEnergyLoss,
    "dice": DiceLossMultiClass,
}


def get_available_losses():
    return loss_dict.keys()


def get_loss_function(loss_name):
    """lib/loss.py:753-757: KeyError on an unknown name."""
    if loss_name in get_available_losses():
        return loss_dict[loss_name]
    raise KeyError("Network {} is not avaiable!\n Choose from: {}".format(loss_name, get_available_losses()))
