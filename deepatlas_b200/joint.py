"""The joint seg+reg training step, assembled from the reference's registry entries.

The reference tree ships no joint loop (README.md:15-19 lists it as TODO); its ingredients are all
present and SURVEY.md 8(d) fixes the step definition used here and by the oracle:

    P_m = seg(I_m) ; P_t = seg(I_t)                      two B=1 calls (BN statistics per volume)
    disp, I_w, phi = reg(I_m, I_t)                        lib/network_factory/voxel_morph.py:62-92
    S_w = grid_sample(softmax(P_m), phi)                  same call as voxel_morph.py:90-91
    L = l_sim * lncc(I_w, I_t) + l_reg * bendingEnergy(disp)
      + l_ana * dice(S_w, onehot(S_t))                    soft-target branch lib/loss.py:435-436
      + l_sup * (dice(P_m, S_m) + dice(P_t, S_t))         softmax=True, train_seg.py:54-55

The lambdas are not in the reference (paper hyper-parameters); they are configuration here.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import ops
from .losses import get_loss_function
from .networks import get_network

DEFAULT_LAMBDAS = dict(sim=1.0, reg=1000.0, ana=1.0, sup=1.0)


class JointModel(nn.Module):
    def __init__(self, n_classes=32, in_channel=1, seg_name="UNet_light", lambdas=None, overlap_reg=False, overlap_seg=False):
        """``overlap_reg``: run the registration network on a side stream next to the two segmentation passes (they
        share no state: the registration net has no BatchNorm, parameters are read-only in the forward, its gradients
        land in its own bucket slots).  Small low-resolution kernels of one branch then fill SMs the other leaves idle;
        inside a captured CUDA graph the two branches become parallel paths.  The backward of each branch runs on the
        stream of its forward (autograd's rule); call ``join_streams()`` after ``backward()`` and before anything on the
        current stream reads the registration net's gradients (an optimizer step, an all-reduce).

        ``overlap_seg``: additionally run the TARGET image's segmentation pass (and its supervised Dice) on a second side
        stream next to the moving image's.  The two passes share one network, so the side pass (a) defers its BatchNorm
        running-statistics updates (``networks.deferred_bn_updates``: applied after the join, i.e. in the reference's
        order: moving, then target) and (b) adds its parameter gradients into the gradient bucket's alternate slots
        (``ops.grad_slot(1)``, ``FlatGradBucket.enable_alt()``; folded in by ``allreduce()``).  Needs the fused head path."""
        super().__init__()
        self.overlap_reg = bool(overlap_reg)
        self.overlap_seg = bool(overlap_seg)
        self._side = None
        self._side2 = None
        self._pending = []      # side streams whose backward has not been joined yet
        self.n_classes = n_classes
        self.seg = get_network(seg_name)(in_channel, n_classes, bias=True, BN=True)
        self.reg = get_network("voxel_morph_cvpr")()
        self.lambdas = dict(DEFAULT_LAMBDAS, **(lambdas or {}))
        self.sim_loss = get_loss_function("lncc")()
        self.reg_loss = get_loss_function("bendingEnergy")()
        self.sup_dice = get_loss_function("dice")(n_class=n_classes, weight_type="Uniform", softmax=True, eps=1e-6)
        self.ana_dice = get_loss_function("dice")(n_class=n_classes, weight_type="Uniform", softmax=False, eps=1e-6)

    def weights_init(self):
        self.seg.weights_init()
        self.reg.weights_init()

    def trainable_parameters(self):
        return list(self.seg.parameters()) + list(self.reg.parameters())

    def _side_stream(self, device):
        if self._side is None or self._side.device != device:
            self._side = torch.cuda.Stream(device=device)
        return self._side

    def join_streams(self):
        """Make the current stream wait for the side streams (after ``backward()``): the registration branch's and the
        weight-gradient stream of ``ops.set_wgrad_overlap``."""
        # (only streams that ran a branch since the last join: see ops.join_wgrad_stream)
        for st in self._pending:
            torch.cuda.current_stream(st.device).wait_stream(st)
        self._pending = []
        ops.join_wgrad_stream()

    def _run_reg(self, I_m, I_t):
        """Registration branch: network, warp, and the two losses that depend on nothing else (similarity, bending
        energy).  Returns (disp, I_w, phi, sim, reg)."""
        def branch():
            disp, I_w, phi = self.reg(I_m, I_t)
            return disp, I_w, phi, self.sim_loss(I_w, I_t), self.reg_loss(disp)
        if not (self.overlap_reg and I_m.is_cuda):
            return branch()
        # the side stream waits for the fork event recorded at the START of joint_loss, not for the segmentation passes
        # that the host has enqueued in between (the host-side call order stays seg, seg, reg)
        side = self._side_stream(I_m.device)
        side.wait_event(self._fork)
        with torch.cuda.stream(side):
            out = branch()
        self._pending.append(side)
        return out

    def _fork_here(self, ref):
        if (self.overlap_reg or self.overlap_seg) and ref.is_cuda:
            self._fork = torch.cuda.Event()
            self._fork.record(torch.cuda.current_stream(ref.device))

    def _join_reg(self, outs):
        if self.overlap_reg and self._side is not None and outs[0].is_cuda:
            main = torch.cuda.current_stream(outs[0].device)
            main.wait_stream(self._side)
            for t in outs:
                t.record_stream(main)

    def joint_loss(self, I_m, S_m, I_t, S_t):
        """I_*: (1,1,D,H,W) fp32 images; S_*: (1,D,H,W) integer label maps (uint8 is read directly)."""
        lam = self.lambdas
        mode = os.environ.get("DA_JOINT_UNFUSED", "0")   # A/B switch: 1 = separate softmax and Dice passes, 2 = unfused head
        fuse_head = mode == "0" and hasattr(self.seg, "forward_features") and not getattr(self.seg, "res", False)
        if fuse_head:
            # the class head, softmax and the Dice sums as one kernel each way: the 32-class logits (629 MB per volume at
            # 160x192x160) and their gradient never reach HBM; for the moving image the same pass writes the
            # probabilities that the anatomy term warps, and the backward takes both of their gradients
            self._fork_here(I_m)
            if self.overlap_seg and I_m.is_cuda:
                # the target image's pass on its own stream: BatchNorm buffer updates deferred, gradients into slot 1
                from .networks import deferred_bn_updates
                main = torch.cuda.current_stream(I_m.device)
                if self._side2 is None or self._side2.device != I_m.device:
                    self._side2 = torch.cuda.Stream(device=I_m.device)
                self._side2.wait_event(self._fork)
                with torch.cuda.stream(self._side2), deferred_bn_updates() as pending, ops.grad_slot(1):
                    F_t = self.seg.forward_features(I_t)
                    sup_t, _ = self.sup_dice.forward_head(F_t, self.seg.head, S_t)
                F_m = self.seg.forward_features(I_m)
                disp, I_w, phi, sim, reg = self._run_reg(I_m, I_t)
                sup_m, prob_m = self.sup_dice.forward_head(F_m, self.seg.head, S_m, want_probs=True)
                main.wait_stream(self._side2)
                self._pending.append(self._side2)
                sup_t.record_stream(main)
                pending.apply()      # after the moving pass's own updates (stream order) and the side pass's statistics (join)
            else:
                F_m = self.seg.forward_features(I_m)
                F_t = self.seg.forward_features(I_t)
                disp, I_w, phi, sim, reg = self._run_reg(I_m, I_t)     # (side stream with overlap_reg)
                sup_m, prob_m = self.sup_dice.forward_head(F_m, self.seg.head, S_m, want_probs=True)
                sup_t, _ = self.sup_dice.forward_head(F_t, self.seg.head, S_t)
            self._join_reg((disp, I_w, phi, sim, reg))
        else:
            self._fork_here(I_m)
            P_m = self.seg(I_m)
            P_t = self.seg(I_t)
            disp, I_w, phi, sim, reg = self._run_reg(I_m, I_t)
            self._join_reg((disp, I_w, phi, sim, reg))
            # softmax(P_m) has two consumers (supervised Dice, anatomy term): one pass yields the Dice sums AND the
            # probabilities, and one backward pass takes both gradients (no separate softmax, no sum of two gradients)
            if mode == "1":
                sup_m, prob_m = self.sup_dice(P_m, S_m), ops.softmax(P_m)
            else:
                sup_m, prob_m = self.sup_dice.forward_with_probs(P_m, S_m)
            sup_t = self.sup_dice(P_t, S_t)
        parts = {
            "sim": sim,
            "reg": reg,
            # dice(grid_sample(softmax(P_m), phi), onehot(S_t)): warp and Dice sums fused, labels stand for the one-hot
            "ana": self.ana_dice.forward_warped(prob_m, phi, S_t),
            "sup": sup_m + sup_t,
        }
        loss = lam["sim"] * parts["sim"] + lam["reg"] * parts["reg"] + lam["ana"] * parts["ana"] + lam["sup"] * parts["sup"]
        return loss, parts


def make_synthetic_pair(size, n_classes, seed=230, device="cpu"):
    """SURVEY.md 8(d) synthetic inputs: images U[0,1) fp32 (1,1,D,H,W); labels randint(0,C) uint8 (1,D,H,W);
    moving and target are independent draws.  Generated on the CPU generator so every device sees the same data."""
    g = torch.Generator().manual_seed(seed)
    D, H, W = size
    I_m = torch.rand((1, 1, D, H, W), generator=g)
    I_t = torch.rand((1, 1, D, H, W), generator=g)
    S_m = torch.randint(0, n_classes, (1, D, H, W), generator=g, dtype=torch.uint8)
    S_t = torch.randint(0, n_classes, (1, D, H, W), generator=g, dtype=torch.uint8)
    return tuple(t.to(device) for t in (I_m, S_m, I_t, S_t))


class RegOnlyModel(nn.Module):
    """BASELINE.json config #3: registration net + trilinear warp + LNCC + bending energy (no segmentation net).
    L = l_sim * lncc(I_w, I_t) + l_reg * bendingEnergy(disp); the ingredients are lib/network_factory/voxel_morph.py:62-92
    and lib/loss.py:589-617,674-730."""

    def __init__(self, lambdas=None):
        super().__init__()
        self.reg = get_network("voxel_morph_cvpr")()
        self.lambdas = dict(DEFAULT_LAMBDAS, **(lambdas or {}))
        self.sim_loss = get_loss_function("lncc")()
        self.reg_loss = get_loss_function("bendingEnergy")()

    def weights_init(self):
        self.reg.weights_init()

    def trainable_parameters(self):
        return list(self.reg.parameters())

    def joint_loss(self, I_m, S_m, I_t, S_t):
        disp, I_w, _ = self.reg(I_m, I_t)
        parts = {"sim": self.sim_loss(I_w, I_t), "reg": self.reg_loss(disp)}
        return self.lambdas["sim"] * parts["sim"] + self.lambdas["reg"] * parts["reg"], parts


class SegOnlyModel(nn.Module):
    """BASELINE.json configs #1/#2: one segmentation net + supervised Dice (models/segmentation.py:152-157 with
    train_seg.py:54-55's loss settings); one volume per step."""

    def __init__(self, n_classes=4, in_channel=1, seg_name="UNet"):
        super().__init__()
        self.n_classes = n_classes
        self.seg = get_network(seg_name)(in_channel, n_classes, bias=True, BN=True)
        self.sup_dice = get_loss_function("dice")(n_class=n_classes, weight_type="Uniform", softmax=True, eps=1e-6)

    def weights_init(self):
        self.seg.weights_init()

    def trainable_parameters(self):
        return list(self.seg.parameters())

    def joint_loss(self, I_m, S_m, I_t=None, S_t=None):
        sup = self.sup_dice(self.seg(I_m), S_m)
        return sup, {"sup": sup}
