"""Generate tests/golden/*.npz from the REAL reference modules.  TEST INFRASTRUCTURE (build container only).

    python oracle/make_golden.py            # needs /root/reference; writes tests/golden/

The reference ships no tests or golden vectors, so these fixtures are frozen outputs of
the reference's own classes -- ``lib.network_factory.get_network(...)``, ``lib.loss.get_loss_function(...)``,
``lib.utils.get_identity_transform`` and ``lib.transforms.mask_to_one_hot`` -- imported where they lie (see
oracle/ref_import.py) and run on CPU in fp32 under the torch build recorded in ``meta.json``.  They travel to
the GPU box, where /root/reference does not exist, and pin both the CPU port (oracle/ref_port.py, ``-m "not
gpu"``) and the CUDA path (``-m gpu``).

Inputs are seeded ``torch.Generator`` draws (seed 230 = the reference's ``random_seed``, train_seg.py:36);
network weights are NOT stored: they come from ``weights_init()`` under ``torch.manual_seed(230)``, which the
mirror classes reproduce bit-for-bit (same construction order, same RNG stream); per-tensor checksums are
stored so a drift of that stream is detected instead of silently mis-compared.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SEED = 230


def _g(seed=SEED):
    return torch.Generator().manual_seed(seed)


def _np(t):
    return t.detach().cpu().numpy()


def _checksums(sd):
    return {k: [float(v.double().sum()), float(v.double().abs().sum())] for k, v in sd.items() if v.is_floating_point()}


def golden_ops(ref):
    out = {}
    # ---- identity grid + warp (lib/utils.py:89-102, voxel_morph.py:88-91) ------------------------------
    g = _g()
    size = (6, 7, 8)
    src = torch.rand((2, 3) + size, generator=g, requires_grad=True)
    disp = (torch.randn((2, 3) + size, generator=g) * 0.3).requires_grad_(True)
    ident = ref.utils.get_identity_transform(size)
    phi = disp + ident
    warped = torch.nn.functional.grid_sample(src, phi.permute(0, 2, 3, 4, 1), mode="bilinear", padding_mode="zeros",
                                             align_corners=True)
    cot = torch.randn(warped.shape, generator=g)
    (warped * cot).sum().backward()
    out.update(warp_src=_np(src), warp_disp=_np(disp), warp_identity=_np(ident), warp_out=_np(warped), warp_cot=_np(cot),
               warp_gsrc=_np(src.grad), warp_gdisp=_np(disp.grad))
    # ---- mask_to_one_hot (lib/transforms.py:675-689) ----------------------------------------------------
    lab = torch.randint(0, 5, (2, 1, 4, 5, 6), generator=g)
    out.update(onehot_labels=_np(lab).astype(np.uint8), onehot_out=_np(ref.transforms.mask_to_one_hot(lab, 5)))
    # ---- Dice (lib/loss.py:397-476): every weighting, hard and soft targets -----------------------------
    C, dsize = 4, (6, 6, 7)
    logits = torch.randn((2, C) + dsize, generator=g)
    labels = torch.randint(0, C, (2,) + dsize, generator=g)
    soft = torch.softmax(torch.randn((2, C) + dsize, generator=g), 1)
    out.update(dice_logits=_np(logits), dice_labels=_np(labels).astype(np.uint8), dice_soft=_np(soft), dice_cot=np.float32(1.0))
    k = 0
    for wt in ("Uniform", "Simple", "Volume"):
        for softmax in (True, False):
            for no_bg in (False, True):
                for tgt_name, tgt in (("hard", labels), ("soft", soft)):
                    x = (logits if softmax else torch.softmax(logits, 1)).clone().requires_grad_(True)
                    crit = ref.get_loss_function("dice")(n_class=C, weight_type=wt, no_bg=no_bg, softmax=softmax, eps=1e-6)
                    loss = crit(x, tgt)
                    loss.backward()
                    out[f"dice_{wt}_{int(softmax)}_{int(no_bg)}_{tgt_name}_loss"] = _np(loss)
                    out[f"dice_{wt}_{int(softmax)}_{int(no_bg)}_{tgt_name}_grad"] = _np(x.grad)
                    k += 1
    # ---- LNCC (lib/loss.py:589-617) -------------------------------------------------------------------------
    lsize = (12, 13, 14)
    I = torch.rand((2, 1) + lsize, generator=g, requires_grad=True)
    J = torch.rand((2, 1) + lsize, generator=g, requires_grad=True)
    crit = ref.get_loss_function("lncc")()
    loss = crit(I, J)
    loss.backward()
    out.update(lncc_I=_np(I), lncc_J=_np(J), lncc_loss=_np(loss), lncc_gI=_np(I.grad), lncc_gJ=_np(J.grad))
    # ---- bending energy (lib/loss.py:674-730) ---------------------------------------------------------------
    for name, bsize, spacing in (("iso", (8, 9, 10), (1, 1, 1)), ("aniso", (10, 8, 12), (1.0, 1.5, 2.0))):
        u = (torch.randn((2, 3) + bsize, generator=g) * 0.1).requires_grad_(True)
        crit = ref.get_loss_function("bendingEnergy")(spacing=spacing)
        loss = crit(u)
        loss.backward()
        out.update({f"bend_{name}_u": _np(u), f"bend_{name}_spacing": np.asarray(spacing, np.float32),
                    f"bend_{name}_loss": _np(loss), f"bend_{name}_grad": _np(u.grad)})
    return out


def golden_nets(ref):
    out, meta = {}, {}
    g = _g()
    # ---- UNet_light(1, 4, bias, BN), train mode, 16^3 (config C1 at reduced size) ---------------------------
    torch.manual_seed(SEED)
    net = ref.get_network("UNet_light")(1, 4, bias=True, BN=True)
    net.weights_init()
    net.train()
    meta["unet_light_checksums"] = _checksums(net.state_dict())
    x = torch.rand((1, 1, 16, 16, 16), generator=g)
    lab = torch.randint(0, 4, (1, 16, 16, 16), generator=g)
    logits = net(x)
    crit = ref.get_loss_function("dice")(n_class=4, weight_type="Uniform", softmax=True, eps=1e-6)
    loss = crit(logits, lab)
    loss.backward()
    out.update(ul_x=_np(x), ul_labels=_np(lab).astype(np.uint8), ul_logits=_np(logits), ul_loss=_np(loss),
               ul_argmax=_np(torch.max(logits, 1)[1]).astype(np.uint8))
    sd = net.state_dict()
    out.update(ul_running_mean0=_np(sd["encoders.0.0.BN.running_mean"]), ul_running_var0=_np(sd["encoders.0.0.BN.running_var"]))
    for k in ("encoders.0.0.conv.weight", "encoders.3.1.conv.weight", "encoders.1.0.BN.weight", "up_samplers.1.deconv.weight",
              "up_samplers.2.deconv.bias", "decoders.decBlock0.0.conv.weight", "decoders.decBlock2.2.weight",
              "decoders.decBlock2.2.bias", "decoders.decBlock1.1.BN.bias"):
        out["ul_grad/" + k] = _np(dict(net.named_parameters())[k].grad)
    # ---- VoxelMorphCVPR2018 at (16, 24, 16) + LNCC + bending ---------------------------------------------
    torch.manual_seed(SEED)
    vm = ref.get_network("voxel_morph_cvpr")()
    vm.weights_init()
    meta["voxelmorph_checksums"] = _checksums(vm.state_dict())
    vs = (16, 24, 16)
    s, t = torch.rand((1, 1) + vs, generator=g), torch.rand((1, 1) + vs, generator=g)
    disp, warped, deform = vm(s, t)
    loss = ref.get_loss_function("lncc")()(warped, t) + 1000.0 * ref.get_loss_function("bendingEnergy")()(disp)
    loss.backward()
    out.update(vm_s=_np(s), vm_t=_np(t), vm_disp=_np(disp), vm_warped=_np(warped), vm_deform=_np(deform), vm_loss=_np(loss))
    for k in ("encoders.0.conv.weight", "encoders.4.conv.bias", "decoders.2.conv.weight", "flow.weight", "flow.bias"):
        out["vm_grad/" + k] = _np(dict(vm.named_parameters())[k].grad)
    # ---- UNet (32 base) forward at 16^3 (weights regenerated from the seed; logits only) ----------------------
    torch.manual_seed(SEED)
    un = ref.get_network("UNet")(1, 4, bias=True, BN=True)
    un.weights_init()
    un.train()
    meta["unet_checksums"] = {k: v for k, v in list(_checksums(un.state_dict()).items())[:12]}
    xu = torch.rand((1, 1, 16, 16, 16), generator=g)
    out.update(un_x=_np(xu), un_logits=_np(un(xu)))
    # ---- the joint step's loss (definition: SURVEY.md 8(d)) from reference modules, C = 4, 16^3 ------------
    torch.manual_seed(SEED)
    seg = ref.get_network("UNet_light")(1, 4, bias=True, BN=True)
    seg.weights_init()
    reg = ref.get_network("voxel_morph_cvpr")()
    reg.weights_init()
    seg.train()
    gj = _g()
    D = (16, 16, 16)
    I_m, I_t = torch.rand((1, 1) + D, generator=gj), torch.rand((1, 1) + D, generator=gj)
    S_m = torch.randint(0, 4, (1,) + D, generator=gj, dtype=torch.uint8)
    S_t = torch.randint(0, 4, (1,) + D, generator=gj, dtype=torch.uint8)
    P_m, P_t = seg(I_m), seg(I_t)
    disp, I_w, phi = reg(I_m, I_t)
    S_w = torch.nn.functional.grid_sample(torch.softmax(P_m, 1), phi.permute(0, 2, 3, 4, 1), mode="bilinear",
                                          padding_mode="zeros", align_corners=True)
    onehot = ref.transforms.mask_to_one_hot(S_t.reshape(1, 1, *D), 4)
    dice_sup = ref.get_loss_function("dice")(n_class=4, weight_type="Uniform", softmax=True, eps=1e-6)
    dice_ana = ref.get_loss_function("dice")(n_class=4, weight_type="Uniform", softmax=False, eps=1e-6)
    parts = dict(sim=ref.get_loss_function("lncc")()(I_w, I_t), reg=ref.get_loss_function("bendingEnergy")()(disp),
                 ana=dice_ana(S_w, onehot), sup=dice_sup(P_m, S_m.long()) + dice_sup(P_t, S_t.long()))
    loss = parts["sim"] + 1000.0 * parts["reg"] + parts["ana"] + parts["sup"]
    loss.backward()
    out.update(joint_loss=_np(loss), **{"joint_part_" + k: _np(v) for k, v in parts.items()})
    out["joint_grad/seg.encoders.0.0.conv.weight"] = _np(dict(seg.named_parameters())["encoders.0.0.conv.weight"].grad)
    out["joint_grad/seg.decoders.decBlock2.2.weight"] = _np(dict(seg.named_parameters())["decoders.decBlock2.2.weight"].grad)
    out["joint_grad/reg.flow.weight"] = _np(reg.flow.weight.grad)
    out["joint_grad/reg.encoders.0.conv.weight"] = _np(reg.encoders[0].conv.weight.grad)
    return out, meta


def main():
    if not ref_import.available():
        raise SystemExit("reference tree not found at %s" % ref_import.REFERENCE_ROOT)
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))
    ref = ref_import.load()
    os.makedirs(OUT, exist_ok=True)
    ops = golden_ops(ref)
    nets, meta = golden_nets(ref)
    np.savez_compressed(os.path.join(OUT, "ops.npz"), **ops)
    np.savez_compressed(os.path.join(OUT, "nets.npz"), **nets)
    meta.update(torch=torch.__version__, numpy=np.__version__, seed=SEED, reference_root=ref_import.REFERENCE_ROOT,
                generated_by="oracle/make_golden.py", dtype="float32 (CPU)")
    with open(os.path.join(OUT, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    for n in ("ops.npz", "nets.npz", "meta.json"):
        print(n, os.path.getsize(os.path.join(OUT, n)), "bytes")


if __name__ == "__main__":
    main()
