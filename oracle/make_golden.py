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


def golden_extra(ref):
    """the remaining registry losses and the UNet_generator variants, from the real modules."""
    out = {}
    g = _g(SEED + 1)
    size = (7, 8, 9)
    a = torch.rand((2, 1) + size, generator=g, requires_grad=True)
    b = torch.rand((2, 1) + size, generator=g, requires_grad=True)
    out.update(pair_a=_np(a), pair_b=_np(b))
    for name, crit, args in (("ncc", ref.get_loss_function("ncc")(), (a, b)), ("mse", ref.get_loss_function("mse")(), (a, b)),
                             ("L2", ref.get_loss_function("L2")(), (a,))):
        a.grad = b.grad = None
        loss = crit(*args)
        loss.backward()
        out[f"{name}_loss"], out[f"{name}_ga"] = _np(loss), _np(a.grad)
        if len(args) == 2:
            out[f"{name}_gb"] = _np(b.grad)
    u = (torch.randn((2, 3, 8, 9, 10), generator=g) * 0.1).requires_grad_(True)
    out["grad_u"] = _np(u)
    for norm in ("L2", "L1"):
        for k, sp in enumerate(((1, 1, 1), (1.0, 1.5, 2.0))):
            u.grad = None
            loss = ref.get_loss_function("gradient")(norm=norm, spacing=sp)(u)
            loss.backward()
            out[f"grad_{norm}_{k}_loss"], out[f"grad_{norm}_{k}_g"] = _np(loss), _np(u.grad)
            out[f"grad_spacing_{k}"] = np.asarray(sp, np.float32)
    C, xs = 5, (6, 7, 8)
    x = torch.randn((2, C) + xs, generator=g, requires_grad=True)
    t = torch.randint(0, C, (2,) + xs, generator=g)
    soft = torch.softmax(torch.randn((2, C) + xs, generator=g), 1).requires_grad_(True)
    w = torch.rand(C, generator=g) + 0.5
    alpha = torch.rand(C, 1, generator=g) + 0.5
    out.update(xent_x=_np(x), xent_t=_np(t).astype(np.uint8), xent_soft=_np(soft), xent_w=_np(w), xent_alpha=_np(alpha))
    cases = {
        "ce": lambda xi: ref.get_loss_function("cross_entropy")()(xi, t),
        "ce_w": lambda xi: ref.get_loss_function("cross_entropy")(weight=w)(xi, t),
        "focal": lambda xi: ref.get_loss_function("focal")(C)(xi, t),
        "focal_a": lambda xi: ref.get_loss_function("focal")(C, alpha=alpha, gamma=1.5, size_average=False)(xi, t),
        "focal_nosm": lambda xi: ref.get_loss_function("focal")(C, soft_max=False)(torch.softmax(xi, 1), t),
        "sce_sm": lambda xi: ref.get_loss_function("soft_cross_entropy")(softmax=True)(xi, soft),
        "sce": lambda xi: ref.get_loss_function("soft_cross_entropy")(softmax=False)(torch.softmax(xi, 1) * 1.0, soft),  # clamp_ is in place there
    }
    for name, fn in cases.items():
        x.grad = soft.grad = None
        loss = fn(x)
        loss.backward()
        out[f"{name}_loss"], out[f"{name}_gx"] = _np(loss), _np(x.grad)
        if soft.grad is not None:
            out[f"{name}_gt"] = _np(soft.grad)
    # UNet_generator variants (lib/network_factory/unets.py:230-241,264,275)
    variants = {"strided": (dict(maxpool=False), [(4, 8), (8, 8, 16)], [(8, 8, 8)], 3),
                "upsample": (dict(upsample=True), [(4, 8), (8, 8, 16)], [(16, 8, 8)], 3),
                "res": (dict(res=True), [(8, 8), (8, 8)], [(8, 8)], 8),
                "all": (dict(maxpool=False, upsample=True, res=True), [(8, 8), (8, 8)], [(8, 8)], 8)}
    xin = torch.rand((1, 1, 8, 12, 8), generator=g)
    out["var_x"] = _np(xin)
    for name, (kw, enc, dec, ncls) in variants.items():
        torch.manual_seed(SEED)
        net = ref.network_factory.unets.UNet_generator(enc, dec, act="LeakyReLU", **kw)(1, ncls, bias=True, BN=True)
        net.weights_init()
        net.train()
        for k, v in net.state_dict().items():
            out[f"var_{name}_sd/{k}"] = _np(v)
        y = net(xin)
        cot = torch.randn(y.shape, generator=g)
        (y * cot).sum().backward()
        out[f"var_{name}_y"], out[f"var_{name}_cot"] = _np(y), _np(cot)
        for k, p_ in net.named_parameters():
            out[f"var_{name}_grad/{k}"] = _np(p_.grad)
    return out


def main():
    if not ref_import.available():
        raise SystemExit("reference tree not found at %s" % ref_import.REFERENCE_ROOT)
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))
    ref = ref_import.load()
    os.makedirs(OUT, exist_ok=True)
    if "--extra-only" in sys.argv:   # leaves ops.npz / nets.npz / meta.json as committed
        np.savez_compressed(os.path.join(OUT, "extra.npz"), **golden_extra(ref))
        print("extra.npz", os.path.getsize(os.path.join(OUT, "extra.npz")), "bytes")
        return
    ops = golden_ops(ref)
    nets, meta = golden_nets(ref)
    np.savez_compressed(os.path.join(OUT, "extra.npz"), **golden_extra(ref))
    np.savez_compressed(os.path.join(OUT, "ops.npz"), **ops)
    np.savez_compressed(os.path.join(OUT, "nets.npz"), **nets)
    meta.update(torch=torch.__version__, numpy=np.__version__, seed=SEED, reference_root=ref_import.REFERENCE_ROOT,
                generated_by="oracle/make_golden.py", dtype="float32 (CPU)")
    with open(os.path.join(OUT, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    for n in ("ops.npz", "nets.npz", "extra.npz", "meta.json"):
        print(n, os.path.getsize(os.path.join(OUT, n)), "bytes")


if __name__ == "__main__":
    main()
