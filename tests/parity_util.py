"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import torch


def rel_err(a, b):
    """max-norm relative error ||a-b||_inf / max(||b||_inf, tiny)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


def cpu_state(module, dtype=torch.float32):
    return {k: (v.detach().cpu().to(dtype) if v.is_floating_point() else v.detach().cpu().clone())
            for k, v in module.state_dict().items()}


def oracle_joint_loss(model, batch, P, dtype=torch.float32):
    """The joint step of deepatlas_b200/joint.py restated on the CPU with oracle/ref_port.py.  Returns
    (loss, {param_name: grad}) with parameter names as in JointModel ('seg.*', 'reg.*')."""
    I_m, S_m, I_t, S_t = [t.detach().cpu() for t in batch]
    I_m, I_t = I_m.to(dtype), I_t.to(dtype)
    C = model.n_classes
    lam = model.lambdas
    seg_sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
              for k, v in cpu_state(model.seg, dtype).items()}
    reg_sd = {k: v.clone().requires_grad_(True) for k, v in cpu_state(model.reg, dtype).items()}
    P_m = P.unet_generator_forward(I_m, seg_sd, 1, True)
    P_t = P.unet_generator_forward(I_t, seg_sd, 1, True)
    disp, I_w, phi = P.voxelmorph_forward(I_m, I_t, reg_sd)
    S_w = P.warp(torch.softmax(P_m, 1), phi)
    onehot = P.mask_to_one_hot(S_t.reshape(1, 1, *S_t.shape[1:]), C, dtype=dtype)
    loss = (lam["sim"] * P.lncc(I_w, I_t) + lam["reg"] * P.bending_energy(disp)
            + lam["ana"] * P.dice_multiclass(S_w, onehot, C, "Uniform", False, False, 1e-6)
            + lam["sup"] * (P.dice_multiclass(P_m, S_m.long(), C, "Uniform", False, True, 1e-6)
                            + P.dice_multiclass(P_t, S_t.long(), C, "Uniform", False, True, 1e-6)))
    loss.backward()
    grads = {}
    for k, v in seg_sd.items():
        if v.is_floating_point() and v.requires_grad and v.grad is not None:
            grads["seg." + k] = v.grad
    for k, v in reg_sd.items():
        if v.grad is not None:
            grads["reg." + k] = v.grad
    return loss.detach(), grads
