"""The CPU port (oracle/ref_port.py) against the committed golden fixtures (tests/golden/, produced from the
real reference modules by oracle/make_golden.py).  Runs anywhere -- in particular on the GPU box, where
/root/reference does not exist.  1e-6 relative for single ops; whole networks get NET_TOL because ATen's CPU reductions
(batch-norm statistics, conv accumulation) change summation order with the host's thread count."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import ref_port as P

NET_TOL, NET_GRAD_TOL = 2e-5, 1e-4
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ops_gold():
    return dict(np.load(os.path.join(GOLD, "ops.npz")))


@pytest.fixture(scope="module")
def nets_gold():
    return dict(np.load(os.path.join(GOLD, "nets.npz")))


@pytest.fixture(scope="module")
def meta():
    return json.load(open(os.path.join(GOLD, "meta.json")))


def _t(a):
    return torch.from_numpy(np.asarray(a))


def _close(a, b, tol=1e-6, floor=1e-30):
    """max-norm relative; `floor` puts analytically-zero quantities (a conv bias in front of a BatchNorm has a
    round-off-only gradient) on an absolute scale."""
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).abs().max()) <= tol * max(float(b.abs().max()), floor)


def test_warp_identity_onehot(ops_gold):
    g = ops_gold
    src, disp = _t(g["warp_src"]).requires_grad_(True), _t(g["warp_disp"]).requires_grad_(True)
    ident = P.identity_transform(src.shape[2:])
    assert torch.equal(ident, _t(g["warp_identity"]))
    out = P.warp(src, disp + ident)
    assert _close(out, g["warp_out"])
    (out * _t(g["warp_cot"])).sum().backward()
    assert _close(src.grad, g["warp_gsrc"]) and _close(disp.grad, g["warp_gdisp"])
    assert _close(P.warp_closed_form(src.detach().double(), (disp.detach() + ident).double()), g["warp_out"], 1e-5)
    lab = _t(g["onehot_labels"]).long()
    assert torch.equal(P.mask_to_one_hot(lab, 5), _t(g["onehot_out"]))


@pytest.mark.parametrize("wt", ["Uniform", "Simple", "Volume"])
@pytest.mark.parametrize("softmax", [1, 0])
@pytest.mark.parametrize("no_bg", [0, 1])
@pytest.mark.parametrize("tgt", ["hard", "soft"])
def test_dice(ops_gold, wt, softmax, no_bg, tgt):
    g = ops_gold
    logits = _t(g["dice_logits"])
    x = (logits if softmax else torch.softmax(logits, 1)).clone().requires_grad_(True)
    target = _t(g["dice_labels"]).long() if tgt == "hard" else _t(g["dice_soft"])
    loss = P.dice_multiclass(x, target, 4, wt, bool(no_bg), bool(softmax), 1e-6)
    loss.backward()
    key = f"dice_{wt}_{softmax}_{no_bg}_{tgt}"
    assert _close(loss, g[key + "_loss"]) and _close(x.grad, g[key + "_grad"])


def test_lncc_bending(ops_gold):
    g = ops_gold
    I, J = _t(g["lncc_I"]).requires_grad_(True), _t(g["lncc_J"]).requires_grad_(True)
    loss = P.lncc(I, J)
    loss.backward()
    assert _close(loss, g["lncc_loss"]) and _close(I.grad, g["lncc_gI"], 1e-5) and _close(J.grad, g["lncc_gJ"], 1e-5)
    for name in ("iso", "aniso"):
        u = _t(g[f"bend_{name}_u"]).requires_grad_(True)
        loss = P.bending_energy(u, tuple(float(s) for s in g[f"bend_{name}_spacing"]))
        loss.backward()
        assert _close(loss, g[f"bend_{name}_loss"]) and _close(u.grad, g[f"bend_{name}_grad"])


def _mirror(name, args, kw, checksums):
    """Weights are not stored in the fixtures: regenerate them through the mirror class under the seed and verify
    the recorded per-tensor checksums (detects RNG-stream drift instead of mis-comparing)."""
    import deepatlas_b200 as da
    torch.manual_seed(230)
    net = da.get_network(name)(*args, **kw)
    net.weights_init()
    sd = net.state_dict()
    for k, (s, a) in checksums.items():
        assert abs(float(sd[k].double().sum()) - s) <= 1e-9 * max(1.0, a), f"weight stream drifted at {k}"
    return net, {k: v.clone() for k, v in sd.items()}


def test_unet_light(nets_gold, meta):
    g = nets_gold
    _, sd = _mirror("UNet_light", (1, 4), dict(bias=True, BN=True), meta["unet_light_checksums"])
    sd = {k: (v.requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
    stats = {}
    logits = P.unet_generator_forward(_t(g["ul_x"]), sd, 1, True, stats_out=stats)
    assert _close(logits, g["ul_logits"], NET_TOL)
    assert np.array_equal(torch.max(logits, 1)[1].numpy().astype(np.uint8), g["ul_argmax"])   # label indices bit-exact
    loss = P.dice_multiclass(logits, _t(g["ul_labels"]).long(), 4, "Uniform", False, True, 1e-6)
    loss.backward()
    assert _close(loss, g["ul_loss"], NET_TOL)
    assert _close(stats["encoders.0.0.BN.running_mean"], g["ul_running_mean0"], NET_TOL) and _close(stats["encoders.0.0.BN.running_var"], g["ul_running_var0"], NET_TOL)
    for k in [k for k in g if k.startswith("ul_grad/")]:
        assert _close(sd[k[len("ul_grad/"):]].grad, g[k], NET_GRAD_TOL, 1e-3), k


def test_voxelmorph_unet32_joint(nets_gold, meta):
    g = nets_gold
    _, sd = _mirror("voxel_morph_cvpr", (), {}, meta["voxelmorph_checksums"])
    sd = {k: v.requires_grad_(True) for k, v in sd.items()}
    disp, warped, deform = P.voxelmorph_forward(_t(g["vm_s"]), _t(g["vm_t"]), sd)
    assert _close(disp, g["vm_disp"], NET_TOL) and _close(warped, g["vm_warped"], NET_TOL) and _close(deform, g["vm_deform"], NET_TOL)
    loss = P.lncc(warped, _t(g["vm_t"])) + 1000.0 * P.bending_energy(disp)
    loss.backward()
    assert _close(loss, g["vm_loss"], NET_TOL)
    for k in [k for k in g if k.startswith("vm_grad/")]:
        assert _close(sd[k[len("vm_grad/"):]].grad, g[k], NET_GRAD_TOL, 1e-3), k
    _, sdu = _mirror("UNet", (1, 4), dict(bias=True, BN=True), meta["unet_checksums"])
    assert _close(P.unet_forward(_t(g["un_x"]), sdu, True), g["un_logits"], NET_TOL)
    # joint step
    from deepatlas_b200.joint import make_synthetic_pair
    torch.manual_seed(230)
    import deepatlas_b200 as da
    seg = da.get_network("UNet_light")(1, 4, bias=True, BN=True); seg.weights_init()
    reg = da.get_network("voxel_morph_cvpr")(); reg.weights_init()
    seg_sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in seg.state_dict().items()}
    reg_sd = {k: v.clone().requires_grad_(True) for k, v in reg.state_dict().items()}
    batch = make_synthetic_pair((16, 16, 16), 4, seed=230)
    loss = P.joint_loss(seg_sd, reg_sd, batch, 4)
    loss.backward()
    assert _close(loss, g["joint_loss"], NET_TOL)
    assert _close(seg_sd["encoders.0.0.conv.weight"].grad, g["joint_grad/seg.encoders.0.0.conv.weight"], NET_GRAD_TOL)
    assert _close(reg_sd["flow.weight"].grad, g["joint_grad/reg.flow.weight"], NET_GRAD_TOL)
