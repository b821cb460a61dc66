"""The CUDA path (through the C ABI) against the committed golden fixtures made from the REAL reference modules
(tests/golden/, oracle/make_golden.py).  Tolerance 1e-4 max-norm relative (north star); label argmax bit-exact."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-4


def _rel(a, b, floor=1e-30):
    a, b = torch.as_tensor(a).detach().double().cpu(), torch.as_tensor(np.asarray(b)).double()
    return float((a - b).abs().max()) / max(float(b.abs().max()), floor)


@pytest.fixture(scope="module")
def og():
    return dict(np.load(os.path.join(GOLD, "ops.npz")))


@pytest.fixture(scope="module")
def ng():
    return dict(np.load(os.path.join(GOLD, "nets.npz")))


def _c(a, cuda, grad=False):
    t = torch.from_numpy(np.asarray(a)).to(cuda)
    return t.requires_grad_(True) if grad else t


def test_warp(cuda, og):
    from deepatlas_b200 import ops
    src, disp = _c(og["warp_src"], cuda, True), _c(og["warp_disp"], cuda, True)
    out, phi = ops.warp3d(src, disp, add_identity=True, want_phi=True)
    assert _rel(out, og["warp_out"]) < TOL
    assert _rel(phi - disp, og["warp_identity"][None].repeat(2, 0)) < 1e-6
    (out * _c(og["warp_cot"], cuda)).sum().backward()
    assert _rel(src.grad, og["warp_gsrc"]) < TOL and _rel(disp.grad, og["warp_gdisp"]) < TOL


@pytest.mark.parametrize("wt", ["Uniform", "Simple", "Volume"])
@pytest.mark.parametrize("softmax", [1, 0])
@pytest.mark.parametrize("no_bg", [0, 1])
@pytest.mark.parametrize("tgt", ["hard", "soft"])
def test_dice(cuda, og, wt, softmax, no_bg, tgt):
    import deepatlas_b200 as da
    logits = torch.from_numpy(og["dice_logits"])
    x = (logits if softmax else torch.softmax(logits, 1)).to(cuda).requires_grad_(True)
    target = _c(og["dice_labels"], cuda) if tgt == "hard" else _c(og["dice_soft"], cuda)
    crit = da.get_loss_function("dice")(n_class=4, weight_type=wt, no_bg=bool(no_bg), softmax=bool(softmax), eps=1e-6)
    loss = crit(x, target)
    loss.backward()
    key = f"dice_{wt}_{softmax}_{no_bg}_{tgt}"
    assert _rel(loss, og[key + "_loss"]) < TOL and _rel(x.grad, og[key + "_grad"]) < TOL


def test_lncc_bending(cuda, og):
    import deepatlas_b200 as da
    I, J = _c(og["lncc_I"], cuda, True), _c(og["lncc_J"], cuda, True)
    loss = da.get_loss_function("lncc")()(I, J)
    loss.backward()
    assert _rel(loss, og["lncc_loss"]) < TOL
    # The reference's cancellation-form variance has its own fp32 noise: judge the gradient on the
    # precision ladder -- our error against the fp64 restatement must not exceed max(TOL, 3 x the reference's own).
    from oracle import ref_port as P
    I64 = torch.from_numpy(og["lncc_I"]).double().requires_grad_(True)
    J64 = torch.from_numpy(og["lncc_J"]).double().requires_grad_(True)
    P.lncc(I64, J64).backward()
    for ours, gold, truth in ((I.grad, og["lncc_gI"], I64.grad), (J.grad, og["lncc_gJ"], J64.grad)):
        assert _rel(ours, truth.numpy()) < max(TOL, 3 * _rel(torch.from_numpy(gold), truth.numpy()))
    for name in ("iso", "aniso"):
        u = _c(og[f"bend_{name}_u"], cuda, True)
        loss = da.get_loss_function("bendingEnergy")(spacing=tuple(float(s) for s in og[f"bend_{name}_spacing"]))(u)
        loss.backward()
        assert _rel(loss, og[f"bend_{name}_loss"]) < TOL and _rel(u.grad, og[f"bend_{name}_grad"]) < TOL


def test_unet_light_and_argmax(cuda, ng):
    import deepatlas_b200 as da
    torch.manual_seed(230)
    net = da.get_network("UNet_light")(1, 4, bias=True, BN=True)
    net.weights_init()
    net = net.to(cuda).train()
    from parity_util import MaskRecorder, MaskReplay
    with MaskRecorder() as rec:
        logits = net(_c(ng["ul_x"], cuda))
    assert _rel(logits, ng["ul_logits"]) < TOL
    assert np.array_equal(torch.max(logits, 1)[1].cpu().numpy().astype(np.uint8), ng["ul_argmax"])
    loss = da.get_loss_function("dice")(n_class=4, weight_type="Uniform", softmax=True, eps=1e-6)(logits, _c(ng["ul_labels"], cuda))
    loss.backward()
    assert _rel(loss, ng["ul_loss"]) < TOL
    sd = net.state_dict()
    assert _rel(sd["encoders.0.0.BN.running_mean"], ng["ul_running_mean0"]) < TOL
    assert _rel(sd["encoders.0.0.BN.running_var"], ng["ul_running_var0"]) < TOL
    # Whole-network gradients sit on the precision ladder: the golden values are the REFERENCE's fp32
    # results, which carry their own round-off (batch-1 BatchNorm, 20 layers); the fp64 restatement is the truth and
    # our error against it must stay within max(1e-3, 3 x the reference's own error) -- same rule as test_gpu_nets.
    from oracle import ref_port as P
    from parity_util import check_grads_vs_truth, cpu_state
    torch.manual_seed(230)
    fresh = da.get_network("UNet_light")(1, 4, bias=True, BN=True)
    fresh.weights_init()
    sd64 = {k: (v.double().requires_grad_(True) if v.is_floating_point() and "running" not in k else (v.double() if v.is_floating_point() else v))
            for k, v in cpu_state(fresh).items()}
    with MaskReplay(rec.masks):   # the fp64 truth takes the CUDA path's activation branches (parity_util)
        l64 = P.dice_multiclass(P.unet_generator_forward(torch.from_numpy(ng["ul_x"]).double(), sd64, 1, True),
                                torch.from_numpy(ng["ul_labels"]).long(), 4, "Uniform", False, True, 1e-6)
    l64.backward()
    keys = [k[len("ul_grad/"):] for k in ng if k.startswith("ul_grad/")]
    params = dict(net.named_parameters())
    check_grads_vs_truth({k: params[k].grad for k in keys}, {k: torch.from_numpy(ng["ul_grad/" + k]) for k in keys},
                         {k: sd64[k].grad for k in keys}, 1e-3)


def test_voxelmorph_unet32_joint(cuda, ng):
    import deepatlas_b200 as da
    from deepatlas_b200.joint import JointModel, make_synthetic_pair
    torch.manual_seed(230)
    vm = da.get_network("voxel_morph_cvpr")()
    vm.weights_init()
    vm = vm.to(cuda)
    disp, warped, deform = vm(_c(ng["vm_s"], cuda), _c(ng["vm_t"], cuda))
    assert _rel(disp, ng["vm_disp"]) < TOL and _rel(warped, ng["vm_warped"]) < TOL and _rel(deform, ng["vm_deform"]) < TOL
    torch.manual_seed(230)
    un = da.get_network("UNet")(1, 4, bias=True, BN=True)
    un.weights_init()
    un = un.to(cuda).train()
    assert _rel(un(_c(ng["un_x"], cuda)), ng["un_logits"]) < TOL
    torch.manual_seed(230)    # same construction / init order as make_golden.py: seg, seg.weights_init, reg, reg.weights_init
    seg = da.get_network("UNet_light")(1, 4, bias=True, BN=True)
    seg.weights_init()
    reg = da.get_network("voxel_morph_cvpr")()
    reg.weights_init()
    model = JointModel(n_classes=4)
    model.seg.load_state_dict(seg.state_dict())
    model.reg.load_state_dict(reg.state_dict())
    model = model.to(cuda)
    batch = make_synthetic_pair((16, 16, 16), 4, seed=230, device=cuda)
    loss, parts = model.joint_loss(*batch)
    assert _rel(loss, ng["joint_loss"]) < TOL
    for k in ("sim", "reg", "ana", "sup"):
        assert _rel(parts[k], ng["joint_part_" + k]) < TOL, k
