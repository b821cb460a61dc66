"""GPU parity tests for whole networks and the joint seg+reg step (CUDA path vs the CPU oracle)."""
import pytest
import torch

from parity_util import MASK_BAND, MaskRecorder, MaskReplay, check_grads_vs_truth, conv_impl, cpu_state, max_flips, oracle_joint_loss, rel_err, report

pytestmark = pytest.mark.gpu
TOL = 1e-4        # forward outputs (north star)
GTOL = 1e-3       # parameter gradients through 15-30 layers with batch-1 BatchNorm (fp32 round-off amplification;
                  # the reference's own fp32-vs-fp64 gradient noise at these depths is of the same order)


def _param_grads(model):
    return {k: p.grad.detach().cpu() for k, p in model.named_parameters() if p.grad is not None}


# impl "umma": every k3 s1 p1 forward, data gradient and weight gradient of the network runs on the tcgen05 kernels
# (conv3d_umma*_kernel, conv3d_wgrad_umma*_kernel) whatever the size heuristics say -- the kernels that carry the
# benchmark step -- including the multi-chunk accumulation, the two-source concatenation and the fused bias/activation
IMPLS = ["auto", "umma", "umma_tf32"]


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("n_classes,size,bn", [(4, (32, 32, 32), True), (32, (16, 24, 16), True), (3, (16, 16, 24), False)])
def test_unet_light(cuda, n_classes, size, bn, impl):
    import deepatlas_b200 as da
    from oracle import ref_port as P
    torch.manual_seed(230)
    net = da.get_network("UNet_light")(1, n_classes, bias=True, BN=bn).to(cuda)
    net.weights_init()
    g = torch.Generator().manual_seed(230)
    x = torch.rand((1, 1) + size, generator=g)
    lab = torch.randint(0, n_classes, (1,) + size, generator=g, dtype=torch.uint8)
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in cpu_state(net).items()}
    sd64 = {k: (v.double().requires_grad_(True) if v.is_floating_point() and "running" not in k else (v.double() if v.is_floating_point() else v))
            for k, v in cpu_state(net).items()}
    crit = da.get_loss_function("dice")(n_class=n_classes, weight_type="Uniform", softmax=True, eps=1e-6)
    with conv_impl(impl):
        with MaskRecorder() as rec:   # the oracle takes the CUDA path's activation branches (parity_util.MaskReplay)
            y = net(x.to(cuda))
        loss = crit(y, lab.to(cuda))
        loss.backward()
    stats = {}
    replay = MaskReplay(rec.masks, MASK_BAND[impl])
    with replay:
        y_ref = P.unet_generator_forward(x, sd, 1, bn, stats_out=stats)
        replay.restart()
        y64 = P.unet_generator_forward(x.double(), sd64, 1, bn)
    assert replay.flips <= max_flips(replay), f"{replay.flips} activation-mask flips"  # each verified to sit within round-off of zero
    P.dice_multiclass(y64, lab.long(), n_classes, "Uniform", False, True, 1e-6).backward()
    y64 = y64.detach()
    loss_ref = P.dice_multiclass(y_ref, lab.long(), n_classes, "Uniform", False, True, 1e-6)
    loss_ref.backward()
    assert y.shape == y_ref.shape
    assert rel_err(y, y_ref) < max(TOL, 2 * rel_err(y_ref, y64)), f"logits rel err {rel_err(y, y_ref):.3e}"
    assert rel_err(loss, loss_ref) < TOL
    # label argmax indices: bit-exact wherever the oracle's own top-2 margin is above fp32 round-off
    top2 = y64.topk(2, dim=1).values
    decided = (top2[:, 0] - top2[:, 1]) > 1e-5 * y64.abs().max()
    assert torch.equal(y.argmax(1).cpu()[decided], y_ref.argmax(1)[decided])
    assert decided.float().mean() > 0.99
    w = check_grads_vs_truth(_param_grads(net), {k: v.grad for k, v in sd.items() if v.is_floating_point() and v.requires_grad},
                             {k: v.grad for k, v in sd64.items() if v.is_floating_point() and v.requires_grad}, GTOL)
    report("unet_light", impl=impl, classes=n_classes, size=list(size), bn=bn, logits_rel_err_vs_fp32=rel_err(y, y_ref),
           logits_rel_err_vs_fp64=rel_err(y, y64), reference_fp32_vs_fp64=rel_err(y_ref, y64), mask_flips=replay.flips, **w)
    if bn:  # running statistics updated exactly like nn.BatchNorm3d
        for k, v in stats.items():
            assert rel_err(net.state_dict()[k], v) < 1e-4, k
        assert int(net.state_dict()["encoders.0.0.BN.num_batches_tracked"]) == 1


def test_unet_32base(cuda):
    import deepatlas_b200 as da
    from oracle import ref_port as P
    torch.manual_seed(230)
    net = da.get_network("UNet")(1, 4, bias=True, BN=True).to(cuda)
    net.weights_init()
    x = torch.rand((1, 1, 16, 16, 16), generator=torch.Generator().manual_seed(230))
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in cpu_state(net).items()}
    sd64 = {k: (v.double().requires_grad_(True) if v.is_floating_point() and "running" not in k else (v.double() if v.is_floating_point() else v))
            for k, v in cpu_state(net).items()}
    with MaskRecorder() as rec:
        y = net(x.to(cuda))
    replay = MaskReplay(rec.masks)
    with replay:
        y_ref = P.unet_forward(x, sd, True)
        replay.restart()
        y64 = P.unet_forward(x.double(), sd64, True)
    assert rel_err(y, y64) < max(TOL, 3 * rel_err(y_ref, y64)), f"UNet logits rel err {rel_err(y, y64):.3e}"
    cot = torch.randn(y_ref.shape, generator=torch.Generator().manual_seed(7))
    (y * cot.to(cuda)).sum().backward()
    (y_ref * cot).sum().backward()
    (y64 * cot.double()).sum().backward()
    check_grads_vs_truth(_param_grads(net), {k: v.grad for k, v in sd.items() if v.is_floating_point() and v.requires_grad},
                         {k: v.grad for k, v in sd64.items() if v.is_floating_point() and v.requires_grad}, GTOL)


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("size", [(32, 48, 32), (16, 16, 16), (24, 20, 36)])
def test_voxelmorph(cuda, size, impl):
    import deepatlas_b200 as da
    from oracle import ref_port as P
    torch.manual_seed(230)
    net = da.get_network("voxel_morph_cvpr")().to(cuda)
    net.weights_init()
    g = torch.Generator().manual_seed(230)
    s, t = torch.rand((1, 1) + size, generator=g), torch.rand((1, 1) + size, generator=g)
    sd = {k: v.clone().requires_grad_(True) for k, v in cpu_state(net).items()}
    with conv_impl(impl), MaskRecorder() as rec:
        out = net(s.to(cuda), t.to(cuda))
    replay = MaskReplay(rec.masks, MASK_BAND[impl])
    with replay:
        ref = P.voxelmorph_forward(s, t, sd)
    for name, a, b in zip(("disp", "warped", "deform"), out, ref):
        assert a.shape == b.shape
        assert rel_err(a, b) < TOL, f"{name}: rel err {rel_err(a, b):.3e}"
    lncc, bend = da.get_loss_function("lncc")(), da.get_loss_function("bendingEnergy")()
    with conv_impl(impl):
        (lncc(out[1], t.to(cuda)) + 1000.0 * bend(out[0])).backward()
    (P.lncc(ref[1], t) + 1000.0 * P.bending_energy(ref[0])).backward()
    sd64 = {k: v.double().requires_grad_(True) for k, v in cpu_state(net).items()}
    replay.restart()
    with replay:
        r64 = P.voxelmorph_forward(s.double(), t.double(), sd64)
    (P.lncc(r64[1], t.double()) + 1000.0 * P.bending_energy(r64[0])).backward()
    w = check_grads_vs_truth(_param_grads(net), {k: v.grad for k, v in sd.items()}, {k: v.grad for k, v in sd64.items()}, GTOL)
    report("voxelmorph", impl=impl, size=list(size), disp_rel_err=rel_err(out[0], ref[0]), warped_rel_err=rel_err(out[1], ref[1]), **w)


# the last case is sized so that the AUTOMATIC selection (what bench.py runs) takes the tcgen05 forward / data-gradient /
# weight-gradient kernels on the full- and half-resolution levels (>= 65 536 voxels, W >= 20/24) and the tiled FFMA
# kernels below, i.e. the same mix of kernels as the 160x192x160 benchmark step
@pytest.mark.parametrize("C,size,impl", [(4, (16, 16, 16), "auto"), (32, (16, 24, 16), "auto"), (4, (16, 16, 16), "umma"),
                                         (32, (16, 24, 16), "umma"), (4, (16, 16, 16), "umma_tf32"), (8, (64, 64, 128), "auto")])
def test_joint_step(cuda, C, size, impl):
    from deepatlas_b200.joint import JointModel, make_synthetic_pair
    from oracle import ref_port as P
    torch.manual_seed(230)
    model = JointModel(n_classes=C).to(cuda)
    model.weights_init()
    batch = make_synthetic_pair(size, C, seed=230, device=cuda)
    with conv_impl(impl):
        with MaskRecorder() as rec:
            loss, parts = model.joint_loss(*batch)
        loss.backward()
    replay = MaskReplay(rec.masks, MASK_BAND[impl])
    ref_loss, ref_grads = oracle_joint_loss(model, batch, P, replay=replay)
    true_loss, true_grads = oracle_joint_loss(model, batch, P, dtype=torch.float64, replay=replay)
    assert rel_err(loss, true_loss) < max(TOL, 3 * rel_err(ref_loss, true_loss))
    ours = _param_grads(model)
    assert len(true_grads) >= 60
    w = check_grads_vs_truth(ours, ref_grads, true_grads, GTOL)
    report("joint_step", impl=impl, classes=C, size=list(size), loss_rel_err_vs_fp64=rel_err(loss, true_loss),
           reference_fp32_loss_rel_err=rel_err(ref_loss, true_loss), mask_flips=replay.flips, **w)


def test_registry_errors_and_install(built_lib):
    import deepatlas_b200 as da
    with pytest.raises(KeyError):
        da.get_network("nope")
    with pytest.raises(KeyError):
        da.get_loss_function("nope")
