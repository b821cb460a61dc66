"""GPU parity for SURVEY.md 8(f) rows 2-4: the remaining registry losses, the UNet_generator variants and the device-side
input stage -- the CUDA path (through the C ABI) against the golden fixtures made from the real reference modules
(tests/golden/extra.npz) and against the CPU oracle on further seeded inputs.  1e-4 max-norm relative."""
import os

import numpy as np
import pytest
import torch

from parity_util import rel_err
from test_oracle_golden_extra import VARIANTS

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-4


@pytest.fixture(scope="module")
def eg():
    return dict(np.load(os.path.join(GOLD, "extra.npz")))


def _rel(a, b, floor=1e-30):
    a, b = torch.as_tensor(a).detach().double().cpu(), torch.as_tensor(np.asarray(b)).double()
    return float((a - b).abs().max()) / max(float(b.abs().max()), floor)


def _c(a, cuda, grad=False):
    t = torch.from_numpy(np.asarray(a)).to(cuda)
    return t.requires_grad_(True) if grad else t


def _g(seed=230):
    return torch.Generator().manual_seed(seed)


def test_pair_losses_golden(cuda, eg):
    import deepatlas_b200 as da
    a, b = _c(eg["pair_a"], cuda, True), _c(eg["pair_b"], cuda, True)
    for name, args in (("ncc", (a, b)), ("mse", (a, b)), ("L2", (a,))):
        a.grad = b.grad = None
        loss = da.get_loss_function(name)()(*args)
        loss.backward()
        assert _rel(loss, eg[f"{name}_loss"]) < TOL and _rel(a.grad, eg[f"{name}_ga"]) < TOL, name
        if len(args) == 2:
            assert _rel(b.grad, eg[f"{name}_gb"]) < TOL, name


def test_ncc_identities_and_sizes(cuda):
    """ncc(a, a) = 0, ncc(a, -a) = 2, affine invariance; odd sizes and a batch with different samples."""
    import deepatlas_b200 as da
    from oracle import ref_port as P
    g = _g()
    crit = da.get_loss_function("ncc")()
    a = torch.rand((3, 1, 9, 11, 13), generator=g)
    b = torch.rand((3, 1, 9, 11, 13), generator=g) * 3 + 7
    ac, bc = a.to(cuda), b.to(cuda)
    assert abs(float(crit(ac, ac))) < 1e-6 and abs(float(crit(ac, -ac)) - 2) < 1e-6
    assert abs(float(crit(ac, 2.5 * ac + 4)) - 0) < 1e-5
    ar, br = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    P.ncc_loss(ar.double(), br.double()).backward()
    ag, bg = ac.clone().requires_grad_(True), bc.clone().requires_grad_(True)
    loss = crit(ag, bg)
    loss.backward()
    assert rel_err(loss, P.ncc_loss(a.double(), b.double())) < TOL
    assert rel_err(ag.grad, ar.grad) < TOL and rel_err(bg.grad, br.grad) < TOL
    # sum reduction of the mse mirror
    mse = da.get_loss_function("mse")(reduction="sum")(ac, bc)
    assert rel_err(mse, ((a.double() - b.double()) ** 2).sum()) < 1e-5


def test_gradient_loss_golden_and_oracle(cuda, eg):
    import deepatlas_b200 as da
    from oracle import ref_port as P
    u = _c(eg["grad_u"], cuda, True)
    for norm in ("L2", "L1"):
        for k in (0, 1):
            u.grad = None
            sp = tuple(float(v) for v in eg[f"grad_spacing_{k}"])
            loss = da.get_loss_function("gradient")(norm=norm, spacing=sp)(u)
            loss.backward()
            assert _rel(loss, eg[f"grad_{norm}_{k}_loss"]) < TOL and _rel(u.grad, eg[f"grad_{norm}_{k}_g"]) < TOL
    # minimal extents and odd sizes against the oracle
    for size in ((3, 3, 3), (5, 9, 17)):
        v = torch.randn((1, 3) + size, generator=_g()) * 0.2
        vr = v.clone().double().requires_grad_(True)
        P.gradient_loss(vr).backward()
        vg = v.to(cuda).requires_grad_(True)
        loss = da.get_loss_function("gradient")()(vg)
        loss.backward()
        assert rel_err(loss, P.gradient_loss(v.double())) < TOL and rel_err(vg.grad, vr.grad) < TOL


def _mirror_xent_cases(da, C, t, soft, w, alpha):
    L = da.get_loss_function
    from deepatlas_b200 import ops
    return {
        "ce": lambda x: L("cross_entropy")()(x, t),
        "ce_w": lambda x: L("cross_entropy")(weight=w)(x, t),
        "focal": lambda x: L("focal")(C)(x, t),
        "focal_a": lambda x: L("focal")(C, alpha=alpha, gamma=1.5, size_average=False)(x, t),
        "focal_nosm": lambda x: L("focal")(C, soft_max=False)(ops.softmax(x), t),
        "sce_sm": lambda x: L("soft_cross_entropy")(softmax=True)(x, soft),
        "sce": lambda x: L("soft_cross_entropy")(softmax=False)(ops.softmax(x), soft),
    }


def test_xent_family_golden(cuda, eg):
    import deepatlas_b200 as da
    x = _c(eg["xent_x"], cuda, True)
    soft = _c(eg["xent_soft"], cuda, True)
    for tdtype in (torch.uint8, torch.int64):
        t = _c(eg["xent_t"], cuda).to(tdtype)
        cases = _mirror_xent_cases(da, 5, t, soft, _c(eg["xent_w"], cuda), _c(eg["xent_alpha"], cuda))
        for name, fn in cases.items():
            x.grad = soft.grad = None
            loss = fn(x)
            loss.backward()
            assert _rel(loss, eg[f"{name}_loss"]) < TOL, name
            assert _rel(x.grad, eg[f"{name}_gx"]) < TOL, name
            if name.startswith("sce"):
                assert _rel(soft.grad, eg[f"{name}_gt"]) < TOL, name


def test_xent_ignore_index_clamp_and_32_classes(cuda):
    import deepatlas_b200 as da
    from oracle import ref_port as P
    g = _g()
    C, size = 32, (5, 6, 7)
    x = torch.randn((2, C) + size, generator=g) * 3
    t = torch.randint(0, C, (2,) + size, generator=g)
    t[0, 0, 0, :4] = -100
    xr = x.clone().double().requires_grad_(True)
    ref = P.cross_entropy(xr, t)
    ref.backward()
    xg = x.to(cuda).requires_grad_(True)
    loss = da.get_loss_function("cross_entropy")()(xg, t.to(cuda))
    loss.backward()
    assert rel_err(loss, ref) < TOL and rel_err(xg.grad, xr.grad) < TOL
    # the 1e-8 clamp of SoftCrossEntropy(softmax=False): zeros in the prediction
    p = torch.softmax(x, 1)
    p[0, :3, 0, 0, 0] = 0.0
    soft = torch.softmax(torch.randn((2, C) + size, generator=g), 1)
    pr = p.clone().double().requires_grad_(True)
    ref = P.soft_cross_entropy(pr, soft.double(), False)
    ref.backward()
    pg = p.to(cuda).requires_grad_(True)
    loss = da.get_loss_function("soft_cross_entropy")(softmax=False)(pg, soft.to(cuda))
    loss.backward()
    assert rel_err(loss, ref) < TOL and rel_err(pg.grad, pr.grad) < TOL
    with pytest.raises(NotImplementedError):
        da.get_loss_function("soft_cross_entropy")()(pg, t.to(cuda))


@pytest.mark.parametrize("size", [(3, 4, 5), (1, 1, 1), (8, 6, 10)])
def test_upsample_trilinear2(cuda, size):
    from deepatlas_b200 import ops
    g = _g()
    x = torch.rand((2, 3) + size, generator=g)
    xr = x.clone().requires_grad_(True)
    yr = torch.nn.functional.interpolate(xr, scale_factor=2, mode="trilinear")
    cot = torch.randn(yr.shape, generator=g)
    (yr * cot).sum().backward()
    xg = x.to(cuda).requires_grad_(True)
    yg = ops.upsample_trilinear2(xg)
    (yg * cot.to(cuda)).sum().backward()
    assert rel_err(yg, yr) < 1e-6 and rel_err(xg.grad, xr.grad) < 1e-5


@pytest.mark.parametrize("cin,cout,bias", [(8, 16, True), (3, 5, False), (16, 16, True)])
def test_conv_k2s2_is_the_deconv_adjoint(cuda, cin, cout, bias):
    from deepatlas_b200 import ops
    g = _g()
    x = torch.randn((2, cin, 6, 8, 12), generator=g)
    w = torch.randn((cout, cin, 2, 2, 2), generator=g) * 0.2
    b = torch.randn(cout, generator=g) if bias else None
    ins = [x, w] + ([b] if bias else [])
    ref_in = [t.clone().double().requires_grad_(True) for t in ins]
    yr = torch.nn.functional.conv3d(ref_in[0], ref_in[1], ref_in[2] if bias else None, stride=2)
    cot = torch.randn(yr.shape, generator=g)
    (yr * cot.double()).sum().backward()
    gi = [t.to(cuda).requires_grad_(True) for t in ins]
    yg = ops.conv_k2s2(gi[0], gi[1], gi[2] if bias else None)
    (yg * cot.to(cuda)).sum().backward()
    assert rel_err(yg, yr) < TOL
    for a, r in zip(gi, ref_in):
        assert rel_err(a.grad, r.grad) < TOL
    with pytest.raises(RuntimeError):
        ops.conv_k2s2(gi[0][:, :, :5], gi[1])


def test_residual_add_broadcast(cuda):
    from deepatlas_b200 import ops
    g = _g()
    a = torch.randn((2, 6, 4, 5, 6), generator=g)
    for cb in (6, 1):
        b = torch.randn((2, cb, 4, 5, 6), generator=g)
        ar, br = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
        cot = torch.randn(a.shape, generator=g)
        ((ar + br) * cot).sum().backward()
        ag, bg = a.to(cuda).requires_grad_(True), b.to(cuda).requires_grad_(True)
        out = ops.add(bg, ag)       # operand order must not matter
        (out * cot.to(cuda)).sum().backward()
        assert rel_err(out, a + b) < 1e-6 and rel_err(ag.grad, ar.grad) < 1e-6 and rel_err(bg.grad, br.grad) < 1e-5
    with pytest.raises(RuntimeError):
        ops.add(a.to(cuda), torch.zeros((2, 3, 4, 5, 6), device=cuda))


@pytest.mark.parametrize("name", list(VARIANTS))
def test_unet_generator_variants_golden(cuda, eg, name):
    """Forward against the golden output of the real reference class; parameter gradients on the precision ladder with
    activation-mask replay (tests/parity_util.py), like the whole-network tests of tests/test_gpu_nets.py."""
    from deepatlas_b200 import networks as M
    from oracle import ref_port as P
    from parity_util import MaskRecorder, MaskReplay, check_grads_vs_truth
    kw, enc, dec, ncls = VARIANTS[name]
    net = M.UNet_generator(enc, dec, act="LeakyReLU", **kw)(1, ncls, bias=True, BN=True)
    sd0 = {k[len(f"var_{name}_sd/"):]: torch.from_numpy(v) for k, v in eg.items() if k.startswith(f"var_{name}_sd/")}
    net.load_state_dict(sd0, strict=True)
    net.to(cuda).train()
    with MaskRecorder() as rec:
        y = net(_c(eg["var_x"], cuda))
    assert _rel(y, eg[f"var_{name}_y"]) < TOL
    cot = torch.from_numpy(eg[f"var_{name}_cot"])
    (y * cot.to(cuda)).sum().backward()
    cfg = dict(encoders=enc, decoders=dec, act="LeakyReLU", **kw)
    x = torch.from_numpy(eg["var_x"])

    def leaf(v, dt):
        v = v.to(dt) if v.is_floating_point() else v
        return v.clone().requires_grad_(True) if v.is_floating_point() else v
    sd32 = {k: (leaf(v, torch.float32) if "running" not in k else v.clone()) for k, v in sd0.items()}
    sd64 = {k: (leaf(v, torch.float64) if "running" not in k else (v.double() if v.is_floating_point() else v)) for k, v in sd0.items()}
    replay = MaskReplay(rec.masks)
    with replay:
        (P.unet_generator_forward(x, sd32, 1, True, cfg=cfg) * cot).sum().backward()
        replay.restart()
        (P.unet_generator_forward(x.double(), sd64, 1, True, cfg=cfg) * cot.double()).sum().backward()
    assert replay.flips <= 16
    ours = {k: p.grad.detach().cpu() for k, p in net.named_parameters()}
    check_grads_vs_truth(ours, {k: v.grad for k, v in sd32.items() if v.is_floating_point() and v.requires_grad},
                         {k: v.grad for k, v in sd64.items() if v.is_floating_point() and v.requires_grad}, 1e-3)


def test_input_stage(cuda):
    from deepatlas_b200 import input_stage, ops
    g = _g()
    img = torch.rand((1, 20, 24, 28), generator=g) * 1.6 - 0.3       # leaves [0, 1] on both sides
    seg = torch.randint(0, 32, (20, 24, 28), generator=g, dtype=torch.uint8)
    for crop in (None, [2, 3, 4], [1, 2, 3, 4, 5, 6]):
        lo, size = input_stage.crop_window(img.shape, crop)
        ref_img = img[:, lo[0]:lo[0] + size[0], lo[1]:lo[1] + size[1], lo[2]:lo[2] + size[2]].clamp(0, 1)
        ref_seg = seg[lo[0]:lo[0] + size[0], lo[1]:lo[1] + size[1], lo[2]:lo[2] + size[2]]
        stage = input_stage.DeviceInputStage(cuda, crop_size=crop)
        stage.submit(img, seg)
        stage.submit(img * 0.5, seg)
        di, ds = stage.get()
        assert di.is_cuda and ds.dtype == torch.uint8
        assert torch.equal(di.cpu(), ref_img) and torch.equal(ds.cpu(), ref_seg)
        di2, _ = stage.get()
        assert torch.equal(di2.cpu(), (img * 0.5)[:, lo[0]:lo[0] + size[0], lo[1]:lo[1] + size[1], lo[2]:lo[2] + size[2]].clamp(0, 1))
    # the CropTensor arithmetic itself (lib/transforms.py:145-157)
    assert input_stage.crop_window((1, 200, 200, 200), [10, 20, 20]) == ((10, 20, 20), (180, 160, 160))
    with pytest.raises(ValueError):
        input_stage.crop_window((1, 8, 8, 8), [1, 2])
    with pytest.raises(RuntimeError):
        ops.crop_clip(torch.zeros(1, 4, 4, 4), (0, 0, 0), (2, 2, 2))
    # prefetch(): same samples, same order, uint8 labels feed the Dice loss directly
    import deepatlas_b200 as da
    loader = [(img[None] * s, seg[None], f"case{i}") for i, s in enumerate((1.0, 0.7, 0.4))]
    got = list(input_stage.prefetch(loader, cuda))
    assert [n for _, _, n in got] == ["case0", "case1", "case2"]
    for (di, ds, _), (hi, hs, _) in zip(got, loader):
        assert torch.equal(di.cpu(), hi.clamp(0, 1)) and torch.equal(ds.cpu(), hs)
    logits = torch.randn((1, 32, 20, 24, 28), generator=g).to(cuda)
    crit = da.get_loss_function("dice")(n_class=32, weight_type="Uniform", softmax=True, eps=1e-6)
    assert torch.equal(crit(logits, got[0][1]), crit(logits, got[0][1].long()))


def test_dice_on_label(cuda):
    import deepatlas_b200 as da
    from deepatlas_b200 import evaluation
    from oracle import ref_port as P
    g = _g()
    a = torch.randint(0, 32, (2, 1, 9, 10, 11), generator=g)
    b = torch.randint(0, 32, (2, 1, 9, 10, 11), generator=g)
    b[0][b[0] == 3] = 0
    counts = evaluation.label_overlap_counts(a.to(cuda), b.to(torch.uint8).to(cuda))
    for c in (0, 3, 31):   # exact integer counts
        assert int(counts[1, 0, c]) == int((a[1] == c).sum()) and int(counts[0, 1, c]) == int((b[0] == c).sum())
        assert int(counts[1, 2, c]) == int(((a[1] == c) & (b[1] == c)).sum())
    for dt in (torch.uint8, torch.int64, torch.float32):
        for wt in ("Uniform", "Simple"):
            ours = da.DiceLossOnLabel()(a.to(dt).to(cuda), b.to(dt).to(cuda), weight_type=wt)
            assert abs(float(ours) - float(P.dice_on_label(a, b, None, 10e-6, wt))) < 1e-6, (dt, wt)
    assert abs(float(da.DiceLossOnLabel()(a.to(cuda), a.to(cuda)))) < 1e-4     # identical maps: loss 0 up to eps


@pytest.mark.parametrize("size", [(12, 14, 16), (66, 70, 68), (132, 130, 136)])
def test_lncc_multiscale(cuda, size):
    """All three branches of the scale schedule.  The reference's fp32 expressions cancel, so the
    bar is the precision ladder: our error against the fp64 restatement <= max(1e-4, 3 x the fp32 restatement's own)."""
    import deepatlas_b200 as da
    from oracle import ref_port as P
    g = _g()
    I, J = torch.rand((1, 1) + size, generator=g), torch.rand((1, 1) + size, generator=g)
    J = 0.6 * J + 0.4 * I                      # correlated pair: lncc well away from 0
    I64, J64 = I.double().requires_grad_(True), J.double().requires_grad_(True)
    truth = P.lncc_multiscale(I64, J64)
    truth.backward()
    I32, J32 = I.clone().requires_grad_(True), J.clone().requires_grad_(True)
    ref32 = P.lncc_multiscale(I32, J32)
    ref32.backward()
    Ig, Jg = I.to(cuda).requires_grad_(True), J.to(cuda).requires_grad_(True)
    loss = da.LNCCLoss()(Ig, Jg)
    loss.backward()
    assert rel_err(loss, truth) < max(TOL, 3 * rel_err(ref32, truth))
    for ours, r32, t64 in ((Ig.grad, I32.grad, I64.grad), (Jg.grad, J32.grad, J64.grad)):
        assert rel_err(ours, t64) < max(TOL, 3 * rel_err(r32, t64))


@pytest.mark.parametrize("C", [4, 32])
@pytest.mark.parametrize("dtype", [torch.uint8, torch.int64])
def test_softmax_dice_with_probs_equals_the_two_pass_form(cuda, C, dtype):
    """forward_with_probs == (forward, softmax) and its single backward pass == the sum of the two separate backwards."""
    import deepatlas_b200 as da
    from deepatlas_b200 import ops
    g = _g()
    size = (9, 10, 12)
    x = torch.randn((2, C) + size, generator=g)
    t = torch.randint(0, C, (2,) + size, generator=g).to(dtype).to(cuda)
    cot = torch.randn((2, C) + size, generator=g).to(cuda)
    crit = da.get_loss_function("dice")(n_class=C, weight_type="Uniform", softmax=True, eps=1e-6)
    xa = x.to(cuda).requires_grad_(True)
    la, pa = crit.forward_with_probs(xa, t)
    (3.0 * la + (pa * cot).sum()).backward()
    xb = x.to(cuda).requires_grad_(True)
    lb, pb = crit(xb, t), ops.softmax(xb)
    (3.0 * lb + (pb * cot).sum()).backward()
    assert torch.equal(la, lb) and torch.equal(pa, pb)
    assert rel_err(xa.grad, xb.grad) < 1e-5
    # ... and against the CPU oracle (softmax + DiceLossMultiClass of the reference), values and the combined gradient
    from oracle import ref_port as P
    xo = x.clone().requires_grad_(True)
    lo, po = P.dice_multiclass(xo, t.cpu().long(), C, "Uniform", False, True, 1e-6), torch.softmax(xo, 1)
    (3.0 * lo + (po * cot.cpu()).sum()).backward()
    assert rel_err(la, lo) < TOL and rel_err(pa, po) < TOL and rel_err(xa.grad, xo.grad) < TOL
    # only one of the two outputs used downstream
    xc = x.to(cuda).requires_grad_(True)
    lc, _ = crit.forward_with_probs(xc, t)
    lc.backward()
    xd = x.to(cuda).requires_grad_(True)
    crit(xd, t).backward()
    assert rel_err(xc.grad, xd.grad) < 1e-6


@pytest.mark.parametrize("C", [4, 32])
@pytest.mark.parametrize("dtype", [torch.uint8, torch.int64])
@pytest.mark.parametrize("size", [(9, 10, 12), (8, 6, 67 * 2)])
def test_head_softmax_dice_vs_oracle(cuda, C, dtype, size):
    """1x1x1 head + softmax + Dice as one kernel each way (DiceLossMultiClass.forward_head) against the CPU oracle's
    conv3d -> DiceLossMultiClass(softmax=True) in fp32 and fp64, with a second consumer of the probabilities:
    loss, probabilities, feature / weight / bias gradients."""
    import torch.nn.functional as F

    import deepatlas_b200 as da
    from deepatlas_b200 import networks as M
    from oracle import ref_port as P
    g = _g()
    N, K = 2, 16
    feat = torch.randn((N, K) + size, generator=g)
    w = torch.randn((C, K, 1, 1, 1), generator=g) * 0.4
    b = torch.randn((C,), generator=g) * 0.2
    t = torch.randint(0, C, (N,) + size, generator=g)
    cot = torch.randn((N, C) + size, generator=g) * 0.01

    def oracle(dt):
        f_, w_, b_ = (x.to(dt).clone().requires_grad_(True) for x in (feat, w, b))
        logits = F.conv3d(f_, w_, b_)
        loss = P.dice_multiclass(logits, t, n_class=C, weight_type="Uniform", softmax=True, eps=1e-6)
        probs = torch.softmax(logits, dim=1)
        (3.0 * loss + (probs * cot.to(dt)).sum()).backward()
        return loss.detach(), probs.detach(), f_.grad, w_.grad, b_.grad

    head = M._Conv1x1(K, C, kernel_size=1, stride=1, padding=0, bias=True).to(cuda)
    with torch.no_grad():
        head.weight.copy_(w.to(cuda))
        head.bias.copy_(b.to(cuda))
    fg = feat.to(cuda).requires_grad_(True)
    crit = da.get_loss_function("dice")(n_class=C, weight_type="Uniform", softmax=True, eps=1e-6)
    from deepatlas_b200 import ops
    assert ops.head_dice_supported(K, C, feat[0, 0].numel())
    loss, probs = crit.forward_head(fg, head, t.to(dtype).to(cuda), want_probs=True)
    (3.0 * loss + (probs * cot.to(cuda)).sum()).backward()
    r32, r64 = oracle(torch.float32), oracle(torch.float64)
    ours = (loss, probs, fg.grad, head.weight.grad, head.bias.grad)
    for name, o, a32, a64 in zip(("loss", "probs", "dfeat", "dweight", "dbias"), ours, r32, r64):
        assert rel_err(o, a64) < max(TOL, 3 * rel_err(a32, a64)), name
    # without the second consumer: the probabilities are not written, the backward takes the Dice part only
    fg2 = feat.to(cuda).requires_grad_(True)
    head.zero_grad()
    loss2, none = crit.forward_head(fg2, head, t.to(dtype).to(cuda))
    assert none is None and rel_err(loss2, r64[0]) < TOL
    loss2.backward()
    f_, w_, b_ = (x.double().clone().requires_grad_(True) for x in (feat, w, b))
    P.dice_multiclass(F.conv3d(f_, w_, b_), t, n_class=C, weight_type="Uniform", softmax=True, eps=1e-6).backward()
    for name, o, a64 in (("dfeat", fg2.grad, f_.grad), ("dweight", head.weight.grad, w_.grad), ("dbias", head.bias.grad, b_.grad)):
        assert rel_err(o, a64) < TOL, name
    # and it equals the unfused path of the same package (separate head, softmax-Dice calls)
    fg3 = feat.to(cuda).requires_grad_(True)
    loss3 = crit(head(fg3), t.to(dtype).to(cuda))
    assert rel_err(loss3, loss2) < 1e-5


@pytest.mark.parametrize("kind", ["convT3_res", "k2s2", "conv_res"])
def test_vm_blocks_residual_and_deconv_vs_reference_semantics(cuda, kind):
    """modules.convBlock(residual=True) (``x += x``, modules.py:59-60) and modules.deconvBlock (ConvTranspose3d k3 s1 p1 with
    ``x += input``, k2 s2; modules.py:65-86) on the GPU against their restatement on torch's CPU ops: values, input and
    parameter gradients."""
    import torch.nn.functional as F

    from deepatlas_b200 import networks as M
    g = _g()
    C, size = 8, (6, 8, 10)
    x = torch.randn((2, C) + size, generator=g)
    if kind == "conv_res":
        blk = M.convBlockVM(C, C, stride=1, bias=True, residual=True).to(cuda)
        w, b = blk.conv.weight, blk.conv.bias
        ref = lambda xx, ww, bb: 2.0 * torch.relu(F.conv3d(xx, ww, bb, padding=1))  # noqa: E731
    elif kind == "convT3_res":
        blk = M.deconvBlockVM(C, C, 3, stride=1, padding=1, bias=True, residual=True).to(cuda)
        w, b = blk.deconv.weight, blk.deconv.bias
        ref = lambda xx, ww, bb: torch.relu(F.conv_transpose3d(xx, ww, bb, stride=1, padding=1)) + xx  # noqa: E731
    else:
        blk = M.deconvBlockVM(C, 12, 2, stride=2, bias=True).to(cuda)
        w, b = blk.deconv.weight, blk.deconv.bias
        ref = lambda xx, ww, bb: torch.relu(F.conv_transpose3d(xx, ww, bb, stride=2))  # noqa: E731
    with torch.no_grad():
        w.copy_((torch.randn(w.shape, generator=g) * 0.2).to(cuda))
        b.copy_((torch.randn(b.shape, generator=g) * 0.1).to(cuda))
    xg = x.to(cuda).requires_grad_(True)
    y = blk(xg)
    xc, wc, bc = x.clone().requires_grad_(True), w.detach().cpu().clone().requires_grad_(True), b.detach().cpu().clone().requires_grad_(True)
    yc = ref(xc, wc, bc)
    cot = torch.randn(yc.shape, generator=g)
    (y * cot.to(cuda)).sum().backward()
    (yc * cot).sum().backward()
    assert rel_err(y, yc) < TOL
    for name, a, r in (("dx", xg.grad, xc.grad), ("dw", w.grad, wc.grad), ("db", b.grad, bc.grad)):
        assert rel_err(a, r) < 2 * TOL, (kind, name)


def test_focal_loss_2d_inputs_vs_oracle(cuda):
    """FocalLoss on (observations, classes) inputs -- the reference's other accepted shape (lib/loss.py:167) -- against the
    oracle: value and gradient."""
    import deepatlas_b200 as da
    from oracle import ref_port as P
    g = _g()
    n, C = 301, 5
    x = torch.randn((n, C), generator=g)
    t = torch.randint(0, C, (n,), generator=g)
    xg = x.to(cuda).requires_grad_(True)
    loss = da.get_loss_function("focal")(C, gamma=2)(xg, t.to(cuda))
    loss.backward()
    xc = x.clone().requires_grad_(True)
    ref = P.focal_loss(xc, t, gamma=2)
    ref.backward()
    assert rel_err(loss, ref) < TOL and rel_err(xg.grad, xc.grad) < TOL
