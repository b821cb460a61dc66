"""GPU parity tests, op by op: the CUDA path (through the C ABI) against the CPU oracle
(oracle/ref_port.py) on the same seeded inputs.  Tolerance: 1e-4 max-norm relative (north star) for
fp32 outputs and gradients unless a test states otherwise; index results bit-exact."""
import pytest
import torch
import torch.nn.functional as F

from parity_util import conv_impl, rel_err, rel_err_quantile

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _g(seed=230):
    return torch.Generator().manual_seed(seed)


def _run_both(fn_gpu, fn_cpu, inputs, cuda, grad_mask=None):
    """inputs: list of CPU tensors; float ones get requires_grad per grad_mask. Returns (out_gpu, grads_gpu, out_cpu, grads_cpu)."""
    grad_mask = grad_mask or [t.is_floating_point() for t in inputs]
    gi = [t.to(cuda).requires_grad_(m) if t.is_floating_point() else t.to(cuda) for t, m in zip(inputs, grad_mask)]
    ci = [t.clone().requires_grad_(m) if t.is_floating_point() else t.clone() for t, m in zip(inputs, grad_mask)]
    og, oc = fn_gpu(*gi), fn_cpu(*ci)
    og_l = og if isinstance(og, (tuple, list)) else [og]
    oc_l = oc if isinstance(oc, (tuple, list)) else [oc]
    gen = _g(7)
    cot = [torch.randn(o.shape, generator=gen) for o in oc_l]
    sum((o * c.to(cuda)).sum() for o, c in zip(og_l, cot)).backward()
    sum((o * c).sum() for o, c in zip(oc_l, cot)).backward()
    gg = [t.grad if (t.is_floating_point() and m) else None for t, m in zip(gi, grad_mask)]
    cg = [t.grad if (t.is_floating_point() and m) else None for t, m in zip(ci, grad_mask)]
    return og_l, gg, oc_l, cg


def _check(og, gg, oc, cg, tol=TOL, what=""):
    for i, (a, b) in enumerate(zip(og, oc)):
        e = rel_err(a, b)
        assert e < tol, f"{what} output {i}: rel err {e:.3e}"
    for i, (a, b) in enumerate(zip(gg, cg)):
        if b is None:
            continue
        assert a is not None, f"{what} grad {i} missing"
        e = rel_err(a, b)
        assert e < tol, f"{what} grad {i}: rel err {e:.3e}"


# ----------------------------------------------------------------------------------------------------------
def test_no_cpu_fallback(built_lib):
    from deepatlas_b200 import ops
    with pytest.raises(RuntimeError):
        ops.softmax(torch.zeros(1, 2, 4, 4, 4))


@pytest.mark.parametrize("C,size", [(1, (12, 14, 16)), (5, (9, 10, 11)), (8, (9, 10, 12))])
@pytest.mark.parametrize("add_id", [True, False])
def test_warp3d(cuda, C, size, add_id):
    from deepatlas_b200 import ops
    from oracle import ref_port as P
    g = _g()
    src = torch.rand((2, C) + size, generator=g)
    field = torch.randn((2, 3) + size, generator=g) * 0.35   # large enough to leave the volume at the borders
    if not add_id:
        field = field + P.identity_transform(size)[None]

    def cpu(s, f):
        phi = f + P.identity_transform(size)[None] if add_id else f
        return P.warp(s, phi), phi

    def gpu(s, f):
        return ops.warp3d(s, f, add_identity=add_id, want_phi=True)

    res = _run_both(gpu, cpu, [src, field], cuda)
    _check(*res, what="warp3d")
    # independent closed form (fp64) as the truth rung
    truth = P.warp_closed_form(src.double(), (field + P.identity_transform(size)[None] if add_id else field).double())
    assert rel_err(res[0][0], truth) < TOL


def test_warp3d_identity_reproduces_input(cuda):
    from deepatlas_b200 import ops
    src = torch.rand((1, 2, 8, 9, 10), generator=_g()).to(cuda)
    out = ops.warp3d(src, torch.zeros((1, 3, 8, 9, 10), device=cuda), add_identity=True)
    assert rel_err(out, src) < 1e-5


@pytest.mark.parametrize("C", [2, 4, 7, 32])
@pytest.mark.parametrize("softmax", [True, False])
@pytest.mark.parametrize("dtype", [torch.uint8, torch.int64])
def test_dice_hard_target(cuda, C, softmax, dtype):
    import deepatlas_b200 as da
    from oracle import ref_port as P
    g = _g()
    size = (10, 12, 14)
    x = torch.randn((2, C) + size, generator=g)
    if not softmax:
        x = torch.softmax(x, 1)
    t = torch.randint(0, C, (2,) + size, generator=g).to(dtype)
    for wt in ("Uniform", "Simple", "Volume"):
        for no_bg in (False, True):
            crit = da.get_loss_function("dice")(n_class=C, weight_type=wt, no_bg=no_bg, softmax=softmax, eps=1e-6)
            res = _run_both(lambda a, b: crit(a, b), lambda a, b: P.dice_multiclass(a, b.long(), C, wt, no_bg, softmax, 1e-6),
                            [x, t], cuda)
            _check(*res, what=f"dice {wt} no_bg={no_bg}")


def test_dice_soft_target_and_identities(cuda):
    import deepatlas_b200 as da
    from oracle import ref_port as P
    g = _g()
    C, size = 4, (8, 8, 8)
    s = torch.softmax(torch.randn((1, C) + size, generator=g), 1)
    t = torch.softmax(torch.randn((1, C) + size, generator=g), 1)
    crit = da.get_loss_function("dice")(n_class=C, weight_type="Uniform", softmax=False, eps=1e-6)
    res = _run_both(lambda a, b: crit(a, b), lambda a, b: P.dice_multiclass(a, b, C, "Uniform", False, False, 1e-6), [s, t], cuda)
    _check(*res, what="dice soft")
    # Dice(one-hot(t), t) == 0 exactly-ish (SURVEY section 4 identity)
    lab = torch.randint(0, C, (1,) + size, generator=g)
    oh = P.mask_to_one_hot(lab.reshape(1, 1, -1), C).reshape((1, C) + size)
    assert abs(float(crit(oh.to(cuda), lab.to(cuda)))) < 1e-6
    with pytest.raises(ValueError):
        crit(s.to(cuda), torch.zeros((1, C + 1) + size, device=cuda))


def test_softmax_op(cuda):
    from deepatlas_b200 import ops
    x = torch.randn((2, 6, 5, 6, 7), generator=_g()) * 3
    res = _run_both(ops.softmax, lambda a: torch.softmax(a, 1), [x], cuda)
    _check(*res, what="softmax")


@pytest.mark.parametrize("size", [(12, 14, 16), (9, 9, 9), (20, 11, 13)])
def test_lncc(cuda, size):
    import deepatlas_b200 as da
    from oracle import ref_port as P
    g = _g()
    I = torch.rand((2, 1) + size, generator=g)
    J = torch.rand((2, 1) + size, generator=g)
    crit = da.get_loss_function("lncc")().to(cuda)
    og, gg, oc, cg = _run_both(lambda a, b: crit(a, b), lambda a, b: P.lncc(a, b), [I, J], cuda)
    # truth rung: fp64 oracle.  Parity rule for LNCC: our error vs fp64 <= max(1e-4, reference fp32 error vs fp64)
    Id, Jd = I.double().requires_grad_(True), J.double().requires_grad_(True)
    ld = P.lncc(Id, Jd)
    cot = torch.randn(oc[0].shape, generator=_g(7))
    (ld * cot.double()).sum().backward()
    assert rel_err(og[0], ld) < max(TOL, rel_err(oc[0], ld))
    for ours, ref32, truth in zip(gg, cg, (Id.grad, Jd.grad)):
        assert rel_err(ours, truth) < max(TOL, rel_err(ref32, truth))
    # identities: LNCC(I,I) == 0, LNCC(I, aI+b) ~ 0
    Ic = I.to(cuda)
    assert abs(float(crit(Ic, Ic))) < 1e-5
    assert abs(float(crit(Ic, 2.0 * Ic + 0.5))) < 1e-4


@pytest.mark.parametrize("size,spacing", [((10, 12, 14), (1, 1, 1)), ((7, 9, 8), (1, 2, 3))])
def test_bending(cuda, size, spacing):
    import deepatlas_b200 as da
    from oracle import ref_port as P
    u = torch.randn((2, 3) + size, generator=_g()) * 0.1
    crit = da.get_loss_function("bendingEnergy")(spacing=spacing)
    res = _run_both(lambda a: crit(a), lambda a: P.bending_energy(a, spacing), [u], cuda)
    _check(*res, what="bending")
    # the reference's non-L2 path (lib/loss.py:721 falls through): means of absolute second differences
    crit1 = da.get_loss_function("bendingEnergy")(norm="L1", spacing=spacing)
    res = _run_both(lambda a: crit1(a), lambda a: P.bending_energy(a, spacing, norm="L1"), [u], cuda)
    _check(*res, what="bending L1")
    # affine field has zero bending energy
    idt = P.identity_transform(size)[None].to(cuda)
    assert float(crit(idt * 0.3 + 0.1)) < 1e-10


CONV_CASES = [
    # C1, C2, Cout, ks, stride, size, transposed
    (1, 0, 8, 3, 1, (8, 10, 12), False),
    (2, 0, 16, 3, 1, (8, 10, 12), False),
    (8, 0, 16, 3, 1, (36, 20, 48), False),      # tiled path (V >= 32768, W >= 16)
    (16, 32, 16, 3, 1, (32, 34, 40), False),    # tiled, two sources, ragged tiles
    (8, 16, 3, 3, 1, (32, 32, 36), False),      # flow-like head Cout=3
    (20, 0, 12, 3, 1, (33, 32, 35), False),     # odd extents, channel counts off the register tiles
    (16, 0, 32, 3, 2, (16, 18, 20), False),     # stride 2
    (32, 0, 32, 3, 2, (9, 11, 13), False),      # stride 2, odd extents
    (16, 0, 32, 3, 2, (20, 24, 40), False),     # stride 2 through the TMA-staged tiled weight gradient (W % 8 == 0)
    (6, 0, 12, 3, 2, (10, 18, 72), False),      # stride 2, ragged channel groups, two x tiles
    (16, 0, 7, 1, 1, (8, 9, 10), False),        # 1x1 head
    (16, 0, 32, 1, 1, (16, 20, 24), False),     # 1x1 head, 32 classes (streaming k1 kernels, 16-channel blocks)
    (8, 8, 5, 1, 1, (16, 20, 24), False),       # 1x1 streaming kernel, two sources, 8-channel block partly filled
    (40, 0, 24, 1, 1, (12, 16, 32), False),     # 1x1 streaming kernel, second channel block partly filled
    (24, 8, 16, 3, 1, (8, 8, 8), True),         # ConvTranspose3d k3 s1 p1, two sources
    (16, 0, 8, 3, 1, (32, 32, 32), True),       # ConvTranspose3d through the tiled kernel
    (6, 0, 16, 3, 1, (32, 32, 32), False),      # TMA path, last channel chunk partial (zero-filled by the TMA unit)
    (8, 6, 8, 3, 1, (32, 36, 32), False),       # TMA path, two sources, second one partial
    (32, 16, 16, 3, 1, (8, 64, 160), False),    # full-width rows of the benchmark volume (decBlock2.0 channels)
    (64, 64, 64, 3, 1, (20, 24, 20), False),    # deepest decoder level of UNet_light at the benchmark size
    (16, 0, 32, 3, 1, (32, 32, 40), False),     # 32-channel output blocks (tensor-core path: CB = 32, KC = 8)
    (24, 0, 48, 3, 1, (30, 33, 44), False),     # ragged everything: partial channel chunk / block, odd extents
    (12, 9, 16, 3, 1, (16, 20, 24), False),     # the concatenation boundary falls inside an 8-channel K chunk of the fp16 tensor path
    (35, 30, 20, 3, 1, (12, 12, 24), False),    # 65 input channels: two 32-channel launches + a 16-channel one holding a single channel
    (96, 0, 32, 3, 1, (8, 12, 40), False),      # three full 32-channel launches accumulate through the output
    (160, 96, 80, 3, 1, (6, 8, 24), False),     # wide layers of the 32-base UNet: eight accumulating launches, five output blocks
]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("impl", ["auto", "direct", "ffma", "umma", "umma_tf32"])
@pytest.mark.parametrize("slope", [None, 0.0])
def test_conv3d(cuda, case, impl, slope):
    from deepatlas_b200 import _lib, ops
    C1, C2, Cout, ks, stride, size, transposed = case
    # activation-mask flips: a pre-activation within fp32 round-off of zero may land on either side
    band, max_flips = 1e-5, 4
    if impl not in ("auto", "umma") and slope is not None:
        pytest.skip("activation epilogue covered by the auto run")
    g = _g()
    x1 = torch.randn((2, C1) + size, generator=g)
    x2 = torch.randn((2, C2) + size, generator=g) if C2 else None
    Cin = C1 + C2
    w = torch.randn((Cin, Cout, ks, ks, ks) if transposed else (Cout, Cin, ks, ks, ks), generator=g) * (2.0 / (Cin * ks ** 3)) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    pad = 1 if ks == 3 else 0

    mask_box = {}

    def cpu(*a):
        a = list(a)
        xx = torch.cat((a[0], a[1]), 1) if C2 else a[0]
        ww, bb = a[-2], a[-1]
        y = F.conv_transpose3d(xx, ww, bb, stride=1, padding=1) if transposed else F.conv3d(xx, ww, bb, stride=stride, padding=pad)
        if slope is None:
            return y
        # The activation is discontinuous in its derivative at 0: a pre-activation within fp32 round-off of zero may
        # land on either side depending on summation order.  Those voxels (checked below to be round-off cases)
        # take the GPU's side so that the gradient comparison is not dominated by a legitimate mask flip.
        pos = y.detach() > 0
        flipped = pos != mask_box["gpu_pos"]
        mask_box["flipped"] = int(flipped.sum())
        if mask_box["flipped"]:
            assert float(y.detach()[flipped].abs().max()) < band * float(y.detach().abs().max()), "mask differs away from zero"
        return torch.where(mask_box["gpu_pos"], y, y * slope)

    def gpu(*a):
        a = list(a)
        out = ops.conv3d(a[0], a[-2], a[-1], x2=a[1] if C2 else None, transposed=transposed, stride=stride, pad=pad, slope=slope)
        mask_box["gpu_pos"] = out.detach().cpu() > 0
        return out

    if impl.startswith("umma") and (ks != 3 or stride != 1):
        pytest.skip("tensor-core path covers k3 s1 p1")
    ins = [x1] + ([x2] if C2 else []) + [w, b]
    with conv_impl(impl):
        res = _run_both(gpu, cpu, ins, cuda)
    _check(*res, what=f"conv3d {case} {impl} slope={slope}")
    if slope is not None:
        assert mask_box["flipped"] <= max_flips, f"{mask_box['flipped']} activation-mask flips"


@pytest.mark.parametrize("slope", [0.01, 0.0, None])
@pytest.mark.parametrize("training", [True, False])
def test_bn_act(cuda, slope, training):
    from deepatlas_b200 import ops
    g = _g()
    C = 6
    x = torch.randn((2, C, 6, 7, 9), generator=g) * 2 + 0.5
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1
    rm0, rv0 = torch.randn(C, generator=g) * 0.1, torch.rand(C, generator=g) + 0.5
    rm_g, rv_g = rm0.clone().to(cuda), rv0.clone().to(cuda)
    rm_c, rv_c = rm0.clone(), rv0.clone()

    def cpu(a, ga, be):
        y = F.batch_norm(a, rm_c, rv_c, ga, be, training=training, momentum=0.1, eps=1e-5)
        return y if slope is None else F.leaky_relu(y, slope)

    def gpu(a, ga, be):
        return ops.bn_act(a, ga, be, rm_g, rv_g, training=training, momentum=0.1, eps=1e-5, slope=slope)

    res = _run_both(gpu, cpu, [x, gamma, beta], cuda)
    _check(*res, what="bn_act")
    assert rel_err(rm_g, rm_c) < 1e-5 and rel_err(rv_g, rv_c) < 1e-5


def test_maxpool2_first_max_tie_rule(cuda):
    from deepatlas_b200 import ops
    g = _g()
    # heavy ties: ReLU-like data with many exact zeros and repeated values
    x = torch.randint(0, 3, (2, 3, 8, 10, 12), generator=g).float()
    res = _run_both(ops.maxpool2, lambda a: F.max_pool3d(a, 2), [x], cuda)
    for a, b in zip(res[0], res[2]):
        assert torch.equal(a.cpu(), b)
    assert torch.equal(res[1][0].cpu(), res[3][0])      # gradient routed to the same (first) maximum
    xo = torch.randn((1, 2, 7, 9, 11), generator=g)      # odd extents (floor)
    res = _run_both(ops.maxpool2, lambda a: F.max_pool3d(a, 2), [xo], cuda)
    assert torch.equal(res[0][0].cpu(), res[2][0]) and torch.equal(res[1][0].cpu(), res[3][0])


@pytest.mark.parametrize("sin,sout", [((5, 6, 5), (10, 12, 10)), ((3, 4, 5), (7, 9, 10)), ((4, 4, 4), (5, 4, 13))])
def test_upsample_nearest(cuda, sin, sout):
    from deepatlas_b200 import ops
    x = torch.randn((2, 3) + sin, generator=_g())
    res = _run_both(lambda a: ops.upsample_nearest(a, sout), lambda a: F.interpolate(a, size=sout), [x], cuda)
    assert torch.equal(res[0][0].cpu(), res[2][0])
    assert rel_err(res[1][0], res[3][0]) < 1e-6


@pytest.mark.parametrize("Cin,Cout,size", [(8, 8, (4, 5, 6)), (32, 32, (6, 6, 8)), (6, 10, (3, 4, 5)), (64, 64, (5, 6, 5)),
                                           (12, 20, (3, 6, 40)), (32, 32, (4, 10, 36)),  # these two: TMA-tiled weight gradient, ragged groups / tiles
                                           # tensor-core (mma.sync 3xTF32) kernels: 32 / 64 input channels, W % 8 == 0 for the weight gradient
                                           (64, 64, (3, 4, 16)), (32, 64, (3, 5, 24)), (64, 32, (2, 3, 8)), (32, 32, (7, 9, 40))])
def test_deconv_k2s2(cuda, Cin, Cout, size):
    from deepatlas_b200 import ops
    g = _g()
    x = torch.randn((2, Cin) + size, generator=g)
    w = torch.randn((Cin, Cout, 2, 2, 2), generator=g) * (1.0 / Cin) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    res = _run_both(lambda a, ww, bb: ops.deconv_k2s2(a, ww, bb), lambda a, ww, bb: F.conv_transpose3d(a, ww, bb, stride=2), [x, w, b], cuda)
    _check(*res, what="deconv_k2s2")


@pytest.mark.parametrize("C,size,dtype", [(4, (9, 10, 12), torch.uint8), (32, (8, 12, 16), torch.int64), (6, (7, 9, 11), torch.uint8)])
def test_warped_dice_fused_vs_oracle(cuda, C, size, dtype):
    """The fused anatomy term against the oracle's grid_sample + Dice, all three weightings, values and both gradients."""
    import deepatlas_b200 as da
    from oracle import ref_port as P
    g = _g()
    prob = torch.softmax(torch.randn((2, C) + size, generator=g), 1)
    phi = P.identity_transform(size)[None] + torch.randn((2, 3) + size, generator=g) * 0.3
    lab = torch.randint(0, C, (2,) + size, generator=g).to(dtype)
    for wt in ("Uniform", "Simple", "Volume"):
        crit = da.get_loss_function("dice")(n_class=C, weight_type=wt, softmax=False, eps=1e-6)
        pg, fg = prob.to(cuda).requires_grad_(True), phi.to(cuda).requires_grad_(True)
        loss = crit.forward_warped(pg, fg, lab.to(cuda))
        loss.backward()
        pc, fc = prob.clone().requires_grad_(True), phi.clone().requires_grad_(True)
        ref = P.dice_multiclass(P.warp(pc, fc), lab.long(), C, wt, False, False, 1e-6)
        ref.backward()
        assert rel_err(loss, ref) < TOL, wt
        assert rel_err(pg.grad, pc.grad) < TOL and rel_err(fg.grad, fc.grad) < TOL, wt
        # and against the unfused CUDA path (warp3d + dice_sums)
        p2, f2 = prob.to(cuda).requires_grad_(True), phi.to(cuda).requires_grad_(True)
        crit(da.ops.warp3d(p2, f2), lab.to(cuda)).backward()
        assert rel_err(pg.grad, p2.grad) < TOL and rel_err(fg.grad, f2.grad) < TOL
    # the scatter formulation (default) and the gather kernel (fixed summation order) give the same sums; the
    # label-driven sums T and I do not depend on the formulation at all
    s_scatter = da.ops.warped_dice_sums(prob.to(cuda), phi.to(cuda), lab.to(cuda))
    s_gather = da.ops.warped_dice_sums(prob.to(cuda), phi.to(cuda), lab.to(cuda), deterministic=True)
    assert rel_err(s_scatter[:, 0], s_gather[:, 0]) < 1e-5
    assert torch.equal(s_scatter[:, 1], s_gather[:, 1])
    assert rel_err(s_scatter[:, 2], s_gather[:, 2]) < 1e-5
    # out-of-range labels (ignored by the one-hot) and a field that leaves the volume: zero padding on both paths
    lab2 = lab.long().clone(); lab2[:, 0] = C + 3
    far = phi * 1.6
    a = da.ops.warped_dice_sums(prob.to(cuda), far.to(cuda), lab2.to(cuda))
    b = da.ops.warped_dice_sums(prob.to(cuda), far.to(cuda), lab2.to(cuda), deterministic=True)
    assert rel_err(a, b) < 1e-5


@pytest.mark.parametrize("C,size,dtype", [(4, (9, 10, 11), torch.uint8), (32, (8, 12, 16), torch.int64)])
def test_eval_argmax_and_dice_bit_exact(cuda, C, size, dtype):
    """Validation step: label argmax indices and per-class Dice (scipy formula) bit-exact against the oracle, including
    ties (first maximum wins) and a class absent from both maps (nan, as scipy)."""
    from deepatlas_b200 import evaluation
    from oracle import ref_port as P
    g = _g()
    logits = torch.randn((2, C) + size, generator=g)
    logits[:, 2] = logits[:, 1]                      # exact ties between classes 1 and 2 everywhere class 1 would win
    logits[:, C - 1] = -50.0                         # last class never predicted ...
    truth = torch.randint(0, C - 1, (2,) + size, generator=g).to(dtype)   # ... and never present
    counts, pred = evaluation.argmax_counts(logits.to(cuda), truth.to(cuda))
    ref_dice, ref_labels = P.eval_dice_per_class(logits, truth.long(), C)
    assert torch.equal(pred.cpu().long(), ref_labels)
    got = evaluation.dice_per_class(logits.to(cuda), truth.to(cuda)).cpu()
    assert torch.equal(torch.isnan(got), torch.isnan(ref_dice)) and bool(torch.isnan(got[:, -1]).all())
    ok = ~torch.isnan(ref_dice)
    assert torch.equal(got[ok], ref_dice[ok])
    assert int(counts[:, 0].sum()) == 2 * logits[0, 0].numel() and int(counts[:, 0, 2].sum()) == 0


@pytest.mark.parametrize("slope", [None, 0.0, 0.01])
def test_bn_act_output_bounds(cuda, slope):
    """The max-abs bounds batch norm attaches to its output and to its input gradient (consumed by the convolutions'
    tensor-core path instead of a pass over the tensor) are true upper bounds, and tight: the forward one is the exact
    maximum up to round-off, the backward one a triangle inequality."""
    from deepatlas_b200 import ops
    g = _g()
    C, size = 6, (9, 10, 12)
    x = (torch.randn((2, C) + size, generator=g) * 3 + 1).to(cuda).requires_grad_(True)
    ga = (torch.randn(C, generator=g)).to(cuda).requires_grad_(True)
    be = (torch.randn(C, generator=g)).to(cuda).requires_grad_(True)
    rm, rv = torch.zeros(C, device=cuda), torch.ones(C, device=cuda)
    y = ops.bn_act(x, ga, be, rm, rv, training=True, slope=slope)
    a = ops._get_amax(y)
    assert a is not None
    ymax = float(y.detach().abs().max())
    assert ymax <= float(a) <= ymax * 1.001
    p = ops.maxpool2(y)
    assert ops._get_amax(p) is a
    seen = {}

    class Probe(torch.autograd.Function):
        @staticmethod
        def forward(ctx, t):
            return t * 1.0

        @staticmethod
        def backward(ctx, gt):
            seen["amax"], seen["max"] = ops._get_amax(gt), float(gt.abs().max())
            return gt

    xp = Probe.apply(x)
    y2 = ops.bn_act(xp, ga, be, rm, rv, training=True, slope=slope)
    (y2 * torch.randn(y2.shape, generator=g).to(cuda)).sum().backward()
    assert seen["amax"] is not None
    assert seen["max"] <= float(seen["amax"]) <= 4 * seen["max"]
    with torch.no_grad():
        y.add_(1.0)   # an in-place change invalidates the bound
    assert ops._get_amax(y) is None


F16X1_TOL = 5e-3   # declared tolerance of the reduced-precision mode: operands rounded to 11 significant bits, fp32 accumulate


@pytest.mark.parametrize("case", [(16, 0, 16, (8, 10, 24)), (32, 16, 16, (6, 9, 40)), (64, 0, 32, (5, 6, 20))])
def test_conv3d_single_pass_mode(cuda, case):
    """da_set_conv_split(2): one tcgen05 MMA per product on the scaled fp16 operands (BASELINE config C2's
    reduced-precision mode).  Forward, data and weight gradients against the fp64 convolution within the declared 5e-3,
    and measurably different from the default 3xFP16 path (i.e. the mode is really taken)."""
    from deepatlas_b200 import ops
    C1, C2, Cout, size = case
    g = _g()
    x1 = torch.randn((1, C1) + size, generator=g)
    x2 = torch.randn((1, C2) + size, generator=g) if C2 else None
    Cin = C1 + C2
    w = torch.randn((Cout, Cin, 3, 3, 3), generator=g) * (2.0 / (Cin * 27)) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    cot = torch.randn((1, Cout) + size, generator=g)
    xx = (torch.cat((x1, x2), 1) if C2 else x1).double().requires_grad_(True)
    wd, bd = w.double().requires_grad_(True), b.double().requires_grad_(True)
    (F.conv3d(xx, wd, bd, padding=1) * cot.double()).sum().backward()
    truth = (F.conv3d(xx, wd, bd, padding=1).detach(), xx.grad, wd.grad, bd.grad)
    errs = {}
    for impl in ("umma", "umma_f16x1"):
        a1 = x1.to(cuda).requires_grad_(True)
        a2 = x2.to(cuda).requires_grad_(True) if C2 else None
        wg, bg = w.to(cuda).requires_grad_(True), b.to(cuda).requires_grad_(True)
        with conv_impl(impl):
            y = ops.conv3d(a1, wg, bg, x2=a2, stride=1, pad=1)
            (y * cot.to(cuda)).sum().backward()
        gx = torch.cat((a1.grad, a2.grad), 1) if C2 else a1.grad
        errs[impl] = [rel_err(o, t) for o, t in zip((y, gx, wg.grad, bg.grad), truth)]
    assert max(errs["umma"][:3]) < TOL, errs
    assert max(errs["umma_f16x1"]) < F16X1_TOL, errs
    # half-precision rounding is visible, i.e. the single-pass branch ran (forward and data gradient; the weight gradient
    # of a narrow volume runs the 3xTF32 kernel, which has no single-pass mode)
    assert min(errs["umma_f16x1"][:2]) > 1e-5, errs
