"""Host-side mirror of the reference interface, checked on CPU by substituting torch restatements for the
device kernels' SUMS (the closing formulas, argument checking and registries are host logic)."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_port as P

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cpu_dice_sums(source, target, apply_softmax=False):
    N, C = source.shape[:2]
    p = torch.softmax(source, 1) if apply_softmax else source
    p = p.reshape(N, C, -1)
    if target.is_floating_point():
        t = target.reshape(N, C, -1)
    else:
        t = P.mask_to_one_hot(target.reshape(N, 1, -1), C, dtype=p.dtype)
    return torch.stack([p.sum(2), t.sum(2), (p * t).sum(2)], 1)


def _cpu_bending_sums(u):
    B, C = u.shape[:2]
    c = u[:, :, 1:-1, 1:-1, 1:-1]
    s = lambda t: (t ** 2).reshape(B, C, -1).sum(2)  # noqa: E731
    terms = [s(u[:, :, 2:, 1:-1, 1:-1] + u[:, :, :-2, 1:-1, 1:-1] - 2 * c),
             s(u[:, :, 1:-1, 2:, 1:-1] + u[:, :, 1:-1, :-2, 1:-1] - 2 * c),
             s(u[:, :, 1:-1, 1:-1, 2:] + u[:, :, 1:-1, 1:-1, :-2] - 2 * c),
             s(u[:, :, 2:, 2:, 1:-1] + u[:, :, :-2, :-2, 1:-1] - u[:, :, 2:, :-2, 1:-1] - u[:, :, :-2, 2:, 1:-1]),
             s(u[:, :, 1:-1, 2:, 2:] + u[:, :, 1:-1, :-2, :-2] - u[:, :, 1:-1, 2:, :-2] - u[:, :, 1:-1, :-2, 2:]),
             s(u[:, :, 2:, 1:-1, 2:] + u[:, :, :-2, 1:-1, :-2] - u[:, :, 2:, 1:-1, :-2] - u[:, :, :-2, 1:-1, 2:])]
    return torch.stack(terms, 2)


@pytest.mark.parametrize("wt", ["Uniform", "Simple", "Volume"])
@pytest.mark.parametrize("softmax", [1, 0])
@pytest.mark.parametrize("no_bg", [0, 1])
@pytest.mark.parametrize("tgt", ["hard", "soft"])
def test_dice_closing_formula_against_golden(monkeypatch, wt, softmax, no_bg, tgt):
    from deepatlas_b200 import losses, ops
    monkeypatch.setattr(ops, "dice_sums", _cpu_dice_sums)
    g = np.load(os.path.join(GOLD, "ops.npz"))
    logits = torch.from_numpy(g["dice_logits"])
    x = logits if softmax else torch.softmax(logits, 1)
    target = torch.from_numpy(g["dice_labels"]) if tgt == "hard" else torch.from_numpy(g["dice_soft"])
    crit = losses.DiceLossMultiClass(n_class=4, weight_type=wt, no_bg=bool(no_bg), softmax=bool(softmax), eps=1e-6)
    got = float(crit(x, target))
    want = float(g[f"dice_{wt}_{softmax}_{no_bg}_{tgt}_loss"])
    assert abs(got - want) <= 2e-6 * max(1.0, abs(want))


def test_dice_argument_errors():
    from deepatlas_b200 import losses
    crit = losses.DiceLossMultiClass(n_class=4, weight_type="Uniform")
    with pytest.raises(ValueError):
        crit(torch.zeros(1, 4, 4, 4, 4), torch.zeros(1, 1, 1, 4, 4, 4))
    with pytest.raises(ValueError):
        losses.DiceLossMultiClass(weight_type="nope")(torch.zeros(1, 4, 4, 4, 4), torch.zeros(1, 4, 4, 4, dtype=torch.long))


@pytest.mark.parametrize("name", ["iso", "aniso"])
def test_bending_coefficients_against_golden(monkeypatch, name):
    from deepatlas_b200 import losses, ops
    monkeypatch.setattr(ops, "bending_sums", _cpu_bending_sums)
    g = np.load(os.path.join(GOLD, "ops.npz"))
    u = torch.from_numpy(g[f"bend_{name}_u"])
    crit = losses.BendingEnergyLoss(spacing=tuple(float(s) for s in g[f"bend_{name}_spacing"]))
    got, want = float(crit(u)), float(g[f"bend_{name}_loss"])
    assert abs(got - want) <= 2e-6 * abs(want)


def test_registries_and_unbuilt_variants():
    import deepatlas_b200 as da
    assert set(da.get_available_networks()) == {"voxel_morph_cvpr", "UNet", "UNet_light"}
    assert set(da.get_available_losses()) == {"ncc", "lncc", "mse", "gradient", "bendingEnergy", "dice", "L2", "focal",
                                              "cross_entropy", "soft_cross_entropy"}
    with pytest.raises(KeyError):
        da.get_network("nope")
    with pytest.raises(KeyError):
        da.get_loss_function("nope")
    with pytest.raises(NotImplementedError):
        da.get_loss_function("cross_entropy")(label_smoothing=0.1)
    assert da.get_loss_function("bendingEnergy")(norm="L1").norm == "L1"     # the reference's non-L2 path is built
    with pytest.raises(RuntimeError):
        da.install()                                       # reference modules not imported -> loud


def test_lncc_keeps_reference_state_dict_key():
    import deepatlas_b200 as da
    crit = da.get_loss_function("lncc")()
    assert list(crit.state_dict().keys()) == ["filter"] and tuple(crit.filter.shape) == (1, 1, 9, 9, 9)


# ---- closing formulas of the remaining registry losses, variant construction, crop arithmetic
def _cpu_pair_moments(a, b=None):
    a = a.double()
    b = torch.zeros_like(a) if b is None else b.double()
    am, bm = a - a.mean(1, keepdim=True), b - b.mean(1, keepdim=True)
    return torch.stack([a.sum(1), b.sum(1), (a * a).sum(1), (b * b).sum(1), (a * b).sum(1), ((a - b) ** 2).sum(1),
                        (am * am).sum(1), (bm * bm).sum(1), (am * bm).sum(1)], dim=1)


def _cpu_gradient_sums(u, l1=False):
    f = (lambda r: r.abs()) if l1 else (lambda r: r * r)
    return torch.stack([f(u[:, :, 2:] - u[:, :, :-2]).sum((2, 3, 4)), f(u[:, :, :, 2:] + u[:, :, :, :-2]).sum((2, 3, 4)),
                        f(u[:, :, :, :, 2:] + u[:, :, :, :, :-2]).sum((2, 3, 4))], dim=2)


def _cpu_xent_sums(x, target, mode, class_weight=None, gamma=0.0, focal_softmax=True, ignore_index=-100):
    lp = torch.log_softmax(x, 1)
    nvox = float(x.numel() // x.shape[1])
    if mode == 2:
        return torch.stack([(-target * lp).sum(), torch.tensor(nvox)])
    if mode == 3:
        return torch.stack([(-target * torch.log(x.clamp(min=1e-8))).sum(), torch.tensor(nvox)])
    t = target.long()
    keep = (t != ignore_index) if mode == 0 else torch.ones_like(t, dtype=torch.bool)
    tt = torch.where(keep, t, torch.zeros_like(t))
    w = torch.ones(x.shape[1]) if class_weight is None else class_weight.reshape(-1)
    lpt = lp.gather(1, tt[:, None]).squeeze(1)
    if mode == 0:
        return torch.stack([(-w[tt] * lpt * keep).sum(), (w[tt] * keep).sum()])
    P = torch.softmax(x, 1) if focal_softmax else x
    Pt = P.gather(1, tt[:, None]).squeeze(1)
    return torch.stack([(-w[tt] * (1 + Pt) ** gamma * lpt).sum(), torch.tensor(nvox)])


@pytest.fixture(scope="module")
def extra_gold():
    return dict(np.load(os.path.join(GOLD, "extra.npz")))


def test_remaining_losses_closing_formulas_against_golden(monkeypatch, extra_gold):
    """The host-side halves of the new registry entries (deepatlas_b200/losses.py) on CPU stand-ins for the CUDA sums,
    against the values of the real reference classes."""
    import deepatlas_b200 as da
    from deepatlas_b200 import ops
    g = extra_gold
    monkeypatch.setattr(ops, "pair_moments", _cpu_pair_moments)
    monkeypatch.setattr(ops, "gradient_sums", _cpu_gradient_sums)
    monkeypatch.setattr(ops, "xent_sums", _cpu_xent_sums)
    a, b = torch.from_numpy(g["pair_a"]), torch.from_numpy(g["pair_b"])
    L = da.get_loss_function
    close = lambda x, y: abs(float(x) - float(y)) <= 2e-6 * max(abs(float(y)), 1e-30)  # noqa: E731
    assert close(L("ncc")()(a, b), g["ncc_loss"]) and close(L("mse")()(a, b), g["mse_loss"]) and close(L("L2")()(a), g["L2_loss"])
    u = torch.from_numpy(g["grad_u"])
    for norm in ("L2", "L1"):
        for k in (0, 1):
            sp = tuple(float(v) for v in g[f"grad_spacing_{k}"])
            assert close(L("gradient")(norm=norm, spacing=sp)(u), g[f"grad_{norm}_{k}_loss"]), (norm, k)
    x, t = torch.from_numpy(g["xent_x"]), torch.from_numpy(g["xent_t"])
    soft, w, alpha = torch.from_numpy(g["xent_soft"]), torch.from_numpy(g["xent_w"]), torch.from_numpy(g["xent_alpha"])
    assert close(L("cross_entropy")()(x, t), g["ce_loss"]) and close(L("cross_entropy")(weight=w)(x, t), g["ce_w_loss"])
    assert close(L("focal")(5)(x, t), g["focal_loss"])
    assert close(L("focal")(5, alpha=alpha, gamma=1.5, size_average=False)(x, t), g["focal_a_loss"])
    assert close(L("focal")(5, soft_max=False)(torch.softmax(x, 1), t), g["focal_nosm_loss"])
    assert close(L("soft_cross_entropy")(softmax=True)(x, soft), g["sce_sm_loss"])
    assert close(L("soft_cross_entropy")(softmax=False)(torch.softmax(x, 1), soft), g["sce_loss"])
    with pytest.raises(NotImplementedError):
        L("soft_cross_entropy")()(x, t)
    with pytest.raises(ValueError):
        L("soft_cross_entropy")()(x, soft[:, :3])


def test_pair_moment_gradient_coefficients():
    """ops.pair_moment_coefs: the affine gradient alpha*a + beta*b + gamma reproduces autograd for a random function of
    the nine moments."""
    from deepatlas_b200 import ops
    gen = torch.Generator().manual_seed(230)
    a = torch.rand((3, 50), generator=gen, dtype=torch.float64, requires_grad=True)
    b = torch.rand((3, 50), generator=gen, dtype=torch.float64, requires_grad=True)
    gup = torch.randn((3, 9), generator=gen, dtype=torch.float64)
    m = _cpu_pair_moments(a, b)
    (m * gup).sum().backward()
    ca, cb = ops.pair_moment_coefs(gup, m.detach(), 50.0)
    ga = ca[:, 0:1] * a.detach() + ca[:, 1:2] * b.detach() + ca[:, 2:3]
    gb = cb[:, 0:1] * b.detach() + cb[:, 1:2] * a.detach() + cb[:, 2:3]
    assert float((ga - a.grad).abs().max()) < 1e-12 and float((gb - b.grad).abs().max()) < 1e-12


def test_variant_construction_and_crop_window():
    import deepatlas_b200 as da
    from deepatlas_b200 import input_stage, networks
    net = da.UNet_generator([(4, 8), (8, 8, 16)], [(16, 8, 8)], act="LeakyReLU", upsample=True, maxpool=False, res=False)(1, 3, bias=True, BN=True)
    keys = list(net.state_dict().keys())
    assert "down_samplers.0.weight" in keys and not any(k.startswith("up_samplers") for k in keys)
    assert tuple(net.down_samplers[0].weight.shape) == (8, 8, 2, 2, 2)
    blk = networks.deconvBlockVM(8, 8, 3, stride=1, padding=1, bias=True, batchnorm=True, residual=True)
    assert set(blk.state_dict()) >= {"deconv.weight", "deconv.bias", "bn.weight", "bn.running_mean"}
    with pytest.raises(NotImplementedError):
        networks.deconvBlockVM(8, 8, 4, stride=2, padding=1)
    assert input_stage.crop_window((1, 200, 200, 200), [10, 20, 20]) == ((10, 20, 20), (180, 160, 160))
    assert input_stage.crop_window((20, 24, 28), [1, 2, 3, 4, 5, 6]) == ((1, 2, 3), (15, 17, 19))
    assert input_stage.crop_window((20, 24, 28), None) == ((0, 0, 0), (20, 24, 28))
    with pytest.raises(ValueError):
        input_stage.crop_window((8, 8, 8), [1, 2])
    with pytest.raises(ValueError):
        input_stage.crop_window((8, 8, 8), [4, 4, 4])
    with pytest.raises(RuntimeError):
        input_stage.DeviceInputStage("cpu")


def test_dice_on_label_closing_formula(monkeypatch):
    import deepatlas_b200 as da
    from deepatlas_b200 import evaluation

    def cpu_counts(a, b, bins=256):
        a, b = a.reshape(a.shape[0], -1).long(), b.reshape(b.shape[0], -1).long()
        oh = lambda t: torch.nn.functional.one_hot(t, bins).sum(1)  # noqa: E731
        both = torch.where(a == b, a, torch.full_like(a, bins))
        inter = torch.nn.functional.one_hot(both, bins + 1).sum(1)[:, :bins]
        return torch.stack([oh(a), oh(b), inter], dim=1)
    monkeypatch.setattr(evaluation, "label_overlap_counts", cpu_counts)
    g = torch.Generator().manual_seed(230)
    a = torch.randint(0, 6, (2, 1, 6, 7, 8), generator=g)
    b = torch.randint(0, 6, (2, 1, 6, 7, 8), generator=g)
    b[0][b[0] == 3] = 0
    for wt in ("Uniform", "Simple"):
        for n_class in (None, 8):
            ours = da.DiceLossOnLabel(n_class=n_class)(a, b, weight_type=wt)
            assert abs(float(ours) - float(P.dice_on_label(a, b, n_class, 10e-6, wt))) < 1e-6, (wt, n_class)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under deepatlas_b200/ may import it (or tests/), and only bench.py's CPU
    legs and __graft_entry__.smoke() may do so at the repo root."""
    import ast
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "deepatlas_b200")
    for fn in sorted(os.listdir(pkg)):
        if not fn.endswith(".py"):
            continue
        tree = ast.parse(open(os.path.join(pkg, fn)).read())
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            for n in names:
                assert not (n == "oracle" or n.startswith("oracle.") or n == "tests" or n.startswith("tests.") or n == "parity_util"), (fn, n)
