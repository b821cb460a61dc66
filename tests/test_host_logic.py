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
    with pytest.raises(NotImplementedError):
        losses.BendingEnergyLoss(norm="L1")


def test_registries_and_unbuilt_variants():
    import deepatlas_b200 as da
    assert set(da.get_available_networks()) == {"voxel_morph_cvpr", "UNet", "UNet_light"}
    assert set(da.get_available_losses()) == {"lncc", "bendingEnergy", "dice"}
    with pytest.raises(KeyError):
        da.get_network("nope")
    with pytest.raises(KeyError):
        da.get_loss_function("nope")
    with pytest.raises(NotImplementedError):
        da.UNet_generator([(8, 16)], [], upsample=True)
    with pytest.raises(RuntimeError):
        da.install()                                       # reference modules not imported -> loud


def test_lncc_keeps_reference_state_dict_key():
    import deepatlas_b200 as da
    crit = da.get_loss_function("lncc")()
    assert list(crit.state_dict().keys()) == ["filter"] and tuple(crit.filter.shape) == (1, 1, 9, 9, 9)
