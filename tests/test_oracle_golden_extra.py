"""The CPU port against tests/golden/extra.npz (remaining registry losses, UNet_generator variants; made from the real
reference modules by `python oracle/make_golden.py --extra-only`).  Runs anywhere."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_port as P

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VARIANTS = {"strided": (dict(maxpool=False), [(4, 8), (8, 8, 16)], [(8, 8, 8)], 3),
            "upsample": (dict(upsample=True), [(4, 8), (8, 8, 16)], [(16, 8, 8)], 3),
            "res": (dict(res=True), [(8, 8), (8, 8)], [(8, 8)], 8),
            "all": (dict(maxpool=False, upsample=True, res=True), [(8, 8), (8, 8)], [(8, 8)], 8)}


@pytest.fixture(scope="module")
def eg():
    return dict(np.load(os.path.join(GOLD, "extra.npz")))


def _t(a, grad=False):
    t = torch.from_numpy(np.asarray(a))
    return t.requires_grad_(True) if grad else t


def _close(a, b, tol=1e-6, floor=1e-30):
    a, b = torch.as_tensor(a).detach().double(), torch.as_tensor(np.asarray(b)).double()
    return float((a - b).abs().max()) <= tol * max(float(b.abs().max()), floor)


def xent_cases(C, t, soft, w, alpha):
    """name -> function of x returning the loss (the same table drives the GPU test with the mirror classes)."""
    return {
        "ce": lambda x: P.cross_entropy(x, t),
        "ce_w": lambda x: P.cross_entropy(x, t, weight=w),
        "focal": lambda x: P.focal_loss(x, t),
        "focal_a": lambda x: P.focal_loss(x, t, alpha, 1.5, size_average=False),
        "focal_nosm": lambda x: P.focal_loss(torch.softmax(x, 1), t, soft_max=False),
        "sce_sm": lambda x: P.soft_cross_entropy(x, soft, True),
        "sce": lambda x: P.soft_cross_entropy(torch.softmax(x, 1), soft, False),
    }


def test_pair_and_gradient_losses(eg):
    a, b = _t(eg["pair_a"], True), _t(eg["pair_b"], True)
    for name, fn in (("ncc", lambda: P.ncc_loss(a, b)), ("mse", lambda: P.mse_loss(a, b)), ("L2", lambda: P.l2_loss(a))):
        a.grad = b.grad = None
        loss = fn()
        loss.backward()
        assert _close(loss, eg[f"{name}_loss"]) and _close(a.grad, eg[f"{name}_ga"]), name
        if name != "L2":
            assert _close(b.grad, eg[f"{name}_gb"]), name
    u = _t(eg["grad_u"], True)
    for norm in ("L2", "L1"):
        for k in (0, 1):
            u.grad = None
            loss = P.gradient_loss(u, norm, tuple(float(v) for v in eg[f"grad_spacing_{k}"]))
            loss.backward()
            assert _close(loss, eg[f"grad_{norm}_{k}_loss"]) and _close(u.grad, eg[f"grad_{norm}_{k}_g"])


def test_xent_family(eg):
    x = _t(eg["xent_x"], True)
    t = _t(eg["xent_t"]).long()
    soft = _t(eg["xent_soft"], True)
    for name, fn in xent_cases(5, t, soft, _t(eg["xent_w"]), _t(eg["xent_alpha"])).items():
        x.grad = soft.grad = None
        loss = fn(x)
        loss.backward()
        assert _close(loss, eg[f"{name}_loss"]) and _close(x.grad, eg[f"{name}_gx"], 2e-6), name
        if name.startswith("sce"):
            assert _close(soft.grad, eg[f"{name}_gt"]), name


@pytest.mark.parametrize("name", list(VARIANTS))
def test_unet_generator_variants(eg, name):
    kw, enc, dec, ncls = VARIANTS[name]
    sd = {k[len(f"var_{name}_sd/"):]: _t(v) for k, v in eg.items() if k.startswith(f"var_{name}_sd/")}
    y = P.unet_generator_forward(_t(eg["var_x"]), sd, 1, True, cfg=dict(encoders=enc, decoders=dec, act="LeakyReLU", **kw))
    assert _close(y, eg[f"var_{name}_y"], 2e-5)
