"""Every C-ABI call the autograd wrappers make, validated WITHOUT a GPU: `_lib.call` is replaced by a checker that
matches the argument list against the ctypes table (arity and per-argument conversion), so a wrong argument count or
type in deepatlas_b200/ops.py fails here instead of on the GPU box.  No kernel runs; outputs are uninitialised."""
import ctypes

import pytest
import torch


@pytest.fixture()
def dry(monkeypatch, built_lib):
    from deepatlas_b200 import _lib, ops
    seen = []

    def call(name, *args):
        codes, ret = _lib.SIGNATURES[name]
        assert ret == "rc", name
        assert len(args) == len(codes), f"{name}: {len(args)} arguments for signature '{codes}'"
        for i, (c, a) in enumerate(zip(codes, args)):
            if c in ("i", "l"):
                assert isinstance(a, int), f"{name} arg {i}: expected int, got {type(a).__name__}"
            elif c in ("f", "d"):
                assert isinstance(a, (int, float)), f"{name} arg {i}: expected float, got {type(a).__name__}"
            else:
                assert a is None or isinstance(a, (ctypes.c_void_p, int)), f"{name} arg {i}: expected pointer, got {type(a).__name__}"
            _lib._T[c].from_param(a)
        seen.append(name)

    monkeypatch.setattr(_lib, "call", call)
    from deepatlas_b200 import evaluation
    monkeypatch.setattr(ops, "_stream", lambda: ctypes.c_void_p(0))
    monkeypatch.setattr(ops, "_f32", lambda t, name: t.contiguous())
    monkeypatch.setattr(evaluation, "_stream", lambda: ctypes.c_void_p(0))
    monkeypatch.setattr(evaluation, "_f32", lambda t, name: t.contiguous())
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))
    return seen


def _r(*shape, grad=True):
    return torch.rand(shape).requires_grad_(grad)


def test_new_ops_argument_lists(dry):
    import deepatlas_b200 as da
    from deepatlas_b200 import ops
    L = da.get_loss_function
    a, b = _r(2, 1, 5, 6, 7), _r(2, 1, 5, 6, 7)
    for loss in (L("ncc")()(a, b), L("mse")()(a, b), L("L2")()(a), L("gradient")()(_r(1, 3, 5, 6, 7)),
                 L("gradient")(norm="L1")(_r(1, 3, 5, 6, 7))):
        loss.backward()
    x = _r(2, 4, 3, 4, 5)
    t = torch.randint(0, 4, (2, 3, 4, 5))
    soft = _r(2, 4, 3, 4, 5)
    for loss in (L("cross_entropy")()(x, t), L("cross_entropy")(weight=torch.ones(4))(x, t.to(torch.uint8)),
                 L("focal")(4)(x, t), L("soft_cross_entropy")(softmax=True)(x, soft), L("soft_cross_entropy")()(x, soft)):
        loss.backward()
    da.LNCCLoss()(_r(1, 1, 12, 14, 16), _r(1, 1, 12, 14, 16)).backward()
    assert "da_lncc_ms_fwd" in dry and "da_lncc_ms_bwd" in dry
    y = ops.upsample_trilinear2(_r(1, 2, 3, 4, 5))
    y.sum().backward()
    w, bias = _r(6, 4, 2, 2, 2), _r(6)
    ops.conv_k2s2(_r(1, 4, 4, 6, 8), w, bias).sum().backward()
    ops.add(_r(1, 4, 3, 3, 3), _r(1, 1, 3, 3, 3)).sum().backward()
    ops.add(_r(1, 4, 3, 3, 3), _r(1, 4, 3, 3, 3)).sum().backward()
    ops.crop_clip(torch.rand(1, 8, 8, 8), (1, 1, 1), (4, 4, 4))
    ops.crop_labels(torch.zeros((8, 8, 8), dtype=torch.uint8), (1, 1, 1), (4, 4, 4))
    from deepatlas_b200 import evaluation
    evaluation.label_overlap_counts(torch.zeros((1, 1, 4, 4, 4), dtype=torch.uint8), torch.zeros((1, 1, 4, 4, 4)))
    evaluation.argmax_counts(torch.rand(1, 3, 4, 4, 4), torch.zeros((1, 4, 4, 4), dtype=torch.int64))
    assert "da_label_overlap_counts" in dry and "da_argmax_counts" in dry
    for name in ("da_pair_moments_fwd", "da_affine2", "da_gradient_loss_fwd", "da_gradient_loss_bwd", "da_xent_fwd", "da_xent_bwd",
                 "da_upsample_trilinear2_fwd", "da_upsample_trilinear2_bwd", "da_deconv_k2s2_dgrad", "da_deconv_k2s2_fwd",
                 "da_deconv_k2s2_wgrad", "da_channel_sum", "da_add_bcast", "da_channel_reduce", "da_crop_clip_f32", "da_crop_u8"):
        assert name in dry, name


def test_variant_networks_argument_lists(dry):
    from deepatlas_b200 import networks as M
    for kw, enc, dec, ncls in ((dict(maxpool=False), [(4, 8), (8, 8, 16)], [(8, 8, 8)], 3),
                               (dict(upsample=True), [(4, 8), (8, 8, 16)], [(16, 8, 8)], 3),
                               (dict(maxpool=False, upsample=True, res=True), [(8, 8), (8, 8)], [(8, 8)], 8)):
        net = M.UNet_generator(enc, dec, act="LeakyReLU", **kw)(1, ncls, bias=True, BN=True)
        y = net(torch.rand(1, 1, 8, 12, 8))
        assert tuple(y.shape) == (1, ncls, 8, 12, 8)
        y.sum().backward()
    blk = M.deconvBlockVM(4, 4, 3, stride=1, padding=1, bias=True, residual=True)
    blk(torch.rand(1, 4, 4, 4, 4).requires_grad_(True)).sum().backward()


def test_hot_path_argument_lists(dry):
    """The joint step's own wrappers through the same checker."""
    from deepatlas_b200.joint import JointModel, make_synthetic_pair
    model = JointModel(n_classes=4)
    model.weights_init()
    batch = make_synthetic_pair((16, 16, 16), 4, seed=230, device="cpu")
    loss, _ = model.joint_loss(*batch)
    loss.backward()
    assert "da_conv3d_fwd_ex" in dry and "da_conv3d_wgrad_ex" in dry and "da_warp_dice_sums_bwd" in dry and "da_lncc_bwd" in dry
