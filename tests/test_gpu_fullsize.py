"""The benchmark extent itself: 160 x 192 x 160 (BASELINE.json config C4), where a CPU oracle run takes minutes.

The CUDA path is checked here through size-independent properties of the operators -- bilinearity / adjointness of the
convolutions (forward, data gradient and weight gradient of the tcgen05 and mma.sync kernels must agree with each other
through Euler's identity), scaling invariance, the analytic identities of SURVEY.md section 4 (warp by the identity,
Dice of a one-hot prediction, LNCC of affinely related images, bending energy of an affine field), exact counts -- on the
very kernels, grids and tile counts the benchmark step launches.  Parity against the oracle proper is established at the
smaller extents of the other -m gpu tests (and at 80 x 96 x 80 inside bench.py)."""
import pytest
import torch

from parity_util import rel_err

pytestmark = pytest.mark.gpu
FULL = (160, 192, 160)
HALF = (80, 96, 80)


def _g(seed=230):
    return torch.Generator().manual_seed(seed)


def _rand(shape, cuda, scale=1.0, seed=230):
    g = torch.Generator(device=cuda).manual_seed(seed)
    return (torch.rand(shape, device=cuda, generator=g) - 0.5) * (2.0 * scale)


def _dot(a, b):
    return float((a.double() * b.double()).sum())


@pytest.mark.parametrize("C1,C2,Cout", [(32, 16, 16), (16, 0, 16), (8, 16, 3), (1, 0, 8)])
def test_conv3d_full_extent_bilinearity(cuda, C1, C2, Cout):
    """L = <conv(cat(x1, x2); w) + b, y> is bilinear in (x, w): <dL/dx, x> = <dL/dw, w> = L - <b, sum y>, dL/db = sum y,
    and conv(2x) = 2 conv(x) - b.  Forward, data gradients and weight / bias gradient of the layer shapes of the
    benchmark step (tcgen05 kernels; the input layer's FFMA kernels) must agree with each other to fp32 round-off."""
    from deepatlas_b200 import ops
    x1 = _rand((1, C1) + FULL, cuda, seed=1).requires_grad_(True)
    x2 = _rand((1, C2) + FULL, cuda, seed=2).requires_grad_(True) if C2 else None
    w = (_rand((Cout, C1 + C2, 3, 3, 3), cuda, 0.1, seed=3)).requires_grad_(True)
    b = _rand((Cout,), cuda, 0.1, seed=4).requires_grad_(True)
    y = _rand((1, Cout) + FULL, cuda, seed=5)
    out = ops.conv3d(x1, w, b, x2=x2, stride=1, pad=1)
    L = _dot(out, y)
    (out * y).sum().backward()
    ysum = y.double().sum(dim=(0, 2, 3, 4))
    lin = L - float((b.detach().double() * ysum).sum())
    scale = float(out.detach().abs().double().mean()) * float(y.abs().double().mean()) * y.numel()   # size of the terms summed
    ex = _dot(x1.grad, x1.detach()) + (_dot(x2.grad, x2.detach()) if C2 else 0.0)
    assert abs(ex - lin) <= 1e-5 * scale, ("dgrad", ex, lin)
    assert abs(_dot(w.grad, w.detach()) - lin) <= 1e-5 * scale, ("wgrad", _dot(w.grad, w.detach()), lin)
    assert rel_err(b.grad, ysum) < 1e-5
    with torch.no_grad():
        out2 = ops.conv3d(2.0 * x1, w, b, x2=(2.0 * x2 if C2 else None), stride=1, pad=1)
        assert rel_err(out2, 2.0 * out - b.view(1, -1, 1, 1, 1)) < 1e-6


def test_deconv_k2s2_full_extent_bilinearity(cuda):
    """The same identities for the up-sampler that produces the full-resolution 32-channel tensor (mma.sync kernels)."""
    from deepatlas_b200 import ops
    x = _rand((1, 32) + HALF, cuda, seed=1).requires_grad_(True)
    w = _rand((32, 32, 2, 2, 2), cuda, 0.2, seed=2).requires_grad_(True)
    b = _rand((32,), cuda, 0.1, seed=3).requires_grad_(True)
    y = _rand((1, 32) + FULL, cuda, seed=4)
    out = ops.deconv_k2s2(x, w, b)
    assert tuple(out.shape) == (1, 32) + FULL
    L = _dot(out, y)
    (out * y).sum().backward()
    ysum = y.double().sum(dim=(0, 2, 3, 4))
    lin = L - float((b.detach().double() * ysum).sum())
    scale = float(out.detach().abs().double().mean()) * float(y.abs().double().mean()) * y.numel()
    assert abs(_dot(x.grad, x.detach()) - lin) <= 1e-5 * scale
    assert abs(_dot(w.grad, w.detach()) - lin) <= 1e-5 * scale
    assert rel_err(b.grad, ysum) < 1e-5


def test_losses_full_extent_identities(cuda):
    """SURVEY.md section 4 identities at the benchmark extent: warp by the identity, Dice of a one-hot prediction (32
    classes, uint8 labels), exact evaluation counts, LNCC of affinely related images, bending energy of an affine field."""
    import deepatlas_b200 as da
    from deepatlas_b200 import evaluation, ops
    C = 32
    V = FULL[0] * FULL[1] * FULL[2]
    img = _rand((1, 1) + FULL, cuda, seed=1) + 0.5
    warped, _ = ops.warp3d(img, torch.zeros((1, 3) + FULL, device=cuda), add_identity=True, want_phi=True)
    # (fp32 round-off of the normalised sampling positions at extents of 160-192, times the slope of a white-noise image)
    assert rel_err(warped, img) < 1e-4
    lab = torch.randint(0, C, (1,) + FULL, generator=_g(2), dtype=torch.uint8).to(cuda)
    onehot = torch.zeros((1, C) + FULL, device=cuda).scatter_(1, lab[:, None].long(), 1.0)
    dice = da.get_loss_function("dice")(n_class=C, weight_type="Uniform", softmax=False, eps=1e-6)
    assert abs(float(dice(onehot, lab))) < 1e-6
    counts, pred = evaluation.argmax_counts(onehot, lab)
    assert torch.equal(pred.reshape(lab.shape), lab)
    assert counts.sum(dim=2).tolist() == [[V, V, V]]      # |pred|, |truth| and their overlap: V each for a perfect prediction
    assert torch.equal(counts[0, 0], torch.bincount(lab.reshape(-1).long(), minlength=C))
    lncc = da.get_loss_function("lncc")().to(cuda)
    assert abs(float(lncc(img, img))) < 1e-5 and abs(float(lncc(img, 2.0 * img + 0.5))) < 1e-4
    from oracle import ref_port as P
    idt = P.identity_transform(FULL)[None].to(cuda)
    assert float(da.get_loss_function("bendingEnergy")()(idt * 0.3 + 0.1)) < 1e-9


def test_joint_step_full_extent_is_finite_and_reproducible(cuda):
    """One whole joint step at 160 x 192 x 160 / 32 classes (the benchmark workload, bench.py's seed): finite loss and
    gradients, every parameter reached, and the same loss from the graphed, branch-overlapped step as from the plain
    eager one (the arithmetic is the same; only the anatomy term's atomics reorder fp32 sums)."""
    from deepatlas_b200 import ops
    from deepatlas_b200.dist import FlatGradBucket
    from deepatlas_b200.graph import GraphedStep
    from deepatlas_b200.joint import JointModel, make_synthetic_pair
    batch = make_synthetic_pair(FULL, 32, seed=230, device=cuda)
    losses = {}
    for mode in ("eager", "graph+overlap"):
        ov = mode != "eager"
        ops.set_wgrad_overlap(ov)
        torch.manual_seed(230)
        model = JointModel(n_classes=32, overlap_reg=ov, overlap_seg=ov).to(cuda)
        model.weights_init()
        bucket = FlatGradBucket(model.trainable_parameters())
        if ov:
            bucket.enable_alt()

        def compute(*b):
            bucket.zero()
            loss, _ = model.joint_loss(*b)
            loss.backward()
            model.join_streams()
            bucket.allreduce(1)
            return loss.detach()

        state = {k: v.clone() for k, v in model.state_dict().items()}
        run = GraphedStep(compute, batch, warmup=1) if ov else compute
        if ov:
            model.load_state_dict(state)
        loss = float(run(*batch))
        torch.cuda.synchronize()
        flat = bucket.flat
        assert torch.isfinite(flat).all() and float(flat.abs().max()) > 0
        off, untouched = 0, []
        for name, p in list(model.seg.named_parameters()) + list(model.reg.named_parameters()):
            n = p.numel()
            if float(flat[off:off + n].abs().max()) == 0.0:
                untouched.append(name)
            off += n
        assert not untouched, untouched
        losses[mode] = (loss, flat.clone())
        del model, bucket, run
        torch.cuda.empty_cache()
    ops.set_wgrad_overlap(False)
    assert abs(losses["eager"][0] - losses["graph+overlap"][0]) <= 1e-5 * abs(losses["eager"][0]), losses
    assert rel_err(losses["graph+overlap"][1], losses["eager"][1]) < 1e-4
