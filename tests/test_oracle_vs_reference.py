"""Pins the CPU port (oracle/ref_port.py) and the host-side mirror classes to the REAL reference modules,
imported from /root/reference where they lie (build container only; skipped on the GPU box, where the golden
fixtures made by oracle/make_golden.py take over).  fp32 on CPU; the port must be BIT-identical because it
restates the reference on top of the same ATen calls."""
import os

import pytest
import torch

from oracle import ref_import
from oracle import ref_port as P

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")


@pytest.fixture(scope="module")
def ref():
    torch.set_num_threads(4)
    return ref_import.load()


def _g(seed=230):
    return torch.Generator().manual_seed(seed)


@pytest.mark.parametrize("name,args", [("UNet_light", (1, 4)), ("UNet_light", (1, 32)), ("UNet", (1, 4)), ("voxel_morph_cvpr", ())])
def test_mirror_state_dict_matches_reference(ref, name, args):
    """Same keys, shapes and -- under the same seed -- the same xavier RNG stream as the reference classes
    (checkpoints are the on-disk contract: models/base.py:98-108 loads with strict=True)."""
    import deepatlas_b200 as da
    kw = dict(bias=True, BN=True) if args else {}
    torch.manual_seed(230)
    r = ref.get_network(name)(*args, **kw)
    r.weights_init()
    torch.manual_seed(230)
    m = da.get_network(name)(*args, **kw)
    m.weights_init()
    rs, ms = r.state_dict(), m.state_dict()
    assert list(rs.keys()) == list(ms.keys())
    for k in rs:
        assert rs[k].shape == ms[k].shape, k
        assert torch.equal(rs[k], ms[k]), k
    m.load_state_dict(rs, strict=True)
    r.load_state_dict(ms, strict=True)


@pytest.mark.parametrize("classes,bn", [(4, True), (3, False)])
def test_port_unet_light(ref, classes, bn):
    torch.manual_seed(230)
    net = ref.get_network("UNet_light")(1, classes, bias=True, BN=bn)
    net.weights_init()
    net.train()
    x = torch.rand((1, 1, 16, 24, 16), generator=_g())
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    stats = {}
    y = P.unet_generator_forward(x, sd, 1, bn, stats_out=stats)
    y_ref = net(x)
    assert torch.equal(y, y_ref)
    after = net.state_dict()
    for k, v in stats.items():   # running statistics move exactly as nn.BatchNorm3d moves them
        assert torch.equal(v, after[k]), k


def test_port_unet32(ref):
    torch.manual_seed(230)
    net = ref.get_network("UNet")(1, 4, bias=True, BN=True)
    net.weights_init()
    net.train()
    x = torch.rand((1, 1, 16, 16, 16), generator=_g())
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    assert torch.equal(P.unet_forward(x, sd, True), net(x))


def test_port_voxelmorph_and_identity(ref):
    torch.manual_seed(230)
    net = ref.get_network("voxel_morph_cvpr")()
    net.weights_init()
    g = _g()
    s, t = torch.rand((1, 1, 16, 24, 20), generator=g), torch.rand((1, 1, 16, 24, 20), generator=g)
    out = P.voxelmorph_forward(s, t, dict(net.state_dict()))
    out_ref = net(s, t)
    for a, b in zip(out, out_ref):
        assert torch.equal(a, b)
    assert torch.equal(P.identity_transform((5, 6, 7)), ref.utils.get_identity_transform((5, 6, 7)))
    # closed-form trilinear formula agrees with the ATen call the reference makes (fp64: 1e-12)
    phi = out_ref[2].double()
    cf = P.warp_closed_form(s.double(), phi)
    assert float((cf - P.warp(s.double(), phi)).abs().max()) < 1e-12


@pytest.mark.parametrize("wt", ["Uniform", "Simple", "Volume"])
@pytest.mark.parametrize("softmax", [True, False])
@pytest.mark.parametrize("no_bg", [True, False])
def test_port_dice(ref, wt, softmax, no_bg):
    g = _g()
    C = 5
    logits = torch.randn((2, C, 6, 7, 8), generator=g)
    labels = torch.randint(0, C, (2, 6, 7, 8), generator=g)
    soft = torch.softmax(torch.randn((2, C, 6, 7, 8), generator=g), 1)
    x = logits if softmax else torch.softmax(logits, 1)
    crit = ref.get_loss_function("dice")(n_class=C, weight_type=wt, no_bg=no_bg, softmax=softmax, eps=1e-6)
    for tgt in (labels, soft):
        assert torch.equal(P.dice_multiclass(x, tgt, C, wt, no_bg, softmax, 1e-6), crit(x, tgt))
    assert torch.equal(P.mask_to_one_hot(labels.reshape(2, 1, -1), C), ref.transforms.mask_to_one_hot(labels.reshape(2, 1, -1), C))
    with pytest.raises(ValueError):
        crit(x, labels[:, None, None])


def test_port_lncc_bending(ref):
    g = _g()
    I, J = torch.rand((2, 1, 12, 13, 14), generator=g), torch.rand((2, 1, 12, 13, 14), generator=g)
    assert torch.equal(P.lncc(I, J), ref.get_loss_function("lncc")()(I, J))
    u = torch.randn((2, 3, 8, 9, 10), generator=g) * 0.1
    for sp in ((1, 1, 1), (1.0, 1.5, 2.0)):
        assert torch.equal(P.bending_energy(u, sp), ref.get_loss_function("bendingEnergy")(spacing=sp)(u))
        assert torch.equal(P.bending_energy(u, sp, norm="L1"), ref.get_loss_function("bendingEnergy")(norm="L1", spacing=sp)(u))
    # analytic identities of SURVEY.md section 4
    assert float(P.lncc(I, I)) < 1e-5
    idt = P.identity_transform((8, 9, 10))[None]
    assert float(P.bending_energy(idt * 0.3 + 0.1)) < 1e-10


def test_registry_seams(ref):
    """install() overwrites the reference's own registries and nothing else; bad names keep raising KeyError."""
    import deepatlas_b200 as da
    keep_n, keep_l = dict(ref.network_dic), dict(ref.loss_dict)
    try:
        replaced = da.install(ref.network_factory, ref.loss)
        assert {"network:UNet_light", "network:UNet", "network:voxel_morph_cvpr", "loss:dice", "loss:lncc", "loss:bendingEnergy"} <= set(replaced)
        assert ref.get_network("UNet_light") is da.UNet_light
        assert ref.get_loss_function("dice") is da.DiceLossMultiClass
        assert set(ref.loss_dict) == set(keep_l)      # the un-replaced reference losses stay registered
        with pytest.raises(KeyError):
            ref.get_network("nope")
    finally:
        ref.network_dic.clear(); ref.network_dic.update(keep_n)
        ref.loss_dict.clear(); ref.loss_dict.update(keep_l)


# ---- the remaining registry losses and the UNet_generator variants ------------------
def test_port_remaining_losses(ref):
    g = _g()
    a, b = torch.rand((2, 1, 7, 8, 9), generator=g), torch.rand((2, 1, 7, 8, 9), generator=g)
    assert torch.equal(P.ncc_loss(a, b), ref.get_loss_function("ncc")()(a, b))
    assert torch.equal(P.mse_loss(a, b), ref.get_loss_function("mse")()(a, b))
    assert torch.equal(P.mse_loss(a, b), ref.loss.MSELoss()(a, b))
    assert torch.equal(P.l2_loss(a), ref.get_loss_function("L2")()(a))
    u = torch.randn((2, 3, 8, 9, 10), generator=g) * 0.1
    for norm in ("L2", "L1"):
        for sp in ((1, 1, 1), (1.0, 1.5, 2.0)):
            assert torch.equal(P.gradient_loss(u, norm, sp), ref.get_loss_function("gradient")(norm=norm, spacing=sp)(u))
    C = 5
    x = torch.randn((2, C, 6, 7, 8), generator=g)
    t = torch.randint(0, C, (2, 6, 7, 8), generator=g)
    soft = torch.softmax(torch.randn((2, C, 6, 7, 8), generator=g), 1)
    w = torch.rand(C, generator=g) + 0.5
    assert torch.equal(P.cross_entropy(x, t), ref.get_loss_function("cross_entropy")()(x, t))
    assert torch.equal(P.cross_entropy(x, t, weight=w), ref.get_loss_function("cross_entropy")(weight=w)(x, t))
    for soft_max in (True, False):
        for size_average in (True, False):
            xin = x if soft_max else torch.softmax(x, 1)
            crit = ref.get_loss_function("focal")(C, gamma=2, size_average=size_average, soft_max=soft_max)
            assert torch.equal(P.focal_loss(xin, t, None, 2, size_average, soft_max), crit(xin, t))
    alpha = (torch.rand(C, 1, generator=g) + 0.5)
    assert torch.equal(P.focal_loss(x, t, alpha, 1.5), ref.get_loss_function("focal")(C, alpha=alpha, gamma=1.5)(x, t))
    x2, t2 = torch.randn((37, C), generator=_g(5)), torch.randint(0, C, (37,), generator=_g(6))   # (observations, classes)
    assert torch.equal(P.focal_loss(x2, t2), ref.get_loss_function("focal")(C)(x2, t2))
    assert torch.equal(P.soft_cross_entropy(x, soft, True), ref.get_loss_function("soft_cross_entropy")(softmax=True)(x, soft))
    p = torch.softmax(x, 1)
    p[0, 0, 0, 0, :3] = 0.0     # exercises the 1e-8 clamp
    assert torch.equal(P.soft_cross_entropy(p, soft, False),
                       ref.get_loss_function("soft_cross_entropy")(softmax=False)(p.clone(), soft))


def test_mirror_registry_is_complete(ref):
    """Every name of the reference's loss registry resolves in the mirror, with the reference's constructor
    argument names (lib/loss.py:739-750)."""
    import inspect

    import deepatlas_b200 as da
    assert set(da.loss_dict) == set(ref.loss_dict)
    for name in ("ncc", "gradient", "L2", "focal", "soft_cross_entropy", "dice", "lncc", "bendingEnergy"):
        rp = list(inspect.signature(ref.loss_dict[name].__init__).parameters)
        mp = list(inspect.signature(da.loss_dict[name].__init__).parameters)
        norm = lambda q: ["self"] if q == ["self", "args", "kwargs"] else q   # nn.Module's default __init__  # noqa: E731
        assert norm(rp) == norm(mp), name
    with pytest.raises(KeyError):
        da.get_loss_function("nope")


VARIANTS = [dict(maxpool=False), dict(upsample=True), dict(res=True), dict(maxpool=False, upsample=True, res=True)]


def _variant_cfg(kw):
    if kw.get("res"):      # channel counts the reference's `+` accepts: equal, or a single input channel
        enc, dec = [(8, 8), (8, 8)], [(8, 8)]
    elif kw.get("upsample"):
        enc, dec = [(4, 8), (8, 8, 16)], [(16, 8, 8)]
    else:
        enc, dec = [(4, 8), (8, 8, 16)], [(8, 8, 8)]
    return enc, dec


@pytest.mark.parametrize("kw", VARIANTS)
def test_port_and_mirror_unet_generator_variants(ref, kw):
    import deepatlas_b200 as da
    from deepatlas_b200 import networks as M
    enc, dec = _variant_cfg(kw)
    n_classes = 8 if kw.get("res") else 3
    torch.manual_seed(230)
    net = ref.network_factory.unets.UNet_generator(enc, dec, act="LeakyReLU", **kw)(1, n_classes, bias=True, BN=True)
    net.weights_init()
    net.train()
    torch.manual_seed(230)
    mir = M.UNet_generator(enc, dec, act="LeakyReLU", **kw)(1, n_classes, bias=True, BN=True)
    mir.weights_init()
    rs, ms = net.state_dict(), mir.state_dict()
    assert list(rs.keys()) == list(ms.keys())
    for k in rs:
        assert torch.equal(rs[k], ms[k]), k
    x = torch.rand((1, 1, 8, 12, 8), generator=_g())
    sd = {k: v.clone() for k, v in rs.items()}
    cfg = dict(encoders=enc, decoders=dec, act="LeakyReLU", **kw)
    assert torch.equal(P.unet_generator_forward(x, sd, 1, True, cfg=cfg), net(x))
    assert da is not None


def test_port_upsample_closed_form():
    x = torch.rand((1, 2, 3, 4, 5), generator=_g(), dtype=torch.float64)
    ref_out = torch.nn.Upsample(scale_factor=2, mode="trilinear")(x)
    assert float((P.upsample_trilinear2_closed_form(x) - ref_out).abs().max()) < 1e-14


def test_port_dice_on_label(ref):
    g = _g()
    a = torch.randint(0, 6, (2, 1, 6, 7, 8), generator=g)
    b = torch.randint(0, 6, (2, 1, 6, 7, 8), generator=g)
    b[0][b[0] == 3] = 0          # an empty target class: Simple weighting hits its inf -> 1 rule
    for wt in ("Uniform", "Simple"):
        assert torch.equal(P.dice_on_label(a, b, None, 10e-6, wt), ref.loss.DiceLossOnLabel()(a, b, weight_type=wt))
        assert torch.equal(P.dice_on_label(a, b, 8, 10e-6, wt), ref.loss.DiceLossOnLabel(n_class=8)(a, b, weight_type=wt))


@pytest.mark.parametrize("size", [(12, 14, 16), (66, 68, 70)])
def test_port_lncc_multiscale(ref, size, monkeypatch):
    """The reference builds its filters with .cuda() (lib/loss.py:539): identity here, the arithmetic is device independent."""
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    g = _g()
    I, J = torch.rand((1, 1) + size, generator=g), torch.rand((1, 1) + size, generator=g)
    assert torch.equal(P.lncc_multiscale(I, J), ref.loss.LNCCLoss()(I, J))
    import deepatlas_b200 as da
    sched = da.LNCCLoss.schedule(list(size))
    crit = ref.loss.LNCCLoss()
    crit(I, J)
    assert [s[0] for s in sched] == crit.scale and [s[1] for s in sched] == crit.dilation
    assert [s[2] for s in sched] == [st[0] for st in crit.step] and [s[3] for s in sched] == crit.scale_weight


def test_checkpoint_round_trip_through_reference_code(tmp_path):
    """The on-disk contract: a '.pth.tar' written by the reference's own save_checkpoint from a
    REFERENCE model restores into the mirror through the reference's own initialize_model (strict=True), optimizer
    state included -- and the other way round (models/base.py:70-120)."""
    import deepatlas_b200 as da
    refm = ref_import.load(with_models=True)
    import models.base as base
    for name, args, kw in (("UNet_light", (1, 4), dict(bias=True, BN=True)), ("voxel_morph_cvpr", (), {})):
        torch.manual_seed(230)
        r = refm.get_network(name)(*args, **kw)
        r.weights_init()
        m = da.get_network(name)(*args, **kw)          # different (default) initialisation
        opt_r = torch.optim.Adam(r.parameters(), lr=1e-3)
        opt_m = torch.optim.Adam(m.parameters(), lr=1e-3)
        state = {"epoch": 3, "model_state_dict": r.state_dict(), "optimizer_state_dict": opt_r.state_dict(), "best_score": 0.5}
        base.BaseExperiment.save_checkpoint(state, True, str(tmp_path), prefix=name)
        epoch, best = base.BaseExperiment.initialize_model(m, opt_m, str(tmp_path / f"{name}_model_best.pth.tar"))
        assert (epoch, best) == (3, 0.5)
        for k, v in r.state_dict().items():
            assert torch.equal(v, m.state_dict()[k]), k
        # and back: a checkpoint of the mirror restores into the reference class
        with torch.no_grad():
            for p in m.parameters():
                p.add_(1.0)
        state = {"epoch": 4, "model_state_dict": m.state_dict(), "optimizer_state_dict": opt_m.state_dict(), "seg_best_score": torch.tensor(0.25)}
        base.BaseExperiment.save_checkpoint(state, False, str(tmp_path), prefix=name + "_mirror")
        epoch, best = base.BaseExperiment.initialize_model(r, opt_r, str(tmp_path / f"{name}_mirror_checkpoint.pth.tar"))
        assert (epoch, best) == (4, 0.25)
        for k, v in m.state_dict().items():
            assert torch.equal(v, r.state_dict()[k]), k


def test_reference_trainer_through_the_registry_seams(tmp_path, monkeypatch):
    """BASELINE config C1 (plumbing): the reference's own SegmentationExperiment.train() runs two epochs on synthetic 16^3
    volumes on the CPU with the harness shims of SURVEY.md Appendix A and writes its checkpoints; after install() the
    SAME trainer code constructs this package's network and loss through its registries (models/segmentation.py:81-88)
    and its own initialize_model() restores the checkpoint it has just written into them (strict=True)."""
    import deepatlas_b200 as da
    refm = ref_import.load(with_models=True)
    seg = refm.segmentation
    n_classes, size = 4, (16, 16, 16)

    class SyntheticDataset(torch.utils.data.Dataset):     # lib/datasets.py:68: [image (1,D,H,W) f32, seg (D,H,W) u8, name]
        def __init__(self, list_file, data_dir, with_seg=True, preload=False, pre_transform=None, n_samples=2, **kw):
            g = torch.Generator().manual_seed(230)
            self.items = [(torch.rand((1,) + size, generator=g), torch.randint(0, n_classes, size, generator=g, dtype=torch.uint8))
                          for _ in range(2)]

        def __len__(self):
            return len(self.items)

        def __getitem__(self, i):
            return [self.items[i][0], self.items[i][1], f"case{i}"]

    monkeypatch.setattr(seg.med_data, "get_seg_dataset", lambda name: SyntheticDataset)
    monkeypatch.setattr(seg.vis, "make_segmentation_image_summary", lambda *a, **k: torch.zeros(3, 8, 8))
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.nn.Module, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.cuda, "manual_seed", lambda *a, **k: None)

    def config():
        return dict(debug_mode=True, resume_dir="", random_seed=230, data="MindBoggle", n_epochs=2, samples_per_epoch=2, batch_size=1,
                    valid_batch_size=1, print_batch_period=50, valid_epoch_period=1, save_ckpts_epoch_period=1, model="UNet_light",
                    model_settings={"in_channel": 1, "n_classes": n_classes, "bias": True, "BN": True}, n_classes=n_classes,
                    class_name={k: str(k) for k in range(1, n_classes)}, crop_size=None, loss="dice",
                    loss_settings={"n_class": n_classes, "weight_type": "Uniform", "no_bg": False, "softmax": True, "eps": 1e-6},
                    learning_rate=1e-3, lr_mode="multiStep", milestones=[0.5, 1], gamma=0.2, num_samples=1, preload=True,
                    data_dir=str(tmp_path / "data"), valid_data_dir=str(tmp_path / "data"), training_list_file="train.txt",
                    validation_list_file="valid.txt", testing_list_file="test.txt", log_dir=str(tmp_path / "logs"))

    exp = seg.SegmentationExperiment(config())
    exp.train()
    ckpt = os.path.join(exp.ckpoint_dir, "checkpoint.pth.tar")
    assert os.path.isfile(ckpt) and os.path.isfile(os.path.join(exp.ckpoint_dir, "train_config.json"))
    assert type(exp.model).__module__.startswith("lib.network_factory")

    keep_n, keep_l = dict(refm.network_dic), dict(refm.loss_dict)
    try:
        da.install(refm.network_factory, refm.loss)
        exp2 = seg.SegmentationExperiment(config())
        exp2.setup_model()
        exp2.setup_loss()
        exp2.setup_optimizer()
        assert isinstance(exp2.model, da.UNet_light) and isinstance(exp2.criterion, da.DiceLossMultiClass)
        epoch, best = exp2.initialize_model(exp2.model, exp2.optimizer, ckpt)      # the reference's own loader, strict=True
        assert epoch == 2
        for k, v in exp.model.state_dict().items():
            assert torch.equal(v, exp2.model.state_dict()[k]), k
        with pytest.raises(RuntimeError):                                          # and there is no CPU path behind it
            exp2.model(torch.rand((1, 1) + size))
    finally:
        refm.network_dic.clear(); refm.network_dic.update(keep_n)
        refm.loss_dict.clear(); refm.loss_dict.update(keep_l)
