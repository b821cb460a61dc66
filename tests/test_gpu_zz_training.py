"""A short TRAINING run on the GPU against the same run restated on the CPU oracle: the loop of
models/segmentation.py:139-160 (zero_grad, forward, Dice on uint8 labels, backward, Adam step) for three steps.
Judged on the loss trajectory: Adam's first updates are lr * sign(g), so parameters whose gradient is analytically zero
(a conv bias in front of a BatchNorm) move by +-lr at random in BOTH implementations without touching any output."""
import pytest
import torch

from parity_util import cpu_state, rel_err

pytestmark = pytest.mark.gpu


def test_three_adam_steps_follow_the_oracle(cuda):
    import deepatlas_b200 as da
    from oracle import ref_port as P
    n_classes, size, steps = 4, (16, 24, 16), 3
    torch.manual_seed(230)
    net = da.get_network("UNet_light")(1, n_classes, bias=True, BN=True).to(cuda)
    net.weights_init()
    net.train()
    g = torch.Generator().manual_seed(230)
    x = torch.rand((1, 1) + size, generator=g)
    lab = torch.randint(0, n_classes, (1,) + size, generator=g, dtype=torch.uint8)
    # the oracle's copy: float64 leaves (the truth rung), same Adam
    sd = {k: (v.double().requires_grad_(True) if v.is_floating_point() and "running" not in k else (v.double() if v.is_floating_point() else v.clone()))
          for k, v in cpu_state(net).items()}
    leaves = [v for v in sd.values() if v.is_floating_point() and v.requires_grad]
    opt_ref = torch.optim.Adam(leaves, lr=1e-3)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    crit = da.get_loss_function("dice")(n_class=n_classes, weight_type="Uniform", no_bg=False, softmax=True, eps=1e-6)
    xg, lg = x.to(cuda), lab.to(cuda)
    ours, ref = [], []
    for _ in range(steps):
        opt.zero_grad()
        loss = crit(net(xg), lg)
        loss.backward()
        opt.step()
        ours.append(float(loss.detach()))
        opt_ref.zero_grad()
        stats = {}
        y = P.unet_generator_forward(x.double(), sd, 1, True, stats_out=stats)
        l_ref = P.dice_multiclass(y, lab.long(), n_classes, "Uniform", False, True, 1e-6)
        l_ref.backward()
        opt_ref.step()
        for k, v in stats.items():      # running statistics move as nn.BatchNorm3d moves them
            sd[k] = v
        ref.append(float(l_ref.detach()))
    assert ours[-1] < ours[0], ours                       # it trains
    for a, b in zip(ours, ref):
        assert abs(a - b) <= 1e-3 * abs(b), (ours, ref)
    # the running statistics after three steps (momentum 0.1, unbiased variance) agree as well
    after = net.state_dict()
    for k in ("encoders.0.0.BN.running_mean", "encoders.0.0.BN.running_var", "decoders.decBlock2.1.BN.running_var"):
        assert rel_err(after[k], sd[k]) < 1e-2, k
    assert int(after["encoders.0.0.BN.num_batches_tracked"]) == steps


def test_bucket_inplace_gradients_equal_autograd_accumulation(cuda):
    """FlatGradBucket(inplace=True): the backward kernels add parameter gradients straight into the bucket views
    (accumulate = 1 inside their fixed-order reduces, autograd gets None).  Same gradients as the plain path, in which
    autograd adds returned tensors -- on the joint step, where the segmentation net's parameters are used twice."""
    from deepatlas_b200.dist import FlatGradBucket
    from deepatlas_b200.joint import JointModel, make_synthetic_pair
    n_classes, size = 4, (16, 24, 16)
    batch = make_synthetic_pair(size, n_classes, seed=231, device=cuda)
    flats, losses = [], []
    for inplace in (True, False):
        torch.manual_seed(230)
        model = JointModel(n_classes=n_classes).to(cuda)
        model.weights_init()
        bucket = FlatGradBucket(model.trainable_parameters(), inplace=inplace)
        assert all(getattr(p.grad, "_da_inplace", False) == inplace for p in bucket.params)
        for _ in range(2):   # the second round starts from zero() like a training loop
            bucket.zero()
            loss, _ = model.joint_loss(*batch)
            loss.backward()
        assert bucket.rebind(copy=True) == 0          # every gradient still lives in its bucket slot
        flats.append(bucket.flat.clone())
        losses.append(float(loss))
    assert losses[0] == losses[1]
    assert float(flats[0].abs().max()) > 0
    assert rel_err(flats[0], flats[1]) < 1e-5


def test_graphed_step_equals_eager(cuda):
    """deepatlas_b200.graph.GraphedStep: the whole joint training step (zero grads, forward, backward, Adam) captured as
    one CUDA graph and replayed on new inputs follows the eager loop: same losses, same weights after three steps --
    with the optimizer inside the graph, and with the update as an eager tail (the multi-process layout, where the NCCL
    all-reduce stays outside the graph)."""
    from deepatlas_b200.dist import FlatGradBucket
    from deepatlas_b200.graph import GraphedStep
    from deepatlas_b200.joint import JointModel, make_synthetic_pair
    n_classes, size = 4, (16, 24, 16)
    batches = [make_synthetic_pair(size, n_classes, seed=300 + i, device=cuda) for i in range(3)]
    results = {}
    for mode in ("eager", "graph", "graph+tail"):
        torch.manual_seed(230)
        model = JointModel(n_classes=n_classes).to(cuda)
        model.weights_init()
        bucket = FlatGradBucket(model.trainable_parameters())
        opt = torch.optim.Adam(bucket.params, lr=1e-3, fused=True, capturable=True)

        def compute(*batch):
            bucket.zero()
            loss, _ = model.joint_loss(*batch)
            loss.backward()
            return loss.detach()

        def update():
            bucket.allreduce(1)
            opt.step()

        def step(*batch):
            loss = compute(*batch)
            update()
            return loss

        if mode == "eager":
            run = step
        else:
            # capture on a copy of the state: the warm-up steps inside GraphedStep move parameters and Adam moments
            state = {k: v.clone() for k, v in model.state_dict().items()}
            run = GraphedStep(step, batches[0], warmup=2) if mode == "graph" else GraphedStep(compute, batches[0], warmup=2, eager_tail=update)
            model.load_state_dict(state)
            for grp in opt.param_groups:
                for p in grp["params"]:
                    st = opt.state[p]
                    st["step"].zero_()
                    st["exp_avg"].zero_()
                    st["exp_avg_sq"].zero_()
            assert run.launches_per_step > 100
        losses, grads = [], None
        for b in batches:
            losses.append(float(run(*b)))
            if grads is None:
                grads = bucket.flat.clone()   # the first step's gradients (Adam reads the bucket, it does not clear it)
        results[mode] = (losses, grads)
    l_e, g_e = results["eager"]
    for mode in ("graph", "graph+tail"):
        l_g, g_g = results[mode]
        for a, b in zip(l_e, l_g):
            assert abs(a - b) <= 1e-5 * abs(a), (mode, l_e, l_g)
        assert rel_err(g_g, g_e) < 1e-4, mode
        # (parameters are not compared: Adam's first updates are lr * sign(g), so every element whose gradient is round-off
        # -- the anatomy term's scatter kernels add with atomics -- moves by +-lr at random in any two runs)


def test_overlapped_branches_equal_serial(cuda):
    """JointModel(overlap_reg=True) runs the registration network on a side stream next to the segmentation passes
    (forward and, by autograd's stream rule, backward); overlap_seg=True also the target image's segmentation pass
    (BatchNorm buffer updates deferred, gradients into the bucket's alternate slots); ops.set_wgrad_overlap puts the
    convolutions' weight gradients on their own stream.  Same loss, gradients and BatchNorm running statistics as the
    serial order, eagerly and inside a captured graph (where the branches become parallel paths)."""
    from deepatlas_b200 import ops
    from deepatlas_b200.dist import FlatGradBucket
    from deepatlas_b200.graph import GraphedStep
    from deepatlas_b200.joint import JointModel, make_synthetic_pair
    n_classes, size = 4, (16, 24, 16)
    batch = make_synthetic_pair(size, n_classes, seed=411, device=cuda)
    out = {}
    modes = ("serial", "reg", "reg+graph", "reg+wgrad", "reg+wgrad+graph", "reg+seg+wgrad", "reg+seg+wgrad+graph")
    for mode in modes:
        ops.set_wgrad_overlap("wgrad" in mode)   # weight gradients on their own side stream
        torch.manual_seed(230)
        model = JointModel(n_classes=n_classes, overlap_reg="reg" in mode, overlap_seg="seg" in mode).to(cuda)
        model.weights_init()
        bucket = FlatGradBucket(model.trainable_parameters())
        if "seg" in mode:
            bucket.enable_alt()

        def compute(*b):
            bucket.zero()
            loss, _ = model.joint_loss(*b)
            loss.backward()
            model.join_streams()
            bucket.allreduce(1)          # (single process: joins the weight-gradient stream, folds the alternate slots)
            return loss.detach()

        state = {k: v.clone() for k, v in model.state_dict().items()}
        if mode.endswith("graph"):
            run = GraphedStep(compute, batch, warmup=2)
            model.load_state_dict(state)     # (the warm-up runs moved the BatchNorm running statistics only)
        else:
            run = compute
        loss = float(run(*batch))
        torch.cuda.synchronize()
        bn = {k: v.clone() for k, v in model.state_dict().items() if "running" in k or "num_batches" in k}
        out[mode] = (loss, bucket.flat.clone(), bn)
    ops.set_wgrad_overlap(False)
    for mode in modes[1:]:
        assert abs(out[mode][0] - out["serial"][0]) <= 1e-6 * abs(out["serial"][0]), mode
        assert rel_err(out[mode][1], out["serial"][1]) < 1e-5, mode
        for k, v in out["serial"][2].items():
            if "num_batches" in k:
                assert int(out[mode][2][k]) == int(v) == 2, (mode, k)     # two passes of the network per step
            else:
                assert rel_err(out[mode][2][k], v) < 1e-5, (mode, k)
