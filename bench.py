#!/usr/bin/env python
"""bench.py -- volumes/sec (fwd+bwd) of the 160x192x160 joint seg+reg step (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W [--config c4]      our arm (CUDA path through the C ABI)
  python bench.py --impl reference --gpus N --steps K --warmup W   reference arm: the CPU restatement of the
        reference's own PyTorch path (oracle/ref_port.py; /root/reference cannot travel to the GPU box), the FULL
        workload per step, all host cores, rank 0 only.
  python bench.py --impl torch_cuda ...                            informational: the same restatement's ATen calls
        on the GPU (stock PyTorch eager: cuDNN / ATen kernels), cudnn.allow_tf32 off and on.

One "step" of the default config (c4) = one joint seg+reg training step on one synthetic volume pair per GPU: two
UNet_light(1,32) passes, VoxelMorph + warp, LNCC + bending + anatomy Dice + 2 supervised Dice, one backward, one
flat-bucket gradient all-reduce (N>1), fused Adam.  1 pair = 2 volumes.  Prints ONE JSON line on rank 0.
--config selects the other BASELINE.json configurations (parity-test cases; the driver benches the default):
  c2  seg-only 32-base UNet, 128^3, 4 classes (one volume per step)       c3  registration only, 160^3 pairs
  c5  joint step at 256^3, 4 classes
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algo_bytes: SURVEY.md 8(d) / BASELINE.md section 4, algorithmic bytes per step per GPU (ideal-fusion model)
WORKLOADS = {
    "c4": dict(name="joint_seg_reg_160x192x160_c32_fp32", kind="joint", size=(160, 192, 160), classes=32, volumes=2, algo_bytes=42.0e9,
               metric="volumes/sec (fwd+bwd) 160x192x160 joint seg+reg"),
    "c5": dict(name="joint_seg_reg_256x256x256_c4_fp32", kind="joint", size=(256, 256, 256), classes=4, volumes=2, algo_bytes=105.8e9,
               metric="volumes/sec (fwd+bwd) 256x256x256 joint seg+reg"),
    "c3": dict(name="reg_only_160x160x160_fp32", kind="reg", size=(160, 160, 160), classes=2, volumes=2, algo_bytes=5.18e9,
               metric="volumes/sec (fwd+bwd) 160x160x160 registration (VoxelMorph + warp + LNCC + bending)"),
    # BASELINE config #2 is the reduced-precision ("bf16") case: its k3 convolutions run the tensor-core kernels in the
    # single-pass mode (da_set_conv_split(2): fp16 operands scaled per tensor, fp32 accumulate and storage); --precision fp32
    # times the exact 3xFP16 path instead
    "c2": dict(name="seg_only_unet32_128x128x128_c4_f16x1", kind="seg", size=(128, 128, 128), classes=4, volumes=1, algo_bytes=20.14e9,
               metric="volumes/sec (fwd+bwd) 128x128x128 seg-only 32-base UNet", precision="f16x1"),
}
PARITY_SIZE = (80, 96, 80)          # extent of the in-bench parity check against the fp32 / fp64 oracle (CPU leg)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), float(d["bf16_tflops"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc, self.idx = None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------------
# The reference's own PyTorch path, restated in oracle/ref_port.py (pinned bit-identical to /root/reference by
# tests/test_oracle_vs_reference.py).  Only the CPU legs below and the informational torch_cuda arm execute it.
# ------------------------------------------------------------------------------------------------------------
def _oracle_state(wl, device):
    """Leaf tensors with the reference's state_dict names; shapes come from our mirror classes (identical state_dict,
    tested).  Returns (loss_fn(batch), leaves)."""
    import torch
    from oracle import ref_port as P
    import deepatlas_b200 as da
    torch.manual_seed(230)
    C = wl["classes"]

    def leaves_of(net):
        return {k: (v.detach().clone().to(device).requires_grad_(True) if v.is_floating_point() and "running" not in k
                    else v.detach().clone().to(device)) for k, v in net.state_dict().items()}

    if wl["kind"] == "joint":
        seg = da.get_network("UNet_light")(1, C, bias=True, BN=True); seg.weights_init()
        reg = da.get_network("voxel_morph_cvpr")(); reg.weights_init()
        seg_sd, reg_sd = leaves_of(seg), leaves_of(reg)
        sds = [seg_sd, reg_sd]
        fn = lambda b: P.joint_loss(seg_sd, reg_sd, b, C)  # noqa: E731
    elif wl["kind"] == "reg":
        reg = da.get_network("voxel_morph_cvpr")(); reg.weights_init()
        reg_sd = leaves_of(reg)
        sds = [reg_sd]

        def fn(b):
            disp, I_w, _ = P.voxelmorph_forward(b[0], b[2], reg_sd)
            return P.lncc(I_w, b[2]) + 1000.0 * P.bending_energy(disp)
    else:
        seg = da.get_network("UNet")(1, C, bias=True, BN=True); seg.weights_init()
        seg_sd = leaves_of(seg)
        sds = [seg_sd]
        fn = lambda b: P.dice_multiclass(P.unet_forward(b[0], seg_sd, True), b[1].long(), C, "Uniform", False, True, 1e-6)  # noqa: E731
    leaves = [v for sd in sds for v in sd.values() if v.is_floating_point() and v.requires_grad]
    return fn, leaves


def oracle_steps(wl, size, steps, warmup, threads=None, device="cpu"):
    """Training steps (forward, backward, Adam, loss read-back) of the restated reference path; seconds per step."""
    import torch
    from deepatlas_b200.joint import make_synthetic_pair
    if threads:
        torch.set_num_threads(threads)
    fn, leaves = _oracle_state(wl, device)
    opt = torch.optim.Adam(leaves, lr=1e-3)
    batch = make_synthetic_pair(size, wl["classes"], seed=230, device=device)
    times = []
    for i in range(warmup + steps):
        if device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss = fn(batch)
        loss.backward()
        opt.step()
        _ = loss.item()
        if device != "cpu":
            torch.cuda.synchronize()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return times


def run_reference(args, wl):
    """Reference arm: CPU, all host threads, the full workload per step (each step is a measured duration)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    # a full step takes 20-80 s depending on the host: the first one is the warm-up and sets how many timed steps fit
    t0 = time.perf_counter()
    warm = min(args.warmup, 1)
    if warm:
        oracle_steps(wl, wl["size"], 0, 1, threads)
    t_warm = time.perf_counter() - t0
    steps = max(1, min(args.steps, 3 if t_warm < 35 else (2 if t_warm < 70 else 1)))
    times = oracle_steps(wl, wl["size"], steps, 0, threads)
    t = sorted(times)[len(times) // 2]
    value = wl["volumes"] / t
    D, H, W = wl["size"]
    sample = f"full {D}x{H}x{W} step ({wl['name']}), median of {len(times)} measured steps after {warm} warm-up, {threads} torch threads"
    line = {"impl": "reference", "metric": wl["metric"], "value": value, "unit": "volumes/s",
            "n_gpus": args.gpus, "steps": len(times), "warmup": warm, "ms_per_step": t * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "pairs_per_gpu": 1, "classes": wl["classes"], "device": "cpu", "torch_threads": threads,
                       "steps_requested": args.steps, "note": "one CPU step takes ~20 s: at most 3 timed steps + 1 warm-up are run"},
            "cpu_baseline": {"value": value, "unit": "volumes/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def torch_cuda_baseline(wl, steps=3, warmup=2):
    """Stock PyTorch-CUDA eager of the same step (the restated reference path's ATen calls on the GPU): the
    "existing Blackwell kernels" (cuDNN / ATen) next to ours.  Informational; both TF32 settings."""
    import torch
    out = {}
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        key = "tf32_on" if tf32 else "tf32_off"
        try:
            times = oracle_steps(wl, wl["size"], steps, warmup, device="cuda")
            t = sorted(times)[len(times) // 2]
            out[key] = {"ms_per_step": t * 1e3, "value": wl["volumes"] / t, "unit": "volumes/s", "steps": len(times),
                        "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
        except Exception as e:  # noqa: BLE001  (out of memory on a busy box must not kill the bench line)
            out[key] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
        torch.cuda.empty_cache()
    torch.backends.cudnn.allow_tf32 = True
    out["what"] = ("oracle/ref_port.py (= the reference's forward, loss, autograd backward, Adam) on cuda:0 through stock ATen/cuDNN "
                   f"{torch.backends.cudnn.version()}, torch {torch.__version__}; wall clock with synchronize, median")
    return out


def run_torch_cuda(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    r = torch_cuda_baseline(wl, steps=max(1, min(args.steps, 5)), warmup=max(1, min(args.warmup, 2)))
    best = r.get("tf32_off", {})
    line = {"impl": "torch_cuda", "metric": wl["metric"], "value": best.get("value"), "unit": "volumes/s", "n_gpus": 1,
            "ms_per_step": best.get("ms_per_step"), "higher_is_better": True, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "classes": wl["classes"]}, "torch_cuda_baseline": r}
    print(json.dumps(line), flush=True)
    return 0


def parity_check(dev):
    """The CUDA path against the fp32 / fp64 oracle on one PARITY_SIZE pair (every conv level of that extent takes the
    same tcgen05 kernels as the benchmark volume): loss error and the worst parameter-gradient ratio on the precision
    ladder of tests/parity_util.py."""
    import torch
    from oracle import ref_port as P
    from deepatlas_b200.joint import JointModel, make_synthetic_pair
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from parity_util import MaskRecorder, MaskReplay, check_grads_vs_truth, oracle_joint_loss, rel_err
    C = 8
    torch.manual_seed(230)
    model = JointModel(n_classes=C).to(dev)
    model.weights_init()
    batch = make_synthetic_pair(PARITY_SIZE, C, seed=230, device=dev)
    with MaskRecorder() as rec:
        loss, _ = model.joint_loss(*batch)
    loss.backward()
    torch.cuda.synchronize()
    replay = MaskReplay(rec.masks)
    ref_loss, ref_grads = oracle_joint_loss(model, batch, P, replay=replay)
    true_loss, true_grads = oracle_joint_loss(model, batch, P, dtype=torch.float64, replay=replay)
    ours = {k: p.grad.detach().cpu() for k, p in model.named_parameters() if p.grad is not None}
    w = check_grads_vs_truth(ours, ref_grads, true_grads, 1e-3, strict=False)
    return {"extent": list(PARITY_SIZE), "classes": C, "loss_rel_err_vs_fp64": rel_err(loss, true_loss),
            "reference_fp32_loss_rel_err_vs_fp64": rel_err(ref_loss, true_loss), "activation_mask_flips": replay.flips,
            "worst_grad_ratio_to_bound": w["worst_ratio_to_bound"], "worst_grad_param": w["worst_ratio_param"],
            "worst_grad_rel_err_vs_fp64": w["worst_rel_err_vs_fp64"], "worst_grad_rel_err_param": w["worst_rel_err_param"],
            "reference_fp32_rel_err_same_param": w["reference_fp32_rel_err_same_param"],
            "within_bound": bool(w["worst_ratio_to_bound"] <= 1.0),
            "bound": "max(1e-3, 3 x the fp32 reference's own error vs fp64), per parameter tensor, max-norm"}


# ------------------------------------------------------------------------------------------------------------
def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    import deepatlas_b200 as da  # noqa: F401
    from deepatlas_b200 import _lib, ops
    from deepatlas_b200.dist import FlatGradBucket, broadcast_parameters, init_from_env
    from deepatlas_b200.joint import JointModel, RegOnlyModel, SegOnlyModel, make_synthetic_pair

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (our arm) needs a CUDA device; there is no CPU fallback")
    rank, local, world = init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.load()
    SIZE, CLASSES = wl["size"], wl["classes"]
    ops.set_wgrad_overlap(not (args.no_overlap or os.environ.get("DA_BENCH_NO_WGRAD_OVERLAP") == "1"))
    precision = args.precision or wl.get("precision", "fp32")
    if precision == "f16x1":
        _lib.call("da_set_conv_split", 2)

    torch.manual_seed(230)
    if wl["kind"] == "joint":
        ov = not (args.no_overlap or os.environ.get("DA_BENCH_NO_OVERLAP") == "1")
        model = JointModel(n_classes=CLASSES, overlap_reg=ov, overlap_seg=ov and os.environ.get("DA_BENCH_NO_SEG_OVERLAP") != "1").to(dev)
    elif wl["kind"] == "reg":
        model = RegOnlyModel().to(dev)
    else:
        model = SegOnlyModel(n_classes=CLASSES, seg_name="UNet").to(dev)
    model.weights_init()
    broadcast_parameters(model)
    bucket = FlatGradBucket(model.trainable_parameters())
    if getattr(model, "overlap_seg", False):
        bucket.enable_alt()
    use_graph = not args.no_graph
    opt = torch.optim.Adam(bucket.params, lr=1e-3, fused=True, capturable=use_graph)

    host = [t.pin_memory() for t in make_synthetic_pair(SIZE, CLASSES, seed=230 + rank)]
    if wl["kind"] == "seg":
        host = host[:2]
    dev_batch = [t.to(dev) for t in host]
    h2d = sum(t.numel() * t.element_size() for t in host)

    def compute(batch):
        bucket.zero()
        loss, _ = model.joint_loss(*batch)
        loss.backward()
        if hasattr(model, "join_streams"):
            model.join_streams()   # the registration branch's backward ran on its side stream
        ops.join_wgrad_stream()    # weight gradients run on theirs
        return loss.detach()

    def update():
        bucket.allreduce(world)
        opt.step()

    def step(batch):
        loss = compute(batch)
        update()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the whole step as one CUDA graph (deepatlas_b200/graph.py): eager, the host side of a step (about a thousand launches
    # through Python) is as long as its GPU side; --no-graph times the eager loop
    overlap_on = wl["kind"] == "joint" and not (args.no_overlap or os.environ.get("DA_BENCH_NO_OVERLAP") == "1")
    overlap_note = "; registration branch, the target image's segmentation pass and the convolution weight gradients on side streams (parallel graph paths)" if overlap_on else ""
    gstep, graph_note = None, "eager"
    run = step
    if use_graph:
        from deepatlas_b200.graph import GraphedStep
        try:
            if world == 1:
                gstep = GraphedStep(lambda *b: step(b), dev_batch, warmup=args.warmup)
                graph_note = "one CUDA graph per step (zero grads, forward, backward, Adam)" + overlap_note
            else:   # the NCCL all-reduce and the optimizer stay outside the graph (torch's NCCL watchdog, see graph.py)
                gstep = GraphedStep(lambda *b: compute(b), dev_batch, warmup=args.warmup, eager_tail=update)
                graph_note = "one CUDA graph per step (zero grads, forward, backward) + eager NCCL all-reduce and Adam" + overlap_note
            run = lambda batch: gstep(*batch)   # noqa: E731
        except Exception as e:   # noqa: BLE001  (keep the bench alive: the eager step is the same arithmetic)
            torch.cuda.synchronize()
            graph_note = f"eager (graph capture failed: {type(e).__name__}: {str(e)[:120]})"
    for _ in range(args.warmup):
        run(dev_batch)
    # ---- device-resident timing -------------------------------------------------------------------------
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    l0 = _lib.size("da_launch_count")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = run(dev_batch)
    e1.record()
    barrier()
    clocks = sampler.stop()
    launches = _lib.size("da_launch_count") - l0   # (launch calls made by the host: none during graph replays)
    if gstep is not None:
        launches = gstep.launches_per_step * args.steps   # library kernels inside the replayed graphs
    ms = e0.elapsed_time(e1) / args.steps
    # ---- end to end: host buffers in, loss scalar out, every step ---------------------------------------
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # Through the package's own input stage (deepatlas_b200/input_stage.py, what a training loop uses): every step's
    # host buffers go through pinned staging and an async copy on a side stream, one step ahead of the arithmetic
    # (the first step's copy cannot overlap anything and is inside the timed region too); the loss is read back to the
    # host every step (the reference's loop logs it every step, models/segmentation.py:157) ...
    from deepatlas_b200.input_stage import DeviceInputStage
    samples = [(host[i], host[i + 1]) for i in range(0, len(host), 2)]
    stage = DeviceInputStage(dev, depth=2 * len(samples))

    def submit_all():
        for img, seg in samples:
            stage.submit(img, seg)

    # ... through a pinned buffer, one step behind: step k's loss is copied to the host right after the step on the same
    # stream and read while step k+1 runs (a blocking .item() per step leaves the GPU idle for the launch latency of the
    # next 1000-node graph); every step's loss is read inside the timed region, the last one after the loop.
    loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ready = [torch.cuda.Event() for _ in range(2)]

    def e2e_steps(nsteps):
        last_ = None
        submit_all()
        for k in range(nsteps):
            batch = [t for _ in samples for t in stage.get()]
            loss_k = run(batch)           # asynchronous: one graph launch (or the eager launches)
            loss_host[k % 2].copy_(loss_k, non_blocking=True)
            loss_ready[k % 2].record()
            if k + 1 < nsteps:
                submit_all()              # the next step's staging + copies run while this step computes
            if k > 0:
                loss_ready[(k - 1) % 2].synchronize()
                last_ = float(loss_host[(k - 1) % 2])
        loss_ready[(nsteps - 1) % 2].synchronize()
        return float(loss_host[(nsteps - 1) % 2])

    # untimed warm-up of this path too (first use of the input stage: lazy loading of its crop kernel, its side stream and
    # device buffers; on a fresh box this stalled one timed step by 130 ms), then EXACTLY args.steps timed steps
    e2e_steps(min(args.warmup, 2))
    barrier()
    f0.record()
    last = e2e_steps(args.steps)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1) / args.steps
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    peak_mem = torch.cuda.max_memory_allocated() / 2 ** 30

    line = None
    if rank == 0:
        hbm_peak, mma_peak, peak_src = _peaks()
        del dev_batch
        model.zero_grad(set_to_none=True)
        torch.cuda.empty_cache()
        # ---- dominant kernels, timed live on the launching stream: decBlock2.0 of UNet_light, cat(32,16) -> 16 at the
        # benchmark extent (the largest layer of the step: 204 GFLOP forward), forward and weight gradient -------------
        import ctypes
        BS = WORKLOADS["c4"]["size"]
        V = BS[0] * BS[1] * BS[2]
        x1 = torch.rand((1, 32, *BS), device=dev)
        x2 = torch.rand((1, 16, *BS), device=dev)
        w = torch.randn((16, 48, 3, 3, 3), device=dev) * 0.03
        dy = torch.rand((1, 16, *BS), device=dev)
        y = torch.empty((1, 16, *BS), device=dev)
        gw, gb = torch.empty_like(w), torch.empty(16, device=dev)
        nbw = _lib.size("da_conv3d_wgrad_workspace_bytes", 48, 16, 3)
        nbf = _lib.size("da_conv3d_pack_bytes", 48, 16, 3)
        wsw = torch.empty(nbw, dtype=torch.uint8, device=dev)
        wsf = torch.empty(nbf, dtype=torch.uint8, device=dev)
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        P = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731

        def wgrad():   # tcgen05 weight gradient (32-channel halo blocks, 3xFP16) + region reduce + bias sum + the two max-abs passes
            _lib.call("da_conv3d_wgrad", P(x1), 32, P(x2), 16, P(dy), 0, P(gw), P(gb), 1, BS[0], BS[1], BS[2], 16, 3, 1, 1, P(wsw), nbw, st)

        def fwd():     # tcgen05 forward: weight image + max-abs pass + a 32-channel and a 16-channel launch (3xFP16)
            _lib.call("da_conv3d_fwd", P(x1), 32, P(x2), 16, P(w), 0, None, P(y), 1, BS[0], BS[1], BS[2], 16, 3, 1, 1, 0,
                      ctypes.c_float(0.0), P(wsf), nbf, st)

        def timeit(fn, reps=8):
            for _ in range(3):
                fn()
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            k0.record()
            for _ in range(reps):
                fn()
            k1.record()
            torch.cuda.synchronize()
            return k0.elapsed_time(k1) / reps

        wg_ms, fw_ms = timeit(wgrad), timeit(fwd)
        k_bytes = 4.0 * (48 * V + 16 * V + 16 * 48 * 27)        # read X (both sources) + dY, write dW
        f_bytes = 4.0 * (48 * V + 16 * V + 16 * 48 * 27)        # read X, W, write Y
        k_flop = 2.0 * 27 * 48 * 16 * V
        del x1, x2, w, dy, y, gw, gb, wsw, wsf
        probe = {"kind::f16 SS M=128 N=256": 1985.0, "kind::f16 SS M=128 N=144 (the shape these kernels issue)": 1382.0,
                 "kind::tf32 SS M=128 N=256": 1003.0, "source": "tools/probes/umma16_probe.cu on a B200, profiles/r02_umma16_probe_and_base.log"}

        def tensor_roofline(kernel, t_ms, alg_bytes, extra):
            ach = k_flop / (t_ms * 1e-3) / 1e12
            d = {"bound": "tensor", "kernel": kernel, "achieved": ach, "peak": mma_peak, "unit": "TFLOP/s", "frac": ach / mma_peak,
                 "peak_source": peak_src + ": dense bf16 cuBLAS throughput; kind::f16 MMAs (fp16 operands, fp32 accumulate) run at that rate",
                 "tensor_pipe_occupancy": 3.0 * ach / mma_peak,
                 "tensor_pipe_occupancy_note": "every fp32 product costs three fp16 MMAs (hi*hi + lo*hi + hi*lo): 3 x frac is the share of the "
                                               "measured dense peak the tensor pipe is busy for",
                 "measured_mma_issue_peaks_tflops": probe,
                 "launch_ms": t_ms, "algorithmic_flop_per_launch": k_flop, "algorithmic_bytes_per_launch": alg_bytes,
                 "hbm_achieved_gbs": alg_bytes / (t_ms * 1e-3) / 1e9, "hbm_frac": alg_bytes / (t_ms * 1e-3) / 1e9 / hbm_peak,
                 "hbm_peak": hbm_peak, "hbm_peak_source": peak_src}
            d.update(extra)
            return d

        roofline = tensor_roofline(
            "conv3d_umma_kernel (decBlock2.0 forward cat(32,16) -> 16 @160x192x160; tcgen05 kind::f16, 3xFP16 split): "
            "absmax + weight image + one 32-channel and one 16-channel launch", fw_ms, f_bytes,
            {"traffic": 1.513e9, "traffic_source": "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum of the two tcgen05 launches of this layer "
                                                  "(profiles/r02_ncu_a_conv_umma_tma_in_32ch.csv: 642 + 287 MB; r02_ncu_r02_umma_16ch.csv: 316 + 267 MB) "
                                                  "against 1.258 GB algorithmic: the 16-channel launch re-reads and re-writes Y",
             "step": {"algorithmic_bytes": wl["algo_bytes"], "achieved": wl["algo_bytes"] / (ms * 1e-3) / 1e9,
                      "frac": wl["algo_bytes"] / (ms * 1e-3) / 1e9 / hbm_peak}})
        roofline_wgrad = tensor_roofline(
            "conv3d_wgrad_umma16_kernel<32> (decBlock2.0 weight gradient, cat(32,16) x dY16 @160x192x160; tcgen05 3xFP16, one launch) "
            "+ absmax + region reduce + bias sum", wg_ms, k_bytes,
            {"traffic": 1.323e9, "traffic_source": "ncu --set full (profiles/r02_ncu_r02_wgrad16.csv): 1.319 GB read + 4.4 MB written against 1.258 GB algorithmic",
             "note": "128-row MMAs carry 96 useful rows (kx = 0..2 x 32 ci) in the 32-channel block and 48 in the padded 16-channel one"})
        cpu = parity = tcuda = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            D, H, W = SIZE
            times = oracle_steps(wl, SIZE, 1, 0, threads)
            cpu = {"value": wl["volumes"] / times[0], "unit": "volumes/s", "cores": threads, "kind": "port",
                   "sample": f"ONE full {D}x{H}x{W} step of the workload (forward, backward, Adam), measured once, no warm-up: "
                             f"{times[0]:.1f} s; `--impl reference` times 3 such steps after a warm-up"}
            if wl["kind"] == "joint" and not args.no_parity:
                parity = parity_check(dev)
        if world == 1 and not args.no_torch_cuda:
            tcuda = torch_cuda_baseline(wl)
        line = {"metric": wl["metric"], "value": wl["volumes"] * world / (ms * 1e-3), "unit": "volumes/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None,
                "dtype": "f32" if precision == "fp32" else "f16 tensor-core operands (single pass), f32 accumulate / storage / everything else",
                "data": "synthetic",
                "config": {"workload": wl["name"], "pairs_per_gpu": 1, "classes": CLASSES,
                           "seg": {"joint": "UNet_light(1,C,bias,BN)", "seg": "UNet(1,C,bias,BN) 32-base", "reg": None}[wl["kind"]],
                           "reg": None if wl["kind"] == "seg" else "VoxelMorphCVPR2018", "optimizer": "Adam(fused)", "parallelism": f"dp{world}",
                           "step_launch": graph_note,
                           "grad_bucket_bytes": bucket.nbytes,
                           "arithmetic": ("fp32 storage and accumulation; k3 convolutions on tcgen05 as 3xFP16 (fp16 hi/lo pairs of operands "
                                          "scaled per tensor by a power of two: fp32-grade products), k2 s2 deconvolutions on mma.sync as 3xTF32, "
                                          "everything else fp32 CUDA cores") if precision == "fp32" else
                                         ("fp32 storage and accumulation; k3 convolutions of the tensor-core levels as ONE fp16 MMA per product "
                                          "(operands scaled per tensor and rounded to 11 significant bits; declared tolerance 5e-3, "
                                          "tests/test_gpu_ops.py::test_conv3d_single_pass_mode), deep small levels exact FFMA, everything else fp32"),
                           "l2_policy": "working set per step (>10 GB of activations) far exceeds the 126 MB L2; no flush needed"},
                "e2e": {"value": wl["volumes"] * world / (ms_e2e * 1e-3), "unit": "volumes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                        "ms_per_step": ms_e2e, "input_path": "pinned host tensors -> DeviceInputStage (side-stream copy one step ahead, clip on device)",
                        "loss_readback": "every step through a pinned buffer, read by the host one step behind"},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "roofline_wgrad": roofline_wgrad, "cpu_baseline": cpu,
                "torch_cuda_baseline": tcuda, "parity": parity, "peak_mem_gb": peak_mem, "loss": last}
        print(json.dumps(line), flush=True)
    if world > 1:
        gstep = run = None   # the graph goes before the process group
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch_cuda"])
    ap.add_argument("--config", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-torch-cuda", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time the eager step instead of the captured CUDA graph")
    ap.add_argument("--no-overlap", action="store_true", help="joint step: registration net on the main stream (no side-stream overlap)")
    ap.add_argument("--precision", default=None, choices=["fp32", "f16x1"],
                    help="k3 convolutions: fp32-grade 3xFP16 (default, all configs but c2) or the single-pass reduced-precision mode (c2)")
    args = ap.parse_args()
    wl = WORKLOADS[args.config]
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args, wl)
    if args.impl == "torch_cuda":
        return run_torch_cuda(args, wl)
    return run_ours(args, wl)


if __name__ == "__main__":
    sys.exit(main())
