#!/usr/bin/env python
"""bench.py -- volumes/sec (fwd+bwd) of the 160x192x160 joint seg+reg step (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W                 our arm (CUDA path through the C ABI)
  python bench.py --impl reference --gpus N --steps K --warmup W   reference arm: the CPU restatement of the
        reference's own PyTorch path (oracle/ref_port.py; /root/reference cannot travel to the GPU box) timed
        on the host cores.

One "step" = one joint seg+reg training step on one synthetic volume pair per GPU: two UNet_light(1,32) passes,
VoxelMorph + warp, LNCC + bending + anatomy Dice + 2 supervised Dice, one backward, one flat-bucket gradient
all-reduce (N>1), fused Adam.  1 pair = 2 volumes.  Prints ONE JSON line on rank 0.
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

WORKLOAD = "joint_seg_reg_160x192x160_c32_fp32"
SIZE = (160, 192, 160)
CLASSES = 32
CPU_SAMPLE_SIZE = (80, 96, 80)      # bounded CPU sample: 1/8 of the voxels of the workload volume
ALGO_BYTES_STEP = 42.0e9            # SURVEY.md 8(d) / BASELINE.md section 4: algorithmic bytes per C4 step per GPU


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


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
def cpu_joint_steps(size, classes, steps, warmup, threads):
    """The reference's CPU path (restated in oracle/ref_port.py) for the same step definition, incl. Adam."""
    import torch
    from oracle import ref_port as P
    from deepatlas_b200.joint import make_synthetic_pair
    torch.set_num_threads(threads)
    # parameter shapes come from our mirror classes (identical state_dict to the reference, see tests)
    import deepatlas_b200 as da
    torch.manual_seed(230)
    seg = da.get_network("UNet_light")(1, classes, bias=True, BN=True); seg.weights_init()
    reg = da.get_network("voxel_morph_cvpr")(); reg.weights_init()
    seg_sd = {k: (v.detach().clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.detach().clone())
              for k, v in seg.state_dict().items()}
    reg_sd = {k: v.detach().clone().requires_grad_(True) for k, v in reg.state_dict().items()}
    leaves = [v for v in list(seg_sd.values()) + list(reg_sd.values()) if v.is_floating_point() and v.requires_grad]
    opt = torch.optim.Adam(leaves, lr=1e-3)
    batch = make_synthetic_pair(size, classes, seed=230)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss = P.joint_loss(seg_sd, reg_sd, batch, classes)
        loss.backward()
        opt.step()
        _ = loss.item()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return times


def run_reference(args):
    """Reference arm: CPU, all host threads, bounded sample of the workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    threads = os.cpu_count() or 1
    scale = (CPU_SAMPLE_SIZE[0] * CPU_SAMPLE_SIZE[1] * CPU_SAMPLE_SIZE[2]) / float(SIZE[0] * SIZE[1] * SIZE[2])
    steps = max(1, min(args.steps, 3))
    times = cpu_joint_steps(CPU_SAMPLE_SIZE, CLASSES, steps, min(args.warmup, 1), threads)
    t = sorted(times)[len(times) // 2]
    value = 2.0 * scale / t
    sample = (f"joint step on one {CPU_SAMPLE_SIZE[0]}x{CPU_SAMPLE_SIZE[1]}x{CPU_SAMPLE_SIZE[2]} pair (1/8 of the workload's voxels), "
              f"median of {len(times)} steps; volumes/s scaled by voxel count to {SIZE[0]}x{SIZE[1]}x{SIZE[2]} volumes")
    line = {"impl": "reference", "metric": "volumes/sec (fwd+bwd) 160x192x160 joint seg+reg", "value": value, "unit": "volumes/s",
            "n_gpus": args.gpus, "steps": len(times), "warmup": min(args.warmup, 1), "ms_per_step": t * 1e3 / scale,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_gpu": 1, "classes": CLASSES, "device": "cpu", "torch_threads": threads},
            "cpu_baseline": {"value": value, "unit": "volumes/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import deepatlas_b200 as da
    from deepatlas_b200 import _lib, ops
    from deepatlas_b200.dist import FlatGradBucket, broadcast_parameters, init_from_env
    from deepatlas_b200.joint import JointModel, make_synthetic_pair

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (our arm) needs a CUDA device; there is no CPU fallback")
    rank, local, world = init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.load()

    torch.manual_seed(230)
    model = JointModel(n_classes=CLASSES).to(dev)
    model.weights_init()
    broadcast_parameters(model)
    bucket = FlatGradBucket(model.trainable_parameters())
    opt = torch.optim.Adam(bucket.params, lr=1e-3, fused=True)

    host = [t.pin_memory() for t in make_synthetic_pair(SIZE, CLASSES, seed=230 + rank)]
    dev_batch = [t.to(dev) for t in host]
    h2d = sum(t.numel() * t.element_size() for t in host)

    def step(batch):
        bucket.zero()
        loss, _ = model.joint_loss(*batch)
        loss.backward()
        bucket.allreduce(world)
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(dev_batch)
    # ---- device-resident timing -------------------------------------------------------------------------
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    l0 = _lib.size("da_launch_count")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step(dev_batch)
    e1.record()
    barrier()
    clocks = sampler.stop()
    launches = _lib.size("da_launch_count") - l0
    ms = e0.elapsed_time(e1) / args.steps
    # ---- end to end: host buffers in, loss scalar out, every step ---------------------------------------
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    last = None
    for _ in range(args.steps):
        batch = [t.to(dev, non_blocking=True) for t in host]
        last = float(step(batch).item())
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1) / args.steps
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    line = None
    if rank == 0:
        peak, peak_src = _peaks()
        # ---- dominant kernels, timed live on the launching stream (decBlock2.0 of UNet_light: cat(32,16) -> 16 @160x192x160) ----
        import ctypes
        V = SIZE[0] * SIZE[1] * SIZE[2]
        x1 = torch.rand((1, 32, *SIZE), device=dev)
        x2 = torch.rand((1, 16, *SIZE), device=dev)
        w = torch.randn((16, 48, 3, 3, 3), device=dev) * 0.03
        dy = torch.rand((1, 16, *SIZE), device=dev)
        gw, gb = torch.empty_like(w), torch.empty(16, device=dev)
        nbw = _lib.size("da_conv3d_wgrad_workspace_bytes", 48, 16, 3)
        wsw = torch.empty(nbw, dtype=torch.uint8, device=dev)
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        P = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731

        def wgrad():   # tcgen05 weight gradient (second largest share of the step)
            _lib.call("da_conv3d_wgrad", P(x1), 32, P(x2), 16, P(dy), 0, P(gw), P(gb), 1, SIZE[0], SIZE[1], SIZE[2], 16, 3, 1, 1, P(wsw), nbw, st)

        def fwd():     # tcgen05 3xTF32 forward (three 16-channel chunks accumulate)
            ops.conv3d(x1, w, None, x2=x2)

        def timeit(fn, reps=5):
            for _ in range(2):
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
        tf32_peak = None
        try:
            tf32_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"]) / 2.0
        except Exception:
            tf32_peak = 1590.0 / 2.0
        del x1, x2, w, dy, gw, gb, wsw
        # dominant kernel of the step (profiles/r01_launches_step_tcgen05_wgrad.csv: conv3d_umma_kernel 28 %, conv3d_wgrad_umma_tma_kernel 22 %)
        roofline = {"bound": "tensor", "kernel": "conv3d_umma_kernel x3 channel-chunk launches (decBlock2.0 forward cat(32,16) -> 16 @160x192x160; tcgen05 kind::tf32, 3xTF32 split)",
                    "achieved": 3.0 * k_flop / (fw_ms * 1e-3) / 1e12, "peak": tf32_peak, "unit": "TFLOP/s",
                    "frac": 3.0 * k_flop / (fw_ms * 1e-3) / 1e12 / tf32_peak,
                    "useful_fp32_equivalent_tflops": k_flop / (fw_ms * 1e-3) / 1e12,
                    "traffic": 3 * 595.0e6,
                    "traffic_source": "ncu --set full, profiles/r01_ncu_c_conv_umma_tcgen05.csv: dram 318 MB read + 277 MB write per chunk launch, three launches per layer",
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops / 2 (kind::tf32 runs at half the bf16 rate); achieved counts the three TF32 MMAs per fp32 product",
                    "launch_ms": fw_ms, "algorithmic_flop_per_launch": k_flop, "algorithmic_bytes_per_launch": f_bytes,
                    "hbm_achieved_gbs": f_bytes / (fw_ms * 1e-3) / 1e9, "hbm_frac": f_bytes / (fw_ms * 1e-3) / 1e9 / peak, "hbm_peak": peak, "hbm_peak_source": peak_src,
                    "step": {"algorithmic_bytes": ALGO_BYTES_STEP, "achieved": ALGO_BYTES_STEP / (ms * 1e-3) / 1e9,
                             "frac": ALGO_BYTES_STEP / (ms * 1e-3) / 1e9 / peak}}
        roofline_wgrad = {"bound": "tensor", "kernel": "conv3d_wgrad_umma_tma_kernel (decBlock2.0 weight gradient, cat(32,16) x dY16 @160x192x160; tcgen05 3xTF32, one launch) + region reduce + bias sum",
                          "achieved": 3.0 * k_flop / (wg_ms * 1e-3) / 1e12, "peak": tf32_peak, "unit": "TFLOP/s",
                          "frac": 3.0 * k_flop / (wg_ms * 1e-3) / 1e12 / tf32_peak,
                          "useful_fp32_equivalent_tflops": k_flop / (wg_ms * 1e-3) / 1e12,
                          "traffic": 1.3485e9,
                          "traffic_source": "ncu --set full, profiles/r01_ncu_d_wgrad_umma_tma.csv (dram read 1343.9 MB + write 4.6 MB; tensor pipe active 79 %)",
                          "launch_ms": wg_ms, "algorithmic_bytes_per_launch": k_bytes,
                          "hbm_achieved_gbs": k_bytes / (wg_ms * 1e-3) / 1e9, "hbm_frac": k_bytes / (wg_ms * 1e-3) / 1e9 / peak,
                          "note": "128-row MMAs carry 48 useful rows (kx = 0..2 x 16 ci): the tensor pipe is 79 % busy while the useful fp32-equivalent rate is what the step sees"}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            scale = (CPU_SAMPLE_SIZE[0] * CPU_SAMPLE_SIZE[1] * CPU_SAMPLE_SIZE[2]) / float(V)
            times = cpu_joint_steps(CPU_SAMPLE_SIZE, CLASSES, 2, 1, threads)
            tc = min(times)
            cpu = {"value": 2.0 * scale / tc, "unit": "volumes/s", "cores": threads, "kind": "port",
                   "sample": f"joint step on one {CPU_SAMPLE_SIZE[0]}x{CPU_SAMPLE_SIZE[1]}x{CPU_SAMPLE_SIZE[2]} pair (1/8 of the voxels), "
                             f"best of 2 after 1 warm-up, scaled by voxel count"}
        line = {"metric": "volumes/sec (fwd+bwd) 160x192x160 joint seg+reg", "value": 2.0 * world / (ms * 1e-3), "unit": "volumes/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "pairs_per_gpu": 1, "classes": CLASSES, "seg": "UNet_light(1,32,bias,BN)",
                           "reg": "VoxelMorphCVPR2018", "optimizer": "Adam(fused)", "parallelism": f"dp{world}",
                           "grad_bucket_bytes": bucket.nbytes,
                           "l2_policy": "working set per step (>10 GB of activations) far exceeds the 126 MB L2; no flush needed"},
                "e2e": {"value": 2.0 * world / (ms_e2e * 1e-3), "unit": "volumes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                        "ms_per_step": ms_e2e},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "roofline_wgrad": roofline_wgrad, "cpu_baseline": cpu,
                "loss": last}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
