"""One conv3d forward + backward at a hot-layer shape, for ncu (single kernels):
  ncu --set full --clock-control none --import-source on -k regex:wgrad_tiled -c 1 -o gpurun_out/wg python tools/profile_conv.py
Env: DA_SHAPE="C1,C2,Cout,D,H,W" (default decBlock2.1: 16,0,16,160,192,160)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from deepatlas_b200 import ops  # noqa: E402

C1, C2, Cout, D, H, W = (int(v) for v in os.environ.get("DA_SHAPE", "16,0,16,160,192,160").split(","))
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(230)
xg = os.environ.get("DA_NO_DGRAD") != "1"   # DA_NO_DGRAD=1: the backward pass is the weight gradient alone
x1 = torch.rand((1, C1, D, H, W), device=dev, generator=g, requires_grad=xg)
x2 = torch.rand((1, C2, D, H, W), device=dev, generator=g, requires_grad=xg) if C2 else None
w = (torch.randn((Cout, C1 + C2, 3, 3, 3), device=dev, generator=g) * 0.05).requires_grad_(True)
b = torch.zeros(Cout, device=dev, requires_grad=True)
reps = int(os.environ.get("DA_REPS", "2"))
for _ in range(reps):
    y = ops.conv3d(x1, w, b, x2=x2)
    y.backward(torch.ones_like(y))
torch.cuda.synchronize()
# timing without a profiler attached (meaningless under ncu): NT calls back to back between two events, so that the
# host-side cost of a call (allocation, ctypes, tensor-map encoding: ~0.1 ms) hides behind the previous call's kernels
NT = int(os.environ.get("DA_NT", "6"))
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
dbg = os.environ.get("DA_UMMA_DEBUG") == "1"
if dbg:
    import ctypes
    from deepatlas_b200 import _lib
    buf = (ctypes.c_int64 * 11)()
    _lib.call("da_umma_debug_read", ctypes.cast(buf, ctypes.c_void_p))   # clear what the warm-up left
ys = []
ev[0].record()
for _ in range(NT):
    ys.append(ops.conv3d(x1, w, b, x2=x2))
ev[1].record()
torch.cuda.synchronize()
fwd_counters = None
if dbg:
    _lib.call("da_umma_debug_read", ctypes.cast(buf, ctypes.c_void_p))   # reads and clears: what follows belongs to the backward pass
    fwd_counters = list(buf)
gy = torch.ones_like(ys[0])
torch.cuda.synchronize()
ev[2].record()
for y in ys:
    y.backward(gy)
ev[3].record()
torch.cuda.synchronize()
fl = 2.0 * 27 * (C1 + C2) * Cout * D * H * W
tf, tb = ev[0].elapsed_time(ev[1]) / NT, ev[2].elapsed_time(ev[3]) / NT
print(f"shape {C1}+{C2}->{Cout} @{D}x{H}x{W}: fwd {tf:.3f} ms ({fl / tf / 1e9:.1f} TFLOP/s), "
      f"{'bwd' if xg else 'wgrad'} {tb:.3f} ms ({(2 if xg else 1) * fl / tb / 1e9:.1f} TFLOP/s)")

if dbg:
    _lib.call("da_umma_debug_read", ctypes.cast(buf, ctypes.c_void_p))
    if not xg:   # weight gradient alone: the counters are those of conv3d_wgrad_umma16_kernel
        mw, mt, braw, bempty, btot, tmaw, tiles, nct = list(buf)[:8]
        if tiles:
            print(f"  wgrad per tile: MMA warp waits for operands {mw / tiles:.0f} of {mt / tiles:.0f} cycles; B producer warp waits for raw tiles {braw / tiles:.0f}, "
                  f"for a free stage {bempty / tiles:.0f}, of {btot / tiles:.0f}; TMA thread waits {tmaw / tiles:.0f} ({tiles} tiles, {nct} CTAs)")
    acc, plane, issue, total, steps, ctas, ew, et, etot, ebar, eout = fwd_counters
    if steps:
        print(f"  umma MMA warp per plane step: wait-accumulator {acc / steps:.0f}, wait-planes {plane / steps:.0f}, issue {issue / steps:.0f}, "
              f"total {total / steps:.0f} cycles; epilogue warp 0: wait-MMA {ew / steps:.0f}, tmem {et / steps:.0f}, edge-barrier {ebar / steps:.0f}, fold+store {eout / steps:.0f}, total {etot / steps:.0f} ({steps} steps, {ctas} CTAs)")
