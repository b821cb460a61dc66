"""Kernel-level timing of one k3 s1 p1 convolution layer through the C ABI with preallocated buffers (no allocator or
autograd in the timed region): forward, data gradient, weight gradient, each as DA_NT back-to-back calls between two
events.  Env: DA_SHAPE="C1,C2,Cout,D,H,W".  With DA_UMMA_DEBUG=1 also prints the kernels' cycle counters."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from deepatlas_b200 import _lib  # noqa: E402

C1, C2, Cout, D, H, W = (int(v) for v in os.environ.get("DA_SHAPE", "16,0,16,160,192,160").split(","))
NT = int(os.environ.get("DA_NT", "6"))
dev = torch.device("cuda:0")
if os.environ.get("DA_IMPL"):   # 0 auto, 1 direct, 2 tiled FFMA, 3 tensor cores forced
    _lib.call("da_set_conv_impl", int(os.environ["DA_IMPL"]))
g = torch.Generator(device=dev).manual_seed(230)
Cin = C1 + C2
x1 = torch.rand((1, C1, D, H, W), device=dev, generator=g)
x2 = torch.rand((1, C2, D, H, W), device=dev, generator=g) if C2 else None
w = torch.randn((Cout, Cin, 3, 3, 3), device=dev, generator=g) * 0.05
b = torch.zeros(Cout, device=dev)
y = torch.empty((1, Cout, D, H, W), device=dev)
dy = torch.rand((1, Cout, D, H, W), device=dev, generator=g)
dx = torch.empty((1, Cin, D, H, W), device=dev)
gw, gb = torch.empty_like(w), torch.empty_like(b)
nf = _lib.size("da_conv3d_pack_bytes", Cin, Cout, 3)
nd = _lib.size("da_conv3d_dgrad_workspace_bytes", 1, Cin, Cout, D, H, W, 3, 1)
nw = _lib.size("da_conv3d_wgrad_workspace_bytes", Cin, Cout, 3)
ws = torch.empty(max(nf, nd, nw), dtype=torch.uint8, device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731


def fwd():
    _lib.call("da_conv3d_fwd", P(x1), C1, P(x2), C2, P(w), 0, P(b), P(y), 1, D, H, W, Cout, 3, 1, 1, 0, ctypes.c_float(0.0), P(ws), nf, st)


def dgrad():
    _lib.call("da_conv3d_dgrad", P(dy), P(w), 0, P(dx), 1, Cin, 0, Cin, Cout, D, H, W, 3, 1, 1, P(ws), nd, st)


def wgrad():
    _lib.call("da_conv3d_wgrad", P(x1), C1, P(x2), C2, P(dy), 0, P(gw), P(gb), 1, D, H, W, Cout, 3, 1, 1, P(ws), nw, st)


def timeit(fn):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(NT):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / NT


dbg = os.environ.get("DA_UMMA_DEBUG") == "1"
buf = (ctypes.c_int64 * 11)()


def counters():
    _lib.call("da_umma_debug_read", ctypes.cast(buf, ctypes.c_void_p))
    return list(buf)


fl = 2.0 * 27 * Cin * Cout * D * H * W
out = [f"shape {C1}+{C2}->{Cout} @{D}x{H}x{W}:"]
for name, fn in (("fwd", fwd), ("dgrad", dgrad), ("wgrad", wgrad)):
    if dbg:
        counters()
    t = timeit(fn)
    out.append(f"{name} {t:.3f} ms ({fl / t / 1e9:.0f} TFLOP/s)")
    if dbg:
        c = counters()
        if name == "wgrad":
            mw, mt, braw, bempty, btot, tmaw, tiles, nct = c[:8]
            if tiles:
                out.append(f"\n   wgrad per tile: MMA warp waits for operands {mw / tiles:.0f} of {mt / tiles:.0f} cycles; B producer waits raw {braw / tiles:.0f}, "
                           f"free stage {bempty / tiles:.0f}, of {btot / tiles:.0f}; TMA thread waits {tmaw / tiles:.0f}\n  ")
        else:
            acc, plane, issue, total, steps, ctas, ew, et, etot, ebar, eout = c
            if steps:
                out.append(f"\n   {name} per plane step: MMA warp wait-acc {acc / steps:.0f}, wait-input {plane / steps:.0f}, total {total / steps:.0f}; "
                           f"epilogue: wait-MMA {ew / steps:.0f}, tmem {et / steps:.0f}, edge-bar {ebar / steps:.0f}, fold+store {eout / steps:.0f}, total {etot / steps:.0f}\n  ")
print(" ".join(out))
