"""Per-layer timing of every k3 convolution of the C4 joint step through the C ABI (forward, data gradient(s), weight
gradient; preallocated buffers, no autograd): where the convolution time of a step goes, layer by layer.
Env: DA_SIZE="160,192,160", DA_NT repeats, DA_ONLY=substring filter on the layer name, DA_IMPL=da_set_conv_impl code."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from deepatlas_b200 import _lib  # noqa: E402

D0, H0, W0 = (int(v) for v in os.environ.get("DA_SIZE", "160,192,160").split(","))
NT = int(os.environ.get("DA_NT", "4"))
ONLY = os.environ.get("DA_ONLY", "")
dev = torch.device("cuda:0")
if os.environ.get("DA_IMPL"):   # 0 auto, 1 direct, 2 tiled FFMA, 3 tensor cores forced
    _lib.call("da_set_conv_impl", int(os.environ["DA_IMPL"]))
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731

# (name, C1, C2, Cout, level of the INPUT (0 = full), stride, needs dx, multiplicity per step)
LAYERS = [
    ("seg.enc0.0   1->8", 1, 0, 8, 0, 1, False, 2),
    ("seg.enc0.1   8->16", 8, 0, 16, 0, 1, True, 2),
    ("seg.enc1.0  16->16", 16, 0, 16, 1, 1, True, 2),
    ("seg.enc1.1  16->32", 16, 0, 32, 1, 1, True, 2),
    ("seg.enc2.0  32->32", 32, 0, 32, 2, 1, True, 2),
    ("seg.enc2.1  32->64", 32, 0, 64, 2, 1, True, 2),
    ("seg.enc3.0  64->64", 64, 0, 64, 3, 1, True, 4),
    ("seg.dec0.0 64+64->64", 64, 64, 64, 2, 1, True, 2),
    ("seg.dec0.1  64->64", 64, 0, 64, 2, 1, True, 2),
    ("seg.dec1.0 64+32->32", 64, 32, 32, 1, 1, True, 2),
    ("seg.dec1.1  32->32", 32, 0, 32, 1, 1, True, 2),
    ("seg.dec2.0 32+16->16", 32, 16, 16, 0, 1, True, 2),
    ("seg.dec2.1  16->16", 16, 0, 16, 0, 1, True, 2),
    ("reg.enc0    1+1->16", 1, 1, 16, 0, 1, False, 1),
    ("reg.enc1 s2 16->32", 16, 0, 32, 0, 2, True, 1),
    ("reg.enc2 s2 32->32", 32, 0, 32, 1, 2, True, 1),
    ("reg.enc3 s2 32->32", 32, 0, 32, 2, 2, True, 1),
    ("reg.dec1    32->32", 32, 0, 32, 3, 1, True, 1),
    ("reg.dec2 32+32->32", 32, 32, 32, 2, 1, True, 1),
    ("reg.dec3 32+32->32", 32, 32, 32, 1, 1, True, 1),
    ("reg.dec4 32+32->8", 32, 32, 8, 1, 1, True, 1),
    ("reg.dec5     8->8", 8, 0, 8, 0, 1, True, 1),
    ("reg.flow  8+16->3", 8, 16, 3, 0, 1, True, 1),
]


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


tot = {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0}
g = torch.Generator(device=dev).manual_seed(230)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
print(f"{'layer':24s} {'x':>3s} {'fwd':>8s} {'dgrad':>8s} {'wgrad':>8s}   GFLOP   TF/s(f,d,w)")
for name, C1, C2, Cout, lvl, stride, need_dx, mult in LAYERS:
    if ONLY and ONLY not in name:
        continue
    D, H, W = D0 >> lvl, H0 >> lvl, W0 >> lvl
    Do, Ho, Wo = (D - 1) // stride + 1, (H - 1) // stride + 1, (W - 1) // stride + 1
    Cin = C1 + C2
    x1 = torch.rand((1, C1, D, H, W), device=dev, generator=g)
    x2 = torch.rand((1, C2, D, H, W), device=dev, generator=g) if C2 else None
    w = torch.randn((Cout, Cin, 3, 3, 3), device=dev, generator=g) * 0.05
    b = torch.zeros(Cout, device=dev)
    y = torch.empty((1, Cout, Do, Ho, Wo), device=dev)
    dy = torch.rand((1, Cout, Do, Ho, Wo), device=dev, generator=g)
    dx1 = torch.empty_like(x1)
    dx2 = torch.empty_like(x2) if C2 else None
    gw, gb = torch.empty_like(w), torch.empty_like(b)
    nf = _lib.size("da_conv3d_pack_bytes", Cin, Cout, 3)
    nd = _lib.size("da_conv3d_dgrad_workspace_bytes", 1, Cin, Cout, D, H, W, 3, stride)
    nw = _lib.size("da_conv3d_wgrad_workspace_bytes", Cin, Cout, 3)
    ws = torch.empty(max(nf, nd, nw), dtype=torch.uint8, device=dev)
    ax = torch.empty(1, device=dev)
    ady = torch.empty(1, device=dev)

    def fwd():
        _lib.call("da_conv3d_fwd_ex", P(x1), C1, P(x2), C2, P(w), 0, P(b), P(y), 1, D, H, W, Cout, 3, stride, 1, 0, 0.0, P(ws), nf,
                  st, P(ax), 0)

    def dgrad():
        _lib.call("da_conv3d_dgrad_ex", P(dy), P(w), 0, P(dx1), 1, Cin, 0, C1, Cout, D, H, W, 3, stride, 1, P(ws), nd, st, P(ady), 0)
        if C2:
            _lib.call("da_conv3d_dgrad_ex", P(dy), P(w), 0, P(dx2), 1, Cin, C1, C2, Cout, D, H, W, 3, stride, 1, P(ws), nd, st,
                      P(ady), 1)

    def wgrad():
        _lib.call("da_conv3d_wgrad_ex", P(x1), C1, P(x2), C2, P(dy), 0, P(gw), P(gb), 1, D, H, W, Cout, 3, stride, 1, P(ws), nw, st,
                  P(ax), 1, P(ady), 1, 0)

    tf = timeit(fwd)
    td = timeit(dgrad) if need_dx else 0.0
    tw = timeit(wgrad)
    fl = 2.0 * 27 * Cin * Cout * Do * Ho * Wo
    r = lambda t: f"{fl / t / 1e9:5.0f}" if t else "    -"  # noqa: E731
    print(f"{name:24s} {mult:3d} {tf:8.3f} {td:8.3f} {tw:8.3f} {fl / 1e9:7.1f}   {r(tf)} {r(td)} {r(tw)}")
    tot["fwd"] += mult * tf
    tot["dgrad"] += mult * td
    tot["wgrad"] += mult * tw
    del x1, x2, w, y, dy, dx1, dx2, ws
print(f"per step: fwd {tot['fwd']:.2f} ms, dgrad {tot['dgrad']:.2f} ms, wgrad {tot['wgrad']:.2f} ms, all {sum(tot.values()):.2f} ms")
