"""Timing of the three k2 s2 deconvolution passes of the C4 U-Net up-samplers through the C ABI (preallocated buffers)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from deepatlas_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
NT = 5


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


tot = [0.0, 0.0, 0.0]
for name, Cin, Cout, (D, H, W) in (("up2 32->32 @80x96x80", 32, 32, (80, 96, 80)), ("up1 64->64 @40x48x40", 64, 64, (40, 48, 40)),
                                     ("up0 64->64 @20x24x20", 64, 64, (20, 24, 20))) if not os.environ.get("DA_C2") else (
        ("dc3 128->128 @64^3", 128, 128, (64, 64, 64)), ("dc6 256->256 @32^3", 256, 256, (32, 32, 32)), ("dc9 512->512 @16^3", 512, 512, (16, 16, 16))):
    x = torch.rand((1, Cin, D, H, W), device=dev)
    w = torch.randn((Cin, Cout, 2, 2, 2), device=dev) * 0.1
    b = torch.zeros(Cout, device=dev)
    y = torch.empty((1, Cout, 2 * D, 2 * H, 2 * W), device=dev)
    dy = torch.rand_like(y)
    dx, gw, gb = torch.empty_like(x), torch.empty_like(w), torch.empty_like(b)
    nb = _lib.size("da_deconv_k2s2_wgrad_workspace_bytes", Cin, Cout)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    tf = timeit(lambda: _lib.call("da_deconv_k2s2_fwd", P(x), P(w), P(b), P(y), 1, Cin, Cout, D, H, W, st))
    td = timeit(lambda: _lib.call("da_deconv_k2s2_dgrad", P(dy), P(w), P(dx), 1, Cin, Cout, D, H, W, st))
    tw = timeit(lambda: _lib.call("da_deconv_k2s2_wgrad", P(x), P(dy), P(gw), P(gb), 1, Cin, Cout, D, H, W, P(ws), nb, st))
    byt = 4.0 * (x.numel() + y.numel())
    print(f"{name}: fwd {tf:.3f} ms ({byt / tf / 1e6:.0f} GB/s) dgrad {td:.3f} ms ({byt / td / 1e6:.0f} GB/s) wgrad {tw:.3f} ms ({byt / tw / 1e6:.0f} GB/s)")
    for i, t in enumerate((tf, td, tw)):
        tot[i] += 2 * t
print(f"per step (two volumes): fwd {tot[0]:.2f} dgrad {tot[1]:.2f} wgrad {tot[2]:.2f} ms, all {sum(tot):.2f} ms")
