"""Per-parameter gradient error of UNet_light(1,4) + Dice at 16^3 against the fp64 oracle (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import deepatlas_b200 as da
from oracle import ref_port as P
from parity_util import cpu_state
ng = dict(np.load("tests/golden/nets.npz"))
cuda = torch.device("cuda:0")
torch.manual_seed(230)
net = da.get_network("UNet_light")(1, 4, bias=True, BN=True); net.weights_init()
sd64 = {k: (v.double().requires_grad_(True) if v.is_floating_point() and "running" not in k else (v.double() if v.is_floating_point() else v)) for k, v in cpu_state(net).items()}
net = net.to(cuda).train()
x = torch.from_numpy(ng["ul_x"]); lab = torch.from_numpy(ng["ul_labels"])
logits = net(x.to(cuda))
loss = da.get_loss_function("dice")(n_class=4, weight_type="Uniform", softmax=True, eps=1e-6)(logits, lab.to(cuda))
loss.backward()
l64 = P.dice_multiclass(P.unet_generator_forward(x.double(), sd64, 1, True), lab.long(), 4, "Uniform", False, True, 1e-6)
l64.backward()
gmax = max(float(v.grad.abs().max()) for k, v in sd64.items() if v.is_floating_point() and v.requires_grad)
print("loss", float(loss), float(l64), "gmax", gmax)
for k, p in net.named_parameters():
    t = sd64[k].grad
    e = float((p.grad.double().cpu() - t).abs().max())
    print(f"{k:45s} |truth| {float(t.abs().max()):.3e} abs err {e:.3e} rel-to-own {e / max(float(t.abs().max()), 1e-30):.2e} rel-to-gmax {e / gmax:.2e}")
