"""Layer-by-layer comparison of the gradient arriving at every activation output (GPU path vs fp64 oracle)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import deepatlas_b200 as da
from deepatlas_b200 import networks, ops
from oracle import ref_port as P
from parity_util import cpu_state
ng = dict(np.load("tests/golden/nets.npz"))
cuda = torch.device("cuda:0")
torch.manual_seed(230)
net = da.get_network("UNet_light")(1, 4, bias=True, BN=True); net.weights_init()
sd64 = {k: (v.double().requires_grad_(True) if v.is_floating_point() and "running" not in k else (v.double() if v.is_floating_point() else v)) for k, v in cpu_state(net).items()}
net = net.to(cuda).train()
ours, truth, ours_fwd, truth_fwd = [], [], [], []
orig_bn = ops.bn_act
def bn_hook(*a, **k):
    y = orig_bn(*a, **k)
    i = len(ours_fwd); ours_fwd.append(y.detach()); ours.append(None)
    y.register_hook(lambda g, i=i: ours.__setitem__(i, g.detach().clone()))
    return y
ops.bn_act = bn_hook
orig_act = P._act
def act_hook(x, act):
    y = orig_act(x, act)
    i = len(truth_fwd); truth_fwd.append(y.detach()); truth.append(None)
    y.register_hook(lambda g, i=i: truth.__setitem__(i, g.detach().clone()))
    return y
P._act = act_hook
x = torch.from_numpy(ng["ul_x"]); lab = torch.from_numpy(ng["ul_labels"])
logits = net(x.to(cuda))
da.get_loss_function("dice")(n_class=4, weight_type="Uniform", softmax=True, eps=1e-6)(logits, lab.to(cuda)).backward()
P.dice_multiclass(P.unet_generator_forward(x.double(), sd64, 1, True), lab.long(), 4, "Uniform", False, True, 1e-6).backward()
print("layers", len(ours), len(truth))
for i, (a, b, fa, fb) in enumerate(zip(ours, truth, ours_fwd, truth_fwd)):
    a = a.double().cpu(); fa = fa.double().cpu()
    d = a - b
    mx = float(b.abs().max())
    dims = tuple(range(2, d.dim()))
    print(f"{i:2d} shape {tuple(b.shape)} fwd err {float((fa - fb).abs().max() / fb.abs().max()):.2e} | grad max|b| {mx:.3e} maxerr/max {float(d.abs().max()) / mx:.2e} "
          f"mean-offset/max {float(d.mean(dims).abs().max()) / mx:.2e} nbad {int((d.abs() > 1e-3 * mx).sum())}/{d.numel()} neg-mask-mismatch {int(((fa > 0) != (fb > 0)).sum())}")
