"""Config C2 sanity (fp32, the bf16 mode is not built): UNet (32-base) forward+backward at DA_SIZE (default 128^3)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepatlas_b200 as da
size = tuple(int(x) for x in os.environ.get("DA_SIZE", "128,128,128").split(","))
dev = torch.device("cuda:0")
torch.manual_seed(230)
net = da.get_network("UNet")(1, 4, bias=True, BN=True).to(dev); net.weights_init()
crit = da.get_loss_function("dice")(n_class=4, weight_type="Uniform", softmax=True, eps=1e-6)
x = torch.rand((1, 1) + size, device=dev); lab = torch.randint(0, 4, (1,) + size, device=dev, dtype=torch.uint8)
def step():
    net.zero_grad(set_to_none=True)
    loss = crit(net(x), lab); loss.backward(); return loss
for _ in range(2): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); 
for _ in range(3): l = step()
e1.record(); torch.cuda.synchronize()
print(f"UNet(32-base) {size} fp32 fwd+bwd: {e0.elapsed_time(e1) / 3:.1f} ms/step, loss {float(l):.5f}, peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
