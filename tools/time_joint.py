"""Joint seg+reg step at an arbitrary extent (configs C3/C5 sanity): DA_SIZE=D,H,W DA_CLASSES=C python tools/time_joint.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepatlas_b200.dist import FlatGradBucket
from deepatlas_b200.joint import JointModel, make_synthetic_pair
size = tuple(int(x) for x in os.environ.get("DA_SIZE", "256,256,256").split(","))
classes = int(os.environ.get("DA_CLASSES", "4"))
dev = torch.device("cuda:0")
torch.manual_seed(230)
model = JointModel(n_classes=classes).to(dev); model.weights_init()
bucket = FlatGradBucket(model.trainable_parameters())
opt = torch.optim.Adam(bucket.params, lr=1e-3, fused=True)
batch = make_synthetic_pair(size, classes, seed=230, device=dev)
def step():
    bucket.zero(); loss, _ = model.joint_loss(*batch); loss.backward(); opt.step(); return loss
for _ in range(2): l = step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): l = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"joint step {size} C={classes}: {ms:.1f} ms/step = {2000.0 / ms:.2f} volumes/s, loss {float(l.detach()):.5f} finite={bool(torch.isfinite(l))}, peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
