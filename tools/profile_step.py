"""One joint step inside a cudaProfilerStart/Stop range, for ncu:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:<kernel> -c 3 -o gpurun_out/prof python tools/profile_step.py
Numbers printed by a run under ncu are never bench values."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from deepatlas_b200.dist import FlatGradBucket  # noqa: E402
from deepatlas_b200.joint import JointModel, make_synthetic_pair  # noqa: E402

size = tuple(int(x) for x in os.environ.get("DA_SIZE", "160,192,160").split(","))
classes = int(os.environ.get("DA_CLASSES", "32"))
dev = torch.device("cuda:0")
torch.manual_seed(230)
model = JointModel(n_classes=classes).to(dev)
model.weights_init()
bucket = FlatGradBucket(model.trainable_parameters())
opt = torch.optim.Adam(bucket.params, lr=1e-3, fused=True)
batch = make_synthetic_pair(size, classes, seed=230, device=dev)


def step():
    bucket.zero()
    loss, _ = model.joint_loss(*batch)
    loss.backward()
    opt.step()
    return loss


for _ in range(int(os.environ.get("DA_WARM", "2"))):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(loss))
