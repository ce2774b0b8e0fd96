"""One eager training step of the bench workload (for `ncu --metrics gpu__time_duration.sum`): warm-up, then 1 step."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
model, crit, wd = bench.build_everything(dev)
st = bench.Step(model, crit, wd, dev, 0, 1, use_graph=False)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for i in range(steps):
    if i == steps - 1:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    st.body()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done, loss", st.loss.item())
