"""Kernel timeline of ONE CUDA-graph replay of the bench step (torch.profiler / CUPTI): start, duration, stream of every kernel.
  python tools/step_timeline.py gpurun_out/timeline.txt
Not a timing source (profiling overhead): it shows ORDER, overlap between streams and where the gaps are."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
model, crit, wd = bench.build_everything(dev)
st = bench.Step(model, crit, wd, dev, 0, 1, use_graph=True)
st.capture()
for _ in range(3):
    st.run()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    st.run()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/timeline.txt"
t0 = evs[0].time_range.start if evs else 0
with open(out, "w") as f:
    f.write("# start_us  dur_us  stream  kernel\n")
    for e in evs:
        f.write(f"{e.time_range.start - t0:10.1f} {e.time_range.end - e.time_range.start:8.1f} {getattr(e, 'device_index', 0)}:{getattr(e, 'stream', -1) if hasattr(e, 'stream') else -1} {e.name[:90]}\n")
print(len(evs), "device events ->", out, "span us:", (evs[-1].time_range.end - t0) if evs else 0)
