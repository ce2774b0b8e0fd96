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

# per-kernel summary + 0.5 ms windows (busy fraction = union of kernel intervals, sum = concurrency) next to the raw timeline
import collections  # noqa: E402
import re  # noqa: E402

summ = out.replace(".txt", "") + "_kernels.txt"
tot, cnt = collections.defaultdict(float), collections.Counter()
for e in evs:
    nm = re.sub(r"void |at::native::|\(.*", "", e.name)[:80]
    tot[nm] += e.time_range.end - e.time_range.start
    cnt[nm] += 1
T = sum(tot.values())
span = (evs[-1].time_range.end - t0) if evs else 0
with open(summ, "w") as f:
    f.write("# torch.profiler (CUPTI) kernel records of ONE CUDA-graph replay of the bench step (tools/step_timeline.py); profiling inflates "
            "durations: shares and order, not absolutes\n")
    f.write(f"# {len(evs)} kernels, summed kernel time {T / 1e3:.2f} ms over a {span / 1e3:.2f} ms span (streams overlap); tdb::* = this repo\n")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:45]:
        f.write(f"{v:9.1f} us {100 * v / T:5.1f}% n={cnt[k]:4d} avg {v / cnt[k]:7.1f} us  {k}\n")
    f.write("# windows of 0.5 ms: busy % (any kernel running), sum % (summed kernel time / window), kernels, top 3 by time\n")
    W = 500.0
    for b in range(int(span // W) + 1):
        lo, hi = b * W, (b + 1) * W
        segs = [(max(e.time_range.start - t0, lo), min(e.time_range.end - t0, hi), e.name) for e in evs
                if e.time_range.start - t0 < hi and e.time_range.end - t0 > lo]
        busy, cur = 0.0, None
        for a, c in sorted((a, c) for a, c, _ in segs):
            if cur is None:
                cur = [a, c]
            elif a <= cur[1]:
                cur[1] = max(cur[1], c)
            else:
                busy += cur[1] - cur[0]
                cur = [a, c]
        if cur:
            busy += cur[1] - cur[0]
        cc = collections.Counter()
        for a, c, n in segs:
            cc[re.sub(r"void |tdb::|at::native::|\(.*", "", n)[:26]] += c - a
        f.write(f"{lo / 1e3:5.1f} ms busy {100 * busy / W:5.1f}% sum {100 * sum(c - a for a, c, _ in segs) / W:5.1f}% n={len(segs):3d} | "
                + ", ".join(f"{k}:{v:.0f}" for k, v in cc.most_common(3)) + "\n")
print("summary ->", summ)
