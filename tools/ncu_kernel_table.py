"""Per-kernel table (time share, achieved DRAM GB/s, tensor-pipe %) of ONE eager training step from an ncu metrics CSV.

On the GPU box:
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
      --clock-control none --profile-from-start off --csv --log-file gpurun_out/step_metrics.csv python tools/ncu_step.py 2
Here:
  python tools/ncu_kernel_table.py gpurun_out/step_metrics.csv profiles/r01_kernel_table.txt
Values are cold-cache and serialised (ncu replays each launch): time SHARES and per-kernel ratios are meaningful, the
absolute step time is not.  GB/s = (dram read + write bytes) / duration summed over the kernel's launches; the tensor-pipe
figure is the duration-weighted mean of sm__pipe_tensor_cycles_active (% of peak sustained active)."""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "ms": 1e3, "s": 1e6, "second": 1e6,
        "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "%": 1.0, "": 1.0}


def short(name):
    name = re.sub(r"\(.*", "", name)
    return re.sub(r"void |at::|native::|<unnamed>::", "", name)[:70]


def main(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    per = collections.defaultdict(dict)           # launch id -> metric -> value
    names = {}
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", "")) * UNIT.get(row["Metric Unit"], 1.0)
        except (ValueError, KeyError):
            continue
        per[row["ID"]][row["Metric Name"]] = v
        names[row["ID"]] = short(row["Kernel Name"])
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])    # n, us, bytes, tensor% * us
    for i, m in per.items():
        us = m.get("gpu__time_duration.sum", 0.0)
        a = agg[names[i]]
        a[0] += 1
        a[1] += us
        a[2] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        a[3] += m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) * us
    total = sum(a[1] for a in agg.values())
    pk = 6555.2
    pj = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pj):
        pk = json.load(open(pj))["hbm_gbs"]
    with open(out, "w") as f:
        f.write("# per-kernel table of ONE eager training step at cfg-2 (ncu, cold cache, serialised replays: shares and ratios, not absolutes)\n")
        f.write(f"# total {total / 1e3:.2f} ms over {sum(a[0] for a in agg.values())} launches; HBM peak used for the fraction: {pk:.0f} GB/s (measured copy)\n")
        f.write(f"# {'kernel':70s} {'n':>5s} {'total us':>10s} {'share':>6s} {'avg us':>8s} {'DRAM GB/s':>10s} {'of HBM':>7s} {'tensor%':>8s}\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            if a[1] / total < 0.001 and "tdb" not in k:
                continue
            gbs = a[2] / (a[1] * 1e-6) / 1e9 if a[1] else 0.0
            f.write(f"{k:72s} {a[0]:5d} {a[1]:10.1f} {100 * a[1] / total:5.1f}% {a[1] / a[0]:8.1f} {gbs:10.0f} {gbs / pk:7.2f} {a[3] / a[1] if a[1] else 0:8.1f}\n")
    print(open(out).read()[:3000])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
