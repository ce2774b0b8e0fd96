"""Join the TDB_GEMM_LOG shape log with an ncu launch list (same run) -> per-shape time table."""
import collections
import csv
import re
import sys

ncu_csv, shape_log = sys.argv[1], sys.argv[2]
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0   # gemm launches logged before profiling started
lines = [l for l in open(ncu_csv) if not l.startswith("==")]
times = []
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum" or "tdb_gemm_kernel" not in row["Kernel Name"]:
        continue
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    times.append(v * {"ns": 1e-3, "us": 1.0, "ms": 1e3}[u])
shapes = [l.strip() for l in open(shape_log)][skip:]
print(len(times), "profiled gemm launches;", len(shapes), "logged shapes after skip")
n = min(len(times), len(shapes))
agg = collections.defaultdict(lambda: [0.0, 0, 0.0])
for t, s in zip(times[-n:], shapes[-n:]):
    d = dict(kv.split("=") for kv in s.split())
    fl = 2.0 * int(d["M"]) * int(d["N"]) * int(d["K"]) * int(d["taps"]) * int(d["nz"])
    a = agg[s]
    a[0] += t
    a[1] += 1
    a[2] += fl
tot = sum(a[0] for a in agg.values())
print("total gemm time %.2f ms" % (tot / 1e3))
for s, (t, c, fl) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:60]:
    print("%8.1f us %5.1f%% n=%3d avg %7.1f us %7.1f TFLOP/s | %s" % (t, 100 * t / tot, c, t / c, fl / t / 1e6, s))
