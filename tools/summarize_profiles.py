"""Turn the raw ncu artefacts under gpurun_out/ into the small tracked summaries under profiles/.
  python tools/summarize_profiles.py <launches.csv> <gemm_shapes.txt> <prof_gemm.ncu-rep> [<prof_xattn.ncu-rep>]"""
import collections
import csv
import re
import subprocess
import sys

OUT = "profiles"


def launch_table(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}[row["Metric Unit"]]
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r"void |at::|native::|<unnamed>::", "", name)[:80]
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    with open(out, "w") as f:
        f.write(f"# per-kernel device time of ONE eager training step (ncu gpu__time_duration.sum, cold cache, serialised)\n")
        f.write(f"# total {T / 1e3:.2f} ms over {sum(cnt.values())} launches; tdb::* = this repo's kernels\n")
        ours = sum(v for k, v in tot.items() if "tdb" in k)
        f.write(f"# share of this repo's kernels: {100 * ours / T:.1f} %\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:60]:
            f.write(f"{v:9.1f} us {100 * v / T:5.1f}% n={cnt[k]:4d} avg {v / cnt[k]:8.1f} us  {k}\n")


def raw_metrics(rep, out, title):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct",
            "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active"]
    with open(out, "w") as f:
        f.write(f"# {title}\n# source: ncu --set full --clock-control none (cold-cache, replayed; shares not absolutes)\n")
        for r in rows[2:]:
            f.write("----\n")
            for w in want:
                if w in idx:
                    f.write(f"{w} = {r[idx[w]]} {units[idx[w]]}\n")


if __name__ == "__main__":
    launch_table(sys.argv[1], f"{OUT}/r01_launches_step.txt")
    if len(sys.argv) > 3:
        raw_metrics(sys.argv[3], f"{OUT}/r01_ncu_gemm.txt", "tdb_gemm_kernel: layer3 3x3 implicit conv (A) and layer3 conv3 1x1 + residual (B), 100 frames")
    if len(sys.argv) > 4:
        raw_metrics(sys.argv[4], f"{OUT}/r01_ncu_xattn.txt", "xattn_fused_kernel: fused KV projection + time-aligned cross-attention, F=100, S=141")
