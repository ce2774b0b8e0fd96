"""gpurun_out/*.ncu-rep / launch CSV of round 2 -> small tracked summaries under profiles/.
  python tools/summarize_r02.py"""
import collections
import csv
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def raw(rep, out, title):
    if not os.path.exists(rep):
        return None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    res = []
    with open(out, "w") as f:
        f.write(f"# {title}\n# source: ncu --set full --clock-control none --import-source on (cold cache, replayed: ratios, not absolutes)\n")
        for r in rows[2:]:
            f.write("----\n")
            rec = {}
            for w in WANT:
                if w in idx:
                    f.write(f"{w} = {r[idx[w]]} {units[idx[w]]}\n")
                    rec[w] = (r[idx[w]], units[idx[w]])
            res.append(rec)
    return res


def launches(path, out, title):
    if not os.path.exists(path):
        return
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}[row["Metric Unit"]]
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r"void |at::|native::|<unnamed>::", "", name)[:80]
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    ours = sum(v for k, v in tot.items() if "tdb" in k)
    with open(out, "w") as f:
        f.write(f"# {title}\n# total {T / 1e3:.2f} ms over {sum(cnt.values())} launches; share of this repo's kernels (tdb::*) {100 * ours / T:.1f} %, "
                f"{sum(c for k, c in cnt.items() if 'tdb' in k)} launches\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:70]:
            f.write(f"{v:9.1f} us {100 * v / T:5.1f}% n={cnt[k]:4d} avg {v / cnt[k]:8.1f} us  {k}\n")


if __name__ == "__main__":
    launches(os.path.join(G, "launches_bench_r02.csv"), os.path.join(P, "r02_launches_bench.txt"),
             "ncu --metrics gpu__time_duration.sum --clock-control none -s <warm-up> -c 2000 python bench.py --steps 2 --warmup 3 --skip-cpu "
             "--no-dedup-probe: consecutive kernel launches of the bench command itself inside the CUDA-graph replays of the timed region "
             "(graph nodes profiled individually; cold cache, serialised: compare SHARES, not absolutes)")
    res = raw(os.path.join(G, "prof_r02.ncu-rep"), os.path.join(P, "r02_ncu_kernels.txt"),
              "round-2 kernels alone at the bench shapes (tools/ncu_probe_r02.py): layer3 3x3 conv (tdb_gemm2), layer3 conv3 1x1 + residual "
              "(tdb_gemm<128,5>), fused stem, K/V projection of all decoder layers (two-tap GEMM), decoder attention core fwd / bwd")
    if res:
        for rec in res:
            if "tdb_gemm2_kernel" in rec.get("Kernel Name", ("",))[0]:
                def num(k):
                    v, u = rec[k]
                    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                tr = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
                head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
                json.dump({"dram_bytes_per_launch": tr, "source": "profiles/r02_ncu_kernels.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of "
                           "the first tdb_gemm2_kernel launch of tools/ncu_probe_r02.py: layer3 3x3 conv, 125 frames)", "captured_at_commit": head},
                          open(os.path.join(P, "r02_ncu_gemm2_traffic.json"), "w"))
                break
    raw(os.path.join(G, "prof_attn_tc_r02.ncu-rep"), os.path.join(P, "r02_ncu_attn_tc.txt"),
        "tcgen05 self-attention (tools/ncu_attn_tc.py): mha_tc_fwd_kernel / mha_tc_bwd_kernel at the encoder shape 25 x 141 x 8 heads, hash dropout on")
