#!/bin/bash
# shared-halo padded grid: parity + step time A/B
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/ -q -m gpu 2>&1 | tail -8
b() { echo "== bench $*"; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu --no-dedup-probe 2>gpurun_out/bench_a.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['gpu_launches'], d['run']['loss'], d['roofline']['frac'])"; }
b A=1
b TDB_HALO1=0
b A=2
b TDB_HALO1=0 A=3
grep -v "Warn\|warn\|_make_text\|run_backward" gpurun_out/bench_a.err | tail -3
