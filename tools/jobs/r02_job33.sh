#!/bin/bash
mkdir -p gpurun_out
b() { echo "== bench $*"; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu --no-dedup-probe 2>gpurun_out/bench_a.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['gpu_launches'], d['run']['loss'])"; }
b A=1
b TDB_TIMING_SKIP_REDUCE=1
b A=2
b TDB_TIMING_SKIP_REDUCE=1 A=3
b TDB_WGRAD_LAG=2
