#!/bin/bash
mkdir -p gpurun_out
for v in "TDB_TEXT_PRIO=0" "TDB_TEXT_PRIO=-1" "TDB_WGRAD_STREAM=0"; do
  echo "== bench $v"; env $v python bench.py --steps 10 --warmup 3 --skip-cpu --no-dedup-probe 2>gpurun_out/bench_x.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['gpu_launches']//d['steps'])"
done
