#!/bin/bash
mkdir -p gpurun_out
for v in "TDB_GEMM2_MIN_K=256" "TDB_GEMM2_MIN_K=128" ""; do
  echo "== bench $v"; env $v python bench.py --steps 10 --warmup 3 --skip-cpu --no-dedup-probe 2>gpurun_out/bench_x.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['gpu_launches']//d['steps'], d['run']['loss'])"
done
echo "== ablation MIN_K=256"; TDB_GEMM2_MIN_K=256 timeout 600 python tools/step_ablation.py 2>/dev/null | sed -n 9,17p
