#!/bin/bash
mkdir -p gpurun_out
run() { n=$1; shift; echo "== $n"; timeout 900 "$@" > gpurun_out/$n.log 2>&1; echo "rc=$?"; tail -${TAILN:-6} gpurun_out/$n.log; }
TAILN=20 run t_pre python -m pytest tests/test_kernels_gpu.py tests/test_text_gpu.py -q -m gpu -k "preprocess or text"
for v in "TDB_TEXT_SIDE=0" ""; do
  echo "== bench $v"; env $v python bench.py --steps 10 --warmup 3 --skip-cpu --no-dedup-probe 2>gpurun_out/bench_x.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['gpu_launches']//d['steps'])"
done
