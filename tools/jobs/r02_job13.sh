#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/model_parity.txt
run() { n=$1; shift; echo "== $n"; timeout 900 "$@" > gpurun_out/$n.log 2>&1; echo "rc=$?"; tail -${TAILN:-6} gpurun_out/$n.log; }
TAILN=40 run t_text python -m pytest tests/test_text_gpu.py -q -m gpu
TAILN=30 run t_eng python -m pytest tests/test_engine_contract_gpu.py -q -m gpu
TAILN=40 run t_model python -m pytest tests/test_model_gpu.py tests/test_fullsize_gpu.py -q -m gpu
grep "full size\|^cfg1:" gpurun_out/model_parity.txt | grep -v "grad "
for v in "" "TDB_OWN_TEXT=0"; do
  echo "== bench $v"; env $v python bench.py --steps 10 --warmup 3 --skip-cpu --no-dedup-probe 2>gpurun_out/bench_x.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['gpu_launches']//d['steps'], d['run'])"
done
tail -3 gpurun_out/bench_x.err
