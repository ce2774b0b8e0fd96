#!/bin/bash
# 2 GPUs, final code: staged all-reduce correctness (dp_check) + bench
mkdir -p gpurun_out
export TDB_OFFLINE_TEXT_ENCODER=1
echo "== dp_check (staged)"; timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/dp_check.py 2>gpurun_out/dp_check.err | tee gpurun_out/dp_check.txt | tail -8
echo "== bench N=2"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 --skip-cpu --no-dedup-probe 2>gpurun_out/bench_n2.err | tail -1 > gpurun_out/bench_n2.json; python -c "import json; d=json.load(open('gpurun_out/bench_n2.json')); print(d['value'], d['ms_per_step'], d['run'])"
grep -v "Warn\|warn\|_make_text\|run_backward\|OMP_NUM\|\*\*\*\*" gpurun_out/bench_n2.err | tail -3 | cut -c1-200
