#!/bin/bash
# 2 GPUs: staged all-reduce correctness (dp_check) + N=2 bench (staged / unstaged)
mkdir -p gpurun_out
export TDB_OFFLINE_TEXT_ENCODER=1
echo "== dp_check (staged)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/dp_check.py 2>gpurun_out/dp_check.err | tee gpurun_out/dp_check.txt | tail -6
tail -5 gpurun_out/dp_check.err
for v in "" "TDB_STAGED_AR=0"; do
  echo "== bench N=2 $v"; env $v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/bench_n2.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['run'])"
done
tail -3 gpurun_out/bench_n2.err
