#!/bin/bash
# 2 GPUs: does work stealing in the backbone backward (next to the staged gradient all-reduces) pay?  + NCCL CTA cap variant
mkdir -p gpurun_out
export TDB_OFFLINE_TEXT_ENCODER=1
b() { echo "== bench N=2 $*"; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 --skip-cpu --no-dedup-probe 2>gpurun_out/bench_n2.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['run'])"; }
b TDB_CLC_BWD=0
b TDB_CLC_BWD=1
b TDB_CLC_BWD=0 NCCL_MAX_CTAS=8
b TDB_CLC_BWD=1 TDB_CLC_FWD=1
grep -v "Warn\|warn\|_make_text\|run_backward" gpurun_out/bench_n2.err | tail -4 | cut -c1-200
