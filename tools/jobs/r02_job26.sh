#!/bin/bash
# 8 GPUs: the staged two-communicator gradient all-reduce at full node scale (hang check + scaling number)
mkdir -p gpurun_out
export TDB_OFFLINE_TEXT_ENCODER=1
echo "== bench N=8"; timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus 8 --steps 10 --warmup 3 --skip-cpu --no-dedup-probe 2>gpurun_out/bench_n8.err | tail -1 > gpurun_out/bench_n8.json; echo "rc=$?"; cut -c1-700 gpurun_out/bench_n8.json
grep -v "Warn\|warn\|_make_text\|run_backward\|OMP_NUM\|\*\*\*\*" gpurun_out/bench_n8.err | tail -6 | cut -c1-200
