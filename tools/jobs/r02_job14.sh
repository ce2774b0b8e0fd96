#!/bin/bash
# 2 GPUs: staged all-reduce correctness (dp_check, short timeout) after the double-claim fix
mkdir -p gpurun_out
export TDB_OFFLINE_TEXT_ENCODER=1
echo "== dp_check (staged)"; timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/dp_check.py 2>gpurun_out/dp_check.err | tee gpurun_out/dp_check.txt | tail -14
grep -v "Warn\|warn" gpurun_out/dp_check.err | tail -4 | cut -c1-200
