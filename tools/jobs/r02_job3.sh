#!/bin/bash
# round 2, GPU job 3: attention microbench under CUDA graphs, step ablation with finer variants
mkdir -p gpurun_out
run() { n=$1; shift; echo "== $n"; timeout 900 "$@" > gpurun_out/$n.log 2>&1; echo "rc=$?"; tail -${TAILN:-6} gpurun_out/$n.log; }
TAILN=15 run t_tc python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "tc_"
run t_bb python -m pytest tests/test_backbone_gpu.py -q -m gpu
echo "== attn bench"; timeout 300 python tools/attn_bench.py 2>&1 | tee gpurun_out/attn_bench.txt
echo "== ablation"; timeout 600 python tools/step_ablation.py 2>gpurun_out/step_ablation.err | tee gpurun_out/step_ablation.txt
echo "== bench tc=2 (default)"; python bench.py --steps 10 --warmup 3 --skip-cpu --no-dedup-probe 2>gpurun_out/bench_a.err | tail -1 | cut -c1-300
