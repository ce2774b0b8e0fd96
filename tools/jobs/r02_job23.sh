#!/bin/bash
# cluster launch control (work stealing) in both GEMM kernels: parity, behaviour under SM contention, step time A/B
mkdir -p gpurun_out
echo "== gemm tests"; timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -15
echo "== hog"; timeout 300 python tools/clc_hog.py 2>&1 | tail -8
echo "== per-shape"; timeout 600 python tools/conv1x1_bench.py 2>&1 | tail -24
b() { echo "== bench $*"; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu --no-dedup-probe 2>gpurun_out/bench_a.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['gpu_launches'], d['run']['loss'], d['decoder_attn']['kv_proj_us_all_layers'], d['roofline']['frac'])"; }
b A=1
b TDB_CLC=0
b A=2
tail -3 gpurun_out/bench_a.err
echo "== backbone + model tests"; timeout 900 python -m pytest tests/test_backbone_gpu.py tests/test_model_gpu.py -x -q 2>&1 | tail -5
