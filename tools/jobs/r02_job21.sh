#!/bin/bash
# pair-kernel box epilogue (TMA store / coalesced copy-out): parity, per-shape microbenchmark, step time A/B
mkdir -p gpurun_out
echo "== gemm tests"; timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -15
echo "== per-shape"; timeout 600 python tools/conv1x1_bench.py 2>&1 | tail -26
b() { echo "== bench $*"; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu --no-dedup-probe 2>gpurun_out/bench_a.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['gpu_launches'], d['run']['loss'], d['decoder_attn']['kv_proj_us_all_layers'])"; }
b A=1
b TDB_GEMM2_TMA=0 TDB_GEMM2_COPYOUT=0
b TDB_GEMM2_COPYOUT_HALO=0
b TDB_GEMM2_COPYOUT=0
tail -3 gpurun_out/bench_a.err
echo "== backbone + model tests"; timeout 900 python -m pytest tests/test_backbone_gpu.py tests/test_model_gpu.py -x -q 2>&1 | tail -5
