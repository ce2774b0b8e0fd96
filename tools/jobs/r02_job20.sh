#!/bin/bash
# pair-kernel TMA epilogue: parity, per-shape microbenchmark, step time with / without
mkdir -p gpurun_out
echo "== gemm tests"; timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -15
echo "== per-shape"; timeout 600 python tools/conv1x1_bench.py 2>&1 | tail -20
echo "== bench default"; timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu --no-dedup-probe 2>gpurun_out/bench_a.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['gpu_launches'], d['run'])"
echo "== bench TDB_GEMM2_TMA=0"; TDB_GEMM2_TMA=0 timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu --no-dedup-probe 2>gpurun_out/bench_b.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['gpu_launches'], d['run'])"
echo "== bench TDB_GEMM2_TMA_MIN_K=256"; TDB_GEMM2_TMA_MIN_K=256 timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu --no-dedup-probe 2>gpurun_out/bench_c.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['gpu_launches'], d['run'])"
tail -3 gpurun_out/bench_a.err
