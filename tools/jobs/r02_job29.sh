#!/bin/bash
# first block of a stage: conv3 + downsample as one GEMM over [y2 | xs]: parity + step time A/B
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_backbone_gpu.py tests/test_model_gpu.py tests/test_fullsize_gpu.py tests/test_kernels_gpu.py tests/test_engine_contract_gpu.py -x -q -m gpu 2>&1 | tail -6
b() { echo "== bench $*"; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu --no-dedup-probe 2>gpurun_out/bench_a.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['gpu_launches'], d['run']['loss'])"; }
b A=1
b TDB_DS_JOINT=0
b A=2
b TDB_DS_JOINT=0 A=3
grep -v "Warn\|warn\|_make_text\|run_backward" gpurun_out/bench_a.err | tail -3
echo "== ablation"; timeout 600 python tools/step_ablation.py 2>gpurun_out/step_ablation.err | grep "backbone\|full step, train mode (bench)"
grep -v "Warn\|warn\|_make_text\|run_backward" gpurun_out/step_ablation.err | tail -3
