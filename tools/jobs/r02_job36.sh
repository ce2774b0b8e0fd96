#!/bin/bash
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q -m gpu 2>&1 | tail -3
b() { echo "== bench $*"; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu --no-dedup-probe 2>gpurun_out/bench_a.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['gpu_launches'], d['run']['loss'], d['decoder_attn']['core_us_per_layer'], d['decoder_attn']['kv_proj_us_all_layers'])"; }
b A=1
b A=2
