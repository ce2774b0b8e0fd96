nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
tools/run_gpu_tests.sh tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_fullsize_gpu.py tests/test_backbone_gpu.py
python bench.py --steps 10 --warmup 3 --skip-cpu 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/bench_r01f.json; python -c "
import json; d=json.load(open('gpurun_out/bench_r01f.json')); print('ms', d['ms_per_step'], 'value', d['value'], 'launches', d['gpu_launches']); print('dedup', d.get('dedup_slow_frames')); print('e2e', d['e2e'])"; grep -v Warn gpurun_out/bench_err.log | tail -3
TDB_WGRAD_STREAM=0 python bench.py --steps 10 --warmup 3 --skip-cpu --no-dedup-probe 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('no wgrad stream: ms', d['ms_per_step'])"
python tools/step_ablation.py 2> gpurun_out/step_ablation.err | head -7 > gpurun_out/step_ablation2.txt; cat gpurun_out/step_ablation2.txt
