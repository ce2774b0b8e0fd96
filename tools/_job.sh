nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
tools/run_gpu_tests.sh tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_fullsize_gpu.py tests/test_backbone_gpu.py
python bench.py --steps 10 --warmup 3 --skip-cpu 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/bench_r01e.json; cat gpurun_out/bench_r01e.json | cut -c1-600; python -c "
import json; d=json.load(open('gpurun_out/bench_r01e.json')); print('dedup', d.get('dedup_slow_frames')); print('e2e', d['e2e']); print('optim', d['optimizer_step'].get('ms'))"; tail -3 gpurun_out/bench_err.log
python tools/step_ablation.py > gpurun_out/step_ablation.txt 2> gpurun_out/step_ablation.err; cat gpurun_out/step_ablation.txt; grep -v Warn gpurun_out/step_ablation.err | tail -3
