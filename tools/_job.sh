tools/run_gpu_tests.sh tests/test_model_gpu.py tests/test_fullsize_gpu.py tests/test_backbone_gpu.py
python bench.py --steps 10 --warmup 3 --skip-cpu 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/bench_r01g.json; python -c "
import json; d=json.load(open('gpurun_out/bench_r01g.json')); print('ms', d['ms_per_step'], 'value', d['value'], 'launches', d['gpu_launches']); print('dedup', d.get('dedup_slow_frames')); print('e2e', d['e2e'])"; grep -v Warn gpurun_out/bench_err.log | tail -3
python tools/step_ablation.py 2> gpurun_out/step_ablation.err | head -7 > gpurun_out/step_ablation2.txt; cat gpurun_out/step_ablation2.txt
