nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
tools/run_gpu_tests.sh tests/test_kernels_gpu.py tests/test_gemm_gpu.py tests/test_optim.py tests/test_fullsize_gpu.py tests/test_model_gpu.py tests/test_backbone_gpu.py
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --skip-cpu 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/bench_r01d.json; cat gpurun_out/bench_r01d.json | cut -c1-900; tail -3 gpurun_out/bench_err.log
python tools/step_ablation.py > gpurun_out/step_ablation.txt 2> gpurun_out/step_ablation.err; cat gpurun_out/step_ablation.txt; tail -3 gpurun_out/step_ablation.err
