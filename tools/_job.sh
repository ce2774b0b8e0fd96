nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
tools/run_gpu_tests.sh tests/test_optim.py tests/test_fullsize_gpu.py tests/test_model_gpu.py
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/bench_r01c.json; cat gpurun_out/bench_r01c.json | cut -c1-2500; tail -3 gpurun_out/bench_err.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --profile-from-start off --csv --log-file gpurun_out/step_metrics.csv python tools/ncu_step.py 2 > gpurun_out/ncu_step.log 2>&1; tail -2 gpurun_out/ncu_step.log; wc -l gpurun_out/step_metrics.csv
