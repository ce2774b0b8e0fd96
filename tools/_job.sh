nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
tools/run_gpu_tests.sh tests/test_gemm_gpu.py tests/test_kernels_gpu.py tests/test_optim.py tests/test_backbone_gpu.py tests/test_model_gpu.py tests/test_fullsize_gpu.py
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/bench_final.json; cut -c1-400 gpurun_out/bench_final.json; grep -v "Warn\|warn" gpurun_out/bench_err.log | tail -3
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_reference.json; cut -c1-300 gpurun_out/bench_reference.json
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --profile-from-start off --csv --log-file gpurun_out/step_metrics.csv python tools/ncu_step.py 2 > gpurun_out/ncu_step.log 2>&1; tail -1 gpurun_out/ncu_step.log; wc -l gpurun_out/step_metrics.csv
timeout 200 ncu --set full --clock-control none --import-source on -k regex:tdb_gemm -o gpurun_out/prof_gemm_r01c python tools/ncu_probe.py > gpurun_out/ncu_probe.log 2>&1; tail -2 gpurun_out/ncu_probe.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:xattn_fused -s 4 -c 2 -o gpurun_out/prof_xattn_r01c python tools/ncu_xattn.py > gpurun_out/ncu_xattn.log 2>&1; tail -1 gpurun_out/ncu_xattn.log
timeout 200 ncu --set full --clock-control none --cache-control none --import-source on -k regex:xattn_fused -s 4 -c 2 -o gpurun_out/prof_xattn_warm_r01c python tools/ncu_xattn.py > gpurun_out/ncu_xattn2.log 2>&1; tail -1 gpurun_out/ncu_xattn2.log
timeout 200 ncu --set full --clock-control none --import-source on -k "regex:mha_|xattn_bwd" -s 12 -c 7 -o gpurun_out/prof_attn_r01c python tools/ncu_attn.py > gpurun_out/ncu_attn.log 2>&1; tail -1 gpurun_out/ncu_attn.log
python tools/xattn_phase.py > gpurun_out/xattn_phase.txt 2>&1; cat gpurun_out/xattn_phase.txt | head -8
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 11500 -c 2300 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --skip-cpu --no-dedup-probe > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log | cut -c1-200; wc -l gpurun_out/launches_bench.csv
ls -la gpurun_out/*.ncu-rep
