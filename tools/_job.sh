tools/run_gpu_tests.sh tests/test_gemm_gpu.py tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_backbone_gpu.py
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_final.json; cat gpurun_out/bench_final.json | cut -c1-1800
ncu --set full --clock-control none --import-source on -k regex:tdb_gemm -o gpurun_out/prof_gemm2_r01b python tools/ncu_probe.py > gpurun_out/ncu_probe.log 2>&1; tail -2 gpurun_out/ncu_probe.log
