tools/run_gpu_tests.sh tests/test_model_gpu.py tests/test_backbone_gpu.py
for c in 0 10 20; do echo "== L2_CHUNK=$c"; TDB_L2_CHUNK=$c python bench.py --steps 10 --warmup 3 --skip-cpu 2>&1 | tail -1 | cut -c1-400; done
ncu --set full --clock-control none --import-source on -k regex:tdb_gemm -o gpurun_out/prof_gemm2_r01 python tools/ncu_probe.py > gpurun_out/ncu_probe.log 2>&1; tail -3 gpurun_out/ncu_probe.log
