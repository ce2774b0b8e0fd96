tools/run_gpu_tests.sh tests/test_gemm_gpu.py tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_backbone_gpu.py
echo "== bench PDL=0"; TDB_PDL=0 python bench.py --steps 10 --warmup 3 --skip-cpu 2>&1 | tail -1 | cut -c1-250
echo "== bench PDL=1"; python bench.py --steps 10 --warmup 3 --skip-cpu 2>&1 | tail -1 | cut -c1-250
