tools/run_gpu_tests.sh tests/test_gemm_gpu.py tests/test_kernels_gpu.py
timeout 300 python tools/halo_diag.py 2>&1 | grep "median"
echo "== bench"; python bench.py --steps 10 --warmup 3 --skip-cpu 2>&1 | tail -1 | cut -c1-330
python tools/gemm_bench.py 2>&1 | tail -15 | cut -c1-150
