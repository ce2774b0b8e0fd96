tools/run_gpu_tests.sh tests/test_model_gpu.py
echo "== bench side-stream"; python bench.py --steps 10 --warmup 3 --skip-cpu 2>&1 | tail -1 | cut -c1-250
