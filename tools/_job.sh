timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "not pair_variant" 2>&1 | tail -3
timeout 200 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "pair_variant" 2>&1 | tail -15
timeout 200 python tools/xattn_phase.py 2>&1 | tee gpurun_out/xattn_phase.txt | tail -20
