timeout 200 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "xattn" 2>&1 | tail -2
python tools/xattn_phase.py > gpurun_out/xattn_phase.txt 2>&1; cat gpurun_out/xattn_phase.txt
