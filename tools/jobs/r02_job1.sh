#!/bin/bash
# round 2, GPU job 1: baseline on this round's box -- all GPU tests (tcgen05 attention tests in their own processes), bench A/B
mkdir -p gpurun_out; rm -f gpurun_out/model_parity.txt
export TDB_OFFLINE_TEXT_ENCODER=1
run() { n=$1; shift; echo "== $n"; timeout 900 "$@" > gpurun_out/$n.log 2>&1; echo "rc=$?"; tail -${TAILN:-6} gpurun_out/$n.log; }
run t_tc_fwd python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "tc_forward"
TAILN=30 run t_tc_bwd python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "tc_backward"
run t_kernels python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "not tc_"
run t_gemm python -m pytest tests/test_gemm_gpu.py tests/test_backbone_gpu.py tests/test_optim.py -q -m gpu
TAILN=40 run t_model python -m pytest tests/test_model_gpu.py -q -m gpu
run t_full python -m pytest tests/test_fullsize_gpu.py -q -m gpu
cat gpurun_out/model_parity.txt
echo "== bench tc=2 (default)"; python bench.py --steps 10 --warmup 3 --skip-cpu --no-dedup-probe 2>gpurun_out/bench_a.err | tail -1 | cut -c1-400
echo "== bench tc=0"; TDB_MHA_TC=0 python bench.py --steps 10 --warmup 3 --skip-cpu --no-dedup-probe 2>gpurun_out/bench_b.err | tail -1 | cut -c1-300
echo "== bench tc=1"; TDB_MHA_TC=1 python bench.py --steps 10 --warmup 3 --skip-cpu --no-dedup-probe 2>gpurun_out/bench_c.err | tail -1 | cut -c1-300
tail -3 gpurun_out/bench_a.err
