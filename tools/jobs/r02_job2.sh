#!/bin/bash
# round 2, GPU job 2: rewritten tcgen05 attention (per-head CTAs, coalesced P traffic, in-kernel hash dropout)
mkdir -p gpurun_out; rm -f gpurun_out/model_parity.txt
run() { n=$1; shift; echo "== $n"; timeout 900 "$@" > gpurun_out/$n.log 2>&1; echo "rc=$?"; tail -${TAILN:-6} gpurun_out/$n.log; }
TAILN=25 run t_tc_fwd python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "tc_forward"
TAILN=25 run t_tc_bwd python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "tc_backward"
run t_kernels python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "not tc_"
run t_gemm python -m pytest tests/test_gemm_gpu.py tests/test_backbone_gpu.py tests/test_optim.py -q -m gpu
TAILN=40 run t_model python -m pytest tests/test_model_gpu.py -q -m gpu
run t_full python -m pytest tests/test_fullsize_gpu.py -q -m gpu
grep "full size\|^cfg1:" gpurun_out/model_parity.txt | grep -v grad
echo "== attn bench"; timeout 300 python tools/attn_bench.py 2>&1 | tee gpurun_out/attn_bench.txt
echo "== bench tc=2 (default)"; python bench.py --steps 10 --warmup 3 --skip-cpu --no-dedup-probe 2>gpurun_out/bench_a.err | tail -1 | cut -c1-300
echo "== bench tc=0"; TDB_MHA_TC=0 python bench.py --steps 10 --warmup 3 --skip-cpu --no-dedup-probe 2>gpurun_out/bench_b.err | tail -1 | cut -c1-300
