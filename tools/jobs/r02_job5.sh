#!/bin/bash
# round 2, GPU job 5: fused stem
mkdir -p gpurun_out; rm -f gpurun_out/model_parity.txt
run() { n=$1; shift; echo "== $n"; timeout 900 "$@" > gpurun_out/$n.log 2>&1; echo "rc=$?"; tail -${TAILN:-6} gpurun_out/$n.log; }
TAILN=30 run t_stem python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "stem or ffn_hidden"
TAILN=30 run t_model python -m pytest tests/test_model_gpu.py tests/test_backbone_gpu.py -q -m gpu
grep "full size\|^cfg1:\|backbone feat" gpurun_out/model_parity.txt | grep -v "grad "
echo "== bench (default)"; python bench.py --steps 10 --warmup 3 --skip-cpu --no-dedup-probe 2>gpurun_out/bench_a.err | tail -1 | cut -c1-300
echo "== bench unfused stem"; TDB_STEM_FUSED=0 python bench.py --steps 10 --warmup 3 --skip-cpu --no-dedup-probe 2>gpurun_out/bench_b.err | tail -1 | cut -c1-300
echo "== bench two-pass backbone, L2 chunk 16"; TDB_JOINT=0 TDB_L2_CHUNK=16 python bench.py --steps 10 --warmup 3 --skip-cpu --no-dedup-probe 2>gpurun_out/bench_c.err | tail -1 | cut -c1-300
echo "== bench two-pass backbone, no chunk"; TDB_JOINT=0 python bench.py --steps 10 --warmup 3 --skip-cpu --no-dedup-probe 2>gpurun_out/bench_d.err | tail -1 | cut -c1-300
echo "== ablation"; timeout 600 python tools/step_ablation.py 2>gpurun_out/step_ablation.err | tee gpurun_out/step_ablation.txt
