#!/bin/bash
mkdir -p gpurun_out
run() { n=$1; shift; echo "== $n"; timeout 900 "$@" > gpurun_out/$n.log 2>&1; echo "rc=$?"; tail -${TAILN:-6} gpurun_out/$n.log; }
TAILN=20 run t_k python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "criterion"
echo "== timeline"; timeout 600 python tools/step_timeline.py gpurun_out/timeline.txt 2>&1 | tail -3
