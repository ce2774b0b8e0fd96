#!/bin/bash
# run on the GPU box: all -m gpu tests with per-file logs under gpurun_out/
mkdir -p gpurun_out; rm -f gpurun_out/model_parity.txt
for f in "$@"; do
  b=$(basename $f .py)
  timeout 900 python -m pytest $f -x -q -m gpu 2>&1 | tail -60 > gpurun_out/$b.log
  echo "== $f"; tail -25 gpurun_out/$b.log
done
