#!/bin/bash
# final code of the round: full GPU suite, smoke, bench line with the CPU arm
mkdir -p gpurun_out
echo "== tests"; timeout 600 python -m pytest tests/ -q -m gpu 2>&1 | tail -3
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench"; timeout 400 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_x.err | tail -1 > gpurun_out/bench_line_r02.json; cut -c1-260 gpurun_out/bench_line_r02.json
