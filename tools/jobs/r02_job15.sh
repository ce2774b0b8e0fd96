#!/bin/bash
mkdir -p gpurun_out
echo "== attn phase"; timeout 120 python tools/attn_phase.py 2>&1 | tail -24
echo "== full gpu suite"; timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench"; python bench.py --steps 10 --warmup 3 --skip-cpu 2>gpurun_out/bench_x.err | tail -1 > gpurun_out/bench_line.json; python -c "
import json; d=json.load(open('gpurun_out/bench_line.json'))
print(d['ms_per_step'], d['e2e'], d['gpu_launches']//d['steps'])
print(json.dumps(d['decoder_attn'])[:900])
print(json.dumps(d['roofline'])[:400])
"
