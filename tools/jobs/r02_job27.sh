#!/bin/bash
# round-2 closing job: full GPU suite, smoke, ablation, ncu captures + launch list of the bench command, bench line with the CPU arm
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== ablation"; timeout 600 python tools/step_ablation.py 2>gpurun_out/step_ablation.err > gpurun_out/step_ablation.txt; cat gpurun_out/step_ablation.txt
echo "== ncu kernels"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tdb_gemm|stem_fused|xattn_core|xattn_bwd" -c 12 -o gpurun_out/prof_r02 -f python tools/ncu_probe_r02.py > gpurun_out/ncu_r02.log 2>&1; tail -2 gpurun_out/ncu_r02.log
echo "== ncu launch list of the bench command"; timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 2000 --csv --log-file gpurun_out/launches_bench_r02.csv python bench.py --steps 2 --warmup 3 --skip-cpu --no-dedup-probe > gpurun_out/ncu_bench_r02.log 2>&1; tail -1 gpurun_out/ncu_bench_r02.log | cut -c1-200; wc -l gpurun_out/launches_bench_r02.csv
echo "== bench"; python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_x.err | tail -1 > gpurun_out/bench_line_r02.json; cut -c1-300 gpurun_out/bench_line_r02.json
echo "== bench reference arm"; TDB_REF_BUDGET_S=60 timeout 400 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
