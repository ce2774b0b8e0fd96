#!/bin/bash
# round 2 profiling job: ncu --set full of the round's kernels, launch list of the bench command, ablation, bench line with CPU baseline
mkdir -p gpurun_out
echo "== ncu kernels"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tdb_gemm|stem_fused|xattn_core|xattn_bwd" -c 12 -o gpurun_out/prof_r02 -f python tools/ncu_probe_r02.py > gpurun_out/ncu_r02.log 2>&1; tail -2 gpurun_out/ncu_r02.log
echo "== ncu launch list of the bench command"; timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 2000 --csv --log-file gpurun_out/launches_bench_r02.csv python bench.py --steps 2 --warmup 3 --skip-cpu --no-dedup-probe > gpurun_out/ncu_bench_r02.log 2>&1; tail -1 gpurun_out/ncu_bench_r02.log | cut -c1-200; wc -l gpurun_out/launches_bench_r02.csv
echo "== ablation"; timeout 600 python tools/step_ablation.py 2>gpurun_out/step_ablation.err > gpurun_out/step_ablation.txt; cat gpurun_out/step_ablation.txt
echo "== bench"; python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_x.err | tail -1 > gpurun_out/bench_line_r02.json; cut -c1-300 gpurun_out/bench_line_r02.json
