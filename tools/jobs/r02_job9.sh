#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mha_tc -s 4 -c 2 -o gpurun_out/prof_attn_tc_r02 -f python tools/ncu_attn_tc.py > gpurun_out/ncu_attn_tc.log 2>&1
tail -3 gpurun_out/ncu_attn_tc.log; ls -la gpurun_out/*.ncu-rep | tail -2
