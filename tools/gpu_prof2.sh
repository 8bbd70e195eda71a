#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k "regex:gemm_tn_kernel" -s 9 -c 4 -o gpurun_out/prof_gemm_r1 -f python tools/profile_decode.py 1 > gpurun_out/prof_gemm.log 2>&1
echo "ncu gemm exit $?"
timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k "regex:attention_tc" -s 5 -c 1 -o gpurun_out/prof_attn_r1 -f python tools/profile_decode.py 1 > gpurun_out/prof_attn2.log 2>&1
echo "ncu attn exit $?"
ls -la gpurun_out/*.ncu-rep
