#!/bin/bash
mkdir -p gpurun_out
K='regex:k_qkv|k_attn|k_res|k_gateup|k_head|k_sample|k_set_state'
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 600 -c 300 --csv --log-file gpurun_out/launches_decode_r1.csv python tools/profile_decode.py 3 > gpurun_out/prof.log 2>&1
echo "ncu list exit $?"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k "regex:k_gateup|k_head|k_res" -s 400 -c 4 -o gpurun_out/prof_decode_r1 -f python tools/profile_decode.py 3 > gpurun_out/prof2.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out/
