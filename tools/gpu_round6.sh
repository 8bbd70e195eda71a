#!/bin/bash
mkdir -p gpurun_out
K='regex:k_stream|k_attn|k_sample'
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 600 -c 300 --csv --log-file gpurun_out/launches_decode_r3.csv python tools/profile_decode.py 3 > gpurun_out/prof.log 2>&1
python tools/agg_launches.py gpurun_out/launches_decode_r3.csv
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k "regex:k_stream" -s 401 -c 5 -o gpurun_out/prof_decode_r3 -f python tools/profile_decode.py 3 > gpurun_out/prof2.log 2>&1
echo "ncu full exit $?"
